// emm_api.cu -- host side of the C ABI declared in include/enzymm_b200.h.
//
// Owns device memory (plain cudaMalloc, no torch types), lays structure blobs out, launches the
// prepare and search kernels on the caller's stream and brings hits back sorted.  There is no
// CPU fallback anywhere: without a CUDA device every entry point fails with EMM_ERR_NO_DEVICE.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <thread>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "emm_device.cuh"

namespace emm {
void launch_prepare(const DevLibrary &L, const DevBatch &B, float cutoff, bool build_cells,
                    unsigned long long *stats, unsigned long long *bad, int sm_count, cudaStream_t stream);
size_t search_smem_bytes(int blob_cap, int levels, bool cells);
size_t search_fixed_smem(int levels);
size_t search_soft_cap(int levels);
cudaError_t configure_search(int smem_bytes);
void launch_skip_snapshot(int n, int mode, const int *any, const int *pass, unsigned char *skip, cudaStream_t stream);
void launch_search(const DevLibrary &L, const DevBatch &B, const SearchParams &P, const SearchOut &O,
                   const unsigned char *skip, const int4 *sched, const int *ids, bool stats, bool staged, int grid,
                   size_t smem, cudaStream_t stream);
}  // namespace emm

using namespace emm;

static thread_local std::string g_error;

// Largest dynamic shared-memory size the search kernels have been opted into, per device (the
// attribute is per function and device, shared by every library handle of the process).
static size_t g_configured_smem[64] = {0};
static std::mutex g_config_mutex;      // sessions of one device may be driven from different host threads

static int fail(emm_status st, const std::string &msg)
{
    g_error = msg;
    return (int)st;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return fail(EMM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
    } while (0)

constexpr int kHeavyAtoms = 15;       // templates with at least this many atoms (5+ residues) form the first phase
constexpr int kSplitBelow = 8192;     // launches with fewer structures split expensive pairs over idle warps by default

struct emm_library {
    int device = 0;
    int sm_count = 0;
    int smem_optin = 0;
    int compat_version = 0;
    std::vector<uint16_t> class_leaders;   // per typing class: leader lists it appears in
    int stats_enabled = 0;
    DevLibrary d{};
    std::vector<void *> allocs;
    std::vector<uint16_t> h_leader_ttype;
    uint32_t *d_compat = nullptr;
    uint32_t *d_class_mask = nullptr;
    double *d_rmsd = nullptr, *d_cut = nullptr, *d_dyn = nullptr, *d_lr_table = nullptr;
    int32_t *d_lr_index = nullptr;
    int lr_capacity = 0;
    // visiting order of the templates of a range [tb, te): most expensive first
    struct Sched { int4 *d_ids = nullptr; int n = 0, n_heavy = 0; };
    std::vector<int> h_tpl_atoms, h_atom_off;
    std::vector<int64_t> h_pair_off;
    std::vector<double> h_cut;
    std::map<std::pair<int, int>, Sched> sched;
    std::mutex sched_mutex;
};

struct emm_session {
    emm_library *lib = nullptr;
    int64_t max_atoms = 0;
    int32_t max_structures = 0;
    int64_t hit_capacity = 0;
    // device columns
    int64_t *d_atom_off = nullptr;
    double *d_xyz = nullptr;
    uint16_t *d_klass = nullptr;
    int32_t *d_residue = nullptr;
    float *d_bfactor = nullptr;
    uint16_t *d_chain = nullptr;
    int32_t *d_atom_id = nullptr;
    unsigned char *d_blob = nullptr;
    int64_t blob_capacity = 0;
    int64_t *d_blob_off = nullptr;
    emm_hit *d_hits = nullptr;
    unsigned long long *d_hit_count = nullptr;
    unsigned int *d_work = nullptr;
    int *d_any = nullptr, *d_pass = nullptr;
    unsigned char *d_skip = nullptr;
    int *d_ids = nullptr;          // structures that can be staged first, then the ones too large for shared memory
    int32_t *d_status = nullptr;   // per structure: BlobHeader.status written by the prepare kernel
    int32_t *d_kept_bound = nullptr;   // per structure: atoms of a class other than 0 (blob sizing)
    std::vector<int32_t> h_kept_bound;
    unsigned long long *d_stats = nullptr;
    // current batch
    int32_t n_structures = 0;
    int64_t n_atoms = 0;
    bool has_bfactor = false, has_chain = false, has_atom_id = false;
    bool prepared = false;
    float prepared_cutoff = -1.f;
    bool prepared_cells = false;    // the uniform grid was built (only runs with cell_threshold > 0 need it)
    int prepared_version = -1;
    std::vector<int64_t> h_blob_off;
    int64_t max_staged = 0;        // largest staged blob prefix among the stageable structures (upper bound)
    int32_t n_small = 0;           // structures searched from shared memory; the rest are read in place
    std::vector<int> h_ids;
    int last_launches = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_prepare, ev_search;
    size_t ev_prepare_used = 0, ev_search_used = 0;
};

template <typename T>
static int dev_copy(emm_library *lib, const T *host, size_t count, T **out)
{
    *out = nullptr;
    if (count == 0) count = 1;
    void *p = nullptr;
    CUDA_TRY(cudaMalloc(&p, count * sizeof(T)));
    lib->allocs.push_back(p);
    if (host) CUDA_TRY(cudaMemcpy(p, host, count * sizeof(T), cudaMemcpyHostToDevice));
    else CUDA_TRY(cudaMemset(p, 0, count * sizeof(T)));
    *out = reinterpret_cast<T *>(p);
    return EMM_OK;
}

// Per typing class: how many leader lists it appears in (blob sizing) and which (class_mask, for
// the prepare kernel).  Call with the device idle: it rewrites a table the kernels read.
static int compute_class_leaders(emm_library *lib, int class_words, const uint32_t *compat)
{
    const size_t n_class = (size_t)lib->d.class_words_cap * 32, mw = (size_t)lib->d.mask_words;
    lib->class_leaders.assign(n_class, 0);
    std::vector<uint32_t> mask(n_class * mw, 0u);
    for (size_t l = 0; l < lib->h_leader_ttype.size(); ++l) {
        const uint32_t *row = compat + (size_t)lib->h_leader_ttype[l] * class_words;
        for (int w = 0; w < class_words; ++w) {
            uint32_t bits = row[w];
            while (bits) {
                const int b = __builtin_ctz(bits);
                const size_t c = (size_t)w * 32 + b;
                lib->class_leaders[c]++;
                mask[c * mw + (l >> 5)] |= 1u << (l & 31);
                bits &= bits - 1;
            }
        }
    }
    CUDA_TRY(cudaMemcpy(lib->d_class_mask, mask.data(), mask.size() * 4, cudaMemcpyHostToDevice));
    lib->compat_version++;
    return EMM_OK;
}

static bool next_events(std::vector<std::pair<cudaEvent_t, cudaEvent_t>> &pool, size_t &used,
                        cudaEvent_t *a, cudaEvent_t *b)
{
    if (used >= 4096) return false;             // bounded: timings are a diagnostic, not a log
    if (used == pool.size()) {
        cudaEvent_t x, y;
        if (cudaEventCreate(&x) != cudaSuccess || cudaEventCreate(&y) != cudaSuccess) return false;
        pool.emplace_back(x, y);
    }
    *a = pool[used].first;
    *b = pool[used].second;
    ++used;
    return true;
}

extern "C" {

int emm_abi_version(void) { return EMM_ABI_VERSION; }

int emm_hit_size(void) { return (int)sizeof(emm_hit); }

const char *emm_last_error(void) { return g_error.c_str(); }

int emm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int emm_stream_create(int device, void **stream)
{
    if (!stream) return fail(EMM_ERR_INVALID, "stream is null");
    *stream = nullptr;
    const int ndev = emm_device_count();
    if (ndev <= 0) return fail(EMM_ERR_NO_DEVICE, "no CUDA device visible: enzymm_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(EMM_ERR_INVALID, "device index out of range");
    CUDA_TRY(cudaSetDevice(device));
    cudaStream_t st;
    CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    *stream = (void *)st;
    return EMM_OK;
}

int emm_stream_destroy(int device, void *stream)
{
    if (!stream) return EMM_OK;
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaStreamDestroy((cudaStream_t)stream));
    return EMM_OK;
}

int emm_library_create(int device, const emm_library_desc *desc, emm_library **out)
{
    if (!out) return fail(EMM_ERR_INVALID, "out is null");
    *out = nullptr;
    if (!desc) return fail(EMM_ERR_INVALID, "desc is null");
    int ndev = emm_device_count();
    if (ndev <= 0) return fail(EMM_ERR_NO_DEVICE, "no CUDA device visible: enzymm_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(EMM_ERR_INVALID, "device index out of range");
    if (desc->n_templates <= 0 || desc->n_atoms <= 0) return fail(EMM_ERR_INVALID, "empty template library");
    if (desc->n_leader <= 0 || desc->n_leader > 1023) return fail(EMM_ERR_INVALID, "n_leader must be in 1..1023");
    if (desc->class_words <= 0 || desc->n_ttype <= 0) return fail(EMM_ERR_INVALID, "empty compat matrix");
    if (desc->class_words > (1 << kClassBits) / 32) return fail(EMM_ERR_INVALID, "more than 1024 typing classes");

    // validate the search plans on the host: the kernels trust them
    int max_m = 0;
    for (int t = 0; t < desc->n_templates; ++t) {
        const int a0 = desc->atom_off[t], m = desc->atom_off[t + 1] - a0;
        if (m <= 0 || m > EMM_MAX_TEMPLATE_ATOMS) return fail(EMM_ERR_INVALID, "template atom count out of range (1..32)");
        max_m = std::max(max_m, m);
        if (desc->pair_off[t + 1] - desc->pair_off[t] != (int64_t)m * (m - 1) / 2)
            return fail(EMM_ERR_INVALID, "pair_off does not match template size");
        unsigned seen = 0;
        for (int k = 0; k < m; ++k) {
            const int src = desc->plan_src[a0 + k];
            if (k == 0 && src >= 0) return fail(EMM_ERR_INVALID, "plan position 0 must be a leader");
            if (src >= k) return fail(EMM_ERR_INVALID, "plan_src must refer to an earlier position");
            if (src < 0 && -1 - src >= desc->n_leader) return fail(EMM_ERR_INVALID, "leader list index out of range");
            if (desc->plan_ttype[a0 + k] >= desc->n_ttype) return fail(EMM_ERR_INVALID, "plan_ttype out of range");
            const int an = desc->plan_anchor[a0 + k];
            if (k > 0 && (an >= k || (src >= 0 && an != src))) return fail(EMM_ERR_INVALID, "plan_anchor must be an earlier position (the leader for same-residue positions)");
            const int pa = desc->plan_atom[a0 + k];
            if (pa >= m || (seen >> pa) & 1u) return fail(EMM_ERR_INVALID, "plan_atom is not a permutation");
            seen |= 1u << pa;
        }
        if (desc->n_residues[t] < 0 || desc->n_residues[t] > EMM_MAX_RESIDUES || desc->n_residues[t] * 3 > m)
            return fail(EMM_ERR_INVALID, "n_residues out of range");
        if (desc->lr_index[t] >= desc->n_lr || desc->lr_index[t] < -2) return fail(EMM_ERR_INVALID, "lr_index out of range");
    }
    for (int l = 0; l < desc->n_leader; ++l)
        if (desc->leader_ttype[l] >= desc->n_ttype) return fail(EMM_ERR_INVALID, "leader_ttype out of range");

    CUDA_TRY(cudaSetDevice(device));
    emm_library *lib = new emm_library();
    lib->device = device;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    lib->sm_count = prop.multiProcessorCount;
    lib->smem_optin = (int)prop.sharedMemPerBlockOptin;
    lib->h_leader_ttype.assign(desc->leader_ttype, desc->leader_ttype + desc->n_leader);
    lib->h_tpl_atoms.resize((size_t)desc->n_templates);
    for (int t = 0; t < desc->n_templates; ++t) lib->h_tpl_atoms[(size_t)t] = desc->atom_off[t + 1] - desc->atom_off[t];
    lib->h_atom_off.assign(desc->atom_off, desc->atom_off + desc->n_templates + 1);
    lib->h_pair_off.assign(desc->pair_off, desc->pair_off + desc->n_templates + 1);
    lib->h_cut.assign(desc->distance_cutoff, desc->distance_cutoff + desc->n_templates);

    DevLibrary &d = lib->d;
    d.n_templates = desc->n_templates;
    d.n_atoms = desc->n_atoms;
    d.n_ttype = desc->n_ttype;
    d.class_words = desc->class_words;
    d.class_words_cap = (1 << kClassBits) / 32;            // room for all 1024 classes without re-allocation
    d.n_leader = desc->n_leader;
    d.max_tpl_atoms = max_m;
    d.n_lr = desc->n_lr;
    const int T = desc->n_templates, A = desc->n_atoms;
    const int64_t npairs = desc->pair_off[T];
    if (npairs >= (int64_t)1 << 31) { delete lib; return fail(EMM_ERR_INVALID, "template pair table exceeds 2^31 entries"); }
    int rc;
#define COPY(field, src, count) if ((rc = dev_copy(lib, src, (size_t)(count), &field)) != EMM_OK) { emm_library_destroy(lib); return rc; }
    int32_t *p_i32; double *p_f64; uint16_t *p_u16; uint8_t *p_u8; int16_t *p_i16; int64_t *p_i64; float *p_f32; uint32_t *p_u32;
    COPY(p_i32, desc->atom_off, T + 1); d.atom_off = p_i32;
    COPY(p_f64, desc->xyz, 3 * (size_t)A); d.xyz = p_f64;
    COPY(p_f64, desc->weight, A); d.weight = p_f64;
    COPY(p_u16, desc->chain, A); d.chain = p_u16;
    COPY(p_u8, desc->plan_atom, A); d.plan_atom = p_u8;
    COPY(p_u16, desc->plan_ttype, A); d.plan_ttype = p_u16;
    COPY(p_i16, desc->plan_src, A); d.plan_src = p_i16;
    COPY(p_u8, desc->plan_anchor, A); d.plan_anchor = p_u8;
    COPY(p_i64, desc->pair_off, T + 1); d.pair_off = p_i64;
    COPY(p_f64, desc->pair_dist, npairs); d.pair_dist = p_f64;
    {
        std::vector<float> p32((size_t)std::max<int64_t>(npairs, 1));
        for (int64_t i = 0; i < npairs; ++i) p32[(size_t)i] = (float)desc->pair_dist[i];
        COPY(p_f32, p32.data(), npairs); d.pair_dist32 = p_f32;
        // band centre of every plan position's cheap filter, so an expansion reads it with one load
        std::vector<float> ad((size_t)A, 0.f);
        for (int t = 0; t < T; ++t) {
            const int a0 = desc->atom_off[t], m = desc->atom_off[t + 1] - a0;
            for (int k = 1; k < m; ++k)
                ad[(size_t)(a0 + k)] = p32[(size_t)(desc->pair_off[t] + (int64_t)k * (k - 1) / 2 + desc->plan_anchor[a0 + k])];
        }
        COPY(p_f32, ad.data(), A); d.anchor_dist32 = p_f32;
    }
    {
        std::vector<uint32_t> wide((size_t)d.n_ttype * d.class_words_cap, 0u);
        for (int r = 0; r < d.n_ttype; ++r)
            memcpy(&wide[(size_t)r * d.class_words_cap], desc->compat + (size_t)r * desc->class_words, 4 * (size_t)desc->class_words);
        COPY(p_u32, wide.data(), wide.size()); d.compat = p_u32; lib->d_compat = p_u32;
    }
    COPY(p_u16, desc->leader_ttype, desc->n_leader); d.leader_ttype = p_u16;
    COPY(p_f64, desc->rmsd_threshold, T); d.rmsd_thr = p_f64; lib->d_rmsd = p_f64;
    COPY(p_f64, desc->distance_cutoff, T); d.dist_cut = p_f64; lib->d_cut = p_f64;
    COPY(p_f64, desc->max_dynamic_distance, T); d.max_dyn = p_f64; lib->d_dyn = p_f64;
    COPY(p_i32, desc->n_residues, T); d.n_residues = p_i32;
    COPY(p_u8, desc->orient_idx, (size_t)T * EMM_MAX_RESIDUES * 2); d.orient_idx = p_u8;
    COPY(p_f64, desc->orient_vec, (size_t)T * EMM_MAX_RESIDUES * 3); d.orient_vec = p_f64;
    COPY(p_i32, desc->lr_index, T); d.lr_index = p_i32; lib->d_lr_index = p_i32;
    {
        lib->lr_capacity = std::max(desc->n_lr, 64);
        std::vector<double> tab((size_t)lib->lr_capacity * EMM_LR_MODELS * 4, 0.0);
        if (desc->n_lr > 0) memcpy(tab.data(), desc->lr_table, sizeof(double) * (size_t)desc->n_lr * EMM_LR_MODELS * 4);
        COPY(p_f64, tab.data(), tab.size()); d.lr_table = p_f64; lib->d_lr_table = p_f64;
    }
#undef COPY
    d.mask_words = ((desc->n_leader + 255) / 256) * 8;
    {
        void *p = nullptr;
        if (cudaMalloc(&p, (size_t)d.class_words_cap * 32 * d.mask_words * 4) != cudaSuccess) {
            emm_library_destroy(lib);
            return fail(EMM_ERR_NOMEM, "cudaMalloc(class mask)");
        }
        lib->allocs.push_back(p);
        lib->d_class_mask = (uint32_t *)p;
        d.class_mask = lib->d_class_mask;
    }
    if (int rc2 = compute_class_leaders(lib, desc->class_words, desc->compat)) { emm_library_destroy(lib); return rc2; }
    const char *env = getenv("EMM_STATS");
    lib->stats_enabled = env && env[0] == '1';
    *out = lib;
    return EMM_OK;
}

int emm_library_set_compat(emm_library *lib, int32_t class_words, const uint32_t *compat)
{
    if (!lib || !compat) return fail(EMM_ERR_INVALID, "null argument");
    if (class_words <= 0 || class_words > lib->d.class_words_cap)
        return fail(EMM_ERR_INVALID, "class_words exceeds the capacity fixed at library creation (1024 classes)");
    CUDA_TRY(cudaSetDevice(lib->device));
    std::vector<uint32_t> wide((size_t)lib->d.n_ttype * lib->d.class_words_cap, 0u);
    for (int r = 0; r < lib->d.n_ttype; ++r)
        memcpy(&wide[(size_t)r * lib->d.class_words_cap], compat + (size_t)r * class_words, 4 * (size_t)class_words);
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(lib->d_compat, wide.data(), wide.size() * 4, cudaMemcpyHostToDevice));
    lib->d.class_words = class_words;
    return compute_class_leaders(lib, class_words, compat);
}

int emm_library_set_thresholds(emm_library *lib, const double *rmsd_threshold, const double *distance_cutoff,
                               const double *max_dynamic_distance)
{
    if (!lib || !rmsd_threshold || !distance_cutoff || !max_dynamic_distance) return fail(EMM_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(lib->device));
    CUDA_TRY(cudaDeviceSynchronize());
    const size_t bytes = sizeof(double) * (size_t)lib->d.n_templates;
    CUDA_TRY(cudaMemcpy(lib->d_rmsd, rmsd_threshold, bytes, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(lib->d_cut, distance_cutoff, bytes, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(lib->d_dyn, max_dynamic_distance, bytes, cudaMemcpyHostToDevice));
    {
        std::lock_guard<std::mutex> guard(lib->sched_mutex);
        lib->h_cut.assign(distance_cutoff, distance_cutoff + lib->d.n_templates);
        for (auto &kv : lib->sched) cudaFree(kv.second.d_ids);
        lib->sched.clear();
    }
    return EMM_OK;
}

// Visiting order for the templates of [tb, te): more atoms first, then the wider cutoff (the two
// things the cost of a pair grows with), cached per range on the device.
static int get_sched(emm_library *lib, int tb, int te, emm_library::Sched *out)
{
    std::lock_guard<std::mutex> guard(lib->sched_mutex);
    auto it = lib->sched.find({tb, te});
    if (it == lib->sched.end()) {
        std::vector<int> ids((size_t)(te - tb));
        for (int t = tb; t < te; ++t) ids[(size_t)(t - tb)] = t;
        std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) {
            if (lib->h_tpl_atoms[(size_t)a] != lib->h_tpl_atoms[(size_t)b]) return lib->h_tpl_atoms[(size_t)a] > lib->h_tpl_atoms[(size_t)b];
            return lib->h_cut[(size_t)a] > lib->h_cut[(size_t)b];
        });
        emm_library::Sched sc;
        sc.n = te - tb;
        int heavy_atoms = kHeavyAtoms;
        if (const char *env = getenv("EMM_HEAVY_ATOMS")) heavy_atoms = atoi(env);                  // tuning knob
        for (int id : ids) sc.n_heavy += lib->h_tpl_atoms[(size_t)id] >= heavy_atoms;
        // one 16-byte record per entry: template, first atom, atoms, first pair (pair table < 2^31 entries, checked at creation)
        std::vector<int4> recs((size_t)std::max(sc.n, 1));
        for (int i = 0; i < sc.n; ++i) {
            const size_t t = (size_t)ids[(size_t)i];
            recs[(size_t)i] = make_int4((int)t, lib->h_atom_off[t], lib->h_tpl_atoms[t], (int)lib->h_pair_off[t]);
        }
        CUDA_TRY(cudaMalloc(&sc.d_ids, sizeof(int4) * (size_t)std::max(sc.n, 1)));
        if (sc.n) CUDA_TRY(cudaMemcpy(sc.d_ids, recs.data(), sizeof(int4) * (size_t)sc.n, cudaMemcpyHostToDevice));
        it = lib->sched.emplace(std::make_pair(tb, te), sc).first;
    }
    *out = it->second;
    return EMM_OK;
}

int emm_library_set_filter(emm_library *lib, const int32_t *lr_index, int32_t n_lr, const double *lr_table)
{
    if (!lib || !lr_index || (n_lr > 0 && !lr_table)) return fail(EMM_ERR_INVALID, "null argument");
    if (n_lr < 0 || n_lr > lib->lr_capacity) return fail(EMM_ERR_INVALID, "n_lr exceeds the capacity fixed at library creation");
    for (int t = 0; t < lib->d.n_templates; ++t)
        if (lr_index[t] >= n_lr || lr_index[t] < -2) return fail(EMM_ERR_INVALID, "lr_index out of range");
    CUDA_TRY(cudaSetDevice(lib->device));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(lib->d_lr_index, lr_index, sizeof(int32_t) * (size_t)lib->d.n_templates, cudaMemcpyHostToDevice));
    if (n_lr > 0)
        CUDA_TRY(cudaMemcpy(lib->d_lr_table, lr_table, sizeof(double) * (size_t)n_lr * EMM_LR_MODELS * 4, cudaMemcpyHostToDevice));
    lib->d.n_lr = n_lr;
    return EMM_OK;
}

void emm_library_destroy(emm_library *lib)
{
    if (!lib) return;
    cudaSetDevice(lib->device);
    for (void *p : lib->allocs) cudaFree(p);
    for (auto &kv : lib->sched) cudaFree(kv.second.d_ids);
    delete lib;
}

int emm_session_create(emm_library *lib, int64_t max_atoms, int32_t max_structures, int64_t hit_capacity,
                       emm_session **out)
{
    if (!out) return fail(EMM_ERR_INVALID, "out is null");
    *out = nullptr;
    if (!lib || max_atoms <= 0 || max_structures <= 0 || hit_capacity <= 0) return fail(EMM_ERR_INVALID, "bad session sizes");
    CUDA_TRY(cudaSetDevice(lib->device));
    emm_session *s = new emm_session();
    s->lib = lib;
    s->max_atoms = max_atoms;
    s->max_structures = max_structures;
    s->hit_capacity = hit_capacity;
#define ALLOC(ptr, bytes) do { cudaError_t _e = cudaMalloc((void **)&(ptr), (size_t)(bytes)); if (_e != cudaSuccess) { emm_session_destroy(s); return fail(EMM_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(_e)); } } while (0)
    ALLOC(s->d_atom_off, 8 * ((size_t)max_structures + 1));
    ALLOC(s->d_xyz, 24 * (size_t)max_atoms);
    ALLOC(s->d_klass, 2 * (size_t)max_atoms);
    ALLOC(s->d_residue, 4 * (size_t)max_atoms);
    ALLOC(s->d_bfactor, 4 * (size_t)max_atoms);
    ALLOC(s->d_chain, 2 * (size_t)max_atoms);
    ALLOC(s->d_atom_id, 4 * (size_t)max_atoms);
    ALLOC(s->d_blob_off, 8 * ((size_t)max_structures + 1));
    ALLOC(s->d_hits, sizeof(emm_hit) * (size_t)hit_capacity);
    ALLOC(s->d_hit_count, 16);
    ALLOC(s->d_work, 4);
    ALLOC(s->d_any, 4 * (size_t)max_structures);
    ALLOC(s->d_pass, 4 * (size_t)max_structures);
    ALLOC(s->d_skip, (size_t)max_structures);
    ALLOC(s->d_ids, 4 * (size_t)max_structures);
    ALLOC(s->d_status, 4 * (size_t)max_structures);
    ALLOC(s->d_kept_bound, 4 * (size_t)max_structures);
    ALLOC(s->d_stats, 8 * 136);
    s->blob_capacity = 64 * max_atoms + (1024 + 4 * (int64_t)lib->d.n_leader) * max_structures;   // grown on demand at upload
    ALLOC(s->d_blob, s->blob_capacity);
#undef ALLOC
    cudaMemset(s->d_hit_count, 0, 16);
    cudaMemset(s->d_stats, 0, 8 * 136);
    cudaMemset(s->d_any, 0, 4 * (size_t)max_structures);
    cudaMemset(s->d_pass, 0, 4 * (size_t)max_structures);
    cudaMemset(s->d_status, 0, 4 * (size_t)max_structures);
    *out = s;
    return EMM_OK;
}

void emm_session_destroy(emm_session *s)
{
    if (!s) return;
    cudaSetDevice(s->lib->device);
    void *ptrs[] = {s->d_atom_off, s->d_xyz, s->d_klass, s->d_residue, s->d_bfactor, s->d_chain, s->d_atom_id,
                    s->d_blob, s->d_blob_off, s->d_hits, s->d_hit_count, s->d_work, s->d_any, s->d_pass,
                    s->d_skip, s->d_stats, s->d_ids, s->d_status, s->d_kept_bound};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (auto &e : s->ev_prepare) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    for (auto &e : s->ev_search) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    delete s;
}

int emm_session_upload(emm_session *s, const emm_batch *b, void *stream_)
{
    if (!s || !b) return fail(EMM_ERR_INVALID, "null argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (b->n_structures < 0 || b->n_structures > s->max_structures) return fail(EMM_ERR_INVALID, "batch has more structures than the session");
    if (b->n_atoms < 0 || b->n_atoms > s->max_atoms) return fail(EMM_ERR_INVALID, "batch has more atoms than the session");
    if (b->n_structures > 0 && (!b->atom_off || b->atom_off[0] != 0 || b->atom_off[b->n_structures] != b->n_atoms))
        return fail(EMM_ERR_INVALID, "atom_off must start at 0 and end at n_atoms");
    if (b->n_atoms > 0 && (!b->xyz || !b->klass || !b->residue)) return fail(EMM_ERR_INVALID, "xyz, klass and residue are required");
    emm_library *lib = s->lib;
    CUDA_TRY(cudaSetDevice(lib->device));
    const int n = b->n_structures;
    s->h_blob_off.assign((size_t)n + 1, 0);
    int64_t max_staged = 0;
    // shared memory left for a staged blob next to the queues of this library's deepest template
    const int64_t stage_cap = (int64_t)((lib->smem_optin - (int)search_fixed_smem(lib->d.max_tpl_atoms + 1) - 1024) & ~15);
    std::vector<int> large;
    s->h_ids.clear();
    const uint16_t *cl = lib->class_leaders.data();
    const size_t n_class = lib->class_leaders.size();
    // exact size of every structure's leader lists (before masking): one table lookup per atom, on a
    // few threads for large batches (this loop is host time inside the end-to-end path)
    std::vector<int64_t> entries_of((size_t)n, 0);
    s->h_kept_bound.assign((size_t)std::max(n, 1), 0);
    std::atomic<int> bad_input(0);
    {
        auto count = [&](int lo, int hi) {
            for (int i = lo; i < hi; ++i) {
                const int64_t a0 = b->atom_off[i], a1 = b->atom_off[i + 1];
                if (a1 < a0) { bad_input.store(1); return; }
                int64_t entries = 0, nonzero = 0;
                for (int64_t a = a0; a < a1; ++a) {
                    const uint16_t k = b->klass[a];
                    if (k >= n_class) { bad_input.store(2); return; }
                    entries += cl[k];
                    nonzero += k != 0;
                }
                entries_of[(size_t)i] = entries;
                s->h_kept_bound[(size_t)i] = (int32_t)std::min<int64_t>(nonzero, 0x7fffffff);
            }
        };
        const int n_threads = b->n_atoms > (int64_t)2000000 ? std::min(4, std::max(1, (int)std::thread::hardware_concurrency())) : 1;
        std::vector<std::thread> pool;
        const int per = (n + n_threads - 1) / n_threads;
        for (int t = 1; t < n_threads; ++t) pool.emplace_back(count, std::min(n, t * per), std::min(n, (t + 1) * per));
        count(0, std::min(n, per));
        for (auto &th : pool) th.join();
    }
    if (bad_input.load() == 1) return fail(EMM_ERR_INVALID, "atom_off must be non-decreasing");
    if (bad_input.load() == 2) return fail(EMM_ERR_INVALID, "typing class out of range");
    // Staging policy.  A blob is staged when it fits next to the queues; but the CTA should also stay
    // within 196 KB of shared memory, beyond which the SM's L1 shrinks from 60 to 28 KB and the whole
    // launch runs ~1.35x slower.  So the few structures between that soft limit and the hard one are
    // searched in place (4x slower for them) unless they are more than a tenth of the batch.
    const int64_t soft_cap = (int64_t)search_soft_cap(lib->d.max_tpl_atoms + 1) & ~int64_t(1023);
    std::vector<int64_t> staged_of((size_t)n, 0);
    int64_t fits = 0, over_soft = 0;
    for (int i = 0; i < n; ++i) {
        const int64_t a0 = b->atom_off[i], a1 = b->atom_off[i + 1];
        int64_t staged = 0;
        const int64_t bytes = blob_bytes(s->h_kept_bound[(size_t)i], is_wide(a1 - a0), lib->d.n_leader, entries_of[(size_t)i], &staged);
        if (bytes >= (int64_t)1 << 31) return fail(EMM_ERR_INPUT, "a structure is too large: its blob would exceed 2 GiB");
        s->h_blob_off[(size_t)i + 1] = s->h_blob_off[(size_t)i] + bytes;
        const int64_t rounded = (staged + 1023) & ~int64_t(1023);
        const bool ok = !is_wide(a1 - a0) && rounded <= stage_cap;
        staged_of[(size_t)i] = ok ? rounded : -1;
        fits += ok;
        over_soft += ok && rounded > soft_cap;
    }
    const bool keep_l1 = soft_cap > 0 && over_soft * 10 <= fits;
    for (int i = 0; i < n; ++i) {
        const int64_t rounded = staged_of[(size_t)i];
        if (rounded >= 0 && !(keep_l1 && rounded > soft_cap)) {
            max_staged = std::max(max_staged, rounded);
            s->h_ids.push_back(i);
        } else {
            large.push_back(i);      // searched in place from global memory, in a launch of their own
        }
    }
    s->max_staged = max_staged;
    s->n_small = (int32_t)s->h_ids.size();
    s->h_ids.insert(s->h_ids.end(), large.begin(), large.end());
    if (s->h_blob_off[(size_t)n] > s->blob_capacity) {
        CUDA_TRY(cudaStreamSynchronize(stream));
        cudaFree(s->d_blob);
        s->d_blob = nullptr;
        s->blob_capacity = s->h_blob_off[(size_t)n] + (s->h_blob_off[(size_t)n] >> 3);
        cudaError_t e = cudaMalloc((void **)&s->d_blob, (size_t)s->blob_capacity);
        if (e != cudaSuccess) return fail(EMM_ERR_NOMEM, "cudaMalloc(blob)");
    }
    const size_t na = (size_t)b->n_atoms;
    if (n > 0) {
        CUDA_TRY(cudaMemcpyAsync(s->d_atom_off, b->atom_off, 8 * ((size_t)n + 1), cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(s->d_blob_off, s->h_blob_off.data(), 8 * ((size_t)n + 1), cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(s->d_kept_bound, s->h_kept_bound.data(), 4 * (size_t)n, cudaMemcpyHostToDevice, stream));
        if (s->n_small != n) CUDA_TRY(cudaMemcpyAsync(s->d_ids, s->h_ids.data(), 4 * (size_t)n, cudaMemcpyHostToDevice, stream));
    }
    if (na > 0) {
        CUDA_TRY(cudaMemcpyAsync(s->d_xyz, b->xyz, 24 * na, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(s->d_klass, b->klass, 2 * na, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaMemcpyAsync(s->d_residue, b->residue, 4 * na, cudaMemcpyHostToDevice, stream));
        if (b->bfactor) CUDA_TRY(cudaMemcpyAsync(s->d_bfactor, b->bfactor, 4 * na, cudaMemcpyHostToDevice, stream));
        if (b->chain) CUDA_TRY(cudaMemcpyAsync(s->d_chain, b->chain, 2 * na, cudaMemcpyHostToDevice, stream));
        if (b->atom_id) CUDA_TRY(cudaMemcpyAsync(s->d_atom_id, b->atom_id, 4 * na, cudaMemcpyHostToDevice, stream));
    }
    s->n_structures = n;
    s->n_atoms = b->n_atoms;
    s->has_bfactor = b->bfactor != nullptr;
    s->has_chain = b->chain != nullptr;
    s->has_atom_id = b->atom_id != nullptr;
    s->prepared = false;
    CUDA_TRY(cudaMemsetAsync(s->d_hit_count + 1, 0, 8, stream));   // bad-structure counter
    return EMM_OK;
}

int emm_session_run(emm_session *s, const emm_query_params *q, void *stream_)
{
    if (!s || !q) return fail(EMM_ERR_INVALID, "null argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    emm_library *lib = s->lib;
    CUDA_TRY(cudaSetDevice(lib->device));
    s->last_launches = 0;
    const int n = s->n_structures;
    int tb = q->template_begin, te = q->template_end;
    if (te <= 0 || te > lib->d.n_templates) te = lib->d.n_templates;
    if (tb < 0) tb = 0;
    if (q->reset_structure_state) {
        CUDA_TRY(cudaMemsetAsync(s->d_hit_count, 0, 8, stream));
        CUDA_TRY(cudaMemsetAsync(s->d_stats, 0, 8 * 136, stream));
        if (n > 0) {
            CUDA_TRY(cudaMemsetAsync(s->d_any, 0, 4 * (size_t)n, stream));
            CUDA_TRY(cudaMemsetAsync(s->d_pass, 0, 4 * (size_t)n, stream));
        }
    }
    if (n == 0 || tb >= te) return EMM_OK;
    if (!q->ignore_chain && !s->has_chain) return fail(EMM_ERR_INVALID, "ignore_chain=0 needs the chain column");

    DevBatch B{};
    B.n_structures = n;
    B.atom_off = s->d_atom_off;
    B.xyz = s->d_xyz;
    B.klass = s->d_klass;
    B.residue = s->d_residue;
    B.bfactor = s->has_bfactor ? s->d_bfactor : nullptr;
    B.chain = s->has_chain ? s->d_chain : nullptr;
    B.atom_id = s->has_atom_id ? s->d_atom_id : nullptr;
    B.blob = s->d_blob;
    B.blob_off = s->d_blob_off;
    B.status = s->d_status;
    B.kept_bound = s->d_kept_bound;

    const float cutoff = q->conservation_cutoff > 0.f ? q->conservation_cutoff : 0.f;
    if (cutoff > 0.f && !s->has_bfactor) return fail(EMM_ERR_INVALID, "conservation_cutoff needs the bfactor column");
    unsigned long long *stats = s->d_stats;
    const bool want_cells = q->cell_threshold > 0;
    if (q->force_prepare || !s->prepared || s->prepared_cutoff != cutoff || s->prepared_version != lib->compat_version ||
        (want_cells && !s->prepared_cells)) {
        cudaEvent_t e0, e1;
        const bool timed = next_events(s->ev_prepare, s->ev_prepare_used, &e0, &e1);
        if (timed) cudaEventRecord(e0, stream);
        if (q->force_prepare) CUDA_TRY(cudaMemsetAsync(s->d_hit_count + 1, 0, 8, stream));
        launch_prepare(lib->d, B, cutoff, want_cells, stats, s->d_hit_count + 1, lib->sm_count, stream);
        if (timed) cudaEventRecord(e1, stream);
        CUDA_TRY(cudaGetLastError());
        s->prepared = true;
        s->prepared_cutoff = cutoff;
        s->prepared_cells = want_cells;
        s->prepared_version = lib->compat_version;
        s->last_launches++;
    }
    const unsigned char *skip = nullptr;
    if (q->skip_mode == 1 || q->skip_mode == 2) {
        launch_skip_snapshot(n, q->skip_mode, s->d_any, s->d_pass, s->d_skip, stream);
        CUDA_TRY(cudaGetLastError());
        skip = s->d_skip;
        s->last_launches++;
    }
    SearchParams P{};
    P.max_candidates = q->max_candidates;
    P.ignore_chain = q->ignore_chain ? 1 : 0;
    P.template_begin = tb;
    P.template_end = te;
    P.skip_mode = q->skip_mode;
    P.levels = lib->d.max_tpl_atoms + 1;
    P.cell_threshold = q->cell_threshold > 0 ? q->cell_threshold : 0;   // opt-in: typed lists measured faster at every tested size
    // pair splitting: explicit request, else decided per launch below (launch_subset)
    const int donate_request = q->donate_after;
    const char *donate_env = getenv("EMM_DONATE_AFTER");                                           // tuning knob
    const int grid = lib->sm_count;
    emm_library::Sched sc;
    if (int rc = get_sched(lib, tb, te, &sc)) return rc;
    P.n_sched = sc.n;
    P.n_heavy = sc.n_heavy;
    const int fixed = (int)search_fixed_smem(P.levels);
    const int stage_cap = (lib->smem_optin - fixed - 1024) & ~15;
    if (stage_cap < 0) return fail(EMM_ERR_INVALID, "not enough shared memory for the search queues");
    SearchOut O{};
    O.hits = s->d_hits;
    O.hit_capacity = s->hit_capacity;
    O.hit_count = s->d_hit_count;
    O.work_counter = s->d_work;
    O.struct_any = s->d_any;
    O.struct_pass = s->d_pass;
    O.stats = s->d_stats;
    // One launch over the structures whose blob fits in shared memory and, if the batch holds
    // larger ones, a second launch that reads those in place from global memory.
    auto launch_subset = [&](const int *ids, int count, bool staged) -> int {
        if (count <= 0) return EMM_OK;
        int chunks = 1;
        if (count < 2 * grid) chunks = (2 * grid + count - 1) / count;
        chunks = std::min(chunks, std::max(1, (te - tb) / 8));
        chunks = std::max(1, std::min(chunks, 256));
        if (const char *env = getenv("EMM_CHUNKS")) chunks = std::max(1, std::min(atoi(env), 256));   // tuning knob
        P.n_structures = count;
        // Splitting a pair over idle warps pays when one pair can be a visible fraction of the launch:
        // 2.3x on 4 096 structures, where a 150-400 ms pair used to set the time.  From ~8 000 structures
        // up the heavy-first schedule hides such a pair anyway and the bookkeeping costs 2-10 % (10 000
        // structures: 375 vs 386 ms on one batch, 450-517 vs 385-406 ms on another), so it is left off.
        P.donate_after = donate_request < 0 ? -1 : (donate_request > 0 ? donate_request : (count >= kSplitBelow ? -1 : 48));
        if (donate_env) P.donate_after = atoi(donate_env);
        // large batch with a mixed library: heavy templates of every structure first (see SearchParams)
        P.two_phase = chunks == 1 && count >= 2 * grid && sc.n_heavy >= 48 && sc.n_heavy <= sc.n - 48;
        if (const char *env = getenv("EMM_TWO_PHASE")) P.two_phase = P.two_phase && env[0] != '0';   // tuning knob
        if (P.two_phase) chunks = 2;
        P.n_chunks = chunks;
        P.n_items = count * chunks;
        // stage no more than the batch needs: every KB not claimed here stays L1 for the template tables
        const int cap = staged ? (int)((s->max_staged + 1023) & ~int64_t(1023)) : 0;
        P.blob_cap = cap;
        const size_t smem = search_smem_bytes(cap, P.levels, P.cell_threshold > 0);
        {
            std::lock_guard<std::mutex> guard(g_config_mutex);
            if (lib->device >= 64 || smem > g_configured_smem[lib->device]) {   // raise only; never per launch
                CUDA_TRY(configure_search((int)smem));
                if (lib->device < 64) g_configured_smem[lib->device] = smem;
            }
        }
        CUDA_TRY(cudaMemsetAsync(s->d_work, 0, 4, stream));
        cudaEvent_t e0, e1;
        const bool timed = next_events(s->ev_search, s->ev_search_used, &e0, &e1);
        if (timed) cudaEventRecord(e0, stream);
        launch_search(lib->d, B, P, O, skip, sc.d_ids, ids, lib->stats_enabled != 0, staged, std::min(grid, P.n_items), smem, stream);
        if (timed) cudaEventRecord(e1, stream);
        CUDA_TRY(cudaGetLastError());
        s->last_launches++;
        return EMM_OK;
    };
    const bool mixed = s->n_small != n;
    if (int rc = launch_subset(mixed ? s->d_ids : nullptr, s->n_small, true)) return rc;
    if (int rc = launch_subset(s->d_ids + s->n_small, n - s->n_small, false)) return rc;
    return EMM_OK;
}

int emm_session_last_launches(const emm_session *s) { return s ? s->last_launches : 0; }

int emm_session_kernel_ms(emm_session *s, int which, float *out_ms, int capacity, int *count)
{
    if (!s || !count || (which != 0 && which != 1)) return fail(EMM_ERR_INVALID, "bad argument");
    auto &pool = which == 0 ? s->ev_prepare : s->ev_search;
    const size_t used = which == 0 ? s->ev_prepare_used : s->ev_search_used;
    *count = (int)used;
    for (size_t i = 0; i < used && (int)i < capacity && out_ms; ++i)
        CUDA_TRY(cudaEventElapsedTime(out_ms + i, pool[i].first, pool[i].second));
    return EMM_OK;
}

/* Debug: per-level sweep / survivor counters of the stats build (EMM_STATS=1); 128 values. */
int emm_session_debug_counters(emm_session *s, unsigned long long *out128)
{
    if (!s || !out128) return fail(EMM_ERR_INVALID, "null argument");
    CUDA_TRY(cudaMemcpy(out128, s->d_stats + 8, 8 * 128, cudaMemcpyDeviceToHost));
    return EMM_OK;
}

int emm_session_clear_timings(emm_session *s)
{
    if (!s) return fail(EMM_ERR_INVALID, "null argument");
    s->ev_prepare_used = 0;
    s->ev_search_used = 0;
    return EMM_OK;
}

int emm_session_download(emm_session *s, emm_hit *hits, int64_t capacity, int64_t *n_hits, emm_stats *stats,
                         void *stream_)
{
    if (!s || !n_hits) return fail(EMM_ERR_INVALID, "null argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    CUDA_TRY(cudaSetDevice(s->lib->device));
    unsigned long long counters[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(counters, s->d_hit_count, 16, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    const unsigned long long count = counters[0], bad = counters[1];
    *n_hits = (int64_t)count;
    if (stats) {
        unsigned long long raw[8];
        CUDA_TRY(cudaMemcpy(raw, s->d_stats, 64, cudaMemcpyDeviceToHost));
        stats->pairs = raw[0]; stats->sweeps = raw[1]; stats->dist_evals = raw[2]; stats->exact_rechecks = raw[3];
        stats->complete = raw[4]; stats->kept_atoms = raw[5]; stats->staged_bytes = raw[6]; stats->global_blobs = raw[7];
    }
    if ((int64_t)count > s->hit_capacity || (int64_t)count > capacity)
        return fail(EMM_ERR_CAPACITY, "hit buffer too small: enlarge hit_capacity and run again");
    if (count > 0) {
        if (!hits) return fail(EMM_ERR_INVALID, "hits is null");
        CUDA_TRY(cudaMemcpyAsync(hits, s->d_hits, sizeof(emm_hit) * (size_t)count, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        // sort by (structure, template): order 12-byte keys, then move every 280-byte record once
        if (count > 0xFFFFFFFFull) return fail(EMM_ERR_CAPACITY, "more than 2^32 hits in one batch");
        std::vector<std::pair<uint64_t, uint32_t>> keys((size_t)count);
        for (size_t i = 0; i < (size_t)count; ++i)
            keys[i] = {((uint64_t)(uint32_t)hits[i].structure << 32) | (uint32_t)hits[i].template_index, (uint32_t)i};
        std::sort(keys.begin(), keys.end());
        bool ordered = true;
        for (size_t i = 0; i < (size_t)count && ordered; ++i) ordered = keys[i].second == i;
        if (!ordered) {
            std::vector<emm_hit> tmp(hits, hits + count);
            for (size_t i = 0; i < (size_t)count; ++i) hits[i] = tmp[keys[i].second];
        }
    }
    // Loud, but not fatal for the rest of the batch: the hits of every other structure were delivered
    // above; emm_session_structure_status names the offenders.
    if (bad) return fail(EMM_ERR_INPUT, "a structure violates the input contract (residue ordinals must be "
                                        "non-decreasing, at most 4194303 kept atoms, at most 1023 kept atoms per "
                                        "residue): its search was skipped; see emm_session_structure_status");
    return EMM_OK;
}

int emm_session_structure_status(emm_session *s, int32_t *status, int32_t capacity)
{
    if (!s || !status || capacity < s->n_structures) return fail(EMM_ERR_INVALID, "status buffer too small");
    CUDA_TRY(cudaSetDevice(s->lib->device));
    if (!s->prepared) return fail(EMM_ERR_INVALID, "no prepared batch: call emm_session_run first");
    if (s->n_structures > 0)
        CUDA_TRY(cudaMemcpy(status, s->d_status, 4 * (size_t)s->n_structures, cudaMemcpyDeviceToHost));
    return EMM_OK;
}

int emm_query_batch(emm_library *lib, const emm_batch *batch, const emm_query_params *params, emm_hit *hits,
                    int64_t capacity, int64_t *n_hits, emm_stats *stats)
{
    if (!lib || !batch || !params || !n_hits) return fail(EMM_ERR_INVALID, "null argument");
    emm_session *s = nullptr;
    int rc = emm_session_create(lib, std::max<int64_t>(batch->n_atoms, 1), std::max(batch->n_structures, 1),
                                std::max<int64_t>(capacity, 1), &s);
    if (rc != EMM_OK) return rc;
    emm_query_params p = *params;
    p.reset_structure_state = 1;
    rc = emm_session_upload(s, batch, nullptr);
    if (rc == EMM_OK) rc = emm_session_run(s, &p, nullptr);
    if (rc == EMM_OK) rc = emm_session_download(s, hits, capacity, n_hits, stats, nullptr);
    emm_session_destroy(s);
    return rc;
}

}  // extern "C"
