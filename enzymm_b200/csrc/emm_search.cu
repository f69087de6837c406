// emm_search.cu -- the (template x structure) candidate search, superposition and filter kernel
// (north_star subsystems 3, 4 and 5, fused).
//
// Replaces pyjess.Jess(templates).query(...) + Match.predicted_correct for a whole batch
// (reference enzymm/jess_run.py:785-843, 298-346, 425-478).
//
// Mapping.  Persistent CTAs (one per SM, kSearchWarps = 24 warps).  A work item is (structure,
// phase of the cost-ordered template schedule): the CTA stages the structure blob into shared
// memory with one TMA bulk copy (cp.async.bulk + mbarrier), then each WARP pulls templates off a
// shared counter and runs a warp-synchronous depth-first search (entered directly at level 1):
//
//   * level k places plan position k of the template.  Partial assignments live in per-level
//     shared-memory queues as 4-byte (parent slot, atom) entries -- a trie, so a partial costs
//     4 bytes whatever its depth; queues are drained depth-first in chunks of <= 32 entries, so
//     memory is bounded however many candidates a loose cutoff produces (config 4);
//   * EXPANDING a chunk is a cheap, dense filter: for leader positions every lane holds up to two
//     candidate atoms of the structure's leader list in registers and the warp loops over the
//     chunk's partials, testing ONE distance constraint against a broadcast "anchor" atom of the
//     partial; for same-residue positions the lanes cover (partial, residue slot) items and test
//     the typing bit + the anchor distance.  Survivors are ballot/popc-compacted into the next
//     level's queue WITHOUT further checks;
//   * ENTERING a level validates its chunk densely, one entry per lane: the lane walks its chain
//     once, tests the new atom against every placed atom (injectivity + all pairwise-distance
//     constraints) and records the anchor the next expansion needs.  So every constraint is
//     checked exactly once per surviving partial, with all 32 lanes busy;
//   * every accept/reject is decided in FP32 on centred coordinates unless it falls inside the
//     guard band eps, in which case the lane re-evaluates in FP64 with separately rounded
//     operations (__d*_rn) -- the same expressions as the CPU oracle, so the set of complete
//     assignments is identical bit for bit;
//   * complete assignments are superposed one per lane in FP64 registers (Horn's quaternion form
//     of Kabsch + cyclic Jacobi), the per-template minimum is kept, and the winning hit gets
//     EnzyMM's orientation + logistic filter before it is written out.
//
// No tensor cores: this is gather/compare-bound FP32 + integer work with a rare FP64 tail.
#include <math_constants.h>

#include <type_traits>

#include "emm_device.cuh"

namespace emm {

// ---- canonical FP64 arithmetic (no FMA contraction; mirrors oracle/jess_oracle.c) ---------------
#define DADD(a, b) __dadd_rn((a), (b))
#define DSUB(a, b) __dsub_rn((a), (b))
#define DMUL(a, b) __dmul_rn((a), (b))
#define DDIV(a, b) __ddiv_rn((a), (b))
#define DSQRT(a) __dsqrt_rn((a))

constexpr unsigned kFull = 0xffffffffu;
// Queue entry: atom (22 bits) | parent slot (8 bits) << 22 | flags.
constexpr uint32_t kEntryValid = 1u << 30;   // entry passed validation
constexpr uint32_t kEntryDead = 1u << 31;    // entry failed validation
__device__ __forceinline__ int entry_atom(uint32_t e) { return (int)(e & kAtomMask); }
__device__ __forceinline__ int entry_parent(uint32_t e) { return (int)((e >> kAtomBits) & 0xffu); }
__device__ __forceinline__ uint32_t make_entry(int parent, int atom) { return ((uint32_t)parent << kAtomBits) | (uint32_t)atom; }
// same-residue payload of an anchor: first atom of the residue | atoms in it << 22
__device__ __forceinline__ int span_payload(int first, int count) { return first | (count << kAtomBits); }

// Resume point of a cell-list expansion (opt-in path): kept outside WarpState so that the default
// kernels do not pay shared memory for it.
struct CellState {
    int row[kMaxAtoms + 1];      // row of the anchor's cell box being scanned ...
    int pos[kMaxAtoms + 1];      // ... and next position inside that row (-1 = row not started)
};

// Per-warp search state.  Every byte here is shared memory taken from the L1 carve-out: with the
// staged blob (~86 KB for 400-residue structures) and the queues (66 KB) the CTA must stay below
// 196 KB, or the SM's L1 shrinks from 60 KB to 28 KB and every template-table read pays for it
// (measured: 1.35x on the whole kernel).  Hence bytes for the small counters and no padding.
struct alignas(16) WarpState {
    float4 anchor[32];           // per valid partial (compacted): anchor atom xyz, w = payload bits
    unsigned long long todo[kMaxAtoms + 1];   // lanes (two candidate rows) of iteration `cur` still to be pushed
    double best_rmsd;
    unsigned long long n_complete;
    CellState *cell;             // this warp's cell-list resume state (kCells kernels only)
    int cur[kMaxAtoms + 1];      // expansion cursor inside the chunk
    uint32_t best_asg[kMaxAtoms];
    int best_valid;
    int overflow;
    int donated;                 // this warp gave parts of its own pair away
    int for_owner;               // searching a donated subtree on behalf of this warp, or -1
    uint8_t n[kMaxAtoms + 1];    // entries queued per level (<= kQueueCap)
    uint8_t chunk[kMaxAtoms + 1];   // size of the level's current chunk (its last `chunk` entries)
    unsigned char vslot[32];     // compacted list of valid chunk slots
};
static_assert(kQueueCap < 256, "queue counters are bytes");

// ---- splitting one (template, structure) pair across the warps of its CTA -------------------------
// Pair cost is heavy-tailed (a 21-atom template at cutoff 2.0 can take 100 ms of warp time against
// one structure while the median pair takes 10 us).  A warp whose pair has run for a while gives
// untouched partial assignments of the chunk it is entering -- whole subtrees of the search -- to
// warps that have run out of templates: a donation is the chain of atoms placed so far; the helper
// rebuilds that chain as one-entry queues and runs the ordinary search below it; results meet in
// the owner's PairSlot (minimum RMSD with the lexicographic tie-break, summed complete-assignment
// counts), and the owner writes the hit once every donation has come back.
constexpr int kDonationRing = 16;
#ifndef EMM_DONATE_STRIDE
#define EMM_DONATE_STRIDE 4
#endif
#ifndef EMM_NAP_NS
#define EMM_NAP_NS 500u
#endif

struct Donation {
    int owner;                   // warp whose pair this subtree belongs to
    int depth;                   // atoms placed: plan positions 0 .. depth-1
    int t;                       // template
    int pad;
    uint32_t chain[kMaxAtoms];   // chain[j] = atom at plan position j
};

struct PairSlot {
    int lock;                    // spin lock (lane 0 of the merging warp)
    int pending;                 // donations handed out and not merged back yet
    int best_valid, overflow;
    unsigned long long n_complete;
    unsigned long long best_rmsd_bits;   // the double, as bits (all traffic through here is atomic, see sh_get)
    uint32_t best_asg[kMaxAtoms];
};

struct CtaShare {
    int ring_lock, ring_count;
    int idle;                    // warps out of templates, waiting for donations
    int in_loop;                 // warps that still own (or may still take) a template of this item
    Donation ring[kDonationRing];
    PairSlot slot[kSearchWarps];
};

// Everything warps tell each other through CtaShare goes through shared-memory atomics -- reads as
// atomicAdd(p, 0), writes as atomicExch -- in addition to the locks and fences that order it: the
// accesses are rare (a donation, a merge, a poll between two naps), and it keeps the protocol
// visible to compute-sanitizer's racecheck, which knows barriers and atomics but not spin locks.
__device__ __forceinline__ int sh_get(int *p) { return atomicAdd(p, 0); }
__device__ __forceinline__ unsigned sh_get(unsigned *p) { return atomicAdd(p, 0u); }
__device__ __forceinline__ unsigned long long sh_get(unsigned long long *p) { return atomicAdd(p, 0ull); }
__device__ __forceinline__ void sh_set(int *p, int v) { atomicExch(p, v); }
__device__ __forceinline__ void sh_set(unsigned *p, unsigned v) { atomicExch(p, v); }
__device__ __forceinline__ void sh_set(unsigned long long *p, unsigned long long v) { atomicExch(p, v); }
__device__ __forceinline__ int peek(int *p) { return sh_get(p); }

__device__ __forceinline__ void spin_lock(int *lock)
{
    while (atomicCAS(lock, 0, 1) != 0) __nanosleep(20);
    __threadfence_block();
}

__device__ __forceinline__ void spin_unlock(int *lock)
{
    __threadfence_block();
    atomicExch(lock, 0);
}

// Dynamic shared memory: [staged blob][per-warp queues][per-warp state].
extern __shared__ __align__(16) unsigned char g_smem[];

// Global-memory side of a structure, used only by the rare FP64 paths; one copy per CTA in
// shared memory so the out-of-line functions do not drag a struct through local memory.
struct Blob {
    const int32_t *orig;      // local atom id -> position inside the structure
    const double *xyz64;      // structure base, FP64 coordinates as uploaded
    const uint16_t *chain;    // may be null
    const int32_t *atom_id;   // may be null
    // uniform grid (cell list) of the structure, see BlobHeader; read in place from global memory
    const void *cell_start, *cell_atoms;
    int wide;                 // 32-bit entries in cell_start / cell_atoms (and res_start / lead)
    int nx, ny, nz;
    float cell, ox, oy, oz;
};

// Hot-path view of the structure blob.  kStaged: the blob sits at the start of dynamic shared
// memory and every access compiles to LDS with 32-bit addressing; otherwise it is read in place
// from global memory (structures too large to stage).
template <bool kStaged>
struct View {
    const unsigned char *gbase;
    int off_atom, off_resstart, off_leadoff, off_lead;
    int res_shift;
    bool wide;                // index arrays hold 32-bit entries (never the case for a staged blob)
    float eps;
    template <typename T>
    __device__ __forceinline__ T ld(int byte_off) const
    {
        if (kStaged) return *reinterpret_cast<const T *>(g_smem + byte_off);
        return __ldg(reinterpret_cast<const T *>(gbase + byte_off));
    }
    // atom record: centred x, y, z and (res_of << 10 | klass) as the bits of w -- one 128-bit load
    __device__ __forceinline__ float4 atom(int i) const { return ld<float4>(off_atom + 16 * i); }
    static __device__ __forceinline__ unsigned klass_of(const float4 &p) { return __float_as_uint(p.w) & kClassMask; }
    static __device__ __forceinline__ int res_of(const float4 &p) { return (int)(__float_as_uint(p.w) >> kClassBits); }
    __device__ __forceinline__ int idx(int off, int i) const
    {
        if (!kStaged && wide) return (int)ld<uint32_t>(off + 4 * i);
        return ld<uint16_t>(off + 2 * i);
    }
    __device__ __forceinline__ int res_start(int r) const { return idx(off_resstart, r); }
    __device__ __forceinline__ int lead_off(int l) const { return (int)ld<uint32_t>(off_leadoff + 4 * l); }
    __device__ __forceinline__ int lead(int i) const { return idx(off_lead, i); }
};

__device__ __forceinline__ int cell_idx(const void *base, int wide, int i)
{
    return wide ? (int)__ldg(reinterpret_cast<const uint32_t *>(base) + i)
                : (int)__ldg(reinterpret_cast<const uint16_t *>(base) + i);
}

struct SearchArgs {
    DevLibrary L;
    DevBatch B;
    SearchParams P;
    SearchOut O;
    const unsigned char *skip;
    const int4 *sched;         // [P.n_sched] visiting order: (template, atom_off, atoms, pair_off) per entry, so a
                               // warp learns all it needs to start a template from one (mostly L1-resident) load
    const int *ids;            // [P.n_structures] structures of this launch, or null for 0..n-1
};

// ---- TMA bulk copy (cp.async.bulk, SASS: UBLKCP) + mbarrier: global -> shared staging ----------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// One thread: order the CTA's earlier generic-proxy accesses to the staging area before the async
// proxy overwrites it, announce the byte count, and start the copy (size: multiple of 16 bytes).
__device__ __forceinline__ void bulk_load(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ float fast_sqrt(float v)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ double exact_dist(const double *xyz, int a, int b)
{
    const double *pa = xyz + 3 * (int64_t)a, *pb = xyz + 3 * (int64_t)b;
    const double dx = DSUB(pa[0], pb[0]), dy = DSUB(pa[1], pb[1]), dz = DSUB(pa[2], pb[2]);
    return DSQRT(DADD(DADD(DMUL(dx, dx), DMUL(dy, dy)), DMUL(dz, dz)));
}

// delta_ij of SURVEY 8c rule 6 for plan positions (k, j) -- same operation order as the oracle,
// which always adds the weight of the LATER template atom first.
__device__ __forceinline__ double pair_delta(const DevLibrary &L, int a0, int k, int j, double cut,
                                             double max_dyn)
{
    if (max_dyn == cut) return cut;
    const int ik = L.plan_atom[a0 + k], ij = L.plan_atom[a0 + j];
    const int hi = ik > ij ? ik : ij, lo = ik > ij ? ij : ik;
    const double d = DADD(DADD(cut, L.weight[a0 + hi]), L.weight[a0 + lo]);
    return d < max_dyn ? d : max_dyn;
}

// Cyclic Jacobi on a symmetric 4x4 (fully unrolled so the matrices stay in registers).
// The scalar part of one Jacobi rotation (divisions and square roots are long software sequences:
// kept out of line so the six unrolled rotations of jacobi4 share one copy).  Returns false when the
// off-diagonal element is negligible and is simply zeroed; otherwise s, tau and h = t * apq.
__device__ __noinline__ bool jacobi_rotation(double apq, double app, double aqq, int sweep, double *s_out,
                                             double *tau_out, double *h_out)
{
    const double g = DMUL(100.0, fabs(apq));
    if (sweep > 3 && DADD(fabs(app), g) == fabs(app) && DADD(fabs(aqq), g) == fabs(aqq)) return false;
    double h = DSUB(aqq, app);
    double t;
    if (DADD(fabs(h), g) == fabs(h)) {
        t = DDIV(apq, h);
    } else {
        const double theta = DDIV(DMUL(0.5, h), apq);
        t = DDIV(1.0, DADD(fabs(theta), DSQRT(DADD(1.0, DMUL(theta, theta)))));
        if (theta < 0.0) t = -t;
    }
    const double c = DDIV(1.0, DSQRT(DADD(1.0, DMUL(t, t))));
    const double s = DMUL(t, c);
    *s_out = s;
    *tau_out = DDIV(s, DADD(1.0, c));
    *h_out = DMUL(t, apq);
    return true;
}

__device__ __forceinline__ void jacobi4(double (&a)[4][4], double (&v)[4][4])
{
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) v[p][q] = (p == q) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = 0.0;
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = p + 1; q < 4; ++q) off = DADD(off, fabs(a[p][q]));
        if (off == 0.0) break;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                const double apq = a[p][q];
                if (apq != 0.0) {
                    double s, tau, h;
                    if (!jacobi_rotation(apq, a[p][p], a[q][q], sweep, &s, &tau, &h)) {
                        a[p][q] = 0.0;
                        a[q][p] = 0.0;
                    } else {
                        a[p][p] = DSUB(a[p][p], h);
                        a[q][q] = DADD(a[q][q], h);
                        a[p][q] = 0.0;
                        a[q][p] = 0.0;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (k != p && k != q) {
                                const double akp = a[k][p], akq = a[k][q];
                                const double nkp = DSUB(akp, DMUL(s, DADD(akq, DMUL(akp, tau))));
                                const double nkq = DADD(akq, DMUL(s, DSUB(akp, DMUL(akq, tau))));
                                a[k][p] = nkp; a[p][k] = nkp;
                                a[k][q] = nkq; a[q][k] = nkq;
                            }
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const double vkp = v[k][p], vkq = v[k][q];
                            v[k][p] = DSUB(vkp, DMUL(s, DADD(vkq, DMUL(vkp, tau))));
                            v[k][q] = DADD(vkq, DMUL(s, DSUB(vkp, DMUL(vkq, tau))));
                        }
                    }
                }
            }
        }
    }
}

// Optimal proper rotation of the matched query atoms onto the template atoms; returns rmsd.
// asg[i] = local atom id bound to template atom i (template order).
template <typename AsgT>
__device__ double superpose(int m, const double *__restrict__ txyz, const Blob &S,
                            const AsgT *asg, double (&rot)[9], double (&qbar)[3],
                            double (&tbar)[3])
{
    const double inv_m = DDIV(1.0, (double)m);
    double sq[3] = {0.0, 0.0, 0.0}, st[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < m; ++i) {
        const double *q = S.xyz64 + 3 * (int64_t)S.orig[asg[i]];
#pragma unroll
        for (int c = 0; c < 3; ++c) { sq[c] = DADD(sq[c], q[c]); st[c] = DADD(st[c], txyz[3 * i + c]); }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { qbar[c] = DMUL(sq[c], inv_m); tbar[c] = DMUL(st[c], inv_m); }
    double M[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int i = 0; i < m; ++i) {
        const double *q = S.xyz64 + 3 * (int64_t)S.orig[asg[i]];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
                M[a][b] = DADD(M[a][b], DMUL(DSUB(q[a], qbar[a]), DSUB(txyz[3 * i + b], tbar[b])));
    }
    double N[4][4], V[4][4];
    N[0][0] = DADD(DADD(M[0][0], M[1][1]), M[2][2]);
    N[1][1] = DSUB(DSUB(M[0][0], M[1][1]), M[2][2]);
    N[2][2] = DSUB(DSUB(M[1][1], M[0][0]), M[2][2]);
    N[3][3] = DSUB(DSUB(M[2][2], M[0][0]), M[1][1]);
    N[0][1] = N[1][0] = DSUB(M[1][2], M[2][1]);
    N[0][2] = N[2][0] = DSUB(M[2][0], M[0][2]);
    N[0][3] = N[3][0] = DSUB(M[0][1], M[1][0]);
    N[1][2] = N[2][1] = DADD(M[0][1], M[1][0]);
    N[1][3] = N[3][1] = DADD(M[2][0], M[0][2]);
    N[2][3] = N[3][2] = DADD(M[1][2], M[2][1]);
    jacobi4(N, V);
    double best = N[0][0];
    double q0 = V[0][0], q1 = V[1][0], q2 = V[2][0], q3 = V[3][0];
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (N[i][i] > best) { best = N[i][i]; q0 = V[0][i]; q1 = V[1][i]; q2 = V[2][i]; q3 = V[3][i]; }
    const double nrm = DSQRT(DADD(DADD(DADD(DMUL(q0, q0), DMUL(q1, q1)), DMUL(q2, q2)), DMUL(q3, q3)));
    q0 = DDIV(q0, nrm); q1 = DDIV(q1, nrm); q2 = DDIV(q2, nrm); q3 = DDIV(q3, nrm);
    rot[0] = DSUB(DSUB(DADD(DMUL(q0, q0), DMUL(q1, q1)), DMUL(q2, q2)), DMUL(q3, q3));
    rot[1] = DMUL(2.0, DSUB(DMUL(q1, q2), DMUL(q0, q3)));
    rot[2] = DMUL(2.0, DADD(DMUL(q1, q3), DMUL(q0, q2)));
    rot[3] = DMUL(2.0, DADD(DMUL(q1, q2), DMUL(q0, q3)));
    rot[4] = DSUB(DADD(DSUB(DMUL(q0, q0), DMUL(q1, q1)), DMUL(q2, q2)), DMUL(q3, q3));
    rot[5] = DMUL(2.0, DSUB(DMUL(q2, q3), DMUL(q0, q1)));
    rot[6] = DMUL(2.0, DSUB(DMUL(q1, q3), DMUL(q0, q2)));
    rot[7] = DMUL(2.0, DADD(DMUL(q2, q3), DMUL(q0, q1)));
    rot[8] = DADD(DSUB(DSUB(DMUL(q0, q0), DMUL(q1, q1)), DMUL(q2, q2)), DMUL(q3, q3));
    double ssd = 0.0;
    for (int i = 0; i < m; ++i) {
        const double *q = S.xyz64 + 3 * (int64_t)S.orig[asg[i]];
        const double x = DSUB(q[0], qbar[0]), y = DSUB(q[1], qbar[1]), z = DSUB(q[2], qbar[2]);
        const double rx = DSUB(DADD(DADD(DMUL(rot[0], x), DMUL(rot[1], y)), DMUL(rot[2], z)), DSUB(txyz[3 * i], tbar[0]));
        const double ry = DSUB(DADD(DADD(DMUL(rot[3], x), DMUL(rot[4], y)), DMUL(rot[5], z)), DSUB(txyz[3 * i + 1], tbar[1]));
        const double rz = DSUB(DADD(DADD(DMUL(rot[6], x), DMUL(rot[7], y)), DMUL(rot[8], z)), DSUB(txyz[3 * i + 2], tbar[2]));
        ssd = DADD(ssd, DADD(DADD(DMUL(rx, rx), DMUL(ry, ry)), DMUL(rz, rz)));
    }
    return DSQRT(DMUL(ssd, inv_m));
}

// Vec3.angle_to (enzymm/template.py:157-181); returns NaN where the reference raises.
__device__ double angle_between(const double (&u)[3], const double (&v)[3])
{
    const double nu = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    const double nv = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    double a[3], b[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { a[c] = nu == 0.0 ? u[c] : u[c] / nu; b[c] = nv == 0.0 ? v[c] : v[c] / nv; }
    const double dot = DADD(DADD(DMUL(a[0], b[0]), DMUL(a[1], b[1])), DMUL(a[2], b[2]));
    if (dot >= -1.0 && dot <= 1.0) return acos(dot);
    if (fabs(dot - 1.0) <= 1e-5 * fmax(fabs(dot), 1.0)) return 0.0;
    if (fabs(dot + 1.0) <= 1e-5 * fmax(fabs(dot), 1.0)) return CUDART_PI;
    return CUDART_NAN;
}


// ---- rare paths: kept out of line so their register needs do not burden the search loop ---------

// Write the winning hit of (structure s, template t): superposition, orientation, logistic vote.
__device__ __noinline__ void emit_hit(const SearchArgs &A, const Blob &S, int s, int t, const WarpState *ws)
{
    const DevLibrary &L = A.L;
    const SearchOut &O = A.O;
    const int a0 = L.atom_off[t], m = L.atom_off[t + 1] - a0;
    const double *txyz = L.xyz + 3 * (int64_t)a0;
    double rot[9], qbar[3], tbar[3];
    const double rmsd = superpose(m, txyz, S, ws->best_asg, rot, qbar, tbar);
    unsigned flags = ws->overflow ? EMM_HIT_OVERFLOW : 0u;
    double orient = CUDART_NAN;
    const int nres = L.n_residues[t];
    if (nres > 0) {
        // Match.match_vector_list / orientation (jess_run.py:425-478) in the template frame
        double sum = 0.0;
        for (int r = 0; r < nres; ++r) {
            double p[3][3];
            for (int k = 0; k < 3; ++k) {
                const double *q = S.xyz64 + 3 * (int64_t)S.orig[ws->best_asg[3 * r + k]];
                const double x = q[0] - qbar[0], y = q[1] - qbar[1], z = q[2] - qbar[2];
                p[k][0] = rot[0] * x + rot[1] * y + rot[2] * z + tbar[0];
                p[k][1] = rot[3] * x + rot[4] * y + rot[5] * z + tbar[1];
                p[k][2] = rot[6] * x + rot[7] * y + rot[8] * z + tbar[2];
            }
            const int i0 = L.orient_idx[((size_t)t * EMM_MAX_RESIDUES + r) * 2];
            const int i1 = L.orient_idx[((size_t)t * EMM_MAX_RESIDUES + r) * 2 + 1];
            double vec[3];
            if (i1 == 9) {
                const int s1 = i0 == 0 ? 1 : 0, s2 = i0 == 2 ? 1 : 2;
#pragma unroll
                for (int c = 0; c < 3; ++c) vec[c] = (p[s1][c] + p[s2][c]) / 2.0 - p[i0][c];
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c) vec[c] = p[i1][c] - p[i0][c];
            }
            const double *tv = L.orient_vec + ((size_t)t * EMM_MAX_RESIDUES + r) * 3;
            const double tvec[3] = {tv[0], tv[1], tv[2]};
            sum += angle_between(tvec, vec);
        }
        orient = sum / (double)nres;
        flags |= EMM_HIT_ORIENTED;
    }
    // Match.predicted_correct (jess_run.py:298-346)
    const int lr = L.lr_index[t];
    bool pass = true;
    if (lr == -2) {
        flags |= EMM_HIT_NO_MODEL;
        pass = false;
    } else if (lr >= 0) {
        int votes = 0;
        const double *tab = L.lr_table + (size_t)lr * EMM_LR_MODELS * 4;
        for (int k = 0; k < EMM_LR_MODELS; ++k) {
            const double z = tab[4 * k + 2] + tab[4 * k] * rmsd + tab[4 * k + 1] * orient;
            const double value = 1.0 / (1.0 + pow(2.718281828459045, -z));
            if (value >= tab[4 * k + 3]) ++votes;
            if (fabs(value - tab[4 * k + 3]) < 1e-9) flags |= EMM_HIT_BORDERLINE;
        }
        pass = votes >= 2;  // round(5 / 2, 0) == 2 under banker's rounding
        if (!(orient == orient)) pass = false;
    }
    if (pass) flags |= EMM_HIT_PASS;
    atomicAdd(O.struct_any + s, 1);
    if (pass) atomicAdd(O.struct_pass + s, 1);
    const unsigned long long slot = atomicAdd(O.hit_count, 1ull);
    if ((long long)slot < O.hit_capacity) {
        emm_hit *h = O.hits + slot;
        h->structure = s;
        h->template_index = t;
        h->n_complete = (uint32_t)(ws->n_complete > 0xffffffffull ? 0xffffffffull : ws->n_complete);
        h->n_atoms = (uint16_t)m;
        h->flags = (uint16_t)flags;
        h->rmsd = rmsd;
        h->orientation = orient;
        for (int i = 0; i < 9; ++i) h->rot[i] = rot[i];
        for (int i = 0; i < 3; ++i) { h->qbar[i] = qbar[i]; h->tbar[i] = tbar[i]; }
        for (int i = 0; i < kMaxAtoms; ++i) {
            int id = -1;
            if (i < m) { id = S.orig[ws->best_asg[i]]; if (S.atom_id) id = S.atom_id[id]; }
            h->atoms[i] = id;
        }
    }
}

// Superpose the validated complete assignments of one chunk of level m (one per lane), keep the best.
// kNarrow: atom ids fit 16 bits (every staged structure).  The per-lane assignment is a local-memory
// array; as 32-bit words it is 128 B per thread, 96 KB for the CTA's 768 threads -- more than the
// SM's 60 KB of L1 -- and the superposition path thrashes (BASELINE config 4: -25 %).
template <bool kNarrow>
__device__ __noinline__ void process_complete(const SearchArgs &A, const Blob &S, int t, const uint32_t *Q,
                                              WarpState *ws, int base, unsigned valid, int lane)
{
    const DevLibrary &L = A.L;
    const int a0 = L.atom_off[t], m = L.atom_off[t + 1] - a0;
    const double *txyz = L.xyz + 3 * (int64_t)a0;
    const double thr = L.rmsd_thr[t];
    bool have = (valid >> lane) & 1u;
    using AsgT = typename std::conditional<kNarrow, uint16_t, uint32_t>::type;
    AsgT asg[kMaxAtoms];
    if (have) {
        uint32_t w = Q[queue_off(m) + base + lane];
        for (int pos = m - 1; pos >= 0; --pos) {
            asg[L.plan_atom[a0 + pos]] = (AsgT)entry_atom(w);
            if (pos > 0) w = Q[queue_off(pos) + entry_parent(w)];
        }
        if (!A.P.ignore_chain && S.chain) {
            // template atoms on equal chains <=> query atoms on equal chains (oracle rule 11)
            for (int i = 0; i < m && have; ++i)
                for (int j = i + 1; j < m; ++j) {
                    const bool st = L.chain[a0 + i] == L.chain[a0 + j];
                    const bool sq = S.chain[S.orig[asg[i]]] == S.chain[S.orig[asg[j]]];
                    if (st != sq) { have = false; break; }
                }
        }
    }
    const unsigned counted = __ballot_sync(kFull, have);
    double rmsd = CUDART_INF;
    bool accept = false;
    if (have) {
        double rot[9], qbar[3], tbar[3];
        rmsd = superpose(m, txyz, S, asg, rot, qbar, tbar);
        accept = rmsd <= thr;
    }
    const double key = accept ? rmsd : CUDART_INF;
    double mn = key;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(kFull, mn, o));
    if (mn < CUDART_INF) {
        unsigned tied = __ballot_sync(kFull, accept && key == mn);
        if (__popc(tied) > 1) {  // exact RMSD tie: lexicographically smallest assignment wins
            for (int i = 0; i < m && __popc(tied) > 1; ++i) {
                const unsigned v = ((tied >> lane) & 1u) ? (unsigned)asg[i] : 0xffffffffu;
                unsigned mv = v;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mv = min(mv, __shfl_xor_sync(kFull, mv, o));
                tied &= __ballot_sync(kFull, v == mv);
            }
        }
        const int win = __ffs(tied) - 1;
        if (lane == win) {
            bool better = !ws->best_valid || mn < ws->best_rmsd;
            if (!better && mn == ws->best_rmsd) {
                for (int i = 0; i < m; ++i) {
                    if (asg[i] != ws->best_asg[i]) { better = asg[i] < ws->best_asg[i]; break; }
                }
            }
            if (better) {
                ws->best_valid = 1;
                ws->best_rmsd = mn;
                for (int i = 0; i < m; ++i) ws->best_asg[i] = asg[i];
            }
        }
    }
    if (lane == 0) {
        ws->n_complete += (unsigned long long)__popc(counted);
        if (A.P.max_candidates > 0 && ws->n_complete >= (unsigned long long)A.P.max_candidates) ws->overflow = 1;
    }
    __syncwarp();
}

// Lane 0: fold this warp's result of (a part of) a pair into the pair's slot.
__device__ __noinline__ void merge_into_slot(PairSlot *ps, const WarpState *ws, int m)
{
    spin_lock(&ps->lock);
    atomicAdd(&ps->n_complete, ws->n_complete);
    if (ws->overflow) sh_set(&ps->overflow, 1);
    if (ws->best_valid) {
        const double have = __longlong_as_double((long long)sh_get(&ps->best_rmsd_bits));
        bool better = !sh_get(&ps->best_valid) || ws->best_rmsd < have;
        if (!better && ws->best_rmsd == have) {
            for (int i = 0; i < m; ++i) {
                const uint32_t theirs = sh_get(&ps->best_asg[i]);
                if (ws->best_asg[i] != theirs) { better = ws->best_asg[i] < theirs; break; }
            }
        }
        if (better) {
            sh_set(&ps->best_valid, 1);
            sh_set(&ps->best_rmsd_bits, (unsigned long long)__double_as_longlong(ws->best_rmsd));
            for (int i = 0; i < m; ++i) sh_set(&ps->best_asg[i], ws->best_asg[i]);
        }
    }
    spin_unlock(&ps->lock);
}

// Lane 0: the merged result of a split pair back into the owner's warp state, for emit_hit.
__device__ __noinline__ void load_from_slot(PairSlot *ps, WarpState *ws, int m, long long max_candidates)
{
    const unsigned long long n = sh_get(&ps->n_complete);
    ws->best_valid = sh_get(&ps->best_valid);
    ws->best_rmsd = __longlong_as_double((long long)sh_get(&ps->best_rmsd_bits));
    ws->n_complete = n;
    ws->overflow = sh_get(&ps->overflow) || (max_candidates > 0 && n >= (unsigned long long)max_candidates);
    for (int i = 0; i < m; ++i) ws->best_asg[i] = sh_get(&ps->best_asg[i]);
}

// Give the untouched tail of the chunk just entered at level k (valid = its live slots, anchors
// compacted in ws) to idle warps, one partial per donation.  `first_free` = first compacted partial
// no expansion has touched.  Returns the remaining valid mask.
__device__ __noinline__ unsigned donate_tail(CtaShare *sh, uint32_t *Q, WarpState *ws, int wid, int t,
                                             int k, int base, unsigned valid, int first_free, int idle, int lane)
{
    const int owner = ws->for_owner >= 0 ? ws->for_owner : wid;     // a helper donates on its owner's behalf
    const bool fresh = ws->for_owner < 0 && !ws->donated;
    const int P = __popc(valid);
    int give = min(P - first_free, idle);          // warp-uniform: idle was read by one lane
    if (give <= 0) return valid;
    int at = 0;
    if (lane == 0) {
        spin_lock(&sh->ring_lock);
        at = sh_get(&sh->ring_count);
        give = min(give, kDonationRing - at);
        if (give <= 0) spin_unlock(&sh->ring_lock);
    }
    give = __shfl_sync(kFull, give, 0);
    at = __shfl_sync(kFull, at, 0);
    if (give <= 0) return valid;
    if (fresh && lane == 0) {        // first donation of this pair: an empty slot
        PairSlot *ps = &sh->slot[owner];
        sh_set(&ps->pending, 0); sh_set(&ps->best_valid, 0); sh_set(&ps->overflow, 0); sh_set(&ps->n_complete, 0ull);
        sh_set(&ps->best_rmsd_bits, (unsigned long long)__double_as_longlong(CUDART_INF));
    }
    unsigned gone = 0u;
    if (lane < give) {
        const int slot = (int)ws->vslot[P - give + lane];
        gone = 1u << slot;
        Donation *d = &sh->ring[at + lane];
        sh_set(&d->owner, owner); sh_set(&d->depth, k); sh_set(&d->t, t);
        uint32_t w = Q[queue_off(k) + base + slot];
        Q[queue_off(k) + base + slot] = w | kEntryDead;          // the owner will not look at it again
        for (int pos = k - 1; pos >= 0; --pos) {
            sh_set(&d->chain[pos], (uint32_t)entry_atom(w));
            if (pos > 0) w = Q[queue_off(pos) + entry_parent(w)];
        }
    }
    gone = __reduce_or_sync(kFull, gone);
    __threadfence_block();
    __syncwarp();
    if (lane == 0) {
        atomicAdd(&sh->slot[owner].pending, give);
        sh_set(&sh->ring_count, at + give);
        ws->donated = 1;
        spin_unlock(&sh->ring_lock);
    }
    __syncwarp();
    return valid & ~gone;
}

// Guard band: decide one entry of level k (new atom at plan position k-1) with the oracle's FP64
// expressions against every placed atom.
__device__ __noinline__ bool exact_validate(const DevLibrary &L, const Blob &S, const uint32_t *Q, int a0,
                                            int64_t p0, int k, uint32_t e, double cut, double dyn)
{
    const int oa = S.orig[entry_atom(e)];
    const double *row64 = L.pair_dist + p0 + ((k - 1) * (k - 2)) / 2;
    uint32_t w = e;
    for (int pos = k - 2; pos >= 0; --pos) {
        w = Q[queue_off(pos + 1) + entry_parent(w)];
        const double d = exact_dist(S.xyz64, oa, S.orig[entry_atom(w)]);
        const double delta = pair_delta(L, a0, k - 1, pos, cut, dyn);
        if (!(fabs(DSUB(d, row64[pos])) <= delta)) return false;
    }
    return true;
}

// Leader expansion through the uniform-grid cell list, used when the leader list is long (mode-100
// "any residue" atoms, very large assemblies): instead of testing every atom the type can bind,
// scan the cells that intersect the anchor's distance shell and test typing + distance there.
// Per partial the box of cells around its anchor is walked row by row (cells along x are
// contiguous in the cell-sorted atom array), 32 atoms per step.  Returns true when the next queue
// filled up; the resume point (partial, row, position, pending lanes) is kept in the warp state.
template <bool kStaged>
__device__ __noinline__ bool expand_cells(const View<kStaged> V, const Blob &S, WarpState *ws, uint32_t *Qn, int cap_next,
                                          int *n_next_io, int base, int P, const uint32_t *crow, float lo2, float hi2,
                                          int k, int lane, bool *done)
{
    const unsigned lt_mask = (1u << lane) - 1u;
    const float hi = sqrtf(hi2) + 1e-3f;
    int n_next = *n_next_io;
    int pidx = ws->cur[k], row = ws->cell->row[k], ipos = ws->cell->pos[k];
    unsigned todo = (unsigned)ws->todo[k];
    bool full = false;
    for (; pidx < P && !full; ++pidx, row = 0, ipos = -1) {
        const float4 an = ws->anchor[pidx];
        const int parent = base + (int)ws->vslot[pidx];
        const int ix0 = max(0, (int)floorf((an.x - hi - S.ox) / S.cell)), ix1 = min(S.nx - 1, (int)floorf((an.x + hi - S.ox) / S.cell));
        const int iy0 = max(0, (int)floorf((an.y - hi - S.oy) / S.cell)), iy1 = min(S.ny - 1, (int)floorf((an.y + hi - S.oy) / S.cell));
        const int iz0 = max(0, (int)floorf((an.z - hi - S.oz) / S.cell)), iz1 = min(S.nz - 1, (int)floorf((an.z + hi - S.oz) / S.cell));
        if (ix1 < ix0 || iy1 < iy0 || iz1 < iz0) continue;
        const int nry = iy1 - iy0 + 1, nrows = nry * (iz1 - iz0 + 1);
        for (; row < nrows && !full; ++row, ipos = -1) {
            const int iz = iz0 + row / nry, iy = iy0 + row % nry;
            const int c0 = (iz * S.ny + iy) * S.nx + ix0;
            const int beg = cell_idx(S.cell_start, S.wide, c0);
            const int end = cell_idx(S.cell_start, S.wide, c0 + ix1 - ix0 + 1);
            if (ipos < 0) ipos = beg;
            for (; ipos < end; ipos += 32) {
                if (n_next >= cap_next) { full = true; break; }
                const int i = ipos + lane;
                bool alive = i < end && (todo == 0u || ((todo >> lane) & 1u));
                todo = 0u;
                int a = 0;
                float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
                if (alive) {
                    a = cell_idx(S.cell_atoms, S.wide, i);
                    p = V.atom(a);
                    const unsigned kl = V.klass_of(p);
                    alive = (__ldg(crow + (kl >> 5)) >> (kl & 31u)) & 1u;
                }
                if (alive) {
                    const float dx = p.x - an.x, dy = p.y - an.y, dz = p.z - an.z;
                    const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                    alive = d2 >= lo2 && d2 <= hi2;
                }
                const unsigned sv = __ballot_sync(kFull, alive);
                if (sv) {
                    const int room = cap_next - n_next, rank = __popc(sv & lt_mask), cnt = __popc(sv);
                    if (alive && rank < room) Qn[n_next + rank] = make_entry(parent, a);
                    if (cnt > room) {
                        n_next = cap_next;
                        todo = __ballot_sync(kFull, alive && rank >= room);
                        full = true;
                        break;
                    }
                    n_next += cnt;
                }
            }
            if (full) break;
        }
        if (full) break;
    }
    __syncwarp();          // every lane has read the resume point above before lane 0 replaces it
    if (lane == 0) { ws->cur[k] = pidx; ws->cell->row[k] = row; ws->cell->pos[k] = ipos; ws->todo[k] = todo; }
    *n_next_io = n_next;
    *done = !full && pidx >= P;
    return full;
}

__device__ __forceinline__ int L_atoms(const DevLibrary &L, int t) { return L.atom_off[t + 1] - L.atom_off[t]; }

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ int log2_bucket_us(unsigned long long ns, int buckets)
{
    const unsigned long long us = ns >> 10;
    return min(buckets - 1, us ? 64 - __clzll(us) : 0);
}

struct LaneStats {
    unsigned long long sweeps, evals, exact;
};

// Enter level k (k atoms placed; the chunk = entries [base, base+size) of queue k, one per lane):
// validate entries that were only filtered so far, and record for every valid entry the anchor
// the expansion of this level needs.  Returns the mask of valid chunk slots.
template <bool kStats, bool kStaged>
__device__ __forceinline__ unsigned enter_level(const SearchArgs &A, const Blob &S, const View<kStaged> &V,
                                                uint32_t *Q, WarpState *ws,
                                                int a0, int p0, int m, int k, int base, int size,
                                                float cut32, bool dynamic, int t, int lane, LaneStats &st)
{
    const DevLibrary &L = A.L;
    const bool have = lane < size;
    uint32_t e = have ? Q[queue_off(k) + base + lane] : kEntryDead;
    bool alive = !(e & kEntryDead);
    const bool check = alive && !(e & kEntryValid);
    const int a = entry_atom(e);
    const int anchor_pos = k < m ? (int)L.plan_anchor[a0 + k] : -1;     // warp-uniform
    const float eps = V.eps;
    const float4 pa = V.atom(a);
    const float xa = pa.x, ya = pa.y, za = pa.z;
    float ax = xa, ay = ya, az = za;
    int aatom = a, ares = V.res_of(pa);
    bool border = false;
    const float *row = L.pair_dist32 + p0 + ((k - 1) * (k - 2)) / 2;
    uint32_t w = e;
    // a chunk that was validated on an earlier visit only needs its anchors again: stop the walk there
    const int stop = __any_sync(kFull, check) ? 0 : (anchor_pos < 0 ? k - 1 : anchor_pos);
    for (int pos = k - 2; pos >= stop; --pos) {
        w = Q[queue_off(pos + 1) + entry_parent(w)];
        const int b = entry_atom(w);
        const float4 pb = V.atom(b);
        const float xb = pb.x, yb = pb.y, zb = pb.z;
        if (pos == anchor_pos) { ax = xb; ay = yb; az = zb; aatom = b; ares = V.res_of(pb); }
        if (check) {
            const float dx = xa - xb, dy = ya - yb, dz = za - zb;
            const float d = fast_sqrt(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
            // (the FP64 thresholds are re-read where they are needed -- rare paths -- instead of
            // living in four registers across the whole search)
            const float delta = dynamic ? (float)pair_delta(L, a0, k - 1, pos, L.dist_cut[t], L.max_dyn[t]) : cut32;
            const float err = fabsf(d - __ldg(row + pos));
            if (err > delta + eps || b == a) alive = false;
            else if (err >= delta - eps) border = true;
            if (kStats) ++st.evals;
        }
    }
    if (check && alive && border) {
        if (kStats) ++st.exact;
        alive = exact_validate(L, S, Q, a0, p0, k, e, L.dist_cut[t], L.max_dyn[t]);
    }
    if (check) Q[queue_off(k) + base + lane] = e | (alive ? kEntryValid : kEntryDead);
    const unsigned valid = __ballot_sync(kFull, alive);
    if (alive && k < m) {
        int payload = aatom;
        if ((int)L.plan_src[a0 + k] >= 0) {        // same-residue level: the anchor's residue span
            const int rs = V.res_start(ares);
            payload = span_payload(rs, V.res_start(ares + 1) - rs);
        }
        const int rank = __popc(valid & ((1u << lane) - 1u));      // anchors are stored compacted
        ws->anchor[rank] = make_float4(ax, ay, az, __int_as_float(payload));
        ws->vslot[rank] = (unsigned char)lane;
    }
    __syncwarp();
    return valid;
}

// Search template t against the staged structure, or -- don_depth > 0 -- only the subtree below a
// donated partial assignment of don_depth atoms (lane j holds the atom of plan position j in
// don_atom) on behalf of warp don_owner.  Returns true when this warp gave parts of ITS pair away:
// its own part is merged into its PairSlot and the caller writes the hit once all parts are back.
template <bool kStats, bool kStaged, bool kCells>
__device__ __forceinline__ bool search_template(const SearchArgs &A, const Blob &S, const View<kStaged> &V, int s,
                                                const int4 rec, uint32_t *Q, WarpState *ws, int lane, LaneStats &st,
                                                CtaShare *sh, int wid, int don_depth, int don_owner, uint32_t don_atom)
{
    const DevLibrary &L = A.L;
    const int t = rec.x, a0 = rec.y, m = rec.z;
    const int p0 = rec.w;                // first pair of the template (the pair table holds < 2^31 entries)

    // a template with an empty leader list cannot match this structure
    if (don_depth == 0) {
        bool empty = false;
        if (lane < m) {
            const int src = L.plan_src[a0 + lane];
            if (src < 0) empty = V.lead_off(-src) == V.lead_off(-1 - src);
        }
        if (__any_sync(kFull, empty)) return false;
    }

    const double cut64 = L.dist_cut[t];
    const bool dynamic = L.max_dyn[t] != cut64;
    const float cut32 = (float)cut64;
    const float eps = V.eps;
    const unsigned lt_mask = (1u << lane) - 1u;

    if (lane <= m) { ws->n[lane] = 0; ws->chunk[lane] = 0; ws->cur[lane] = 0; ws->todo[lane] = 0ull; if (kCells) { ws->cell->row[lane] = 0; ws->cell->pos[lane] = -1; } }
    if (lane == 0) {
        ws->best_valid = 0; ws->overflow = 0; ws->n_complete = 0ull; ws->best_rmsd = CUDART_INF;
        ws->donated = 0; ws->for_owner = don_depth > 0 ? don_owner : -1;
    }
    __syncwarp();

    int k = 0, base = 0;
    unsigned valid = 1u;          // level 0: the empty partial
    bool entered = true;
    int visits = 0;               // level entries so far: a pair may donate once it has run for a while
    if (don_depth > 0) {
        // A donated subtree: the chain of placed atoms becomes one validated entry per level, every
        // ancestor marked as fully expanded, so the walk below never leaves the subtree.
        if (lane < don_depth) Q[queue_off(lane + 1)] = don_atom | kEntryValid;          // parent: slot 0
        if (lane >= 1 && lane <= don_depth) { ws->n[lane] = 1; ws->chunk[lane] = 1; ws->cur[lane] = lane < don_depth ? -1 : 0; }
        if (lane == 0) ws->cur[0] = -1;
        __syncwarp();
        k = don_depth;
        entered = false;
        valid = 0u;
    } else
    // Fast start.  Level 0 has nothing to test: its expansion copies the first leader list into
    // queue 1 and entering level 1 only derives anchors.  When the list fits one chunk (the usual
    // case: plans start with the rarest type) do both here, without the level machinery.
    if (m > 1) {
        const int src0 = L.plan_src[a0];
        const int lbase = V.lead_off(-1 - src0), B0 = V.lead_off(-src0) - lbase;
        if (B0 <= 32) {
            if (lane < B0) {
                const int a = V.lead(lbase + lane);
                const float4 p = V.atom(a);
                Q[queue_off(1) + lane] = (uint32_t)a | kEntryValid;          // parent: slot 0 of level 0
                int payload = a;
                if ((int)L.plan_src[a0 + 1] >= 0) {                           // position 1 shares this atom's residue
                    const int r = V.res_of(p), rs = V.res_start(r);
                    payload = span_payload(rs, V.res_start(r + 1) - rs);
                }
                ws->anchor[lane] = make_float4(p.x, p.y, p.z, __int_as_float(payload));
                ws->vslot[lane] = (unsigned char)lane;
            }
            if (lane == 0) { ws->n[1] = B0; ws->chunk[1] = B0; ws->cur[0] = -1; }
            if (kStats && lane == 0) {
                atomicAdd(A.O.stats + 40, (unsigned long long)B0);
                atomicAdd(A.O.stats + 72 + 1, (unsigned long long)B0);
                atomicAdd(A.O.stats + 104 + 1, 1ull);
            }
            __syncwarp();
            k = 1;
            valid = B0 >= 32 ? 0xffffffffu : (1u << B0) - 1u;
        }
    }
    for (;;) {
        if (ws->overflow) break;
        if (!entered) {
            const int size = ws->chunk[k];
            base = ws->n[k] - size;
            entered = true;
            if (ws->cur[k] < 0) {
                // back from the children of a chunk whose expansion had finished: nothing left to
                // do with it, so do not re-derive its anchors -- fall through and pop it
                valid = 0u;
            } else {
                valid = k == 0 ? 1u : enter_level<kStats, kStaged>(A, S, V, Q, ws, a0, p0, m, k, base, size, cut32, dynamic, t, lane, st);
            }
            if (kStats && lane == 0 && k > 0 && ws->cur[k] >= 0) {
                atomicAdd(A.O.stats + 72 + k, (unsigned long long)__popc(valid));
                atomicAdd(A.O.stats + 104 + k, 1ull);
            }
            // Split the pair: once it has run for a while and other warps of the CTA sit idle, give
            // them the partials of this chunk that no expansion has touched yet (the tail of the
            // compacted order, so the cursor of this level stays meaningful).
            int idle = 0;
#ifndef EMM_NO_DONATE
            // looked at every EMM_DONATE_STRIDE-th level entry only: the look is a shared-memory load
            // and a shuffle on the critical path of exactly the pairs that have many level entries
            if (!kCells && A.P.donate_after >= 0 && k > 0 && k < m && valid && ++visits > A.P.donate_after &&
                (visits & (EMM_DONATE_STRIDE - 1)) == 0) {
                if (lane == 0) idle = peek(&sh->idle);         // one reader: the branch below must be warp-uniform
                idle = __shfl_sync(kFull, idle, 0);
            }
#endif
            if (idle > 0) {
                const int cur = ws->cur[k];
                const int pend = ws->todo[k] != 0ull ? 1 : 0;
                int first_free;
                if ((int)L.plan_src[a0 + k] >= 0) first_free = (((cur + pend) << 5) + (1 << V.res_shift) - 1) >> V.res_shift;
                else first_free = (cur >> 5) == 0 ? (cur & 31) + 1 : 32;       // leader level: during its first row pair only
                const unsigned before = valid;
                valid = donate_tail(sh, Q, ws, wid, t, k, base, valid, max(first_free, 1), idle, lane);
                if (kStats && lane == 0 && valid != before) atomicAdd(A.O.stats + 12, (unsigned long long)__popc(before ^ valid));
            }
            if (k == m && valid) process_complete<kStaged>(A, S, t, Q, ws, base, valid, lane);
        }
        if (k < m && valid) {
            // ---------------- expand the chunk of level k: a cheap dense filter ----------------
            const int src = L.plan_src[a0 + k];
            const int P = __popc(valid);
            int n_next = ws->n[k + 1];
            int cur = ws->cur[k];
            unsigned long long todo = ws->todo[k];
            uint32_t *Qn = Q + queue_off(k + 1);
            const int cap_next = queue_cap(k + 1);
            // squared acceptance band of the anchor constraint (widened by eps and FP32 rounding)
            float lo2 = 0.f, hi2 = 3.0e38f;
            if (k > 0) {
                const float dt = __ldg(L.anchor_dist32 + a0 + k);
                const float rej = (dynamic ? (float)pair_delta(L, a0, k, (int)L.plan_anchor[a0 + k], L.dist_cut[t], L.max_dyn[t]) : cut32) + eps;
                const float lo = fmaxf(dt - rej, 0.f), hi = dt + rej;
                lo2 = lo * lo * 0.999999f;
                hi2 = hi * hi * 1.000001f;
            }
            bool full = false, done = false;
            const int n_before = n_next;
            // push the lanes of `sv` (ballot of `alive`) into the next queue; returns the lanes that
            // did not fit (0 = all pushed)
            auto push = [&](unsigned sv, bool alive, int a, int parent) -> unsigned {
                const int room = cap_next - n_next, rank = __popc(sv & lt_mask), cnt = __popc(sv);
                if (alive && rank < room) Qn[n_next + rank] = make_entry(parent, a);
                if (cnt > room) {
                    n_next = cap_next;
                    return __ballot_sync(kFull, alive && rank >= room);
                }
                n_next += cnt;
                return 0u;
            };
            if (src < 0) {
                // leader position: every lane holds up to two candidates of the leader list in
                // registers (rows r, r+1) and the warp loops over the chunk's partials
                const int lbase = V.lead_off(-1 - src);
                const int B = V.lead_off(-src) - lbase;
                const int nrows = (B + 31) >> 5;
                int r = cur >> 5, pidx = cur & 31;         // cursor = candidate row * 32 + partial (P <= 32)
                const bool by_cells = kCells && k > 0 && B >= A.P.cell_threshold;
                if (by_cells) {
                    const uint32_t *crow = L.compat + (size_t)L.plan_ttype[a0 + k] * L.class_words_cap;
                    __syncwarp();
                    full = expand_cells<kStaged>(V, S, ws, Qn, cap_next, &n_next, base, P, crow, lo2, hi2, k, lane, &done);
                    __syncwarp();
                    cur = ws->cur[k];
                    todo = ws->todo[k];
                    r = nrows;                              // skip the list loop below
                }
                while (r < nrows) {
                    const int c0 = (r << 5) + lane, c1 = c0 + 32;
                    const bool have0 = c0 < B, have1 = c1 < B;
                    const bool two = ((r + 1) << 5) < B;       // warp-uniform: a second row exists
                    const int a_0 = have0 ? V.lead(lbase + c0) : 0, a_1 = have1 ? V.lead(lbase + c1) : 0;
                    const float4 p_0 = V.atom(a_0), p_1 = V.atom(a_1);
                    const float x0 = p_0.x, y0 = p_0.y, z0 = p_0.z;
                    const float x1 = p_1.x, y1 = p_1.y, z1 = p_1.z;
                    // (row pair, partial) iterations, specialised at compile time on: a second candidate
                    // row exists / level 0 (no anchor) / resuming a partly pushed iteration.
                    // test: the squared-distance band of both rows against one partial's anchor
                    auto test = [&](auto TWO, const float4 an, bool &alive0, bool &alive1) {
                        float dx = x0 - an.x, dy = y0 - an.y, dz = z0 - an.z;
                        const float d0 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        alive0 = alive0 && d0 >= lo2 && d0 <= hi2;
                        if (TWO()) {
                            dx = x1 - an.x; dy = y1 - an.y; dz = z1 - an.z;
                            const float d1 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                            alive1 = alive1 && d1 >= lo2 && d1 <= hi2;
                        }
                        if (kStats) st.evals += (int)have0 + (int)have1;
                    };
                    // commit: push the survivors of partial `pidx`; true when the next queue could not
                    // take all of them (the rest is remembered in `todo`)
                    auto commit = [&](auto TWO, auto ROOT, bool alive0, bool alive1) -> bool {
                        if (kStats && lane == 0) ++st.sweeps;
                        const unsigned sv0 = __ballot_sync(kFull, alive0);
                        const unsigned sv1 = TWO() ? __ballot_sync(kFull, alive1) : 0u;
                        if (sv0 | sv1) {
                            const int parent = ROOT() ? 0 : base + (int)ws->vslot[pidx];
                            unsigned left0 = 0u, left1 = 0u;
                            if (sv0) left0 = push(sv0, alive0, a_0, parent);
                            if (TWO() && sv1) left1 = left0 ? sv1 : push(sv1, alive1, a_1, parent);
                            if (left0 | left1) {
                                todo = (unsigned long long)left0 | ((unsigned long long)left1 << 32);
                                return true;
                            }
                        }
                        return false;
                    };
                    auto one = [&](auto TWO, auto ROOT, auto TODO) -> bool {
                        bool alive0 = have0, alive1 = TWO() && have1;
                        if (TODO()) {
                            alive0 = alive0 && ((todo >> lane) & 1ull);
                            alive1 = alive1 && ((todo >> (32 + lane)) & 1ull);
                            todo = 0ull;
                        }
                        if (!ROOT()) test(TWO, ws->anchor[pidx], alive0, alive1);
                        return commit(TWO, ROOT, alive0, alive1);
                    };
                    auto sweep = [&](auto TWO, auto ROOT) -> bool {
                        if (todo) {                       // finish the iteration a full queue interrupted
                            if (n_next >= cap_next || one(TWO, ROOT, std::true_type{})) return true;
                            ++pidx;
                        }
                        if (!ROOT()) {
                            // two partials per iteration: independent dependency chains, and one vote
                            // rejects both when (as usual) nothing survives
                            while (pidx + 1 < P) {
                                if (n_next >= cap_next) return true;
                                const float4 an_a = ws->anchor[pidx], an_b = ws->anchor[pidx + 1];
                                bool a0 = have0, a1 = TWO() && have1, b0 = have0, b1 = TWO() && have1;
                                test(TWO, an_a, a0, a1);
                                test(TWO, an_b, b0, b1);
                                if (__any_sync(kFull, a0 | a1 | b0 | b1)) {
                                    if (commit(TWO, ROOT, a0, a1)) return true;
                                    ++pidx;
                                    if (n_next >= cap_next) return true;      // resumes by re-testing this partial
                                    if (commit(TWO, ROOT, b0, b1)) return true;
                                    ++pidx;
                                } else {
                                    if (kStats && lane == 0) st.sweeps += 2;
                                    pidx += 2;
                                }
                            }
                        }
                        for (; pidx < P; ++pidx)
                            if (n_next >= cap_next || one(TWO, ROOT, std::false_type{})) return true;
                        return false;
                    };
                    if (k == 0) full = two ? sweep(std::true_type{}, std::true_type{}) : sweep(std::false_type{}, std::true_type{});
                    else full = two ? sweep(std::true_type{}, std::false_type{}) : sweep(std::false_type{}, std::false_type{});
                    if (full) break;
                    pidx = 0;
                    r += 2;
                }
                if (!by_cells) {
                    cur = (r << 5) + pidx;
                    done = r >= nrows;
                }
            } else {
                // same-residue position: lanes cover (partial, residue slot) items
                const int shift = V.res_shift;
                const int n_it = ((P << shift) + 31) >> 5;
                const uint32_t *crow = L.compat + (size_t)L.plan_ttype[a0 + k] * L.class_words_cap;
                int it = cur;
                for (; it < n_it; ++it) {
                    if (n_next >= cap_next) { full = true; break; }
                    const int item = (it << 5) + lane;
                    const int pidx = item >> shift, sidx = item & ((1 << shift) - 1);
                    bool alive = pidx < P && (todo == 0ull || ((todo >> lane) & 1ull));
                    const float4 an = ws->anchor[alive ? pidx : 0];
                    const int payload = __float_as_int(an.w);
                    const int a = (payload & (int)kAtomMask) + sidx;
                    alive = alive && sidx < (int)((unsigned)payload >> kAtomBits);
                    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (alive) {
                        p = V.atom(a);
                        const unsigned kl = V.klass_of(p);
                        alive = (__ldg(crow + (kl >> 5)) >> (kl & 31u)) & 1u;
                    }
                    if (alive) {
                        const float dx = p.x - an.x, dy = p.y - an.y, dz = p.z - an.z;
                        const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                        alive = d2 >= lo2 && d2 <= hi2;
                        if (kStats) ++st.evals;
                    }
                    if (kStats && lane == 0) ++st.sweeps;
                    const unsigned sv = __ballot_sync(kFull, alive);
                    todo = 0ull;
                    if (sv) {
                        const int parent = alive ? base + (int)ws->vslot[pidx] : 0;
                        const unsigned left = push(sv, alive, a, parent);
                        if (left) { todo = left; full = true; break; }
                    }
                }
                cur = it;
                done = it >= n_it;
            }
            __syncwarp();          // every lane has read the level state before lane 0 updates it
            if (lane == 0) { ws->n[k + 1] = n_next; ws->cur[k] = done ? -1 : cur; ws->todo[k] = todo; }   // -1: finished
            if (kStats && lane == 0) atomicAdd(A.O.stats + 40 + k, (unsigned long long)(n_next - n_before));
            if (full || (done && n_next > 0)) {
                // descend: the last <= 32 entries of the next level form its chunk
                ++k;
                if (lane == 0) { ws->chunk[k] = min(32, n_next); ws->cur[k] = 0; ws->todo[k] = 0ull; if (kCells) { ws->cell->row[k] = 0; ws->cell->pos[k] = -1; } }
                __syncwarp();
                entered = false;
                continue;
            }
            __syncwarp();
        }
        // the chunk of level k is exhausted (or dead, or its complete assignments are processed):
        // pop it, and with it every ancestor chunk whose own expansion had already finished
        bool finished = false;
        for (;;) {
            if (k == 0) { finished = true; break; }
            if (lane == 0) {
                ws->n[k] = base;
                ws->chunk[k] = min(32, base);
                ws->cur[k] = 0;
                ws->todo[k] = 0ull;
                if (kCells) { ws->cell->row[k] = 0; ws->cell->pos[k] = -1; }
            }
            __syncwarp();
            if (base != 0) break;                 // more entries at this level: enter its next chunk
            --k;                                  // back to the parent: its chunk, cursor and todo are intact
            if (ws->cur[k] >= 0) break;           // the parent still has candidates to try: re-enter it
            if (k == 0) { finished = true; break; }       // the root was expanded completely
            base = ws->n[k] - ws->chunk[k];
            __syncwarp();                         // every lane has read level k before lane 0 rewrites it above
        }
        if (finished) break;
        entered = false;
    }
    if (kStats && lane == 0 && ws->n_complete) atomicAdd(A.O.stats + 4, ws->n_complete);
    const int for_owner = ws->for_owner;
    if (for_owner >= 0) {                    // a helper: the result goes to the owner's slot
        if (lane == 0) {
            merge_into_slot(&sh->slot[for_owner], ws, m);
            __threadfence_block();
            atomicSub(&sh->slot[for_owner].pending, 1);
        }
        __syncwarp();
        return false;
    }
    if (ws->donated) {                       // the owner of a split pair: its own part joins the others
        if (lane == 0) merge_into_slot(&sh->slot[wid], ws, m);
        __syncwarp();
        return true;
    }
    if (ws->best_valid && lane == 0) emit_hit(A, S, s, t, ws);
    __syncwarp();
    return false;
}

template <bool kStats, bool kStaged, bool kCells>
__global__ void __launch_bounds__(kSearchThreads, 1)
emm_search_kernel(const __grid_constant__ SearchArgs A)
{
    __shared__ int s_item, s_next_pos;
    __shared__ Blob s_blob;
    __shared__ CtaShare s_share;                 // static: addressed by a constant, costs no register
    __shared__ __align__(8) uint64_t s_bar;      // mbarrier the TMA bulk copy of a blob completes on

    const int tid = threadIdx.x;
    int lane = tid & 31, wid = tid >> 5;
    // through a shuffle, so that ptxas keeps both in registers instead of re-reading SR_TID (S2R is
    // slow) wherever register pressure tempts it to rematerialise them
    lane = __shfl_sync(0xffffffffu, lane, lane);
    wid = __shfl_sync(0xffffffffu, wid, lane);
    unsigned stage_parity = 0;
    if (kStaged) {
        if (tid == 0) mbar_init(&s_bar, 1);
        __syncthreads();
    }
    const SearchParams &P = A.P;
    const int qwords = queue_off(P.levels);      // queue words per warp
    uint32_t *Q = reinterpret_cast<uint32_t *>(g_smem + P.blob_cap) + (size_t)wid * qwords;
    WarpState *ws = reinterpret_cast<WarpState *>(g_smem + P.blob_cap + (size_t)kSearchWarps * qwords * 4) + wid;
    CtaShare *const sh = &s_share;
    if (kCells) {
        if (lane == 0)
            ws->cell = reinterpret_cast<CellState *>(g_smem + P.blob_cap + (size_t)kSearchWarps * qwords * 4 +
                                                     (size_t)kSearchWarps * sizeof(WarpState)) + wid;
        __syncwarp();
    }

    LaneStats st = {0ull, 0ull, 0ull};
    unsigned long long st_pairs = 0, st_staged = 0, st_global = 0;

    for (;;) {
        if (tid == 0) s_item = (int)atomicAdd(A.O.work_counter, 1u);
        __syncthreads();
        const int item = s_item;
        if (item >= P.n_items) break;
        int s, pos0, pos1, stride;
        if (P.two_phase) {
            const bool heavy = item < P.n_structures;
            s = heavy ? item : item - P.n_structures;
            pos0 = heavy ? 0 : P.n_heavy;
            pos1 = heavy ? P.n_heavy : P.n_sched;
            stride = 1;
        } else {
            s = item / P.n_chunks;
            pos0 = item - s * P.n_chunks;
            pos1 = P.n_sched;
            stride = P.n_chunks;
        }
        if (A.ids) s = __ldg(A.ids + s);
        const unsigned char *gblob = A.B.blob + A.B.blob_off[s];
        const BlobHeader hdr = *reinterpret_cast<const BlobHeader *>(gblob);
        const bool run = hdr.status == 0 && hdr.n_kept > 0 && pos0 < pos1 && !(A.skip && A.skip[s]);
        unsigned long long item_t0 = 0;
        if (kStats && tid == 0) item_t0 = global_ns();
        if (run) {
            if (kStaged) {
                // host guarantees staged_bytes <= blob_cap for every structure of the batch; one
                // thread hands the copy to the TMA engine, everyone waits on the mbarrier below
                if (tid == 0) bulk_load(g_smem, gblob, (unsigned)hdr.staged_bytes, &s_bar);
                if (kStats && tid == 0) st_staged += hdr.staged_bytes;
            } else if (kStats && tid == 0) {
                ++st_global;
            }
            if (tid == 0) {
                s_next_pos = pos0;
                sh->ring_lock = 0; sh->ring_count = 0; sh->idle = 0; sh->in_loop = kSearchWarps;
                for (int w = 0; w < kSearchWarps; ++w) { sh->slot[w].lock = 0; sh->slot[w].pending = 0; }
                const int64_t abase = A.B.atom_off[s];
                s_blob.orig = reinterpret_cast<const int32_t *>(gblob + hdr.off_orig);
                s_blob.xyz64 = A.B.xyz + 3 * abase;
                s_blob.chain = A.B.chain ? A.B.chain + abase : nullptr;
                s_blob.atom_id = A.B.atom_id ? A.B.atom_id + abase : nullptr;
                s_blob.cell_start = gblob + hdr.off_cellstart;
                s_blob.cell_atoms = gblob + hdr.off_cellatoms;
                s_blob.wide = hdr.wide;
                s_blob.nx = hdr.nx; s_blob.ny = hdr.ny; s_blob.nz = hdr.nz;
                s_blob.cell = hdr.cell; s_blob.ox = hdr.ox; s_blob.oy = hdr.oy; s_blob.oz = hdr.oz;
            }
            if (kStaged) {
                mbar_wait(&s_bar, stage_parity);
                stage_parity ^= 1u;
            }
            __syncthreads();
            View<kStaged> V;
            V.gbase = gblob;
            V.off_atom = hdr.off_atom;
            V.off_resstart = hdr.off_resstart; V.off_leadoff = hdr.off_leadoff;
            V.off_lead = hdr.off_lead; V.res_shift = hdr.res_shift; V.eps = hdr.eps; V.wide = hdr.wide != 0;
            // Per-warp work loop: take templates while there are any; a warp that split its pair helps
            // with donated subtrees until all parts of the pair are back and then writes the hit; a
            // warp out of templates helps until every warp of the CTA has left the loop.
            bool fetching = true, counted_idle = false;
            int owning = -1;
            unsigned nap = EMM_NAP_NS;        // ns between two looks at the shared state while waiting; backs off
            for (;;) {
                int t = -1, depth = 0, owner = 0;
                int4 rec = make_int4(-1, 0, 0, 0);
                uint32_t atom = 0u;
                if (owning >= 0 || !fetching) {
                    // lane 0 looks at the shared state and decides for the warp:
                    // >= 0 take ring record idx | -1 write my split pair's hit | -2 leave | -3 wait
                    int idx = -3;
                    if (lane == 0) {
                        if (peek(&sh->ring_count) > 0) {
                            spin_lock(&sh->ring_lock);
                            if (sh_get(&sh->ring_count) > 0) idx = atomicSub(&sh->ring_count, 1) - 1;   // the lock is kept until the record is copied
                            else spin_unlock(&sh->ring_lock);
                        }
                        if (idx < 0) {
                            if (owning >= 0) { if (peek(&sh->slot[wid].pending) == 0) idx = -1; }
                            else if (peek(&sh->in_loop) == 0) idx = -2;
                        }
                    }
                    idx = __shfl_sync(kFull, idx, 0);
                    if (idx >= 0) {
                        Donation *d = &sh->ring[idx];
                        t = sh_get(&d->t); depth = sh_get(&d->depth); owner = sh_get(&d->owner); atom = sh_get(&d->chain[lane]);
                        rec = make_int4(t, A.L.atom_off[t], L_atoms(A.L, t), (int)A.L.pair_off[t]);
                        __syncwarp();
                        if (lane == 0) {
                            spin_unlock(&sh->ring_lock);
                            if (counted_idle) atomicSub(&sh->idle, 1);
                        }
                        counted_idle = false;
                        nap = EMM_NAP_NS;
                    } else if (idx == -1) {
                        // every part of my split pair is merged: write its hit, go back to the templates
                        __threadfence_block();
                        const int mo = L_atoms(A.L, owning);
                        if (lane == 0) {
                            if (counted_idle) atomicSub(&sh->idle, 1);
                            load_from_slot(&sh->slot[wid], ws, mo, P.max_candidates);
                            if (ws->best_valid) emit_hit(A, s_blob, s, owning, ws);
                        }
                        counted_idle = false;
                        owning = -1;
                        nap = EMM_NAP_NS;
                        __syncwarp();
                        continue;
                    } else if (idx == -2) {
                        if (lane == 0 && counted_idle) atomicSub(&sh->idle, 1);
                        break;                       // every pair of this item is finished
                    } else {
                        // Waiting costs issue slots the searching warps need: sleep, and sleep longer the
                        // longer nothing happens (a donation waits at most a few microseconds; it is only
                        // made by pairs that have already run for tens of microseconds).
                        // Only a warp that is OUT of templates advertises itself as idle: an owner waiting
                        // for the parts of its pair resumes in a moment, and counting it would make every
                        // other long pair split as well (a cascade of one-partial searches for the rest of
                        // the item: measured 5 % on the whole bench).  It still helps while it waits.
                        if (!counted_idle && owning < 0) { if (lane == 0) atomicAdd(&sh->idle, 1); counted_idle = true; }
                        if (kStats && lane == 0 && owning < 0) atomicAdd(A.O.stats + 14, (unsigned long long)nap);    // warp-ns spent waiting
                        // (ncu, round 2: with a fixed 200 ns nap the owners of split pairs alone issued 5 % of
                        // the kernel's instructions while waiting for parts that run for milliseconds)
                        __nanosleep(nap);
                        nap = min(nap * 2u, owning >= 0 ? 4000u : 8000u);
                        continue;
                    }
                } else {
                    int pos = pos1;
                    if (lane == 0) pos = atomicAdd(&s_next_pos, stride);
                    pos = __shfl_sync(kFull, pos, 0);
                    if (pos < pos1) rec = __ldg(A.sched + pos);        // same address in every lane: one broadcast load
                    t = rec.x;
                    if (t < 0) {
                        // No splitting in this launch: nobody will ever need help, so park at the barrier
                        // below like a plain loop would (a waiting warp must not compete for issue slots
                        // with the warps that are still searching).
#ifdef EMM_NO_DONATE
                        break;
#endif
                        if (P.donate_after < 0) break;
                        fetching = false;
                        if (lane == 0) atomicSub(&sh->in_loop, 1);
                        continue;
                    }
                }
                unsigned long long pair_t0 = 0;
                if (kStats) pair_t0 = global_ns();
                const bool split = search_template<kStats, kStaged, kCells>(A, s_blob, V, s, rec, Q, ws, lane, st, sh, wid,
                                                                            depth, owner, atom);
                if (split) { owning = t; nap = 200u; }
                if (kStats && lane == 0) {
                    if (depth == 0) ++st_pairs;
                    const unsigned long long dt = global_ns() - pair_t0;
                    atomicMax(A.O.stats + 13, ((dt >> 10) << 24) | (unsigned long long)t);     // slowest pair (or part of one): us, template
                    atomicAdd(A.O.stats + 28 + log2_bucket_us(dt, 12), 1ull);
                }
            }
        }
        unsigned long long warp_done = 0;
        if (kStats) warp_done = global_ns();
        __syncthreads();
        if (kStats && lane == 0 && run) atomicAdd(A.O.stats + 14, global_ns() - warp_done);    // warp-ns spent waiting
        if (kStats && tid == 0 && run) {
            const unsigned long long dt = global_ns() - item_t0;
            atomicMax(A.O.stats + 8, ((dt >> 10) << 24) | (unsigned long long)s);              // slowest structure: us, index
            atomicAdd(A.O.stats + 15, dt * kSearchWarps);                                      // warp-ns available
            atomicAdd(A.O.stats + 16 + min(11, (int)(dt >> 21)), 1ull);                        // ~2 ms buckets
        }
    }
    if (kStats && tid == 0) {
        const unsigned long long now = global_ns();
        atomicMax(A.O.stats + 11, now);                                   // last CTA out
        atomicMax(A.O.stats + 10, ~now);                                  // first CTA out, as max of the complement
    }
    if (kStats) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            st.evals += __shfl_xor_sync(kFull, st.evals, o);
            st.exact += __shfl_xor_sync(kFull, st.exact, o);
        }
        if (lane == 0) {
            atomicAdd(A.O.stats + 0, st_pairs);
            atomicAdd(A.O.stats + 1, st.sweeps);
            atomicAdd(A.O.stats + 2, st.evals);
            atomicAdd(A.O.stats + 3, st.exact);
            atomicAdd(A.O.stats + 6, st_staged);
            atomicAdd(A.O.stats + 7, st_global);
        }
    }
}

__global__ void emm_skip_snapshot_kernel(int n, int mode, const int *any, const int *pass, unsigned char *skip)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) skip[i] = (mode == 1 ? pass[i] : any[i]) > 0;
}

size_t search_smem_bytes(int blob_cap, int levels, bool cells)
{
    return (size_t)blob_cap + (size_t)kSearchWarps * queue_off(levels) * 4 + (size_t)kSearchWarps * sizeof(WarpState) +
           (cells ? (size_t)kSearchWarps * sizeof(CellState) : 0);
}

// shared memory a CTA needs besides the staged blob, whichever kernel runs: the dynamic part above
// (with the cell-list state) + the static CtaShare
size_t search_fixed_smem(int levels) { return search_smem_bytes(0, levels, true) + sizeof(CtaShare); }

// largest staged blob that keeps a CTA of the default kernels within 196 KB of shared memory (the
// carve-out step below which the SM keeps 60 KB of L1); 0 if the queues alone do not fit
size_t search_soft_cap(int levels)
{
    const size_t keep = 196 * 1024, used = search_smem_bytes(0, levels, false) + sizeof(CtaShare) + 512;
    return keep > used ? keep - used : 0;
}

cudaError_t configure_search(int smem_bytes)
{
    const auto attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
    cudaError_t e = cudaFuncSetAttribute(emm_search_kernel<false, true, false>, attr, smem_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(emm_search_kernel<false, false, false>, attr, smem_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(emm_search_kernel<true, true, false>, attr, smem_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(emm_search_kernel<true, false, false>, attr, smem_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(emm_search_kernel<false, true, true>, attr, smem_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(emm_search_kernel<false, false, true>, attr, smem_bytes);
    return e;
}

void launch_skip_snapshot(int n, int mode, const int *any, const int *pass, unsigned char *skip, cudaStream_t stream)
{
    if (n > 0) emm_skip_snapshot_kernel<<<(n + 255) / 256, 256, 0, stream>>>(n, mode, any, pass, skip);
}

void launch_search(const DevLibrary &L, const DevBatch &B, const SearchParams &P, const SearchOut &O,
                   const unsigned char *skip, const int4 *sched, const int *ids, bool stats, bool staged, int grid, size_t smem, cudaStream_t stream)
{
    if (P.n_items <= 0) return;
    SearchArgs A;
    A.L = L; A.B = B; A.P = P; A.O = O; A.skip = skip; A.sched = sched; A.ids = ids;
    const bool cells = P.cell_threshold > 0;       // the cell-list path is a separate instantiation
    if (cells && staged) emm_search_kernel<false, true, true><<<grid, kSearchThreads, smem, stream>>>(A);
    else if (cells) emm_search_kernel<false, false, true><<<grid, kSearchThreads, smem, stream>>>(A);
    else if (stats && staged) emm_search_kernel<true, true, false><<<grid, kSearchThreads, smem, stream>>>(A);
    else if (stats) emm_search_kernel<true, false, false><<<grid, kSearchThreads, smem, stream>>>(A);
    else if (staged) emm_search_kernel<false, true, false><<<grid, kSearchThreads, smem, stream>>>(A);
    else emm_search_kernel<false, false, false><<<grid, kSearchThreads, smem, stream>>>(A);
}

}  // namespace emm
