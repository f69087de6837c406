// emm_prepare.cu -- per-structure preparation kernel (north_star subsystems 1 and 2).
//
// One CTA per structure turns the uploaded SoA columns into a compact "structure blob":
//   * drops masked atoms (bfactor < conservation cutoff, i.e. pyjess Molecule.conserved, call
//     site enzymm/jess_run.py:541-542) and atoms whose typing class binds no template atom;
//   * centres coordinates on the bounding-box centre and stores them as FP32 (the search kernel
//     decides in FP32 inside a rigorous guard band `eps` and re-evaluates in FP64 otherwise);
//   * builds the residue CSR (same-residue rule, SURVEY.md 8c rule 4) and, per leader type, the
//     ascending list of atoms that type can bind (this replaces Jess's per-template kd-tree /
//     annulus candidate search: typing is ~50x more selective than geometry at this scale).
//
// Roofline: HBM-bound streaming pass.  Algorithmic bytes per atom: 24 (xyz f64) + 2 (class) +
// 4 (residue) + 4 (bfactor) read, <= 22 + 2*lists written.
#include "emm_device.cuh"

namespace emm {

__device__ __forceinline__ int warp_excl_prefix(unsigned ballot, int lane)
{
    return __popc(ballot & ((1u << lane) - 1u));
}

// Exclusive prefix of `flag` over the block (in thread order) + block total.
__device__ __forceinline__ int block_excl_scan_flag(bool flag, int *warp_tot, int *total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned b = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) warp_tot[wid] = __popc(b);
    __syncthreads();
    int before = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kPrepThreads / 32; ++w) {
        const int c = warp_tot[w];
        if (w < wid) before += c;
        tot += c;
    }
    __syncthreads();
    *total = tot;
    return before + warp_excl_prefix(b, lane);
}

// Index arrays of a blob hold 16-bit entries, or 32-bit ones for "wide" structures (> 65 535 atoms).
__device__ __forceinline__ void put_idx(void *base, bool wide, int64_t i, unsigned v)
{
    if (wide) reinterpret_cast<uint32_t *>(base)[i] = v;
    else reinterpret_cast<uint16_t *>(base)[i] = (uint16_t)v;
}

__device__ __forceinline__ unsigned get_idx(const void *base, bool wide, int64_t i)
{
    return wide ? reinterpret_cast<const uint32_t *>(base)[i] : (unsigned)reinterpret_cast<const uint16_t *>(base)[i];
}

__device__ __forceinline__ double block_reduce_minmax(double v, bool is_max, double *scratch)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmax(v, other) : fmin(v, other);
    }
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double r = scratch[0];
#pragma unroll
    for (int w = 1; w < kPrepThreads / 32; ++w) r = is_max ? fmax(r, scratch[w]) : fmin(r, scratch[w]);
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kPrepThreads)
emm_prepare_kernel(DevLibrary L, DevBatch B, float cutoff, int build_cells,
                   unsigned long long *stats, unsigned long long *bad)
{
    __shared__ int s_warp_tot[kPrepThreads / 32];
    __shared__ double s_scratch[kPrepThreads / 32];
    __shared__ int s_bad, s_maxres;
    __shared__ unsigned s_lead_cnt[1024];   // counts, then offsets (n_leader <= 1023 checked on host)
    __shared__ int s_hist[1024];            // kept atoms per typing class (at most 1024 classes)
    __shared__ int s_cell[kMaxCells + 1];   // uniform grid: counts, then start offsets / fill cursors

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const bool use_mask = (B.bfactor != nullptr) && (cutoff > 0.0f);

    for (int s = blockIdx.x; s < B.n_structures; s += gridDim.x) {
        const int64_t base = B.atom_off[s];
        const int N = (int)(B.atom_off[s + 1] - base);
        unsigned char *blob = B.blob + B.blob_off[s];
        const int64_t bound = B.blob_off[s + 1] - B.blob_off[s];
        const int Nb = B.kept_bound[s];          // atoms that can stay at all: what the host sized this blob for
        const int off_orig = (int)(bound - align16(4 * (int64_t)Nb));
        int32_t *orig = reinterpret_cast<int32_t *>(blob + off_orig);
        // compact copy of the kept atoms' classes: pass D scans it once per leader type
        uint16_t *bklass = reinterpret_cast<uint16_t *>(blob + off_orig - align16(2 * (int64_t)Nb));
        const uint16_t *klass_in = B.klass + base;
        const int32_t *res_in = B.residue + base;
        const double *xyz = B.xyz + 3 * base;

        if (tid == 0) { s_bad = 0; s_maxres = 0; }
        __syncthreads();

        // ---- pass A: ordered compaction of kept atoms -----------------------------------------
        int n_kept = 0;
        for (int chunk = 0; chunk < N; chunk += kPrepThreads) {
            const int a = chunk + tid;
            bool keep = false;
            if (a < N) {
                keep = klass_in[a] != 0;
                if (keep && use_mask) keep = B.bfactor[base + a] >= cutoff;
            }
            int tot;
            const int pre = block_excl_scan_flag(keep, s_warp_tot, &tot);
            if (keep) orig[n_kept + pre] = a;
            n_kept += tot;
        }
        __syncthreads();
        int status = 0;
        const bool wide = is_wide(N);
        const int ib = wide ? 4 : 2;
        if (n_kept > (int)kAtomMask) { status = 2; n_kept = 0; }

        // ---- pass B: bounding box -> centre, extent, guard band ---------------------------------
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int i = tid; i < n_kept; i += kPrepThreads) {
            const double *p = xyz + 3 * (int64_t)orig[i];
#pragma unroll
            for (int c = 0; c < 3; ++c) { lo[c] = fmin(lo[c], p[c]); hi[c] = fmax(hi[c], p[c]); }
        }
        double ctr[3], ext[3], half = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double mn = block_reduce_minmax(lo[c], false, s_scratch);
            const double mx = block_reduce_minmax(hi[c], true, s_scratch);
            ctr[c] = n_kept ? 0.5 * (mn + mx) : 0.0;
            ext[c] = n_kept ? mx - mn : 0.0;
            half = fmax(half, 0.5 * ext[c]);
        }
        // |d_fp32 - d_exact| <= ~36 * 2^-24 * half (DESIGN.md "guard band"); 64 gives ~1.8x margin
        const float eps = (float)(64.0 * 5.9604644775390625e-08 * fmax(half, 16.0) + 1e-5);

        // ---- layout of the staged part ------------------------------------------------------------
        // one 16-byte record per atom -- x, y, z and (res_of << 10 | klass) in w -- so that the search
        // kernel gets everything it needs about an atom with one 128-bit shared-memory load
        const int off_atom = (int)sizeof(BlobHeader);
        const int off_resstart = off_atom + 16 * n_kept;
        float4 *atoms = reinterpret_cast<float4 *>(blob + off_atom);
        void *res_start = blob + off_resstart;

        // ---- pass C: coordinates, classes, residue CSR ---------------------------------------------
        int n_res = 0;
        for (int chunk = 0; chunk < n_kept; chunk += kPrepThreads) {
            const int i = chunk + tid;
            bool starts = false;
            int a = 0;
            if (i < n_kept) {
                a = orig[i];
                if (i == 0) {
                    starts = true;
                } else {
                    const int rp = res_in[orig[i - 1]], rc = res_in[a];
                    starts = rc != rp;
                    if (rc < rp) s_bad = 1;
                }
            }
            int tot;
            const int pre = block_excl_scan_flag(starts, s_warp_tot, &tot);
            // residue id of atom i = (#starts at or before i) - 1
            if (i < n_kept) {
                const int rid = n_res + pre + (starts ? 1 : 0) - 1;
                const double *p = xyz + 3 * (int64_t)a;
                atoms[i] = make_float4((float)(p[0] - ctr[0]), (float)(p[1] - ctr[1]), (float)(p[2] - ctr[2]),
                                       __uint_as_float(((uint32_t)rid << kClassBits) | ((uint32_t)klass_in[a] & kClassMask)));
                bklass[i] = klass_in[a];
                if (starts) put_idx(res_start, wide, rid, (unsigned)i);
            }
            n_res += tot;
        }
        __syncthreads();
        if (tid == 0) put_idx(res_start, wide, n_res, (unsigned)n_kept);
        __syncthreads();
        if (s_bad) status = 1;
        int local_max = 0;
        for (int r = tid; r < n_res; r += kPrepThreads)
            local_max = max(local_max, (int)get_idx(res_start, wide, r + 1) - (int)get_idx(res_start, wide, r));
        if (local_max) atomicMax(&s_maxres, local_max);
        __syncthreads();
        if (s_maxres > kMaxResidueAtoms && status == 0) status = 3;
        int res_shift = 0;
        while ((1 << res_shift) < s_maxres) ++res_shift;

        // ---- pass D: leader candidate lists -----------------------------------------------------------
        const int off_leadoff = off_resstart + (int)align16(ib * (int64_t)(n_res + 1));
        const int off_lead = off_leadoff + (int)align16(4 * (int64_t)(L.n_leader + 1));
        uint32_t *lead_off = reinterpret_cast<uint32_t *>(blob + off_leadoff);
        void *lead = blob + off_lead;
        // Counts: a class histogram of the kept atoms, then per list the sum over its classes.
        const int n_class = L.class_words * 32, mw = L.mask_words;
        for (int c = tid; c < n_class; c += kPrepThreads) s_hist[c] = 0;
        __syncthreads();
        for (int i = tid; i < n_kept; i += kPrepThreads) atomicAdd(&s_hist[bklass[i]], 1);
        __syncthreads();
        for (int l = wid; l < L.n_leader; l += kPrepThreads / 32) {
            unsigned cnt = 0;
            for (int c = lane; c < n_class; c += 32) {
                const unsigned h = (unsigned)s_hist[c];
                if (h) cnt += h * ((__ldg(L.class_mask + (size_t)c * mw + (l >> 5)) >> (l & 31)) & 1u);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            if (lane == 0) s_lead_cnt[l] = cnt;
        }
        __syncthreads();
        if (tid == 0) {
            unsigned run = 0;
            for (int l = 0; l < L.n_leader; ++l) { const unsigned c = s_lead_cnt[l]; s_lead_cnt[l] = run; lead_off[l] = run; run += c; }
            s_lead_cnt[L.n_leader] = run;
            lead_off[L.n_leader] = run;
        }
        __syncthreads();
        // Fill: one pass over the atoms per group of 256 lists.  A lane holds its atom's 256-bit
        // membership mask in registers; warp w owns the lists whose bit position within a mask word is
        // w, w + 8, w + 16, w + 24 (32 lists per group), so every test is a shift of a register and
        // every list receives its atoms in ascending order through ballot + popc.
        for (int g = 0; g < mw / 8; ++g) {
            unsigned pos[8][4];
#pragma unroll
            for (int wd = 0; wd < 8; ++wd)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int l = 256 * g + 32 * wd + wid + 8 * j;
                    pos[wd][j] = l < L.n_leader ? s_lead_cnt[l] : 0u;
                }
            for (int i0 = 0; i0 < n_kept; i0 += 32) {
                const int i = i0 + lane;
                uint4 m0 = make_uint4(0u, 0u, 0u, 0u), m1 = m0;
                if (i < n_kept) {
                    const uint4 *row = reinterpret_cast<const uint4 *>(L.class_mask + (size_t)bklass[i] * mw + 8 * g);
                    m0 = __ldg(row);
                    m1 = __ldg(row + 1);
                }
                const unsigned words[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
                for (int wd = 0; wd < 8; ++wd) {
                    const unsigned mine = words[wd] >> wid;
                    if (__any_sync(0xffffffffu, mine & 0x01010101u) == 0) continue;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool ok = (mine >> (8 * j)) & 1u;
                        const unsigned bal = __ballot_sync(0xffffffffu, ok);
                        if (ok) put_idx(lead, wide, pos[wd][j] + warp_excl_prefix(bal, lane), (unsigned)i);
                        pos[wd][j] += __popc(bal);
                    }
                }
            }
        }
        __syncthreads();

        // ---- pass E: uniform grid (cell list) over the kept atoms -------------------------------------
        // Counting sort by cell; atoms inside a cell end up in ascending order, so the layout is
        // deterministic.  The cell edge grows until the grid has at most kMaxCells cells.
        const int off_cellstart = (int)align16(off_lead + ib * (int64_t)s_lead_cnt[L.n_leader]);
        float cell = kMinCell;
        int nx = 0, ny = 0, nz = 0;
        float ox = 0.f, oy = 0.f, oz = 0.f;
        int off_cellatoms = off_cellstart;
        if (build_cells) {       // opt-in (cell_threshold > 0): the default search never reads the grid
        for (;;) {
            nx = (int)(ext[0] / cell) + 1; ny = (int)(ext[1] / cell) + 1; nz = (int)(ext[2] / cell) + 1;
            if ((long long)nx * ny * nz <= kMaxCells) break;
            cell *= 1.25f;
        }
        const int n_cells = nx * ny * nz;
        ox = (float)(-0.5 * ext[0]) - 1e-3f;
        oy = (float)(-0.5 * ext[1]) - 1e-3f;
        oz = (float)(-0.5 * ext[2]) - 1e-3f;
        off_cellatoms = off_cellstart + (int)align16(ib * (int64_t)(n_cells + 1));
        void *cell_start = blob + off_cellstart;
        void *cell_atoms = blob + off_cellatoms;
        auto cell_of = [&](int i) {
            const float4 p = atoms[i];
            const int ix = min(nx - 1, max(0, (int)((p.x - ox) / cell)));
            const int iy = min(ny - 1, max(0, (int)((p.y - oy) / cell)));
            const int iz = min(nz - 1, max(0, (int)((p.z - oz) / cell)));
            return (iz * ny + iy) * nx + ix;
        };
        for (int c = tid; c <= n_cells; c += kPrepThreads) s_cell[c] = 0;
        __syncthreads();
        for (int i = tid; i < n_kept; i += kPrepThreads) atomicAdd(&s_cell[cell_of(i)], 1);
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int c = 0; c < n_cells; ++c) { const int cnt = s_cell[c]; s_cell[c] = run; put_idx(cell_start, wide, c, (unsigned)run); run += cnt; }
            s_cell[n_cells] = run;
            put_idx(cell_start, wide, n_cells, (unsigned)run);
        }
        __syncthreads();
        for (int i = tid; i < n_kept; i += kPrepThreads) put_idx(cell_atoms, wide, atomicAdd(&s_cell[cell_of(i)], 1), (unsigned)i);
        __syncthreads();
        for (int c = tid; c < n_cells; c += kPrepThreads) {      // insertion sort inside each (small) cell
            const int b0 = (int)get_idx(cell_start, wide, c), b1 = (int)get_idx(cell_start, wide, c + 1);
            for (int i = b0 + 1; i < b1; ++i) {
                const unsigned v = get_idx(cell_atoms, wide, i);
                int j = i - 1;
                while (j >= b0 && get_idx(cell_atoms, wide, j) > v) { put_idx(cell_atoms, wide, j + 1, get_idx(cell_atoms, wide, j)); --j; }
                put_idx(cell_atoms, wide, j + 1, v);
            }
        }
        __syncthreads();

        }
        if (tid == 0) {
            BlobHeader h;
            h.n_kept = n_kept; h.n_res = n_res; h.res_shift = res_shift; h.status = status; h.eps = eps;
            // the cell list stays in global memory: it is consulted only for very long leader lists
            h.staged_bytes = off_cellstart;
            h.off_cellstart = off_cellstart; h.off_cellatoms = off_cellatoms; h.nx = nx; h.ny = ny; h.nz = nz;
            h.cell = cell; h.ox = ox; h.oy = oy; h.oz = oz;
            for (int i = 0; i < 8; ++i) h.pad[i] = 0;
            h.off_atom = off_atom;
            h.wide = wide ? 1 : 0;
            for (int i = 0; i < 3; ++i) h.reserved[i] = 0;
            h.off_resstart = off_resstart; h.off_leadoff = off_leadoff;
            h.off_lead = off_lead; h.off_orig = off_orig;
            *reinterpret_cast<BlobHeader *>(blob) = h;
            if (stats) atomicAdd(stats + 5, (unsigned long long)n_kept);
            if (B.status) B.status[s] = status;
            if (status != 0 && bad) atomicAdd(bad, 1ull);
        }
        __syncthreads();
    }
}

void launch_prepare(const DevLibrary &L, const DevBatch &B, float cutoff, bool build_cells,
                    unsigned long long *stats, unsigned long long *bad, int sm_count, cudaStream_t stream)
{
    if (B.n_structures <= 0) return;
    int grid = B.n_structures < sm_count * 8 ? B.n_structures : sm_count * 8;
    emm_prepare_kernel<<<grid, kPrepThreads, 0, stream>>>(L, B, cutoff, build_cells ? 1 : 0, stats, bad);
}

}  // namespace emm
