// emm_device.cuh -- device-side tables shared by the prepare and search kernels (sm_100a).
//
// Data layout in HBM (DESIGN.md "Data layout"):
//   * DevLibrary  : the compiled template library, read-only, ~5 MB for the shipped 6780 active
//                   templates -> L2 resident, broadcast-read through L1 by every warp.
//   * DevBatch    : the uploaded query structures as SoA columns.
//   * structure blob (one per structure, written by emm_prepare_kernel, read by
//                   emm_search_kernel): header + FP32 centred coordinates + residue CSR + typing
//                   classes + per-leader-type candidate lists + uniform-grid cell list.  The
//                   staged part is copied into shared memory once per work item.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/enzymm_b200.h"

namespace emm {

constexpr int kMaxAtoms = EMM_MAX_TEMPLATE_ATOMS;
#ifndef EMM_SEARCH_THREADS
#define EMM_SEARCH_THREADS 768
#endif
constexpr int kSearchThreads = EMM_SEARCH_THREADS;
constexpr int kSearchWarps = kSearchThreads / 32;
// Per-warp trie queues: the first kWideLevels levels hold kQueueCap entries, deeper (rarely
// populated) levels kDeepCap.  Parent indices are 8 bits.
#ifndef EMM_QUEUE_CAP
#define EMM_QUEUE_CAP 48
#endif
#ifndef EMM_DEEP_CAP
#define EMM_DEEP_CAP 16
#endif
#ifndef EMM_WIDE_LEVELS
#define EMM_WIDE_LEVELS 9
#endif
constexpr int kQueueCap = EMM_QUEUE_CAP;
constexpr int kDeepCap = EMM_DEEP_CAP;
constexpr int kWideLevels = EMM_WIDE_LEVELS;
// = k <= kWideLevels ? k * kQueueCap : kWideLevels * kQueueCap + (k - kWideLevels) * kDeepCap, branch-free
__host__ __device__ constexpr int queue_off(int k) { return k * kDeepCap + (k < kWideLevels ? k : kWideLevels) * (kQueueCap - kDeepCap); }
static_assert(queue_off(kWideLevels + 3) == kWideLevels * kQueueCap + 3 * kDeepCap && queue_off(4) == 4 * kQueueCap, "queue layout");
__host__ __device__ constexpr int queue_cap(int k) { return k < kWideLevels ? kQueueCap : kDeepCap; }
constexpr int kPrepThreads = 256;

struct DevLibrary {
    int n_templates, n_atoms, n_ttype, class_words, class_words_cap, n_leader, max_tpl_atoms, n_lr;
    const int32_t *atom_off;
    const double *xyz;
    const double *weight;
    const uint16_t *chain;
    const uint8_t *plan_atom;
    const uint16_t *plan_ttype;
    const int16_t *plan_src;
    const uint8_t *plan_anchor;
    const int64_t *pair_off;
    const double *pair_dist;
    const float *pair_dist32;
    const float *anchor_dist32;   // [n_atoms] template distance between plan position k and its anchor position
    const uint32_t *compat;
    const uint16_t *leader_ttype;
    // per typing class: bit l set = the class belongs to leader list l ([classes][mask_words] words,
    // mask_words a multiple of 8); lets the prepare kernel build all leader lists in one pass
    const uint32_t *class_mask;
    int mask_words;
    const double *rmsd_thr;
    const double *dist_cut;
    const double *max_dyn;
    const int32_t *n_residues;
    const uint8_t *orient_idx;
    const double *orient_vec;
    const int32_t *lr_index;
    const double *lr_table;
};

struct DevBatch {
    int n_structures;
    const int64_t *atom_off;
    const double *xyz;
    const uint16_t *klass;
    const int32_t *residue;
    const float *bfactor;    // may be null
    const uint16_t *chain;   // may be null
    const int32_t *atom_id;  // may be null
    unsigned char *blob;     // all structure blobs
    const int64_t *blob_off; // [n_structures+1]
    int32_t *status;         // [n_structures] BlobHeader.status of every structure, written by prepare
    const int32_t *kept_bound;  // [n_structures] atoms whose class is not 0: an upper bound of n_kept the
                                // host sizes the blob with (class-0 atoms, ~17 % of a protein, never stay)
};

constexpr int kMaxCells = 4096;        // uniform-grid cells per structure (cell edge grows to fit)
constexpr float kMinCell = 6.0f;       // cell edge in Angstrom for ordinary structures

// Atom ids.  A queue entry of the search packs (flags, parent slot, atom) into 32 bits with 22 bits
// for the atom, and an atom record packs (residue, typing class) into 32 bits with 22 bits for the
// residue and 10 for the class, so a structure may keep up to 4 194 303 atoms in as many residues
// (the largest PDB / mmCIF assemblies are well below that).  Index arrays inside the blob (residue
// starts, leader lists, cell list) are 16-bit for structures of at most 65 535 input atoms -- every
// structure that can be staged into shared memory -- and 32-bit ("wide") above.
constexpr int kAtomBits = 22;
constexpr uint32_t kAtomMask = (1u << kAtomBits) - 1u;
constexpr int kClassBits = 10;                         // typing classes per library: at most 1024
constexpr uint32_t kClassMask = (1u << kClassBits) - 1u;
constexpr int kMaxResidueAtoms = 1023;                 // kept atoms of one residue (10-bit span length)
constexpr int64_t kNarrowAtoms = 65535;
__host__ __device__ constexpr bool is_wide(int64_t n_input_atoms) { return n_input_atoms > kNarrowAtoms; }

// Blob header (128 bytes).  All off_* are byte offsets from the blob base, 16-byte aligned.
struct BlobHeader {
    int32_t n_kept;        // atoms kept (mask + class != 0), local ids 0..n_kept-1 in input order
    int32_t n_res;         // residues holding at least one kept atom
    int32_t res_shift;     // log2 of res_stride, res_stride = pow2 >= largest residue
    int32_t status;        // 0 ok, 1 residue order violated, 2 too many atoms kept, 3 residue too large
    float eps;             // FP32 guard band (Angstrom) for this structure
    int32_t staged_bytes;  // prefix of the blob the search kernel stages into shared memory
    int32_t off_atom;      // float4[n_kept] atom records: centred x, y, z; w = bits (res_of << 10 | klass)
    int32_t wide;          // 1: res_start / lead / cell arrays hold 32-bit entries, 0: 16-bit
    int32_t reserved[3];
    int32_t off_resstart;  // idx res_start[n_res+1]
    int32_t off_leadoff;   // uint32 lead_off[n_leader+1]
    int32_t off_lead;      // idx lead[...]
    int32_t off_orig;      // int32 orig[n_kept]: position of the atom inside its structure (NOT staged)
    // uniform grid over the kept atoms (cell list): cell c = (iz*ny + iy)*nx + ix holds
    // cell_atoms[cell_start[c] .. cell_start[c+1]) in ascending atom order
    int32_t off_cellstart; // idx cell_start[nx*ny*nz + 1]
    int32_t off_cellatoms; // idx cell_atoms[n_kept]
    int32_t nx, ny, nz;
    float cell;            // cell edge
    float ox, oy, oz;      // grid origin in centred coordinates
    int32_t pad[8];
};
static_assert(sizeof(BlobHeader) == 128, "blob header is 128 bytes");

struct SearchParams {
    long long max_candidates;
    int ignore_chain;
    int template_begin, template_end;
    int skip_mode;
    int n_chunks;          // template chunks per structure (work item = structure x chunk)
    int n_items;
    // Templates are visited in the order of `sched` (most expensive first).  Two-phase mode (large
    // batches): items [0, n_structures) run the first n_heavy entries for every structure, items
    // [n_structures, 2 n_structures) the rest -- a pair that takes 100x the average starts early and
    // cannot become the tail of the launch.  Otherwise chunk c of a structure takes entries
    // c, c + n_chunks, ... of the order.
    int two_phase;
    int n_structures;
    int n_sched, n_heavy;
    int blob_cap;          // shared-memory bytes available for a staged blob
    int levels;            // queue levels per warp (max template atoms + 1)
    int cell_threshold;    // leader lists at least this long are searched through the cell list
    int donate_after;      // level entries after which a pair may hand subtrees to idle warps; < 0: never
};

struct SearchOut {
    emm_hit *hits;
    long long hit_capacity;
    unsigned long long *hit_count;
    unsigned int *work_counter;
    int *struct_any;       // [n_structures] hits so far
    int *struct_pass;      // [n_structures] PASSing hits so far
    unsigned long long *stats;  // emm_stats as 8 counters
};

inline __host__ __device__ int64_t align16(int64_t v) { return (v + 15) & ~int64_t(15); }

// Size of a structure blob that keeps at most n atoms (n = atoms of a class other than 0; `wide` from
// the INPUT atom count) whose leader lists hold lead_entries entries in total (the host computes this
// exactly at upload, so blobs are laid out without a size pass); *staged_bound = upper bound of the
// prefix the search kernel stages (header .. leader lists).
inline __host__ __device__ int64_t blob_bytes(int64_t n, bool wide, int n_leader, int64_t lead_entries, int64_t *staged_bound = nullptr)
{
    const int64_t ib = wide ? 4 : 2;         // bytes per index entry
    int64_t b = sizeof(BlobHeader);
    b += 16 * n;                             // atom records (x, y, z, res_of | klass)
    b += align16(ib * (n + 1));              // res_start
    b += align16(4 * (int64_t)(n_leader + 1));
    b += align16(ib * lead_entries);
    if (staged_bound) *staged_bound = b;
    b += align16(ib * (int64_t)(kMaxCells + 1)) + align16(ib * n);   // cell_start, cell_atoms
    b += align16(2 * n);                     // compact klass copy (prepare-kernel scratch, not staged)
    b += align16(4 * n);                     // orig
    return b;
}

}  // namespace emm
