/*
 * enzymm_b200.h -- C ABI of the B200-native geometric matching engine.
 *
 * This is the drop-in boundary for ONE path of RayHackett/enzymm: what
 *     pyjess.Jess(templates).query(molecule, rmsd_threshold, distance_cutoff,
 *                                  max_dynamic_distance, max_candidates, best_match=True,
 *                                  ignore_chain=True)
 * computes inside enzymm.jess_run.Matcher._run_jess (reference enzymm/jess_run.py:785-843), plus
 * EnzyMM's RMSD/orientation logistic filter (jess_run.py:298-346, 425-478), batched over many
 * query structures.  The reference has no C plugin ABI for this path -- its boundary is the
 * Python API of the un-vendored `pyjess` wheel (pyproject.toml:30) -- so each entry point below
 * names the pyjess / EnzyMM call it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions: plain pointers and sizes only; every function returns EMM_OK (0) or a negative
 * emm_status; no exceptions cross the boundary; the caller owns all host buffers; the library
 * owns device memory behind opaque handles; one handle is bound to one CUDA device; calls on one
 * handle must be serialised by the caller.  `stream` arguments are cudaStream_t passed as void*
 * (NULL = the legacy default stream).
 */
#ifndef ENZYMM_B200_H
#define ENZYMM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMM_ABI_VERSION 2
#define EMM_MAX_TEMPLATE_ATOMS 32   /* shipped library: 6..24 atoms per template            */
#define EMM_MAX_RESIDUES 10         /* orientation residues (atom triplets) per template     */
#define EMM_LR_MODELS 5             /* logistic models per (size, distance) cell             */

typedef enum emm_status {
    EMM_OK = 0,
    EMM_ERR_INVALID = -1,       /* bad argument / inconsistent table                          */
    EMM_ERR_CUDA = -2,          /* CUDA runtime error (see emm_last_error)                    */
    EMM_ERR_NO_DEVICE = -3,     /* no usable CUDA device: there is NO CPU fallback            */
    EMM_ERR_CAPACITY = -4,      /* hit buffer too small; *n_hits holds the required count     */
    EMM_ERR_INPUT = -5,         /* a structure violates the input contract (see emm_batch)    */
    EMM_ERR_NOMEM = -6
} emm_status;

/* hit.flags */
#define EMM_HIT_OVERFLOW   0x01u  /* max_candidates reached: result depends on enumeration order */
#define EMM_HIT_BORDERLINE 0x02u  /* rmsd threshold or a logistic vote decided within 1e-9       */
#define EMM_HIT_PASS       0x04u  /* Match.predicted_correct (jess_run.py:298-346)               */
#define EMM_HIT_NO_MODEL   0x08u  /* size has models but none for this distance -> host KeyError */
#define EMM_HIT_ORIENTED   0x10u  /* orientation was computed (template has residue triplets)    */

/*
 * Compiled template library.  Replaces pyjess.Jess(templates) (jess_run.py:800) together with
 * pyjess.Template / pyjess.TemplateAtom (template.py:535, 679) and the constants EnzyMM derives
 * per template residue (Residue.calc_residue_orientation, template.py:213-304).
 *
 * Atoms are stored twice: in TEMPLATE order for superposition/output (xyz), and as a SEARCH PLAN
 * (plan_*) -- a permutation chosen by the host so that selective atoms are placed first.  Any
 * permutation yields the same match set; only speed changes.
 *
 * Typing is data driven: template atom i carries a row index `ttype` into the bit matrix
 * compat[n_ttype][class_words]; a query atom of class c may bind it iff bit c of the row is set
 * (SURVEY.md 8c rules 2-3 are evaluated on the host when the matrix is built).  Class 0 is the
 * "binds nothing" class.
 */
typedef struct emm_library_desc {
    int32_t n_templates;
    int32_t n_atoms;               /* total template atoms                                      */
    const int32_t *atom_off;       /* [n_templates+1] CSR into per-atom arrays                  */
    const double *xyz;             /* [n_atoms][3] template coordinates, template order         */
    const double *weight;          /* [n_atoms] distance_weight, template order                 */
    const uint16_t *chain;         /* [n_atoms] chain id code, template order (ignore_chain=0)  */
    /* search plan, plan position k of template t lives at atom_off[t]+k */
    const uint8_t *plan_atom;      /* [n_atoms] template-order index of the atom placed at k    */
    const uint16_t *plan_ttype;    /* [n_atoms] compat row of that atom                         */
    const int16_t *plan_src;       /* [n_atoms] >=0: same residue as plan position src;         */
                                   /*           <0 : leader, candidates = leader list (-1-src)  */
    const uint8_t *plan_anchor;    /* [n_atoms] earlier plan position (< k) whose distance constraint */
                                   /*           filters candidates first; position 0: unused;   */
                                   /*           same-residue positions: must equal plan_src      */
    const int64_t *pair_off;       /* [n_templates+1] CSR into pair_dist (k*(k-1)/2 + j, j<k)   */
    const double *pair_dist;       /* template distances between plan positions, FP64           */
    /* typing */
    int32_t n_ttype;
    int32_t class_words;           /* 32-bit words per compat row                               */
    const uint32_t *compat;        /* [n_ttype][class_words]                                    */
    int32_t n_leader;              /* distinct leader types                                     */
    const uint16_t *leader_ttype;  /* [n_leader] compat row each leader list is built from      */
    /* per-template thresholds (jess_run.py:564-571, 724-736) */
    const double *rmsd_threshold;  /* [n_templates]                                             */
    const double *distance_cutoff; /* [n_templates]                                             */
    const double *max_dynamic_distance; /* [n_templates]                                        */
    /* orientation + logistic filter (jess_run.py:298-346, 425-478; template.py:213-304) */
    const int32_t *n_residues;     /* [n_templates] atom triplets; 0 = no orientation           */
    const uint8_t *orient_idx;     /* [n_templates][EMM_MAX_RESIDUES][2] (first, second|9)      */
    const double *orient_vec;      /* [n_templates][EMM_MAX_RESIDUES][3] template vectors       */
    const int32_t *lr_index;       /* [n_templates] row of lr_table; -1 pass always; -2 no model */
    int32_t n_lr;
    const double *lr_table;        /* [n_lr][EMM_LR_MODELS][4] = coef_rmsd, coef_orient, intercept, threshold */
} emm_library_desc;

typedef struct emm_library emm_library;

int emm_abi_version(void);
int emm_hit_size(void);             /* sizeof(emm_hit), for binding sanity checks */
const char *emm_last_error(void);
int emm_device_count(void);
/* A non-blocking CUDA stream on `device` for the `stream` arguments below (NULL = the default
 * stream).  Two sessions on two streams overlap one batch's upload with another's search. */
int emm_stream_create(int device, void **stream);
int emm_stream_destroy(int device, void *stream);

int emm_library_create(int device, const emm_library_desc *desc, emm_library **out);
/* Replace the compat matrix (same n_ttype; class_words may grow up to the value at creation). */
int emm_library_set_compat(emm_library *lib, int32_t class_words, const uint32_t *compat);
/* Replace per-template thresholds (the rmsd/distance/max_dynamic triple of Jess.query). */
int emm_library_set_thresholds(emm_library *lib, const double *rmsd_threshold,
                               const double *distance_cutoff, const double *max_dynamic_distance);
/* Replace the logistic-filter assignment (it is keyed by the distance cutoff, jess_run.py:309-311).
 * n_lr may not exceed max(64, n_lr at creation). */
int emm_library_set_filter(emm_library *lib, const int32_t *lr_index, int32_t n_lr, const double *lr_table);
void emm_library_destroy(emm_library *lib);

/*
 * A batch of query structures as SoA columns.  Replaces pyjess.Molecule / pyjess.Atom on the
 * matching path (jess_run.py:538-548): the host parses and classifies, the device gets numbers.
 *
 * Input contract (violations -> EMM_ERR_INPUT): within one structure `residue` is non-decreasing
 * (atoms of a residue are contiguous; the host reorders odd files and reports original indices
 * through atom_id); at most 4 194 303 atoms of a structure survive masking and class-0 removal, at
 * most 1 023 of them in one residue.  A residue is a distinct (chain_id, residue_number) (SURVEY.md
 * 8c rule 4).  A structure that breaks the contract is skipped, not searched: the batch's other
 * structures are searched and their hits delivered, emm_session_download still returns EMM_ERR_INPUT
 * and emm_session_structure_status names the offenders.
 */
typedef struct emm_batch {
    int32_t n_structures;
    int64_t n_atoms;
    const int64_t *atom_off;   /* [n_structures+1]                                              */
    const double *xyz;         /* [n_atoms][3] exactly the doubles the PDB text parses to        */
    const uint16_t *klass;     /* [n_atoms] typing class (column of compat); 0 binds nothing     */
    const int32_t *residue;    /* [n_atoms] residue ordinal inside its structure                 */
    const float *bfactor;      /* [n_atoms] temperature factor / pLDDT, or NULL                  */
    const uint16_t *chain;     /* [n_atoms] chain id code, or NULL (needed for ignore_chain=0)   */
    const int32_t *atom_id;    /* [n_atoms] index reported in hits, or NULL = position in structure */
} emm_batch;

typedef struct emm_query_params {
    int64_t max_candidates;       /* Jess.query(max_candidates=...), jess_run.py:808; <=0 = unlimited */
    int32_t ignore_chain;         /* Jess.query(ignore_chain=...),   jess_run.py:810              */
    float conservation_cutoff;    /* keep atoms with bfactor >= cutoff (Molecule.conserved); 0 = all */
    int32_t template_begin;       /* templates [begin, end) of the library                        */
    int32_t template_end;
    int32_t skip_mode;            /* 0 search all; 1 skip structures that already hold a PASSing  */
                                  /* hit; 2 ... that hold any hit (--skip-smaller-hits, jess_run.py:951-958) */
    int32_t reset_structure_state;/* 1: clear per-structure hit counters before this run          */
    int32_t force_prepare;        /* 1: run the prepare kernel even if this batch is already prepared */
    int32_t cell_threshold;       /* > 0: leader candidate lists at least this long are searched through  */
                                  /* the uniform-grid cell list instead of scanned; <= 0: never (default) */
    int32_t donate_after;         /* splitting of one expensive (template, structure) pair over the warps */
                                  /* of its CTA: 0 default (launches below 8192 structures: after 48 level */
                                  /* visits, when warps sit idle; larger launches: never); n > 0 after n   */
                                  /* level visits; < 0 never.  Results do not depend on it.               */
} emm_query_params;

typedef struct emm_hit {
    int32_t structure;            /* index in the batch                                           */
    int32_t template_index;
    uint32_t n_complete;          /* complete assignments examined ("candidates")                 */
    uint16_t n_atoms;
    uint16_t flags;
    double rmsd;                  /* Hit.rmsd                                                     */
    double orientation;           /* Match.orientation (radians); NaN if not EMM_HIT_ORIENTED     */
    double rot[9];                /* row-major R: q' = R (q - qbar) + tbar (query -> template)    */
    double qbar[3];
    double tbar[3];
    int32_t atoms[EMM_MAX_TEMPLATE_ATOMS]; /* matched query atoms in TEMPLATE order (Hit.atoms)   */
} emm_hit;

typedef struct emm_stats {
    uint64_t pairs;               /* (template, structure) pairs searched                         */
    uint64_t sweeps;              /* 32-lane candidate sweeps                                     */
    uint64_t dist_evals;          /* FP32 pairwise-distance constraint evaluations                */
    uint64_t exact_rechecks;      /* candidates re-evaluated in FP64 (guard band)                 */
    uint64_t complete;            /* complete assignments superposed                              */
    uint64_t kept_atoms;          /* atoms surviving mask + class-0 removal                       */
    uint64_t staged_bytes;        /* bytes of structure blobs staged into shared memory           */
    uint64_t global_blobs;        /* work items whose blob did not fit shared memory              */
} emm_stats;

/*
 * Session = device buffers for batches up to the given sizes, on the library's device.
 * upload / run / download are asynchronous on `stream` except where noted, so a caller can time
 * the resident path (run only) and the end-to-end path (all three) separately.
 */
typedef struct emm_session emm_session;

int emm_session_create(emm_library *lib, int64_t max_atoms, int32_t max_structures,
                       int64_t hit_capacity, emm_session **out);
void emm_session_destroy(emm_session *s);

/* Host -> device copy of a batch (async; host buffers must stay valid until the stream reaches
 * this point -- use pinned memory for true overlap). */
int emm_session_upload(emm_session *s, const emm_batch *batch, void *stream);

/* Launch prepare + search kernels on the uploaded batch (async).  Replaces the loop
 * `for mol in molecules: Jess(templates).query(mol, ...)` of Matcher.run (jess_run.py:930-986). */
int emm_session_run(emm_session *s, const emm_query_params *params, void *stream);

/* Device -> host copy of the hits produced since the last reset; synchronises `stream`.
 * Hits come back sorted by (structure, template_index).  stats may be NULL. */
int emm_session_download(emm_session *s, emm_hit *hits, int64_t capacity, int64_t *n_hits,
                         emm_stats *stats, void *stream);

/* Per-structure outcome of the last prepare pass of the uploaded batch (synchronous; call after
 * emm_session_download): 0 searched; 1 residue ordinals decrease; 2 too many kept atoms; 3 a residue
 * keeps more than 1023 atoms.  capacity >= n_structures of the batch. */
int emm_session_structure_status(emm_session *s, int32_t *status, int32_t capacity);

/* Number of kernel launches issued by the last emm_session_run (for bench accounting). */
int emm_session_last_launches(const emm_session *s);

/* Device time of the kernels launched since emm_session_clear_timings, measured with CUDA events
 * recorded on the launching stream around each kernel (call after the stream was synchronised).
 * which = 0: prepare kernel, 1: search kernel.  Writes up to capacity durations (ms), returns the
 * number of launches recorded through *count. */
int emm_session_kernel_ms(emm_session *s, int which, float *out_ms, int capacity, int *count);
int emm_session_clear_timings(emm_session *s);

/* Diagnostics of the EMM_STATS=1 build path: sweeps[64] then survivors[64], indexed by search level
 * (+32 for same-residue levels). */
int emm_session_debug_counters(emm_session *s, unsigned long long *out128);

/* Convenience: upload + run + download on the default stream (the end-to-end call). */
int emm_query_batch(emm_library *lib, const emm_batch *batch, const emm_query_params *params,
                    emm_hit *hits, int64_t capacity, int64_t *n_hits, emm_stats *stats);

/*
 * Native PDB ingest (host only; needs no GPU).  Replaces pyjess.Molecule.load on the matching path
 * (jess_run.py:538): ATOM and HETATM records in file order up to the first ENDMDL; coordinates are
 * the doubles strtod / Python float() produce.  Strings are blank-stripped, NUL padded fixed-width
 * fields: name[4], resname[4], chain[2], segment[4], element[2]; altloc / icode are single chars.
 *
 * Every reader below takes PDB or mmCIF text and tells them apart by content, as
 * pyjess.Molecule.load(format="detect") does: a text whose first token is a data_ block header is
 * read through its _atom_site category -- the rows of the first pdbx_PDB_model_num, in file order,
 * identifiers from the label_* items (label_atom_id, label_comp_id, label_asym_id, label_seq_id;
 * an item that is absent or '.' / '?' falls back to its auth_* twin) unless EMM_PDB_CIF_AUTHOR asks
 * for the auth_* items first (PyJess's use_author).  Names longer than the fields are cut; a chain
 * id longer than two characters is EMM_ERR_INPUT.  header_id receives the data block name when it
 * fits four characters.  The readers that take PATHS also accept gzip-compressed files (.gz by
 * content, not by name).  No reference test holds an mmCIF or gzip input: this part is unpinned.
 */
#define EMM_PDB_CIF_AUTHOR 1    /* flags of the _ex readers: prefer auth_* over label_* identifiers */
#define EMM_PDB_SKIP_BAD 2      /* emm_pdb_pack_files_ex: a file that cannot be opened, read, inflated or parsed
                                 * does not fail the call; it stays in the batch as a structure without atoms
                                 * and emm_pdb_batch_file_status / _file_message say what happened to it */
int emm_pdb_count_atoms(const char *text, int64_t len, int64_t *n_atoms);
int emm_pdb_parse(const char *text, int64_t len, int64_t capacity, int32_t *serial, char *name, char *altloc,
                  char *resname, char *chain, int32_t *resnum, char *icode, double *xyz, double *occupancy,
                  double *bfactor, char *segment, char *element, int8_t *charge, char *header_id /* [5] */,
                  int64_t *n_atoms);
/* As emm_pdb_parse with reader flags (emm_pdb_parse = flags 0). */
int emm_pdb_parse_ex(const char *text, int64_t len, int32_t flags, int64_t capacity, int32_t *serial, char *name,
                     char *altloc, char *resname, char *chain, int32_t *resnum, char *icode, double *xyz,
                     double *occupancy, double *bfactor, char *segment, char *element, int8_t *charge,
                     char *header_id /* [5] */, int64_t *n_atoms);
const char *emm_pdb_last_error(void);

/* Many files at once on n_threads threads into one SoA batch owned by the library. */
typedef struct emm_pdb_batch emm_pdb_batch;
typedef struct emm_pdb_columns {
    int32_t n_files;
    int64_t n_atoms;
    const int64_t *atom_off;      /* [n_files+1] */
    const int32_t *serial; const char *name; const char *altloc; const char *resname; const char *chain;
    const int32_t *resnum; const char *icode; const double *xyz; const double *occupancy; const double *bfactor;
    const char *segment; const char *element; const int8_t *charge;
    const char *header_id;        /* [n_files][5] HEADER idCode or empty */
} emm_pdb_columns;
int emm_pdb_load_files(const char *const *paths, int32_t n_files, int32_t n_threads, emm_pdb_batch **out);
int emm_pdb_load_files_ex(const char *const *paths, int32_t n_files, int32_t n_threads, int32_t flags, emm_pdb_batch **out);
int emm_pdb_batch_columns(const emm_pdb_batch *batch, emm_pdb_columns *out);
void emm_pdb_batch_free(emm_pdb_batch *batch);

/*
 * Files -> the columns of emm_batch in one native pass, without per-atom host objects: what
 * jess_run.py:538-548 (Molecule.load per file) plus the packing of the batched upload do together.
 * Typing needs the (residue name, atom name) of every atom; the reader returns a per-atom index
 * into a table of the distinct kinds it saw (first-appearance order, files in argument order), and
 * the caller maps kinds to typing classes of its compiled library (klass = class_of_kind[kind]).
 * residue / atom_id follow emm_batch: residues numbered by first appearance of (chain, resSeq);
 * atom_id is NULL unless some file had a residue split over several runs (then it holds, for every
 * file, the original index of each packed atom).
 */
typedef struct emm_pdb_packed {
    int32_t n_files;
    int64_t n_atoms;
    const int64_t *atom_off;      /* [n_files+1] */
    const double *xyz;            /* [n_atoms][3] */
    const uint32_t *kind;         /* [n_atoms] index into kind_names */
    const int32_t *residue;       /* [n_atoms] */
    const float *bfactor;         /* [n_atoms] */
    const uint16_t *chain;        /* [n_atoms] chain id bytes: b0 | b1 << 8 */
    const int32_t *atom_id;       /* [n_atoms] or NULL */
    const uint16_t *klass;        /* [n_atoms] typing classes, NULL until emm_pdb_batch_classify */
    int32_t n_kinds;
    const char *kind_names;       /* [n_kinds][8]: resname[4] name[4], blank-stripped, NUL padded */
    const char *header_id;        /* [n_files][5] */
    /* what the results table needs besides the hits (Match.dump, jess_run.py:185-284) */
    const int64_t *res_off;       /* [n_files+1] CSR into res_key */
    const uint64_t *res_key;      /* per residue ordinal: chain id code << 32 | (uint32) residue number */
    const int32_t *residue_count; /* [n_files] Match.query_residue_count (jess_run.py:487-496) */
} emm_pdb_packed;
int emm_pdb_pack_files(const char *const *paths, int32_t n_files, int32_t n_threads, emm_pdb_batch **out);
int emm_pdb_pack_files_ex(const char *const *paths, int32_t n_files, int32_t n_threads, int32_t flags, emm_pdb_batch **out);
/* Per file of a packed batch: 0 fine, 1 cannot open, 2 cannot read, 3 malformed, 4 cannot inflate (only
 * EMM_PDB_SKIP_BAD leaves non-zero entries); the message is "" for a file that is fine and stays valid
 * until the batch is freed. */
int emm_pdb_batch_file_status(const emm_pdb_batch *batch, int32_t *status /* [n_files] */, int32_t capacity);
const char *emm_pdb_batch_file_message(const emm_pdb_batch *batch, int32_t file);
int emm_pdb_batch_packed(const emm_pdb_batch *batch, emm_pdb_packed *out);
/* The same packing from per-structure columns that are already in memory (what Matcher.run has after
 * Molecule.load, jess_run.py:538-548): name4 / resname4 / chain2 are blank-stripped NUL-padded fixed
 * width bytes ([n][4], [n][4], [n][2]) per structure.  Result: as emm_pdb_pack_files. */
int emm_pack_columns(int32_t n_structures, const int64_t *sizes, const uint8_t *const *name4,
                     const uint8_t *const *resname4, const uint8_t *const *chain2, const int32_t *const *resnum,
                     const double *const *xyz, const double *const *bfactor, int32_t n_threads, emm_pdb_batch **out);
/* klass[a] = class_of_kind[kind[a]] for every atom, on the batch's thread pool */
int emm_pdb_batch_classify(emm_pdb_batch *batch, const uint16_t *class_of_kind, int32_t n_kinds);

/*
 * Native writer for the rows of the results table (host only).  Replaces Match.dump per match
 * (jess_run.py:185-284) when whole batches are written: the caller chooses and orders the rows
 * (Matcher semantics: size groups, completeness, filter verdict, match indices) and gathers, per row,
 * the residue name / chain / residue number of the matched atoms; this formats them byte for byte as
 * Python does -- str(round(x, 5)), str(bool), csv QUOTE_MINIMAL with tabs.
 */
typedef struct emm_tsv_rows {
    int64_t n_rows;
    const int32_t *n_atoms;            /* [n_rows] matched atoms, three per template residue            */
    const char *resname4;              /* [n_rows][EMM_MAX_TEMPLATE_ATOMS][4] blank-stripped, NUL padded */
    const char *chain2;                /* [n_rows][EMM_MAX_TEMPLATE_ATOMS][2]                            */
    const int32_t *resnum;             /* [n_rows][EMM_MAX_TEMPLATE_ATOMS]                               */
    const double *rmsd;                /* [n_rows] Hit.rmsd                                              */
    const double *log_evalue;          /* [n_rows] Hit.log_evalue (NaN: not reproducible)     */
    const double *orientation;         /* [n_rows] Match.orientation                                     */
    const int32_t *match_index;        /* [n_rows] Match.index                                           */
    const uint8_t *complete;           /* [n_rows] Match.complete                                        */
    const uint8_t *predicted;          /* [n_rows] Match.predicted_correct: 0 False, 1 True, 2 left empty */
    const int32_t *template_index;     /* [n_rows] */
    const int32_t *structure;          /* [n_rows] */
    const char *const *query_id;       /* [n_structures] Molecule.id                                     */
    const int32_t *query_atom_count;   /* [n_structures] */
    const int32_t *query_residue_count;/* [n_structures] */
    const char *const *tpl_distance;   /* [n_templates] str(pairwise_distance)                           */
    const char *const *tpl_static;     /* [n_templates] columns template_pdb_id .. template_cath, tabbed  */
    const uint8_t *tpl_multimeric;     /* [n_templates] Template.multimeric                              */
    const int32_t *tpl_order_off;      /* [n_templates+1] CSR into tpl_order                             */
    const int32_t *tpl_order;          /* Template.relative_order                                        */
    const char *const *tpl_annotation; /* [n_templates] the trailing annotation columns, tabbed          */
} emm_tsv_rows;
int emm_tsv_format(const emm_tsv_rows *rows, char **text, int64_t *len);   /* *text: free with emm_tsv_free */
void emm_tsv_free(char *text);
/* str(round(v, 5)) as Python prints it (exposed for the formatter's own tests) */
int emm_tsv_repr_round5(double v, char *out, int32_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* ENZYMM_B200_H */
