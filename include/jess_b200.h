/* jess_b200.h -- C ABI of the B200-native template-query engine.
 *
 * This header is the drop-in boundary for the PyJess hot path
 *   Jess(templates).query(molecule, rmsd_threshold, distance_cutoff,
 *                         max_dynamic_distance, ...) -> Hit ...
 * It replaces the subset of the reference C library that PyJess's Cython
 * module binds through include/jess/{jess,super,molecule,template,
 * tess_template,tess_atom,res_index}.pxd.  Every entry point below cites the
 * reference interface it stands in for (paths relative to the reference
 * checkout; "eff:" = file after applying patches/<file>.patch).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch/Python types.
 *   - every function returns JB_OK (0) or a negative jb_status; the text of
 *     the last error on the calling thread is available via jb_last_error().
 *   - all *_desc structs are HOST memory, borrowed for the duration of the
 *     call; handles own their device (HBM) memory.
 *   - one jb_context per GPU per process (one process per GPU under
 *     torchrun); all work of a context is issued on one CUDA stream which
 *     the caller may supply (e.g. torch.cuda.current_stream().cuda_stream).
 *   - there is NO CPU fallback: if no CUDA device is usable, every entry
 *     point fails with JB_ERR_CUDA.
 *
 * Tokens.  Atom names are the reference's 4-char tokens (blanks -> '_',
 * eff:Atom.c:17-30), residue names its 3-char tokens followed by NUL
 * (compared over 4 bytes, eff:TessAtom.h:81-91).  Both travel as uint32,
 * byte i of the token in bits [8i, 8i+8) (little endian memcpy of char[4]).
 */
#ifndef JESS_B200_H
#define JESS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JB_ABI_VERSION 3
/* hard limit of the GPU path; larger templates are rejected loudly */
#define JB_MAX_TEMPLATE_ATOMS 32

typedef enum jb_status {
    JB_OK = 0,
    JB_ERR_INVALID = -1,   /* bad argument / unsupported template            */
    JB_ERR_CUDA = -2,      /* CUDA runtime error or no device                */
    JB_ERR_NOMEM = -3,     /* host or device allocation failure              */
    JB_ERR_CAPACITY = -4   /* a single structure overflowed the frontier cap */
} jb_status;

/* IgnoreType of the reference (eff:Template.h:24-29) */
typedef enum jb_ignore_chain {
    JB_IGNORE_NONE = 0,
    JB_IGNORE_RESIDUES = 1,
    JB_IGNORE_ATOMS = 2
} jb_ignore_chain;

/* ---- flat template table --------------------------------------------------
 * What Template.__init__ builds per template (src/pyjess/_jess.pyx:1825-1912):
 * n TessAtom records (eff:TessAtom.h:36-49) + the n x n FP64 distance matrix.
 * Templates are listed in INSERTION order; the engine reports template
 * indices in that numbering.  TA = atom_offset[n_templates].                 */
typedef struct jb_template_table {
    int32_t n_templates;
    const int32_t *atom_offset;      /* [T+1]  first atom of template t       */
    const int32_t *match_mode;       /* [TA]   TessAtom.code                  */
    const int32_t *residue_number;   /* [TA]   TessAtom.resSeq                */
    const uint8_t *chain_id2;        /* [TA]   TessAtom.chainID2              */
    const double *position;          /* [TA*3] TessAtom.pos                   */
    const double *distance_weight;   /* [TA]   TessAtom.distWeight            */
    const int32_t *name_offset;      /* [TA+1] into atom_names                */
    const uint32_t *atom_names;      /* 4-char tokens                         */
    const int32_t *resname_offset;   /* [TA+1] into residue_names             */
    const uint32_t *residue_names;   /* 3-char tokens + NUL                   */
    const int64_t *distance_offset;  /* [T+1]  into distances, n*n per tpl    */
    const double *distances;         /* TessTemplate.distance, row major      */
} jb_template_table;

/* ---- flat molecule batch ---------------------------------------------------
 * Structure-of-arrays image of M reference Molecules (eff:Molecule.h:24-30,
 * struct _Atom eff:Atom.h:29-47).  Only the fields the query path reads.
 * NA = atom_offset[n_molecules].                                             */
typedef struct jb_molecule_batch {
    int32_t n_molecules;
    const int64_t *atom_offset;      /* [M+1]                                 */
    const double *x;                 /* [NA] Atom.x[0]                        */
    const double *y;                 /* [NA] Atom.x[1]                        */
    const double *z;                 /* [NA] Atom.x[2]                        */
    const uint32_t *name;            /* [NA] Atom.name[0..3]                  */
    const uint32_t *residue_name;    /* [NA] Atom.resName[0..3]               */
    const int32_t *residue_number;   /* [NA] Atom.resSeq                      */
    const uint8_t *chain_id2;        /* [NA] Atom.chainID2                    */
    const int32_t *serial;           /* [NA] Atom.serial (orders the residue index,
                                      * eff:ResIndex.c:18-27); NULL = 1..n in atom
                                      * order.  Only read for JB_ORDER_REFERENCE.   */
} jb_molecule_batch;

/* order of the hit records (jb_hits) inside one (molecule, template) run      */
typedef enum jb_hit_order {
    JB_ORDER_CANONICAL = 0,  /* atom tuple ascending (set comparison, default)  */
    JB_ORDER_REFERENCE = 1,  /* the order in which the reference's depth-first
                              * Scanner_next yields them (eff:Scanner.c:193-304)
                              * when run with reorder=True: first scan level in
                              * candidate-list order (eff:TessTemplate.c:141-201),
                              * deeper levels in the reverse leaf order of the k-d
                              * tree built over the candidate set
                              * (eff:KdTree.c:238-296,343-423, qselect.h.patch).
                              * Independent of the engine's own `reorder`.        */
    JB_ORDER_REFERENCE_NOREORDER = 2  /* same for the reference run with
                              * reorder=False (template atom order)               */
} jb_hit_order;

/* arguments of Jess.query (src/pyjess/_jess.pyx:2758-2769) that reach the
 * search.  max_candidates (order dependent, :2263-2269) is applied by the host
 * layer on an all_tuples run in JB_ORDER_REFERENCE.                          */
typedef struct jb_query_params {
    double rmsd_threshold;
    double distance_cutoff;          /* Jess_query t  (eff:Jess.c:119)        */
    double max_dynamic_distance;     /* Jess_query s                          */
    int32_t ignore_chain;            /* jb_ignore_chain                       */
    int32_t reorder;                 /* scan order: 0 template order, 1 the
                                      * reference's (Scanner_reuse reorder flag,
                                      * eff:Scanner.c:88-99), 2 the engine's own
                                      * fan-out heuristic; the hit SET is the same */
    int32_t best_match;              /* keep min-RMSD hit per (mol, template) */
    int32_t hit_order;               /* jb_hit_order                          */
    int32_t all_tuples;              /* 1: report EVERY full tuple with its
                                      * superposition (no RMSD filter, NaN
                                      * rotations kept): what JessQuery_next
                                      * yields before Query.__next__ filters   */
    int32_t flags;                   /* JB_QUERY_* bits                       */
} jb_query_params;

/* jb_query_params.flags */
#define JB_QUERY_COORDS 1            /* also return, per hit, the coordinates of the matched atoms
                                      * transformed into the template frame on the device: what
                                      * Hit.atoms(transform=True) computes atom by atom
                                      * (src/pyjess/_jess.pyx:2454-2482 with Hit.transformation
                                      * :2343-2377 and Atom._transform :1293-1307, same operation
                                      * order) -> jb_hits.coords                                */

/* compacted hits, host memory owned by the jb_result they were fetched from.
 * Order: (molecule, template) ascending, then jb_query_params.hit_order.
 * Record i: atoms[i*stride + k] = index (within its molecule) of the atom
 * matched to template atom k (JessQuery_atoms, eff:Jess.c:169-174); -1 pad.
 * rotation/centre0/centre1 as Superposition_rotation / _centroid(0|1)
 * (vendor/jess/src/Super.c:222-233): x' = R (x - centre0) + centre1.         */
typedef struct jb_hits {
    int64_t count;
    int32_t stride;                  /* max template size of the set          */
    int32_t reserved;
    const int32_t *molecule;         /* [count]                               */
    const int32_t *template_index;   /* [count]                               */
    const int32_t *atoms;            /* [count*stride]                        */
    const double *rmsd;              /* [count]                               */
    const double *rotation;          /* [count*9] row major                   */
    const double *centre0;           /* [count*3] centroid of matched atoms   */
    const double *centre1;           /* [count*3] centroid of template atoms  */
    const double *coords;            /* [count*stride*3] JB_QUERY_COORDS: transformed coordinates of
                                      * atoms[i*stride + k] (zeros in the padding); NULL otherwise */
} jb_hits;

/* device-side counters of one query run (define the algorithmic bytes of the
 * roofline): everything is summed over the pairs of the batch.    */
typedef struct jb_stats {
    int64_t pairs;                   /* M*T                                   */
    int64_t live_pairs;              /* pairs with no empty candidate set     */
    int64_t candidate_atoms;         /* sum_k |C_k| over live pairs           */
    int64_t candidate_records;       /* entries in the per-(mol,spec) lists   */
    int64_t frontier_items;          /* sum_l F_l  (accepted partial matches) */
    int64_t frontier_ints;           /* sum_l l*F_l                           */
    int64_t candidate_tests;         /* (partial match, candidate) evaluations*/
    int64_t scan_tests;              /* FP32 box tests executed after x-range narrowing */
    int64_t box_survivors;           /* ... passing the FP32 join-box filter  */
    int64_t tuples;                  /* full tuples superposed                */
    int64_t hits;                    /* tuples passing the RMSD threshold     */
    int64_t nan_dropped;             /* hits dropped for a NaN rotation       */
    int64_t chunks;                  /* pair chunks processed                 */
    int64_t retries;                 /* chunks re-run after a capacity split  */
    int64_t launches;                /* kernels launched by the run           */
    double kernel_ms;                /* device time of the whole run (events) */
    double classify_ms;              /* ... of the candidate-list kernel      */
    double pairs_ms;                 /* ... of the pair set-up kernels        */
    double expand_ms;                /* ... of the frontier sweep (all levels)*/
    double superpose_ms;             /* ... of the superposition kernel       */
} jb_stats;

typedef struct jb_context jb_context;
typedef struct jb_templates jb_templates;
typedef struct jb_molecules jb_molecules;
typedef struct jb_result jb_result;

/* ---- library -------------------------------------------------------------- */
int jb_abi_version(void);
const char *jb_last_error(void);
int jb_device_count(void);

/* Page-locked ("pinned") host memory for the arrays of a jb_molecule_batch: uploads from it are
 * asynchronous DMA transfers.  Usable with every context of the process.                      */
int jb_host_alloc(int64_t bytes, void **out);
void jb_host_free(void *ptr);

/* One context per GPU.  `stream` is a cudaStream_t (0/NULL = a private
 * non-blocking stream).  `workspace_bytes` bounds the frontier workspace in
 * HBM (0 = default, 1/16 of the device memory).                              */
int jb_context_create(int device, void *stream, int64_t workspace_bytes,
                      jb_context **out);
void jb_context_destroy(jb_context *ctx);
int jb_context_synchronize(jb_context *ctx);
/* The context's device buffers (candidate lists, frontiers, hit tables, ordering workspace,
 * staging of jb_query_host) only grow, to the largest workload and hit volume seen -- the
 * counterpart of the reference's per-query malloc/free (Scanner_create .. JessQuery_free,
 * eff:Jess.c:119-167).  jb_context_trim waits for the stream and frees them all; they are
 * re-allocated on demand.  Runs whose hits were not fetched yet are discarded (their
 * jb_result_fetch fails with JB_ERR_INVALID).  jb_query_run / jb_query_host /
 * jb_molecules_upload do this themselves, once, before they report JB_ERR_NOMEM.          */
int jb_context_trim(jb_context *ctx);

/* ---- templates: replaces Jess_create + Jess_addTemplate (jess.pxd:18-21,
 * eff:Jess.c:79-117) and TessTemplate_copy (eff:TessTemplate.c:388-460).
 * De-duplicates atom specs and uploads the replicated tables to the GPU.     */
int jb_templates_create(jb_context *ctx, const jb_template_table *table,
                        jb_templates **out);
void jb_templates_free(jb_templates *tpl);
int jb_templates_spec_count(const jb_templates *tpl);

/* ---- molecules: replaces Molecule_* (molecule.pxd:17-21) and
 * ResIndex_create (res_index.pxd:13, eff:ResIndex.c:28-113).  Copies the SoA
 * host image (pinned or pageable) to HBM on the context stream.  One upload
 * (and one jb_query_host / jb_query_records batch) holds fewer than 2^32 atoms:
 * the sweep addresses atoms with 32-bit offsets; JB_ERR_INVALID beyond.     */
int jb_molecules_upload(jb_context *ctx, const jb_molecule_batch *batch,
                        jb_molecules **out);
void jb_molecules_free(jb_molecules *mol);

/* ---- query: replaces Jess_query + the JessQuery_next loop +
 * JessQuery_superposition + Superposition_rmsd/rotation/centroid +
 * the RMSD/NaN filter of Query.__next__ (jess.pxd:21-29, super.pxd:8-14,
 * src/pyjess/_jess.pyx:2197-2283) for EVERY (molecule, template) pair of the
 * batch.  Launches the kernels on the context stream and returns without
 * waiting for them; hits stay in HBM until jb_result_fetch.                  */
int jb_query_run(jb_context *ctx, const jb_templates *tpl,
                 const jb_molecules *mol, const jb_query_params *params,
                 jb_result **out);
/* Waits for the run, copies the compacted hits to host memory (device->host)
 * and sorts them canonically.                                                */
int jb_result_fetch(jb_result *res, jb_hits *out);
int jb_result_stats(jb_result *res, jb_stats *out);
void jb_result_free(jb_result *res);

/* End-to-end convenience: upload + run + fetch for host buffers.             */
int jb_query_host(jb_context *ctx, const jb_templates *tpl,
                  const jb_molecule_batch *batch,
                  const jb_query_params *params, jb_result **out,
                  jb_hits *hits);

/* ---- structure ingest (HOST code): replaces the per-line
 * PDBMoleculeParser.load loop (src/pyjess/_io.py:68-103) + Atom.loads
 * (src/pyjess/_jess.pyx:931-961) + Atom_parse (vendor/jess/src/Atom.c:32-72 as
 * patched) with ONE pass over the file image.  A record is the reference's
 * struct _Atom (eff:Atom.h:29-47) without padding, so that it maps 1:1 onto the
 * host layer's numpy record dtype.                                            */
#pragma pack(push, 1)
typedef struct jb_atom_record {
    double x[3];
    float occupancy;
    float temp_factor;
    int32_t serial;
    char name[4];                    /* blanks -> '_'                         */
    uint8_t alt_loc;
    char res_name[4];                /* 3 chars + NUL                         */
    uint8_t chain_id1;
    uint8_t chain_id2;
    int32_t res_seq;
    uint8_t i_code;
    char seg_id[4];
    char element[2];
    int8_t charge;
} jb_atom_record;
#pragma pack(pop)

#define JB_PDB_SKIP_HETATM 1         /* PDBMoleculeParser(skip_hetatm=True)   */
#define JB_PDB_IGNORE_ENDMDL 2       /* PDBMoleculeParser(ignore_endmdl=True) */

typedef struct jb_pdb_info {
    int64_t n_atoms;                 /* ATOM/HETATM records found (may exceed capacity) */
    int64_t header_offset;           /* first HEADER line, -1 if none         */
    int64_t header_length;
    int64_t stop_offset;             /* where the scan stopped (ENDMDL / end) */
    int32_t mmcif_tags;              /* a data_/loop_ line was met: not PDB   */
    int32_t n_headers;               /* HEADER lines before stop_offset       */
} jb_pdb_info;

/* Parses the ATOM/HETATM lines of `text` up to the first ENDMDL into `records`.
 * JB_ERR_CAPACITY: info->n_atoms tells the capacity to retry with.           */
int jb_pdb_parse(const char *text, int64_t n_bytes, int32_t flags,
                 jb_atom_record *records, int64_t capacity, jb_pdb_info *info);

/* ---- mmCIF ingest (HOST code): replaces gemmi's reader + the row loop of
 * CIFMoleculeParser._load_block (src/pyjess/_io.py:167-219) for well-formed files: one pass over
 * the file image, the rows of the loop holding `_atom_site.id` -> `records`.  `flags`: the two
 * JB_PDB_* flags (skip_hetatm; ignore_endmdl = every pdbx_PDB_model_num) and JB_CIF_USE_AUTHOR.
 * Whatever the reference answers with an exception, and number spellings beyond plain decimals,
 * are left to the host layer's generic loader: info->fallback != 0 (no records are valid).      */
#define JB_CIF_MAX_META 16
#define JB_CIF_USE_AUTHOR 4          /* CIFMoleculeParser(use_author=True): auth_* columns      */

typedef enum jb_cif_fallback {
    JB_CIF_NONE = 0,
    JB_CIF_SYNTAX = 1,               /* unterminated text field, value without tag, ragged loop */
    JB_CIF_BLOCKS = 2,               /* not exactly one data_ block (Document.sole_block)       */
    JB_CIF_LAYOUT = 3,               /* no / duplicate / missing _atom_site columns, or pairs   */
    JB_CIF_VALUE = 4                 /* a field Atom() rejects or int()/float() must judge      */
} jb_cif_fallback;

typedef struct jb_cif_info {
    int64_t n_atoms;                 /* records found (may exceed capacity)                     */
    int64_t loop_offset;             /* byte range of the _atom_site loop (`loop_` .. last      */
    int64_t loop_end;                /* value): the metadata reader skips it                    */
    int32_t n_blocks;
    int32_t hetatm_unnumbered;       /* HETATM rows numbered from the running maximum: one
                                      * UserWarning each in the reference (_io.py:185-199)      */
    int32_t fallback;                /* jb_cif_fallback                                         */
    int32_t n_meta;
    int64_t name_offset;             /* the block name (after `data_`)                          */
    int64_t name_length;
    int64_t meta[JB_CIF_MAX_META][2];/* byte ranges [begin, end) of the pairs / loops holding
                                      * _entry.id, _struct_keywords.pdbx_keywords, _struct.title,
                                      * _pdbx_audit_revision_history.revision_date: all the host
                                      * layer still has to read (_io.py:221-243)                */
} jb_cif_info;

/* JB_ERR_CAPACITY: info->n_atoms tells the capacity to retry with.                            */
int jb_cif_parse(const char *text, int64_t n_bytes, int32_t flags, jb_atom_record *records,
                 int64_t capacity, jb_cif_info *info);

/* ---- packing (HOST code): the per-molecule _Atom arrays of M Molecules -> the
 * SoA staging image of jb_molecule_batch (what Molecule.__init__,
 * src/pyjess/_jess.pyx:632-684, hands to the C library one struct at a time).
 * Output arrays hold sum(counts) entries (atom_offset: M+1); serial may be NULL. */
int jb_molecules_pack(const jb_atom_record *const *molecules, const int64_t *counts,
                      int32_t n_molecules, int64_t *atom_offset, double *x, double *y,
                      double *z, uint32_t *name, uint32_t *residue_name,
                      int32_t *residue_number, uint8_t *chain_id2, int32_t *serial);

/* End to end from the host layer's atom records (bj_molecules_pack + bj_query_run in one call;
 * what Jess.query hands to Jess_query, src/pyjess/_jess.pyx:2855-2869, is exactly such an array of struct _Atom
 * per Molecule).  Large batches run in slices: while the GPU searches slice k, host threads pack slice k+1 into
 * page-locked staging and the copy engine uploads it, so the packing pass disappears behind the kernels.
 * Atom.serial is shipped only when params->hit_order needs it.                                           */
int jb_query_records(jb_context *ctx, const jb_templates *tpl,
                     const jb_atom_record *const *molecules, const int64_t *counts,
                     int32_t n_molecules, const jb_query_params *params, jb_result **out,
                     jb_hits *hits);

/* ---- hit output (HOST code): the per-atom sprintf loop of Hit.dump
 * (src/pyjess/_jess.pyx:2589-2620) for n records at once.  `xyz` (n*3, may be
 * NULL) overrides the coordinates (the transformed ones).  Needs 96 bytes of
 * `out` per record; returns the bytes written or a negative jb_status.        */
int64_t jb_format_atoms_pdb(const jb_atom_record *atoms, int64_t n,
                            const double *xyz, char *out, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* JESS_B200_H */
