/* include/oswald_cuda.h - C ABI of the B200 (sm_100a) Smith-Waterman scoring library.
 *
 * This is the drop-in boundary for OSWALD's data-parallel hot path.  The reference has no
 * plugin/FFI layer: its device side is reached through OpenCL calls inlined in
 * hybrid_search_avx2()/fpga_search() (reference host/src/HybridSearch.c:76-122, 640-754;
 * FPGAsearch.c:132-274) and its host SIMD team (HybridSearch.c:756-1141).  A maintainer of
 * the reference replaces those regions by the five calls below (INTEGRATION.md shows the
 * patch); everything is plain C: pointers, sizes, no C++ or torch types, no exceptions, and
 * the library never calls exit().
 *
 * Conventions
 *  - residue codes are the reference's (sequences.c:163-175): 0..22 = ABCDEFGHIKLMNPQRSTVWXYZ,
 *    23 = J/O/U/padding;
 *  - a substitution matrix is the reference's 24x32 int8 table (submat.c:4-227), m[r*32+c];
 *  - database sequences are given in the reference's canonical order (stable ascending
 *    length sort of FASTA order, sequences.c:1130-1225); "index" always means that order;
 *  - every score is the exact int32 Gotoh local score the reference's 8->16->32-bit cascade
 *    ends with (HybridSearch.c:831-1134); hits are ordered like utils.c:3-86 leaves them:
 *    score descending, ties by higher index first.
 */
#ifndef OSWALD_CUDA_H
#define OSWALD_CUDA_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define OSW_MATRIX_ROWS 24
#define OSW_MATRIX_COLS 32          /* SUBMAT_COLS, reference submat.h:5 */
#define OSW_PAD_CODE 23             /* PREPROCESSED_DUMMY_ELEMENT, reference sequences.h:17 */
#define OSW_MAX_QUERY_LEN 65535     /* query lengths are u16 in the reference (sequences.c:226) */

enum {
    OSW_OK = 0,
    OSW_E_ARG = -1,        /* bad argument */
    OSW_E_NODEV = -2,      /* no usable CUDA device / device index out of range */
    OSW_E_CUDA = -3,       /* a CUDA runtime call failed (osw_last_error has the text) */
    OSW_E_NOMEM = -4,      /* host or device allocation failed */
    OSW_E_STATE = -5,      /* call order wrong (e.g. search before db_load) */
    OSW_E_ARCH = -6,       /* device is not sm_100 */
    OSW_E_IO = -7,         /* a file cannot be opened / written */
    OSW_E_FORMAT = -8      /* not an X.osw file of this format version, or truncated / corrupt */
};

typedef struct osw_ctx osw_ctx;

/* One line of the report (reference HybridSearch.c:1219-1222: "%d\t%s" score, title). */
typedef struct osw_hit {
    int32_t  score;
    uint32_t index;        /* canonical database index */
} osw_hit;

/* Which kernels osw_search may use (mask). Default = all. */
enum {
    OSW_K_U16 = 1,         /* packed 16-bit DPX inter-task kernel (first stage) */
    OSW_K_I32 = 2,         /* 32-bit kernel: re-score of flagged pairs; alone = score everything at 32 bit */
    OSW_K_DEFAULT = 3,
    /* how the two 16-bit halves of the first stage are used (default: chosen per query set) */
    OSW_K_TWO_TRACK = 4,   /* always two query tracks against one database sequence */
    OSW_K_PAIR_DB = 8,     /* always one query track against two database sequences */
    OSW_K_TRANSPOSED = 16  /* always the transposed form (database residues as the rows of the array, the query as
                              the column stream; chosen by itself for short queries when it is faster), whenever the
                              queries fit it */
};

typedef struct osw_timing {
    double   device_ms;        /* max over this context's GPUs: scoring + re-score + top-r, CUDA events */
    double   score_ms;         /* first-stage scoring kernels only (same clock) */
    double   rescore_ms;       /* 32-bit re-score */
    double   topr_ms;          /* device top-r selection */
    double   h2d_ms;           /* query/profile upload inside osw_search (wall) */
    double   wall_ms;          /* whole osw_search call (wall) */
    uint64_t cells;            /* sum(query lengths) * sum(database residues of this context) */
    uint64_t padded_cells;     /* cell updates actually issued by the first-stage kernels */
    uint64_t rescored_pairs;   /* (query, sequence) pairs that went to the 32-bit kernel */
    uint64_t launches;         /* kernels launched by this call (all GPUs) */
    uint64_t sm_cycles;        /* busy SM-cycles of the first-stage kernels summed over the SMs of GPU 0 (clock64) */
    uint64_t db_stream_bytes;  /* database bytes read by the first-stage kernels (algorithmic) */
    uint64_t bound_bytes;      /* bottom-row bytes they read and wrote between passes (algorithmic: 8 per column each way) */
    uint64_t score_launches;   /* first-stage scoring launches (all GPUs) */
} osw_timing;

/* ---- device discovery: replaces display_device_info(), reference utils.c:175-253 -------- */
int osw_device_count(int *count);
int osw_device_info(int device, char *text, size_t text_size);

/* ---- context: replaces init()/cleanup(), reference utils.c:99-173, 255-262 -------------
 * n_devices GPUs (devices[i], or 0..n-1 when devices is NULL); one stream set per GPU. */
int  osw_init(int n_devices, const int *devices, osw_ctx **out);
void osw_free(osw_ctx *ctx);

/* ---- database: replaces assemble_db_chunks() + clCreateBuffer/clEnqueueWriteBuffer
 * (reference sequences.c:828-1094, HybridSearch.c:76-122, 694-707).
 * The caller passes the canonical database (read-only, caller-owned): n_seqs sequences,
 * sequence i = residues[offsets[i] .. offsets[i+1]).  The library builds the length-binned
 * chunk streams, deals chunk c to shard (c mod (shard_count * n_devices)) and uploads the
 * chunks of shards [shard_rank*n_devices, (shard_rank+1)*n_devices) to its GPUs.  A
 * single-process run uses shard_rank 0, shard_count 1.  max_chunk_residues is the reference's
 * -k (arguments.c:113-117); 0 = default. */
int osw_db_load(osw_ctx *ctx, const uint8_t *residues, const uint64_t *offsets, uint64_t n_seqs,
                int shard_rank, int shard_count, uint64_t max_chunk_residues);
/* ---- the preprocessed database in its device layout, on disk: X.osw ----------------------------
 * The reference's `-O preprocess` writes X.info / X.seq / X.desc (sequences.c:177-208) and every
 * search re-lays X.seq out for the devices (assemble_db_chunks, sequences.c:828-1094).  Here the
 * layout can be written once: osw_db_write_file builds the length-binned chunk streams of the whole
 * canonical database (same arguments as osw_db_load; host only, needs no GPU) and stores them in a
 * versioned little-endian file (oswald_b200/csrc/host/dbformat.h); osw_db_load_file maps it, copies
 * the chunks of this context's shards (every n-th chunk) straight into pinned memory and uploads
 * them - no per-residue work at search time.  One file serves any number of GPUs / ranks.
 * OSW_E_IO: cannot open / write; OSW_E_FORMAT: not such a file, another version, or corrupt - the
 * caller then falls back to osw_db_load from X.seq. */
int osw_db_write_file(const char *path, const uint8_t *residues, const uint64_t *offsets, uint64_t n_seqs,
                      uint64_t max_chunk_residues);
int osw_db_load_file(osw_ctx *ctx, const char *path, int shard_rank, int shard_count);
/* Header of an X.osw file (every output pointer nullable). */
int osw_db_file_info(const char *path, uint64_t *n_seqs, uint64_t *n_residues, uint64_t *n_chunks,
                     uint32_t *max_len, uint32_t *version);

/* Device window for the database, the GPU meaning of the reference's -k ("maximum chunk size in
 * FPGA (bytes)", arguments.c:113-117): when a GPU's share of the column stream is larger than
 * `bytes`, it is not kept resident; every search streams it from pinned host memory through two
 * windows of that size, the copy of the next segment overlapping the scoring of the current one.
 * 0 (default) = keep the database resident.  Call before osw_db_load. */
int osw_set_device_window(osw_ctx *ctx, uint64_t bytes);
/* Copies the chunk streams (kept in pinned host memory by osw_db_load) to the GPUs again:
 * the host->device leg of a cold search, reference clEnqueueWriteBuffer HybridSearch.c:694-707.
 * bytes (nullable) receives the bytes copied. */
int osw_db_upload(osw_ctx *ctx, uint64_t *bytes);
/* Sequences / residues held by this context after osw_db_load. */
int osw_db_stats(const osw_ctx *ctx, uint64_t *n_seqs_local, uint64_t *residues_local,
                 uint64_t *n_chunks_local);

/* ---- search: replaces the per-query kernel launches, score read-back, overflow fix-up and
 * sort_scores() (reference HybridSearch.c:640-754, 790-1178, 1213-1224; utils.c:71-86).
 * queries: nq sequences of residue codes, query q = queries[q_off[q] .. q_off[q+1]).
 * hits:    caller-owned, nq*top_r entries; query q's hits start at hits[q*top_r]; n_hits[q]
 *          (nullable) receives min(top_r, sequences held by this context).
 * all_scores: nullable; caller-owned nq*n_seqs int32 (n_seqs = the canonical database size);
 *          entries of sequences held by this context are written at [q*n_seqs + index],
 *          others are left untouched.
 * Residue codes must be 0..23 and matrix entries within -32..31 (the reference's tables span
 * -17..17), otherwise OSW_E_ARG.  Blocking; not re-entrant per context. */
int osw_search(osw_ctx *ctx, const uint8_t *queries, const uint32_t *q_off, int nq,
               const int8_t *matrix, int gap_open, int gap_extend, int top_r,
               osw_hit *hits, uint32_t *n_hits, int32_t *all_scores, osw_timing *timing);

/* Kernel selection mask (OSW_K_*), for parity tests and profiling. */
int osw_set_kernels(osw_ctx *ctx, int mask);

/* ---- host-side top-r merge of several shards' hit lists (reference order, utils.c:3-69).
 * lists[s] has counts[s] entries already in reference order; writes min(top_r, total). */
size_t osw_merge_hits(const osw_hit *const *lists, const uint32_t *counts, int n_lists,
                      uint32_t top_r, osw_hit *out);

/* ---- roofline calibration: issue rate of the packed-16-bit DPX instructions (SURVEY.md
 * section 8(d)).  Runs dependent-chain-free micro-kernels on `device` and reports warp
 * instructions per SM-cycle * 32 = thread instructions per SM-cycle for: [0] VIADDMNMX.U16x2,
 * [1] VIMNMX3.U16x2, [2] the 6-instruction cell-pair step as issued by the u16 kernel
 * (cell updates per SM-cycle), [3] IMAD, [4] SM clock in MHz during the run, [5..9] variants of
 * the cell-pair step (see calib.cu). */
int osw_calibrate(int device, double out[12]);
/* Co-issue probe: for 12 instruction classes (VIADDMNMX.U16x2, VIMNMX3.U16x2, VIADD, IMAD, HMNMX2,
 * VIMNMX.U16x2, VIMNMX.U32, LOP3, FMNMX, PRMT, SHF, HADD2) out[3k..3k+2] = thread instructions per
 * SM-cycle of the class alone, of 8 VIADDMNMX + 8 of it, of 8 VIADDMNMX + 4 of it.  n_out >= 36; with
 * n_out >= 45 three more classes follow (IMAD.HI, LEA.HI, IMAD by a run-time 65536). */
int osw_calibrate_mix(int device, double *out, int n_out);

/* ---- substitution matrices: the reference's eight tables (submat.c:4-227), selected by the
 * -s name (arguments.c:94-111).  out[24*32], m[r*32+c].  0 on success, -1 unknown name. */
int osw_matrix_count(void);
const char *osw_matrix_name(int k);
int osw_matrix_by_name(const char *name, int8_t *out);

const char *osw_strerror(int code);
const char *osw_last_error(void);      /* text of the last CUDA failure on this thread */

#ifdef __cplusplus
}
#endif
#endif
