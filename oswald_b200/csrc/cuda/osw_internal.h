// osw_internal.h - shared declarations of the CUDA side (not part of the C ABI).
#ifndef OSW_INTERNAL_H
#define OSW_INTERNAL_H
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../../include/oswald_cuda.h"
#include "../host/dbformat.h"

#define OSW_SCORE_FLAGGED 0x7fffffff   // first-stage marker: pair must be re-scored at 32 bit

// ---- 32-bit kernel (sw_i32.cu): exact scores; re-score of flagged pairs, or everything ----
struct I32Params {
    const uint8_t  *stream;      // shard column stream
    const uint64_t *seq_off;     // [n_seqs] stream offset of each sequence
    const uint32_t *seq_len;     // [n_seqs]
    const uint8_t  *queries;     // residue codes, all queries back to back
    const uint32_t *q_off;       // [nq+1]
    const int8_t   *matrix;      // [24*32]
    const uint2    *pairs;       // (query, local sequence) list, or nullptr = all pairs
    const uint64_t *task_off;    // optional: stream offset of each task's sequence (replaces seq_off[s]; staged re-score)
    uint64_t        n_tasks;     // pairs in the list (an upper bound when n_tasks_dev is set), or nq*n_seqs
    const uint32_t *n_tasks_dev; // optional: the list's length, read on the device (min with n_tasks) - lets the
                                 // re-score be enqueued before the host knows how many pairs the first stage flagged
    uint64_t        n_seqs;
    int32_t        *scores;      // [nq][n_seqs]
    int2           *scratch;     // [warps_in_grid][max_len] (H,F) of a pass's bottom row
    uint32_t        max_len;
    int             gap_open_extend;   // go+ge
    int             gap_extend;
    unsigned long long *task_counter;  // dynamic task queue
};
void osw_launch_i32(const I32Params &p, int n_blocks, cudaStream_t st);
int  osw_i32_block_threads();

// ---- packed 16-bit DPX kernel (sw_u16.cu) ------------------------------------------------
// A launch ("pass") runs G lanes x R rows per database sequence in each 16-bit half of the
// packed words.  The rows of a half are a stretch of that half's TRACK: the queries assigned
// to the half, laid end to end with every query starting on a lane boundary.  A lane therefore
// belongs to exactly one query per half.
#define OSW_PROFILE_BYTES (232u * 1024u)   // room for the largest profile image
#define OSW_LANE_START 1u    // the lane holds the query's first row: its input from above is "no row" (zeros)
#define OSW_LANE_EMIT  2u    // the lane is the last one of its query in this pass: it publishes the maximum
struct OswLaneDesc {
    uint32_t query;      // index of the query (score row), or 0xffffffff for an idle lane
    uint32_t q_len;      // its length (0 for an idle lane)
    uint32_t row0;       // first row of the query held by this lane in this pass
    uint32_t flags;      // OSW_LANE_*
};
struct OswPass {
    int G, R;
    OswLaneDesc lane[2][32];   // [half][lane of group]
    int has_in;          // some half continues a query from the previous pass (bottom-row hand-over)
    int has_out;         // some half's last lane holds a query that continues in the next pass
    int pair_db;         // pair-database mode: lane[1] == lane[0]; the halves score two database sequences
};
// Lays the queries out on two tracks and cuts the tracks into passes.  Returns the number of
// passes (<= max_passes), or -1 if they do not fit.  Exported for the tests.
#define OSW_PLAN_AUTO 0
#define OSW_PLAN_TWO_TRACK 1
#define OSW_PLAN_PAIR_DB 2
// min_g: smallest group width allowed (4 = no restriction; 32 = one sequence per warp).
extern "C" int osw_plan_passes(const uint32_t *q_len, int nq, OswPass *out, int max_passes, int mode, int min_g);
// rmax: rows per lane of the full-height passes (0 = default 40; experiments).
int osw_plan_passes_ex(const uint32_t *q_len, int nq, OswPass *out, int max_passes, int mode, int min_g, int rmax);

struct U16Params {
    const uint8_t   *stream;      // the columns [stream_col0, ...) of the shard's stream (all of it, or a window)
    const uint8_t   *pair_stream; // two bytes per column (pair-database mode); same window convention
    uint64_t         stream_col0;
    const osw_chunk *chunks;     // descending-length order
    uint32_t         chunk_first, chunk_end;   // the launch works on chunks [chunk_first, chunk_end)
    const uint8_t   *queries;    // all queries back to back (codes)
    const uint32_t  *q_off;      // [nq+1]
    const int8_t    *matrix;
    unsigned char   *profile;    // global image of the pass's profile table (>= OSW_PROFILE_BYTES), private to the stream
    int32_t         *scores;     // [nq][n_seqs]
    uint64_t         n_seqs;
    uint2           *bound;      // (H,F) bottom rows handed from pass to pass, in place, for the columns
                                 // [bound_col0, bound_col0 + capacity) of the (pair) stream; or nullptr
    uint64_t         bound_col0;
    int              gap_open_extend, gap_extend;
    uint32_t        *chunk_counter;
    uint32_t         static_first;   // deal every warp's first chunk statically, through the table (set by the launcher: few chunks per warp)
    uint32_t         express_ctas;   // CTAs that give the longest chunks a scheduler each (0 = none)
    uint32_t         all_express;    // the launch consists of express CTAs only (the long-chunk launch of a small database)
    uint32_t        *first_table;    // [CTAs x warps] first chunk group of every warp (filled by profile_build_kernel), >= 148 * 16 entries
    uint32_t         dyn_base;       // first group handed out by the counter (set by the launcher)
    unsigned long long *cycle_acc;   // sum over CTAs of their elapsed clock64 cycles (one CTA per SM), or nullptr
    // Pipelined passes over the longest chunks (api.cu): the launches of consecutive passes run at the same
    // time on different SMs; progress_out[k] = columns of chunk chunk_first + k whose bottom row this pass has
    // written, progress_in[k] = the same of the pass before, which this launch waits for block by block.
    uint32_t        *progress_in, *progress_out;
};
// n_ctas CTAs (normally one per SM; a long-chunk launch and the launch beside it share the SMs).
int osw_launch_u16(const U16Params &p, const OswPass &pass, int n_ctas, cudaStream_t st);

// ---- device top-r (topr.cu) --------------------------------------------------------------
struct TopRWork {            // per-device scratch, sized for nq_max queries
    uint32_t *hist;          // [nq][8 rounds][256]
    unsigned long long *prefix;   // [nq][8] key prefix after each round
    uint32_t *remaining;     // [nq][8] rank still to find inside the prefix after each round
    uint32_t *out_count;     // [nq]
    unsigned long long *out_keys;  // [nq][r]
};
#define OSW_TOPR_SMALL_MAX 24576      // sequences per shard up to which top-r is one launch (keys in shared memory)
// Selects for each of nq rows the top_r largest keys (score<<32 | canonical index) into
// w.out_keys (unordered).  Returns number of kernels launched.
// flags (optional): the first scan over the scores also lists every (query, local sequence) whose
// score is OSW_SCORE_FLAGGED - up to flags->capacity pairs, *flags->count counts all of them.  When
// that count comes back non-zero the selection is void: the caller re-scores the pairs and selects again.
struct FlagList { uint2 *pairs; uint32_t *count; uint32_t capacity; };
int osw_topr_select(const int32_t *scores, const uint32_t *canon, uint64_t n_seqs, uint64_t n_canon, int nq,
                    uint32_t top_r, const TopRWork &w, const FlagList *flags, cudaStream_t st);
// Marks flagged scores: appends (q, seq) of every score == OSW_SCORE_FLAGGED to pairs.
int osw_collect_flagged(const int32_t *scores, uint64_t n_seqs, int nq, uint2 *pairs,
                        uint32_t *count, uint32_t capacity, cudaStream_t st);

#endif
