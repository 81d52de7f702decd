// sw_t16.h - the TRANSPOSED form of the first stage (sw_t16.cu): the database sequences are the rows
// of the array, the (short) query is the column stream.  Not part of the C ABI.
#ifndef OSW_SW_T16_H
#define OSW_SW_T16_H
#include "osw_internal.h"

#define OSW_T16_CLASSES 4              // gangs of 16, 8, 4 warps and lone warps (8-warp CTAs: 8, 4, -, 1)
#define OSW_T16_MAX_ROWS32 2048        // ceil(65535 / 32)

// What a search needs to know to run in this form (made by osw_t16_plan from the shard's pair lengths).
struct OswT16Plan {
    int      warps;                            // warps per CTA: 16 or 8 (0 = the queries do not fit this form)
    int      rmax;                             // rows per lane at most: 4 (16 warps) or 8
    uint32_t n_pairs;                          // pair p = the shard's sequences (2p, 2p + 1)
    uint32_t class_begin[OSW_T16_CLASSES + 1]; // ranks (0 = longest pair) [class_begin[i], class_begin[i+1]) run with the gang size of class i
    uint32_t q_cols;                           // entries of the column table: sum over the queries of (length + 64)
    uint32_t m_max;                            // longest query
    size_t   smem_bytes;
    double   est_cycles;                       // model: SM cycles of the launch
    uint64_t padded_cells;                     // cell updates including padding rows and the skew of the array
};

struct OswT16Params {
    const uint8_t  *stream;        // the shard's column stream (resident); a sequence's residues are consecutive bytes
    const uint64_t *seq_off;       // [n_seqs]
    const uint32_t *seq_len;       // [n_seqs]
    const uint8_t  *queries;       // residue codes, all queries back to back
    const uint32_t *q_off;         // [nq + 1]
    int             nq;
    const int8_t   *matrix;        // [24 * 32]
    int32_t        *scores;        // [nq][n_seqs], zeroed
    uint64_t        n_seqs;
    uint32_t       *counters;      // [OSW_T16_CLASSES] task counters, zeroed
    unsigned long long *cycle_acc; // sum over CTAs of their elapsed clock64 cycles, or nullptr
    int             gap_open_extend, gap_extend;
};

// Histogram of the shard's pairs by ceil(length of the longer sequence / 32): hist[OSW_T16_MAX_ROWS32 + 1].
void osw_t16_histogram(const uint32_t *seq_len, uint64_t n_seqs, uint32_t *hist);
// Plans the launch for the given queries; plan->warps == 0 when this form cannot take them.  gang_fraction: a pair
// goes to the smallest gang that scores it within this fraction of the launch's estimated time (1.2 measured best on BASELINE.json config 1).
void osw_t16_plan(const uint32_t *hist, uint64_t n_seqs, const uint32_t *q_off, int nq, int n_sms, double gang_fraction, OswT16Plan *plan);
int  osw_launch_t16(const OswT16Params &p, const OswT16Plan &plan, int n_ctas, cudaStream_t st);
#endif
