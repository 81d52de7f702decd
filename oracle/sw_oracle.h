/* oracle/sw_oracle.h - CPU restatement of OSWALD's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg as the checker.  The product (oswald_b200/) never links, loads or calls
 * anything in this directory.
 *
 * Parity pin: the reference ships no golden vectors or tests, so this
 * restatement is pinned against outputs of the reference itself: oracle/_ref/oswald_ref is
 * the reference's own host/src compiled in place (oracle/Makefile) and tests/golden/ holds
 * score rows and top-r lists it printed (tests/golden/make_golden.py), plus the known-answer
 * table of SURVEY.md section 8(c).
 */
#ifndef OSWALD_SW_ORACLE_H
#define OSWALD_SW_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_PAD 23          /* PREPROCESSED_DUMMY_ELEMENT, reference sequences.h:17 */
#define ORC_ROWS 24
#define ORC_COLS 32         /* SUBMAT_COLS, reference submat.h:5 */

/* Residue code of one FASTA letter (reference sequences.c:163-175, :361-371). */
uint8_t orc_encode_letter(char c);
void orc_encode(const char *letters, size_t n, uint8_t *codes);

/* Fills out[24*32] with the named matrix in the reference's layout (submat.c:4-227):
 * out[r*32+c], rows/cols in alphabet order, row 23 and cols 23..31 zero.  0 on success. */
int orc_matrix(const char *name, int8_t *out);

/* perm[k] = input position of the k-th sequence after the reference's stable ascending
 * length sort (sequences.c:1130-1225: '<=' takes left, swap only if '>'). */
void orc_sort_by_length(const uint32_t *lengths, size_t n, uint32_t *perm);

/* Exact affine-gap local score, int32, gap of length k costs go + k*ge, all borders 0
 * (recurrence of HybridSearch.c:842-913; widths of :937-1134 make it exact). */
int32_t orc_sw_score(const uint8_t *a, int m, const uint8_t *b, int n,
                     const int8_t *matrix, int go, int ge);

/* scores[q*n_seqs + s] for every query q and database sequence s (canonical order).
 * q_off has nq+1 entries, db_off has n_seqs+1 entries.  Uses `threads` OpenMP threads. */
void orc_search(const uint8_t *queries, const uint32_t *q_off, int nq,
                const uint8_t *db, const uint64_t *db_off, size_t n_seqs,
                const int8_t *matrix, int go, int ge, int32_t *scores, int threads);

/* Top-r of one score row in the reference's order (utils.c:3-69 run to completion):
 * score descending, ties by HIGHER canonical index first.  Writes min(r,n) entries. */
size_t orc_top_r(const int32_t *scores, size_t n, size_t r, uint32_t *idx_out, int32_t *score_out);

/* The reference's merge sort itself (same comparisons, same recursion) on (score, index);
 * used to show that orc_top_r's closed-form order equals what utils.c actually does. */
void orc_ref_mergesort(int32_t *scores, uint32_t *idx, size_t n);

#ifdef __cplusplus
}
#endif
#endif
