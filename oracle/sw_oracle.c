/* oracle/sw_oracle.c - see sw_oracle.h.  TEST INFRASTRUCTURE ONLY (parity checker). */
#include "sw_oracle.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- residue codes: reference sequences.c:163-175 ------------------------------------
 * J, O, U become the dummy 'Z'+1; then subtract 'A' plus one for each of J/O/U below it. */
uint8_t orc_encode_letter(char c) {
    int x = (unsigned char)c;
    if (x == 'J' || x == 'O' || x == 'U') x = 'Z' + 1;
    int d = 'A';
    if (x > 'J') d++;
    if (x > 'O') d++;
    if (x > 'U') d++;
    return (uint8_t)(x - d);
}
void orc_encode(const char *s, size_t n, uint8_t *out) {
    for (size_t i = 0; i < n; ++i) out[i] = orc_encode_letter(s[i]);
}

/* ---- matrices: reference submat.c:4-227 ------------------------------------------------ */
static const struct { const char *name; int8_t v[23][23]; } orc_tables[] = {
#include "submat_rows.inc"
};
int orc_matrix(const char *name, int8_t *out) {
    for (size_t t = 0; t < sizeof orc_tables / sizeof orc_tables[0]; ++t) {
        if (strcmp(orc_tables[t].name, name) != 0) continue;
        memset(out, 0, ORC_ROWS * ORC_COLS);
        for (int r = 0; r < 23; ++r)
            for (int c = 0; c < 23; ++c) out[r * ORC_COLS + c] = orc_tables[t].v[r][c];
        return 0;
    }
    return -1;
}

/* ---- stable length sort: reference sequences.c:1130-1225 ------------------------------- */
static void sort_rec(uint32_t *perm, uint32_t *tmp, const uint32_t *len, size_t n) {
    if (n < 2) return;
    size_t h = n / 2;
    sort_rec(perm, tmp, len, h);
    sort_rec(perm + h, tmp, len, n - h);
    size_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = (len[perm[i]] <= len[perm[j]]) ? perm[i++] : perm[j++];
    while (i < h) tmp[k++] = perm[i++];
    while (j < n) tmp[k++] = perm[j++];
    memcpy(perm, tmp, n * sizeof *perm);
}
void orc_sort_by_length(const uint32_t *lengths, size_t n, uint32_t *perm) {
    uint32_t *tmp = (uint32_t *)malloc((n ? n : 1) * sizeof *tmp);
    for (size_t i = 0; i < n; ++i) perm[i] = (uint32_t)i;
    sort_rec(perm, tmp, lengths, n);
    free(tmp);
}

/* ---- Gotoh local score: reference HybridSearch.c:842-913 -------------------------------
 * Per cell (query row i, database column j), exactly the reference's order of operations:
 *   cur  = max(0, H[i-1][j-1] + M[a_i][b_j], maxRow[i], maxCol[j])
 *   maxRow[i] = max(maxRow[i] - ge, cur - (go+ge));  maxCol[j] likewise
 * with every boundary and every maxRow/maxCol starting at 0.  The column blocking of the
 * reference (cpu_block_size) only reorders independent work, so it is not restated.  Here
 * the loops run column-outer so that the state is O(m). */
int32_t orc_sw_score(const uint8_t *a, int m, const uint8_t *b, int n,
                     const int8_t *M, int go, int ge) {
    if (m <= 0 || n <= 0) return 0;
    const int goe = go + ge;
    int32_t *Hcol = (int32_t *)calloc((size_t)m + 1, sizeof(int32_t));  /* H[i][j-1] */
    int32_t *E = (int32_t *)calloc((size_t)m + 1, sizeof(int32_t));     /* maxRow[i] */
    int32_t best = 0;
    for (int j = 0; j < n; ++j) {
        const int bj = b[j];
        int32_t F = 0;            /* maxCol[j] */
        int32_t diag = 0;         /* H[i-1][j-1] */
        for (int i = 0; i < m; ++i) {
            int32_t cur = diag + M[a[i] * ORC_COLS + bj];
            if (cur < E[i]) cur = E[i];
            if (cur < F) cur = F;
            if (cur < 0) cur = 0;
            int32_t open = cur - goe;
            int32_t e = E[i] - ge; E[i] = e > open ? e : open;
            int32_t f = F - ge;    F = f > open ? f : open;
            diag = Hcol[i];
            Hcol[i] = cur;
            if (cur > best) best = cur;
        }
    }
    free(Hcol); free(E);
    return best;
}

void orc_search(const uint8_t *queries, const uint32_t *q_off, int nq,
                const uint8_t *db, const uint64_t *db_off, size_t n_seqs,
                const int8_t *matrix, int go, int ge, int32_t *scores, int threads) {
    long total = (long)nq * (long)n_seqs;
    (void)threads;
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads > 0 ? threads : 1)
    for (long t = 0; t < total; ++t) {
        int q = (int)(t / (long)n_seqs);
        size_t s = (size_t)(t % (long)n_seqs);
        scores[t] = orc_sw_score(queries + q_off[q], (int)(q_off[q + 1] - q_off[q]),
                                 db + db_off[s], (int)(db_off[s + 1] - db_off[s]), matrix, go, ge);
    }
}

/* ---- ranking: reference utils.c:3-69 ----------------------------------------------------
 * merge takes the RIGHT element unless left > right; the size-2 base swaps on <=.  Both
 * put the later (higher-index) element first among equals. */
static void rank_merge(int32_t *s, uint32_t *x, size_t n) {
    size_t i1 = 0, i2 = n / 2, it = 0;
    int32_t *ts = (int32_t *)malloc(n * sizeof *ts);
    uint32_t *tx = (uint32_t *)malloc(n * sizeof *tx);
    while (i1 < n / 2 && i2 < n) {
        if (s[i1] > s[i2]) { ts[it] = s[i1]; tx[it] = x[i1]; i1++; }
        else               { ts[it] = s[i2]; tx[it] = x[i2]; i2++; }
        it++;
    }
    while (i1 < n / 2) { ts[it] = s[i1]; tx[it] = x[i1]; i1++; it++; }
    while (i2 < n)     { ts[it] = s[i2]; tx[it] = x[i2]; i2++; it++; }
    memcpy(s, ts, n * sizeof *ts); memcpy(x, tx, n * sizeof *tx);
    free(ts); free(tx);
}
void orc_ref_mergesort(int32_t *s, uint32_t *x, size_t n) {
    if (n == 2) {
        if (s[0] <= s[1]) {
            int32_t a = s[0]; s[0] = s[1]; s[1] = a;
            uint32_t b = x[0]; x[0] = x[1]; x[1] = b;
        }
    } else if (n > 2) {
        orc_ref_mergesort(s, x, n / 2);
        orc_ref_mergesort(s + n / 2, x + n / 2, n - n / 2);
        rank_merge(s, x, n);
    }
}

static int key_desc(const void *pa, const void *pb) {
    uint64_t a = *(const uint64_t *)pa, b = *(const uint64_t *)pb;
    return a < b ? 1 : (a > b ? -1 : 0);
}
size_t orc_top_r(const int32_t *scores, size_t n, size_t r, uint32_t *idx_out, int32_t *score_out) {
    /* key = score:index; descending key order == (score desc, index desc) for score >= 0 */
    uint64_t *k = (uint64_t *)malloc((n ? n : 1) * sizeof *k);
    for (size_t i = 0; i < n; ++i) k[i] = ((uint64_t)(uint32_t)scores[i] << 32) | (uint32_t)i;
    qsort(k, n, sizeof *k, key_desc);
    if (r > n) r = n;
    for (size_t i = 0; i < r; ++i) { idx_out[i] = (uint32_t)k[i]; score_out[i] = (int32_t)(k[i] >> 32); }
    free(k);
    return r;
}
