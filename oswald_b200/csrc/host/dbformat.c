/* dbformat.c - see dbformat.h. */
#include "dbformat.h"
#include <stdlib.h>
#include <string.h>

/* Walks the canonical sequence list and reports chunk boundaries through `cb`.
 * A chunk closes when adding the next sequence would pass chunk_cols (and it is non-empty). */
typedef void (*chunk_cb)(void *user, uint64_t chunk_index, uint64_t first_seq, uint64_t n_seqs, uint64_t n_cols);
static uint64_t walk_chunks(const uint64_t *off, uint64_t n, uint32_t chunk_cols, chunk_cb cb, void *user) {
    uint64_t c = 0, first = 0, cols = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t len = off[i + 1] - off[i];
        /* empty sequences (they lead the ascending order) have no column, so they cannot share a
         * chunk with real ones: the kernel counts sequences by their LAST columns */
        if (i > first && (cols + len > chunk_cols || (cols == 0 && len > 0))) {
            if (cb) cb(user, c, first, i - first, cols);
            ++c; first = i; cols = 0;
        }
        cols += len;
    }
    if (n > first) { if (cb) cb(user, c, first, n - first, cols); ++c; }
    return c;
}

uint64_t osw_count_chunks(const uint64_t *offsets, uint64_t n_seqs, uint32_t chunk_cols) {
    if (!chunk_cols) chunk_cols = OSW_CHUNK_COLS_DEFAULT;
    return walk_chunks(offsets, n_seqs, chunk_cols, NULL, NULL);
}

#define OSW_PAIR_ALIGN 64      /* columns */
typedef struct { uint32_t shard, n_shards; uint64_t seqs, cols, bytes, chunks, pair_cols; uint32_t max_len;
                 const uint64_t *off; } tally_t;
/* columns of a chunk in the pair stream: sum over pairs of the longer length (ascending order:
 * the second sequence of a pair, or the single last one) */
static uint64_t pair_columns(const uint64_t *off, uint64_t first, uint64_t ns) {
    uint64_t cols = 0;
    for (uint64_t k = 0; k < ns; k += 2) {
        uint64_t i = first + (k + 1 < ns ? k + 1 : k);
        cols += off[i + 1] - off[i];
    }
    return cols;
}
static void tally_cb(void *u, uint64_t c, uint64_t first, uint64_t ns, uint64_t cols) {
    tally_t *t = (tally_t *)u;
    if (c % t->n_shards != t->shard) return;
    t->seqs += ns; t->cols += cols; t->chunks++;
    t->bytes += (cols + OSW_CHUNK_ALIGN - 1) / OSW_CHUNK_ALIGN * OSW_CHUNK_ALIGN;
    t->pair_cols += (pair_columns(t->off, first, ns) + OSW_PAIR_ALIGN - 1) / OSW_PAIR_ALIGN * OSW_PAIR_ALIGN;
    uint64_t last = t->off[first + ns] - t->off[first + ns - 1];   /* longest: order is ascending */
    if (last > t->max_len) t->max_len = (uint32_t)last;
}

typedef struct { uint32_t shard, n_shards; const uint8_t *res; const uint64_t *off; osw_shard *s;
                 uint64_t seq_cursor, byte_cursor, pair_cursor; uint32_t chunk_cursor; int bad_residue; } fill_t;
static void fill_cb(void *u, uint64_t c, uint64_t first, uint64_t ns, uint64_t cols) {
    fill_t *f = (fill_t *)u;
    if (c % f->n_shards != f->shard) return;
    osw_shard *s = f->s;
    /* chunks are stored in reverse so that index 0 is the longest-sequence chunk */
    osw_chunk *ck = &s->chunks[s->n_chunks - 1 - f->chunk_cursor++];
    ck->stream_off = f->byte_cursor; ck->n_cols = (uint32_t)cols; ck->n_seqs = (uint32_t)ns;
    ck->seq0 = (uint32_t)f->seq_cursor; ck->canon0 = (uint32_t)first;
    uint8_t *p = s->stream + f->byte_cursor;
    for (uint64_t i = first; i < first + ns; ++i) {
        uint64_t len = f->off[i + 1] - f->off[i];
        const uint8_t *src = f->res + f->off[i];
        uint64_t l = f->seq_cursor++;
        s->canon[l] = (uint32_t)i; s->seq_off[l] = (uint64_t)(p - s->stream); s->seq_len[l] = (uint32_t)len;
        for (uint64_t k = 0; k < len; ++k) {
            if (src[k] > 23) f->bad_residue = 1;               /* codes are 0..23 (sequences.c:163-175) */
            p[k] = (uint8_t)(src[k] & OSW_COL_CODE);
        }
        if (len) { p[0] |= OSW_COL_FIRST; p[len - 1] |= OSW_COL_LAST; }
        p += len;
    }
    uint64_t padded = (cols + OSW_CHUNK_ALIGN - 1) / OSW_CHUNK_ALIGN * OSW_CHUNK_ALIGN;
    memset(p, OSW_COL_PADBYTE, padded - cols);
    f->byte_cursor += padded;
    /* pair stream of the same chunk */
    ck->pair_off = f->pair_cursor; ck->reserved = 0;
    uint8_t *q = s->pair_stream + 2 * f->pair_cursor;
    uint64_t pc = 0;
    for (uint64_t k = 0; k < ns; k += 2) {
        const uint64_t ia = first + k, ib = first + k + 1;
        const int has_b = k + 1 < ns;
        const uint64_t la = f->off[ia + 1] - f->off[ia], lb = has_b ? f->off[ib + 1] - f->off[ib] : 0;
        const uint64_t n = la > lb ? la : lb;
        const uint8_t *a = f->res + f->off[ia], *b = has_b ? f->res + f->off[ib] : NULL;
        for (uint64_t j = 0; j < n; ++j) {
            q[2 * j] = (uint8_t)(j < la ? (a[j] & OSW_COL_CODE) : OSW_COL_PADBYTE);
            q[2 * j + 1] = (uint8_t)(j < lb ? (b[j] & OSW_COL_CODE) : OSW_COL_PADBYTE);
        }
        if (n) { q[0] |= OSW_COL_FIRST; q[2 * (n - 1)] |= OSW_COL_LAST; }
        q += 2 * n; pc += n;
    }
    ck->n_pair_cols = (uint32_t)pc;
    const uint64_t pc_padded = (pc + OSW_PAIR_ALIGN - 1) / OSW_PAIR_ALIGN * OSW_PAIR_ALIGN;
    memset(q, OSW_COL_PADBYTE, 2 * (pc_padded - pc));
    f->pair_cursor += pc_padded;
}

int osw_shard_build(const uint8_t *residues, const uint64_t *offsets, uint64_t n_seqs,
                    uint32_t shard, uint32_t n_shards, uint32_t chunk_cols, osw_shard *out) {
    if (!out || !n_shards || shard >= n_shards || (n_seqs && (!residues || !offsets))) return -1;
    if (!chunk_cols) chunk_cols = OSW_CHUNK_COLS_DEFAULT;
    memset(out, 0, sizeof *out);
    tally_t t; memset(&t, 0, sizeof t);
    t.shard = shard; t.n_shards = n_shards; t.off = offsets;
    walk_chunks(offsets, n_seqs, chunk_cols, tally_cb, &t);
    out->n_seqs = t.seqs; out->n_residues = t.cols; out->stream_bytes = t.bytes;
    out->n_chunks = (uint32_t)t.chunks; out->max_len = t.max_len;
    out->pair_cols = t.pair_cols;
    out->stream  = (uint8_t *)malloc(t.bytes ? t.bytes : 1);
    out->pair_stream = (uint8_t *)malloc(t.pair_cols ? 2 * t.pair_cols : 1);
    out->chunks  = (osw_chunk *)malloc((t.chunks ? t.chunks : 1) * sizeof(osw_chunk));
    out->canon   = (uint32_t *)malloc((t.seqs ? t.seqs : 1) * sizeof(uint32_t));
    out->seq_off = (uint64_t *)malloc((t.seqs ? t.seqs : 1) * sizeof(uint64_t));
    out->seq_len = (uint32_t *)malloc((t.seqs ? t.seqs : 1) * sizeof(uint32_t));
    if (!out->stream || !out->pair_stream || !out->chunks || !out->canon || !out->seq_off || !out->seq_len) {
        osw_shard_free(out); return -1;
    }
    fill_t f; memset(&f, 0, sizeof f);
    f.shard = shard; f.n_shards = n_shards; f.res = residues; f.off = offsets; f.s = out;
    walk_chunks(offsets, n_seqs, chunk_cols, fill_cb, &f);
    if (f.bad_residue) { osw_shard_free(out); return -2; }
    return 0;
}

void osw_shard_free(osw_shard *s) {
    if (!s) return;
    free(s->stream); free(s->pair_stream); free(s->chunks); free(s->canon); free(s->seq_off); free(s->seq_len);
    memset(s, 0, sizeof *s);
}
