/* dbformat.c - see dbformat.h.
 *
 * Two steps: a serial walk over the sequence lengths that lays out this shard's chunks (offsets
 * into both streams, first sequence), then a fill of the chunks, which are independent of each
 * other (OpenMP).  The two big buffers can come from a caller-supplied allocator (the CUDA
 * library passes pinned memory so that the streams are copied to the GPU straight from here). */
#include "dbformat.h"
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#define OSW_PAIR_ALIGN 64      /* columns */

static int three_levels(void) {          /* OSW_CHUNK_LEVELS=2: the two-level grading (A/B experiments) */
    const char *e = getenv("OSW_CHUNK_LEVELS");
    return !(e && atoi(e) == 2);
}

/* Walks the canonical sequence list and reports chunk boundaries through `cb`.
 * A chunk closes when adding the next sequence would pass chunk_cols (and it is non-empty). */
typedef void (*chunk_cb)(void *user, uint64_t chunk_index, uint64_t first_seq, uint64_t n_seqs, uint64_t n_cols);
static uint64_t walk_chunks(const uint64_t *off, uint64_t n, uint32_t chunk_cols, chunk_cb cb, void *user) {
    uint64_t c = 0, first = 0, cols = 0;
    /* The walk goes from the shortest sequences to the longest; the GPU takes chunks in the opposite
     * order.  The chunks taken LAST are smaller - the first twelfth of the residues here a quarter
     * of the size, the first forty-eighth a sixteenth: fine-grained work to even out the end of a
     * launch, coarse work (less pipeline fill) before it.  Measured at config 2: SMs busy 98.8 % of
     * a launch with two levels, 99.6 % with three (+ 0.4 % GCUPS). */
    const uint64_t fine_until = n ? off[n] / 12 : 0, finest_until = n && three_levels() ? off[n] / 48 : 0;
    const uint32_t coarse = chunk_cols, fine = chunk_cols >= 1024 ? chunk_cols / 4 : chunk_cols;
    const uint32_t finest = chunk_cols >= 4096 ? chunk_cols / 16 : fine;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t len = off[i + 1] - off[i];
        chunk_cols = off[i] < finest_until ? finest : off[i] < fine_until ? fine : coarse;
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

typedef struct { uint32_t shard, n_shards; uint64_t seqs, cols, bytes, chunks; uint32_t max_len;
                 const uint64_t *off; } tally_t;
static void tally_cb(void *u, uint64_t c, uint64_t first, uint64_t ns, uint64_t cols) {
    tally_t *t = (tally_t *)u;
    (void)cols;
    if (c % t->n_shards != t->shard) return;
    t->chunks++; t->seqs += ns;
}

/* layout pass: directory entries in walk order (ascending length), offsets assigned in that order */
typedef struct { uint32_t shard, n_shards; const uint64_t *off; osw_chunk *dir; uint64_t *first_seq;
                 uint64_t n, seq_cursor, byte_cursor, cols; uint32_t max_len; } layout_t;
static void layout_cb(void *u, uint64_t c, uint64_t first, uint64_t ns, uint64_t cols) {
    layout_t *l = (layout_t *)u;
    if (c % l->n_shards != l->shard) return;
    osw_chunk *ck = &l->dir[l->n];
    l->first_seq[l->n] = first;
    ck->stream_off = l->byte_cursor; ck->n_cols = (uint32_t)cols; ck->n_seqs = (uint32_t)ns;
    ck->seq0 = (uint32_t)l->seq_cursor; ck->canon0 = (uint32_t)first;
    ck->pair_off = 0; ck->n_pair_cols = 0; ck->reserved = 0;          /* (pair directory only) */
    l->byte_cursor += (cols + OSW_CHUNK_ALIGN - 1) / OSW_CHUNK_ALIGN * OSW_CHUNK_ALIGN;
    l->seq_cursor += ns; l->cols += cols;
    uint64_t last = l->off[first + ns] - l->off[first + ns - 1];   /* longest: order is ascending */
    if (last > l->max_len) l->max_len = (uint32_t)last;
    l->n++;
}

/* fills one chunk (both streams and the per-sequence tables); returns 1 if a residue code is invalid */
static int fill_chunk(const uint8_t *res, const uint64_t *off, const osw_chunk *ck, uint64_t first, osw_shard *s) {
    int bad = 0;
    const uint64_t ns = ck->n_seqs;
    uint8_t *p = s->stream + ck->stream_off;
    for (uint64_t k = 0; k < ns; ++k) {
        const uint64_t i = first + k, len = off[i + 1] - off[i];
        const uint8_t *src = res + off[i];
        const uint64_t l = ck->seq0 + k;
        s->canon[l] = (uint32_t)i; s->seq_off[l] = (uint64_t)(p - s->stream); s->seq_len[l] = (uint32_t)len;
        for (uint64_t j = 0; j < len; ++j) {
            if (src[j] > 23) bad = 1;                           /* codes are 0..23 (sequences.c:163-175) */
            p[j] = (uint8_t)(src[j] & OSW_COL_CODE);
        }
        if (len) { p[0] |= OSW_COL_FIRST; p[len - 1] |= OSW_COL_LAST; }
        p += len;
    }
    const uint64_t padded = ((uint64_t)ck->n_cols + OSW_CHUNK_ALIGN - 1) / OSW_CHUNK_ALIGN * OSW_CHUNK_ALIGN;
    memset(p, OSW_COL_PADBYTE, padded - ck->n_cols);
    return bad;
}

/* The pair directory: the shard's sequences (local order = ascending length) taken two at a time,
 * pairs grouped into chunks of about chunk_cols / 2 pair columns (the same work as a plain chunk;
 * graded like the plain ones).  Its own walk rather than the plain chunks zipped, so that every
 * pair chunk holds whole PAIRS whatever the plain chunk size is - with small chunks (small
 * databases) most plain chunks hold a single sequence.  Pairs without columns (two empty
 * sequences) get a chunk of their own, like empty sequences in the plain directory. */
static int build_pair_directory(osw_shard *s, uint32_t chunk_cols) {
    const uint64_t n_pairs = (s->n_seqs + 1) / 2;
    osw_chunk *dir = (osw_chunk *)malloc((n_pairs ? n_pairs : 1) * sizeof(osw_chunk));
    if (!dir) return -1;
    const uint32_t coarse = chunk_cols / 2 ? chunk_cols / 2 : 1, fine = chunk_cols >= 1024 ? coarse / 4 : coarse;
    const uint32_t finest = chunk_cols >= 4096 ? coarse / 16 : fine;
    const uint64_t fine_until = s->n_residues / 12, finest_until = three_levels() ? s->n_residues / 48 : 0;
    uint64_t n = 0, cursor = 0, seen = 0;
    uint64_t first = 0, cols = 0;          /* the open chunk: first sequence, pair columns so far */
    for (uint64_t p = 0; p <= n_pairs; ++p) {
        const uint64_t i = p < n_pairs ? 2 * p : s->n_seqs;
        const uint64_t plen = p < n_pairs ? s->seq_len[i + 1 < s->n_seqs ? i + 1 : i] : 0;
        const uint32_t target = seen < finest_until ? finest : seen < fine_until ? fine : coarse;
        if (i > first && (p == n_pairs || cols + plen > target || (cols == 0 && plen > 0))) {
            osw_chunk *ck = &dir[n++];
            memset(ck, 0, sizeof *ck);
            ck->n_seqs = (uint32_t)(i - first); ck->seq0 = (uint32_t)first; ck->canon0 = s->canon[first];
            ck->pair_off = cursor; ck->n_pair_cols = (uint32_t)cols;
            cursor += (cols + OSW_PAIR_ALIGN - 1) / OSW_PAIR_ALIGN * OSW_PAIR_ALIGN;
            first = i; cols = 0;
        }
        if (p < n_pairs) {
            cols += plen;
            seen += s->seq_len[i] + (i + 1 < s->n_seqs ? s->seq_len[i + 1] : 0);
        }
    }
    s->pair_chunks = (osw_chunk *)malloc((n ? n : 1) * sizeof(osw_chunk));
    if (!s->pair_chunks) { free(dir); return -1; }
    for (uint64_t k = 0; k < n; ++k) s->pair_chunks[n - 1 - k] = dir[k];       /* longest first */
    s->n_pair_chunks = (uint32_t)n; s->pair_cols = cursor;
    free(dir);
    return 0;
}

int osw_shard_build_ex(const uint8_t *residues, const uint64_t *offsets, uint64_t n_seqs,
                       uint32_t shard, uint32_t n_shards, uint32_t chunk_cols, int with_pair,
                       osw_alloc_fn alloc, void *alloc_user, osw_shard *out) {
    if (!out || !n_shards || shard >= n_shards || (n_seqs && (!residues || !offsets))) return -1;
    if (!chunk_cols) chunk_cols = OSW_CHUNK_COLS_DEFAULT;
    memset(out, 0, sizeof *out);
    tally_t t; memset(&t, 0, sizeof t);
    t.shard = shard; t.n_shards = n_shards; t.off = offsets;
    walk_chunks(offsets, n_seqs, chunk_cols, tally_cb, &t);
    osw_chunk *dir = (osw_chunk *)malloc((t.chunks ? t.chunks : 1) * sizeof(osw_chunk));
    uint64_t *first_seq = (uint64_t *)malloc((t.chunks ? t.chunks : 1) * sizeof(uint64_t));
    if (!dir || !first_seq) { free(dir); free(first_seq); return -1; }
    layout_t l; memset(&l, 0, sizeof l);
    l.shard = shard; l.n_shards = n_shards; l.off = offsets; l.dir = dir; l.first_seq = first_seq;
    walk_chunks(offsets, n_seqs, chunk_cols, layout_cb, &l);
    out->n_seqs = l.seq_cursor; out->n_residues = l.cols; out->stream_bytes = l.byte_cursor;
    out->n_chunks = (uint32_t)l.n; out->max_len = l.max_len;
    out->external_streams = alloc != NULL;
    if (alloc) out->stream = (uint8_t *)alloc(out->stream_bytes ? out->stream_bytes : 1, alloc_user);
    else out->stream = (uint8_t *)malloc(out->stream_bytes ? out->stream_bytes : 1);
    out->chunks  = (osw_chunk *)malloc((l.n ? l.n : 1) * sizeof(osw_chunk));
    out->canon   = (uint32_t *)malloc((out->n_seqs ? out->n_seqs : 1) * sizeof(uint32_t));
    out->seq_off = (uint64_t *)malloc((out->n_seqs ? out->n_seqs : 1) * sizeof(uint64_t));
    out->seq_len = (uint32_t *)malloc((out->n_seqs ? out->n_seqs : 1) * sizeof(uint32_t));
    if (!out->stream || !out->chunks || !out->canon || !out->seq_off || !out->seq_len) {
        free(dir); free(first_seq);
        osw_shard_free(out);
        return -1;
    }
    int bad = 0;
    const long long n_chunks = (long long)l.n;
#pragma omp parallel for schedule(dynamic, 64) reduction(| : bad)
    for (long long k = 0; k < n_chunks; ++k) bad |= fill_chunk(residues, offsets, &dir[k], first_seq[k], out);
    /* the directory is stored in reverse so that index 0 is the longest-sequence chunk */
    for (uint64_t k = 0; k < l.n; ++k) out->chunks[l.n - 1 - k] = dir[k];
    free(dir); free(first_seq);
    if (bad) { osw_shard_free(out); return -2; }
    if (build_pair_directory(out, chunk_cols) != 0) { osw_shard_free(out); return -1; }
    if (with_pair) {
        const size_t bytes = out->pair_cols ? 2 * out->pair_cols : 1;
        out->pair_stream = (uint8_t *)(alloc ? alloc(bytes, alloc_user) : malloc(bytes));
        if (!out->pair_stream) { osw_shard_free(out); return -1; }
        osw_shard_fill_pair(out, out->stream, out->pair_stream);
    }
    return 0;
}

int osw_shard_build(const uint8_t *residues, const uint64_t *offsets, uint64_t n_seqs,
                    uint32_t shard, uint32_t n_shards, uint32_t chunk_cols, osw_shard *out) {
    return osw_shard_build_ex(residues, offsets, n_seqs, shard, n_shards, chunk_cols, 1, NULL, NULL, out);
}

/* Pair stream from the plain one (same residues, flags stripped): sequences 2p and 2p+1 of a pair
 * chunk zipped column by column.  Also used when the pair stream is first needed after the
 * caller's database arrays are gone. */
void osw_shard_fill_pair(const osw_shard *s, const uint8_t *stream, uint8_t *pair) {
    const long long n_chunks = (long long)s->n_pair_chunks;
#pragma omp parallel for schedule(dynamic, 64)
    for (long long c = 0; c < n_chunks; ++c) {
        const osw_chunk *ck = &s->pair_chunks[c];
        uint8_t *q = pair + 2 * ck->pair_off;
        for (uint32_t k = 0; k < ck->n_seqs; k += 2) {
            const uint64_t la = s->seq_len[ck->seq0 + k];
            const int has_b = k + 1 < ck->n_seqs;
            const uint64_t lb = has_b ? s->seq_len[ck->seq0 + k + 1] : 0;
            const uint8_t *a = stream + s->seq_off[ck->seq0 + k];
            const uint8_t *b = has_b ? stream + s->seq_off[ck->seq0 + k + 1] : NULL;
            const uint64_t n = la > lb ? la : lb;
            for (uint64_t j = 0; j < n; ++j) {
                q[2 * j] = (uint8_t)(j < la ? (a[j] & OSW_COL_CODE) : OSW_COL_PADBYTE);
                q[2 * j + 1] = (uint8_t)(j < lb ? (b[j] & OSW_COL_CODE) : OSW_COL_PADBYTE);
            }
            if (n) { q[0] |= OSW_COL_FIRST; q[2 * (n - 1)] |= OSW_COL_LAST; }
            q += 2 * n;
        }
        const uint64_t pc_padded = ((uint64_t)ck->n_pair_cols + OSW_PAIR_ALIGN - 1) / OSW_PAIR_ALIGN * OSW_PAIR_ALIGN;
        memset(q, OSW_COL_PADBYTE, 2 * (pc_padded - ck->n_pair_cols));
    }
}

/* ---- X.osw ------------------------------------------------------------------------------------ */
static uint64_t fnv1a(uint64_t h, const void *data, size_t n) {
    const uint8_t *p = (const uint8_t *)data;
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}

int osw_dbfile_write(const char *path, const uint8_t *residues, const uint64_t *offsets, uint64_t n_seqs, uint32_t chunk_cols) {
    if (!chunk_cols) chunk_cols = OSW_CHUNK_COLS_DEFAULT;
    osw_shard s;
    const int rc = osw_shard_build_ex(residues, offsets, n_seqs, 0, 1, chunk_cols, 0, NULL, NULL, &s);      /* all chunks = shard 0 of 1 */
    if (rc) return rc == -2 ? -2 : -3;
    osw_dbfile_header h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, OSW_DBFILE_MAGIC, 8);
    h.version = OSW_DBFILE_VERSION; h.chunk_cols = chunk_cols;
    h.n_seqs = s.n_seqs; h.n_residues = s.n_residues; h.n_chunks = s.n_chunks; h.stream_bytes = s.stream_bytes;
    h.max_len = s.max_len; h.chunk_align = OSW_CHUNK_ALIGN;
    h.off_chunks = sizeof h;
    h.off_lengths = h.off_chunks + h.n_chunks * sizeof(osw_chunk);
    h.off_stream = (h.off_lengths + h.n_seqs * sizeof(uint32_t) + 4095) / 4096 * 4096;      /* page aligned: the stream is copied from the mapping */
    h.file_bytes = h.off_stream + h.stream_bytes;
    h.checksum = fnv1a(fnv1a(0xcbf29ce484222325ull, s.chunks, s.n_chunks * sizeof(osw_chunk)), s.seq_len, s.n_seqs * sizeof(uint32_t));
    int ok = 0;
    FILE *f = fopen(path, "wb");
    if (f) {
        static const uint8_t zeros[4096];
        const size_t gap = (size_t)(h.off_stream - (h.off_lengths + h.n_seqs * sizeof(uint32_t)));
        ok = fwrite(&h, sizeof h, 1, f) == 1 &&
             (s.n_chunks == 0 || fwrite(s.chunks, sizeof(osw_chunk), s.n_chunks, f) == s.n_chunks) &&
             (s.n_seqs == 0 || fwrite(s.seq_len, sizeof(uint32_t), s.n_seqs, f) == s.n_seqs) &&
             (gap == 0 || fwrite(zeros, 1, gap, f) == gap) &&
             (s.stream_bytes == 0 || fwrite(s.stream, 1, s.stream_bytes, f) == s.stream_bytes);
        ok = (fclose(f) == 0) && ok;
    }
    osw_shard_free(&s);
    return ok ? 0 : -1;
}

int osw_dbfile_open(const char *path, osw_dbfile *f) {
    memset(f, 0, sizeof *f);
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return -1;
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return -1; }
    if ((size_t)st.st_size < sizeof(osw_dbfile_header)) { close(fd); return -2; }
    void *map = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) return -1;
    memcpy(&f->h, map, sizeof f->h);
    const osw_dbfile_header *h = &f->h;
    int rc = 0;
    if (memcmp(h->magic, OSW_DBFILE_MAGIC, 8) != 0 || h->version != OSW_DBFILE_VERSION || h->chunk_align != OSW_CHUNK_ALIGN) rc = -2;
    else if (h->file_bytes > (uint64_t)st.st_size || h->off_chunks != sizeof *h ||
             h->off_lengths != h->off_chunks + h->n_chunks * sizeof(osw_chunk) ||
             h->off_stream < h->off_lengths + h->n_seqs * sizeof(uint32_t) || h->file_bytes != h->off_stream + h->stream_bytes ||
             h->n_seqs > 0xffffffffull || h->n_chunks > 0xffffffffull || h->stream_bytes % OSW_CHUNK_ALIGN) rc = -3;
    if (rc == 0) {
        f->chunks = (const osw_chunk *)((const uint8_t *)map + h->off_chunks);
        f->lengths = (const uint32_t *)((const uint8_t *)map + h->off_lengths);
        f->stream = (const uint8_t *)map + h->off_stream;
        if (fnv1a(fnv1a(0xcbf29ce484222325ull, f->chunks, h->n_chunks * sizeof(osw_chunk)), f->lengths, h->n_seqs * sizeof(uint32_t)) != h->checksum) rc = -3;
        /* every chunk lies inside the stream and inside the sequence list */
        for (uint64_t k = 0; rc == 0 && k < h->n_chunks; ++k) {
            const osw_chunk *ck = &f->chunks[k];
            const uint64_t padded = ((uint64_t)ck->n_cols + OSW_CHUNK_ALIGN - 1) / OSW_CHUNK_ALIGN * OSW_CHUNK_ALIGN;
            if (ck->stream_off % OSW_CHUNK_ALIGN || ck->stream_off + padded > h->stream_bytes || (uint64_t)ck->canon0 + ck->n_seqs > h->n_seqs || ck->seq0 != ck->canon0) rc = -3;
        }
    }
    if (rc) { munmap(map, (size_t)st.st_size); memset(f, 0, sizeof *f); return rc; }
    f->map = map; f->map_size = (size_t)st.st_size;
    return 0;
}

void osw_dbfile_close(osw_dbfile *f) {
    if (f && f->map) munmap(f->map, f->map_size);
    if (f) memset(f, 0, sizeof *f);
}

int osw_shard_from_file(const osw_dbfile *f, uint32_t shard, uint32_t n_shards, osw_alloc_fn alloc, void *alloc_user, osw_shard *out) {
    if (!f || !f->map || !out || !n_shards || shard >= n_shards) return -1;
    memset(out, 0, sizeof *out);
    const uint64_t nc = f->h.n_chunks;
    /* the file's directory is in descending order; the walk (= dealing) order is ascending: walk index c = nc-1-k */
    uint64_t mine = 0, seqs = 0, bytes = 0, cols = 0;
    for (uint64_t c = shard; c < nc; c += n_shards) {
        const osw_chunk *ck = &f->chunks[nc - 1 - c];
        ++mine; seqs += ck->n_seqs; cols += ck->n_cols;
        bytes += ((uint64_t)ck->n_cols + OSW_CHUNK_ALIGN - 1) / OSW_CHUNK_ALIGN * OSW_CHUNK_ALIGN;
    }
    out->n_seqs = seqs; out->n_residues = cols; out->stream_bytes = bytes; out->n_chunks = (uint32_t)mine;
    out->external_streams = alloc != NULL;
    out->stream = (uint8_t *)(alloc ? alloc(bytes ? bytes : 1, alloc_user) : malloc(bytes ? bytes : 1));
    out->chunks  = (osw_chunk *)malloc((mine ? mine : 1) * sizeof(osw_chunk));
    out->canon   = (uint32_t *)malloc((seqs ? seqs : 1) * sizeof(uint32_t));
    out->seq_off = (uint64_t *)malloc((seqs ? seqs : 1) * sizeof(uint64_t));
    out->seq_len = (uint32_t *)malloc((seqs ? seqs : 1) * sizeof(uint32_t));
    if (!out->stream || !out->chunks || !out->canon || !out->seq_off || !out->seq_len) { osw_shard_free(out); return -1; }
    /* layout (serial, cheap), then the copies (parallel) */
    uint64_t byte_cursor = 0, seq_cursor = 0, n = 0;
    for (uint64_t c = shard; c < nc; c += n_shards, ++n) {
        const osw_chunk *src = &f->chunks[nc - 1 - c];
        osw_chunk *ck = &out->chunks[mine - 1 - n];                 /* stored descending */
        *ck = *src;
        ck->stream_off = byte_cursor; ck->seq0 = (uint32_t)seq_cursor; ck->pair_off = 0; ck->n_pair_cols = 0; ck->reserved = 0;
        byte_cursor += ((uint64_t)src->n_cols + OSW_CHUNK_ALIGN - 1) / OSW_CHUNK_ALIGN * OSW_CHUNK_ALIGN;
        seq_cursor += src->n_seqs;
        if (src->n_seqs) {
            const uint32_t last = f->lengths[(uint64_t)src->canon0 + src->n_seqs - 1];      /* longest: order is ascending */
            if (last > out->max_len) out->max_len = last;
        }
    }
    const long long n_mine = (long long)mine;
#pragma omp parallel for schedule(dynamic, 64)
    for (long long k = 0; k < n_mine; ++k) {
        const osw_chunk *ck = &out->chunks[k];
        const uint64_t c = (uint64_t)shard + (uint64_t)(n_mine - 1 - k) * n_shards;
        const osw_chunk *src = &f->chunks[nc - 1 - c];
        const uint64_t padded = ((uint64_t)ck->n_cols + OSW_CHUNK_ALIGN - 1) / OSW_CHUNK_ALIGN * OSW_CHUNK_ALIGN;
        memcpy(out->stream + ck->stream_off, f->stream + src->stream_off, padded);
        uint64_t off = ck->stream_off;
        for (uint32_t i = 0; i < ck->n_seqs; ++i) {
            const uint32_t canon = ck->canon0 + i, len = f->lengths[canon];
            out->canon[ck->seq0 + i] = canon; out->seq_len[ck->seq0 + i] = len; out->seq_off[ck->seq0 + i] = off;
            off += len;
        }
    }
    if (build_pair_directory(out, f->h.chunk_cols) != 0) { osw_shard_free(out); return -1; }
    return 0;
}

void osw_shard_free(osw_shard *s) {
    if (!s) return;
    if (!s->external_streams) { free(s->stream); free(s->pair_stream); }   /* external ones belong to the allocator's owner */
    free(s->chunks); free(s->pair_chunks); free(s->canon); free(s->seq_off); free(s->seq_len);
    memset(s, 0, sizeof *s);
}
