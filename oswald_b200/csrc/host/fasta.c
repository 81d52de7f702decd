/* fasta.c - FASTA reader and the canonical (stable, ascending length) order.
 *
 * The reference reads the file three times with fgets and one malloc per sequence
 * (sequences.c:28-119); here the file is mapped once and scanned twice (count, fill) by -c
 * threads, each on a block of whole lines.  Semantics kept:
 * a record starts at a line beginning with '>', its title is that line, its residues are
 * all following lines up to the next '>' with the line ends removed; letters are encoded
 * like sequences.c:163-175; the order is a stable ascending sort by length
 * (sequences.c:1130-1225), done here as a counting sort on the 16-bit length. */
#include "oswald_host.h"
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

/* One block of the file (a whole number of lines) scanned by one thread. */
typedef struct { size_t begin, end; uint64_t recs, title_bytes, residues; } fasta_block;

static void block_count(const char *buf, fasta_block *b) {
    uint64_t recs = 0, tb = 0, res = 0;
    for (size_t p = b->begin; p < b->end;) {
        const char *nl = (const char *)memchr(buf + p, '\n', b->end - p);
        const size_t e = nl ? (size_t)(nl - buf) : b->end;
        size_t len = e - p;
        if (len && buf[e - 1] == '\r') --len;
        if (buf[p] == '>') { ++recs; tb += len; }             /* title = the line without '>' (+ NUL) */
        else res += len;
        p = e + 1;
    }
    b->recs = recs; b->title_bytes = tb; b->residues = res;
}

int osw_fasta_read_mt(const char *path, osw_fasta *out, int n_threads) {
    memset(out, 0, sizeof *out);
    int fd = open(path, O_RDONLY);
    if (fd < 0) return -1;
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return -1; }
    size_t size = (size_t)st.st_size;
    const char *buf = size ? (const char *)mmap(NULL, size, PROT_READ, MAP_PRIVATE, fd, 0) : "";
    close(fd);
    if (size && buf == MAP_FAILED) return -1;
    /* whatever precedes the first record is not a sequence (the reference would misread it) */
    size_t start = 0;
    while (start < size && buf[start] != '>') {
        const char *nl = (const char *)memchr(buf + start, '\n', size - start);
        start = nl ? (size_t)(nl - buf) + 1 : size;
    }
    /* blocks of whole lines, one per thread (at least 1 MiB each) */
    if (n_threads < 1) n_threads = 1;
    int nb = (int)((size - start) / (1u << 20)) + 1;
    if (nb > n_threads) nb = n_threads;
    fasta_block *blk = (fasta_block *)calloc((size_t)nb + 1, sizeof *blk);
    if (!blk) { if (size) munmap((void *)buf, size); return -2; }
    for (int k = 0; k <= nb; ++k) {
        size_t cut = k == nb ? size : start + (size - start) / (size_t)nb * (size_t)k;
        if (k && k < nb) {                          /* move the cut to the next line start */
            const char *nl = (const char *)memchr(buf + cut, '\n', size - cut);
            cut = nl ? (size_t)(nl - buf) + 1 : size;
        }
        if (k < nb) blk[k].begin = cut;
        if (k) blk[k - 1].end = cut;
    }
    /* pass A: count records, title bytes and residues per block */
#pragma omp parallel for schedule(static, 1) num_threads(nb)
    for (int k = 0; k < nb; ++k) block_count(buf, &blk[k]);
    uint64_t n = 0, title_bytes = 0, n_res = 0;
    for (int k = 0; k < nb; ++k) { n += blk[k].recs; title_bytes += blk[k].title_bytes; n_res += blk[k].residues; }
    out->titles = (char **)malloc((n ? n : 1) * sizeof(char *));
    out->offsets = (uint64_t *)malloc((n + 1) * sizeof(uint64_t));
    out->residues = (uint8_t *)malloc(n_res ? n_res : 1);
    out->title_pool = (char *)malloc(title_bytes + 1);
    if (!out->titles || !out->offsets || !out->residues || !out->title_pool) {
        if (size) munmap((void *)buf, size);
        free(blk);
        osw_fasta_free(out);
        return -2;
    }
    /* pass B: fill; residues keep their file order, so a block's residues go to one contiguous stretch
     * (a record that spans blocks continues there) */
    uint64_t *rec0 = (uint64_t *)malloc(((size_t)nb + 1) * 3 * sizeof(uint64_t));
    if (!rec0) { if (size) munmap((void *)buf, size); free(blk); osw_fasta_free(out); return -2; }
    uint64_t *res0 = rec0 + nb + 1, *tit0 = res0 + nb + 1;
    rec0[0] = res0[0] = tit0[0] = 0;
    for (int k = 0; k < nb; ++k) { rec0[k + 1] = rec0[k] + blk[k].recs; res0[k + 1] = res0[k] + blk[k].residues; tit0[k + 1] = tit0[k] + blk[k].title_bytes; }
#pragma omp parallel for schedule(static, 1) num_threads(nb)
    for (int k = 0; k < nb; ++k) {
        uint64_t rec = rec0[k], nres = res0[k];
        char *tp = out->title_pool + tit0[k];
        for (size_t p = blk[k].begin; p < blk[k].end;) {
            const char *nl = (const char *)memchr(buf + p, '\n', blk[k].end - p);
            const size_t e = nl ? (size_t)(nl - buf) : blk[k].end;
            size_t len = e - p;
            if (len && buf[e - 1] == '\r') --len;
            if (buf[p] == '>') {
                out->offsets[rec] = nres;
                out->titles[rec] = tp;
                memcpy(tp, buf + p + 1, len - 1);
                tp += len - 1;
                *tp++ = 0;
                ++rec;
            } else {
                for (size_t j = 0; j < len; ++j) out->residues[nres + j] = osw_encode_letter((unsigned char)buf[p + j]);
                nres += len;
            }
            p = e + 1;
        }
    }
    out->offsets[n] = n_res;
    out->n = n; out->n_residues = n_res;
    free(rec0); free(blk);
    if (size) munmap((void *)buf, size);
    return 0;
}

int osw_fasta_read(const char *path, osw_fasta *out) { return osw_fasta_read_mt(path, out, 1); }

void osw_fasta_free(osw_fasta *f) {
    free(f->titles); free(f->offsets); free(f->residues); free(f->title_pool);
    memset(f, 0, sizeof *f);
}

uint64_t *osw_length_order(const osw_fasta *f) {
    /* counting sort on length: stable, O(n + 65536) */
    uint64_t *perm = (uint64_t *)malloc((f->n ? f->n : 1) * sizeof(uint64_t));
    uint64_t *start = (uint64_t *)calloc(65537 + 1, sizeof(uint64_t));
    if (!perm || !start) { free(perm); free(start); return NULL; }
    for (uint64_t i = 0; i < f->n; ++i) {
        uint64_t len = f->offsets[i + 1] - f->offsets[i];
        if (len > 65535) len = 65536;                 /* rejected by the caller */
        start[len + 1]++;
    }
    for (int l = 0; l < 65537; ++l) start[l + 1] += start[l];
    for (uint64_t i = 0; i < f->n; ++i) {
        uint64_t len = f->offsets[i + 1] - f->offsets[i];
        if (len > 65535) len = 65536;
        perm[start[len]++] = i;
    }
    free(start);
    return perm;
}
