/* fasta.c - single-pass FASTA reader and the canonical (stable, ascending length) order.
 *
 * The reference reads the file three times with fgets and one malloc per sequence
 * (sequences.c:28-119); here the file is mapped once and scanned once.  Semantics kept:
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

int osw_fasta_read(const char *path, osw_fasta *out) {
    memset(out, 0, sizeof *out);
    int fd = open(path, O_RDONLY);
    if (fd < 0) return -1;
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return -1; }
    size_t size = (size_t)st.st_size;
    const char *buf = size ? (const char *)mmap(NULL, size, PROT_READ, MAP_PRIVATE, fd, 0) : "";
    close(fd);
    if (size && buf == MAP_FAILED) return -1;
    /* pass A: count records and title bytes (a '>' at a line start) */
    uint64_t n = 0, title_bytes = 0;
    for (size_t p = 0; p < size;) {
        const char *nl = (const char *)memchr(buf + p, '\n', size - p);
        size_t e = nl ? (size_t)(nl - buf) : size;
        if (buf[p] == '>') { ++n; title_bytes += e - p; }
        p = e + 1;
    }
    out->titles = (char **)malloc((n ? n : 1) * sizeof(char *));
    out->offsets = (uint64_t *)malloc((n + 1) * sizeof(uint64_t));
    out->residues = (uint8_t *)malloc(size ? size : 1);          /* upper bound */
    out->title_pool = (char *)malloc(title_bytes + n + 1);
    if (!out->titles || !out->offsets || !out->residues || !out->title_pool) {
        if (size) munmap((void *)buf, size);
        osw_fasta_free(out);
        return -2;
    }
    /* pass B: fill */
    uint64_t rec = 0, nres = 0;
    char *tp = out->title_pool;
    for (size_t p = 0; p < size;) {
        const char *nl = (const char *)memchr(buf + p, '\n', size - p);
        size_t e = nl ? (size_t)(nl - buf) : size;
        size_t len = e - p;
        if (len && buf[e - 1] == '\r') --len;
        if (buf[p] == '>') {
            out->offsets[rec] = nres;
            out->titles[rec] = tp;
            memcpy(tp, buf + p + 1, len ? len - 1 : 0);
            tp += len ? len - 1 : 0;
            *tp++ = 0;
            ++rec;
        } else if (rec) {
            for (size_t k = 0; k < len; ++k) out->residues[nres + k] = osw_encode_letter((unsigned char)buf[p + k]);
            nres += len;
        }
        p = e + 1;
    }
    out->offsets[rec] = nres;
    out->n = rec; out->n_residues = nres;
    if (size) munmap((void *)buf, size);
    return 0;
}

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
