/* preprocess.c - `-O preprocess`: FASTA -> X.info / X.seq / X.desc, byte-compatible with the
 * reference (sequences.c:4-220):
 *   X.info  ASCII "N D max_title_length"            (:187; title length = line incl. '>' and
 *                                                     '\n', plus one, :35)
 *   X.seq   N little-endian u16 lengths (ascending), then D residue codes in that order (:202-205)
 *   X.desc  N lines ">title", in that order          (:137-138)
 * and the loader of that triple for `-O search`. */
#include "oswald_host.h"
#include "oswald_cuda.h"
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

int preprocess_db(const char *input_filename, const char *out_filename, int n_procs) {
    osw_fasta fa;
    int rc = osw_fasta_read_mt(input_filename, &fa, n_procs);
    if (rc == -1) { printf("OSWALD: An error occurred while opening input sequence file.\n"); return 2; }
    if (rc) { printf("OSWALD: An error occurred while allocating memory for sequences.\n"); return 1; }
    for (uint64_t i = 0; i < fa.n; ++i)
        if (fa.offsets[i + 1] - fa.offsets[i] > 65535) {
            printf("OSWALD: sequence %lu is longer than 65535 residues (lengths are 16-bit in the database format).\n", (unsigned long)i + 1);
            osw_fasta_free(&fa);
            return 1;
        }
    uint64_t *perm = osw_length_order(&fa);
    /* the canonical database in memory: lengths, offsets, residues in sorted order */
    uint16_t *lens = (uint16_t *)malloc((fa.n ? fa.n : 1) * sizeof(uint16_t));
    uint64_t *coff = (uint64_t *)malloc((fa.n + 1) * sizeof(uint64_t));
    uint8_t *cres = (uint8_t *)malloc(fa.n_residues ? fa.n_residues : 1);
    if (!perm || !lens || !coff || !cres) {
        printf("OSWALD: An error occurred while allocating memory.\n");
        free(perm); free(lens); free(coff); free(cres); osw_fasta_free(&fa);
        return 1;
    }
    coff[0] = 0;
    for (uint64_t k = 0; k < fa.n; ++k) {
        lens[k] = (uint16_t)(fa.offsets[perm[k] + 1] - fa.offsets[perm[k]]);
        coff[k + 1] = coff[k] + lens[k];
    }
#pragma omp parallel for schedule(static) num_threads(n_procs > 0 ? n_procs : 1)
    for (long long k = 0; k < (long long)fa.n; ++k) memcpy(cres + coff[k], fa.residues + fa.offsets[perm[k]], lens[k]);
    int ret = 0;
    char filename[4096];
    snprintf(filename, sizeof filename, "%s.desc", out_filename);
    FILE *titles_file = fopen(filename, "w");
    if (!titles_file) { printf("OSWALD: An error occurred while opening sequence header file.\n"); ret = 2; goto done; }
    int max_title_length = 0;
    for (uint64_t k = 0; k < fa.n; ++k) {
        const char *t = fa.titles[perm[k]];
        int tl = (int)strlen(t) + 3;                     /* '>' + title + '\n', plus one */
        if (tl > max_title_length) max_title_length = tl;
        fputc('>', titles_file); fputs(t, titles_file); fputc('\n', titles_file);
    }
    fclose(titles_file);
    snprintf(filename, sizeof filename, "%s.info", out_filename);
    FILE *info_file = fopen(filename, "w");
    if (!info_file) { printf("OSWALD: An error occurred while opening info file.\n"); ret = 2; goto done; }
    fprintf(info_file, "%ld %ld %d", (long)fa.n, (long)fa.n_residues, max_title_length);
    fclose(info_file);
    snprintf(filename, sizeof filename, "%s.seq", out_filename);
    FILE *bin_file = fopen(filename, "wb");
    if (!bin_file) { printf("OSWALD: An error occurred while opening sequence file.\n"); ret = 2; goto done; }
    fwrite(lens, sizeof(uint16_t), fa.n, bin_file);
    fwrite(cres, 1, fa.n_residues, bin_file);
    fclose(bin_file);
    /* the same database in its device layout (chunk column streams, directories): X.osw.  A search
     * uses it when it is there and falls back to X.seq otherwise. */
    snprintf(filename, sizeof filename, "%s.osw", out_filename);
    if ((rc = osw_db_write_file(filename, cres, coff, fa.n, 0)) != OSW_OK) {
        printf("OSWALD: cannot write %s: %s (%s).\n", filename, osw_strerror(rc), osw_last_error());
        ret = 2;
    }
done:
    free(lens); free(coff); free(cres); free(perm);
    osw_fasta_free(&fa);
    return ret;
}

int load_database(const char *prefix, osw_database *db) {
    memset(db, 0, sizeof *db);
    char filename[4096];
    snprintf(filename, sizeof filename, "%s.info", prefix);
    FILE *info = fopen(filename, "r");
    if (!info) { printf("OSWALD: An error occurred while opening info file.\n"); return 2; }
    long n = 0, d = 0; int mt = 0;
    if (fscanf(info, "%ld %ld %d", &n, &d, &mt) != 3 || n < 0 || d < 0) { fclose(info); printf("OSWALD: info file is malformed.\n"); return 2; }
    fclose(info);
    snprintf(filename, sizeof filename, "%s.seq", prefix);
    int fd = open(filename, O_RDONLY);
    if (fd < 0) { printf("OSWALD: An error occurred while opening sequence file.\n"); return 2; }
    struct stat st;
    fstat(fd, &st);
    size_t need = (size_t)n * 2 + (size_t)d;
    if ((size_t)st.st_size < need) { close(fd); printf("OSWALD: sequence file is shorter than the info file says.\n"); return 2; }
    void *map = need ? mmap(NULL, need, PROT_READ, MAP_PRIVATE, fd, 0) : NULL;
    close(fd);
    if (need && map == MAP_FAILED) { printf("OSWALD: An error occurred while mapping sequence file.\n"); return 2; }
    db->map = map; db->map_size = need;
    db->n_seqs = (uint64_t)n; db->n_residues = (uint64_t)d; db->max_title_length = mt;
    db->offsets = (uint64_t *)malloc(((size_t)n + 1) * sizeof(uint64_t));
    if (!db->offsets) { printf("OSWALD: An error occurred while allocating memory.\n"); return 1; }
    const uint16_t *lens = (const uint16_t *)map;
    uint64_t acc = 0;
    for (long i = 0; i < n; ++i) { db->offsets[i] = acc; acc += lens[i]; if (lens[i] > db->max_len) db->max_len = lens[i]; }
    db->offsets[n] = acc;
    if (acc != (uint64_t)d) { printf("OSWALD: sequence lengths do not add up to the residue count.\n"); return 2; }
    db->residues = (uint8_t *)map + (size_t)n * 2;
    return 0;
}

void free_database(osw_database *db) {
    if (db->map) munmap(db->map, db->map_size);
    free(db->offsets);
    memset(db, 0, sizeof *db);
}

static const uint32_t *g_sort_indices;          /* (qsort has no context argument; the tool is single-threaded here) */
static int cmp_request(const void *a, const void *b) {
    const size_t x = *(const size_t *)a, y = *(const size_t *)b;
    const uint32_t ix = g_sort_indices[x], iy = g_sort_indices[y];
    return ix != iy ? (ix < iy ? -1 : 1) : (x < y ? -1 : x > y);
}

/* The reference loads all N titles (sequences.c:1096-1127); only the printed ones are needed:
 * one sequential scan of X.desc picks the requested lines. */
int load_database_headers(const char *prefix, const uint32_t *indices, size_t n, char **result) {
    char filename[4096];
    snprintf(filename, sizeof filename, "%s.desc", prefix);
    FILE *f = fopen(filename, "r");
    if (!f) { printf("OSWALD: An error occurred while opening sequence description file.\n"); return 3; }
    /* order the requests by index (n = nq * top: up to a few hundred thousand with many queries) */
    size_t *order = (size_t *)malloc((n ? n : 1) * sizeof(size_t));
    if (!order) { fclose(f); printf("OSWALD: An error occurred while allocating memory.\n"); return 1; }
    for (size_t i = 0; i < n; ++i) order[i] = i;
    g_sort_indices = indices;
    qsort(order, n, sizeof(size_t), cmp_request);
    for (size_t i = 0; i < n; ++i) result[i] = NULL;
    char *line = NULL; size_t cap = 0; ssize_t len;
    uint64_t lineno = 0; size_t k = 0;
    while (k < n && (len = getline(&line, &cap, f)) >= 0) {
        while (k < n && indices[order[k]] == lineno) {
            result[order[k]] = strdup(line);
            ++k;
        }
        ++lineno;
    }
    free(line); free(order);
    fclose(f);
    for (size_t i = 0; i < n; ++i) if (!result[i]) result[i] = strdup(">?\n");
    return 0;
}
