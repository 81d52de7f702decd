/* search.c - `-O search` and `-O info`: the GPU counterpart of hybrid_search_avx2()
 * (reference HybridSearch.c:4-1277) behind the same main() dispatch (main.c:46-62).
 * Loads queries (sorted by length like sequences.c:342) and the preprocessed database, hands
 * them to the CUDA library through the C ABI, prints the reference's report. */
#include "oswald_host.h"
#include "submat.h"
#include "oswald_cuda.h"
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <sys/time.h>

static double wall_s(void) { struct timeval tv; gettimeofday(&tv, NULL); return tv.tv_sec + tv.tv_usec * 1e-6; }
#define PHASE(name) do { if (trace) { double t_ = wall_s(); fprintf(stderr, "oswald trace: %-28s %.3f s\n", name, t_ - t_phase); t_phase = t_; } } while (0)

int gpu_info(void) {
    int n = 0;
    if (osw_device_count(&n) != OSW_OK || n == 0) { printf("OSWALD: no CUDA device found (%s).\n", osw_last_error()); return 1; }
    char text[1024];
    for (int i = 0; i < n; ++i)
        if (osw_device_info(i, text, sizeof text) == OSW_OK) fputs(text, stdout);
    return 0;
}

int gpu_search(const osw_options *opt) {
    const int trace = getenv("OSW_TRACE") != NULL;
    double t_phase = wall_s();
    printf("\nOSWALD v%s \n\n", OSWALD_VERSION);
    printf("Database file:\t\t\t%s\n", opt->sequences_filename);

    /* queries: FASTA order -> stable ascending length (reference sequences.c:342) */
    osw_fasta qf;
    int rc = osw_fasta_read(opt->queries_filename, &qf);
    if (rc == -1) { printf("OSWALD: An error occurred while opening input sequence file.\n"); return 2; }
    if (rc) { printf("OSWALD: An error occurred while allocating memory.\n"); return 1; }
    if (qf.n == 0) { printf("OSWALD: the query file holds no sequence.\n"); return 2; }
    uint64_t *qperm = osw_length_order(&qf);
    int nq = (int)qf.n;
    uint8_t *a = (uint8_t *)malloc(qf.n_residues ? qf.n_residues : 1);
    uint32_t *a_disp = (uint32_t *)malloc(((size_t)nq + 1) * sizeof(uint32_t));
    uint64_t Q = 0;
    for (int k = 0; k < nq; ++k) {
        uint64_t i = qperm[k], len = qf.offsets[i + 1] - qf.offsets[i];
        if (len > OSW_MAX_QUERY_LEN) { printf("OSWALD: query %d is longer than %d residues.\n", k + 1, OSW_MAX_QUERY_LEN); return 2; }
        a_disp[k] = (uint32_t)Q;
        memcpy(a + Q, qf.residues + qf.offsets[i], len);
        Q += len;
    }
    a_disp[nq] = (uint32_t)Q;

    PHASE("load queries");
    osw_database db;
    if ((rc = load_database(opt->sequences_filename, &db)) != 0) return rc;
    PHASE("map database");
    printf("Database size:\t\t\t%ld sequences (%ld residues) \n", (long)db.n_seqs, (long)db.n_residues);
    printf("Longest database sequence: \t%d residues\n", (int)db.max_len);
    printf("Substitution matrix:\t\t%s\n", opt->submat_name);
    printf("Gap open penalty:\t\t%d\n", opt->open_gap);
    printf("Gap extend penalty:\t\t%d\n", opt->extend_gap);
    printf("Query filename:\t\t\t%s\n", opt->queries_filename);

    int8_t matrix[24 * 32];
    osw_matrix_by_name(opt->submat_arg, matrix);
    unsigned long top = db.n_seqs < opt->top ? db.n_seqs : opt->top;     /* HybridSearch.c:64 */

    osw_ctx *ctx = NULL;
    if ((rc = osw_init((int)opt->num_devices, NULL, &ctx)) != OSW_OK) {
        printf("OSWALD: cannot initialise %u GPU(s): %s (%s).\n", opt->num_devices, osw_strerror(rc), osw_last_error());
        return 1;
    }
    PHASE("osw_init");
    if (opt->max_chunk_size_given) osw_set_device_window(ctx, opt->max_chunk_size);      /* -k, arguments.c:113-117 */
    if ((rc = osw_db_load(ctx, db.residues, db.offsets, db.n_seqs, 0, 1, 0)) != OSW_OK) {
        printf("OSWALD: cannot load the database on the GPU(s): %s (%s).\n", osw_strerror(rc), osw_last_error());
        osw_free(ctx);
        return 1;
    }
    PHASE("osw_db_load (layout + H2D)");
    osw_hit *hits = (osw_hit *)malloc(((size_t)nq * (top ? top : 1)) * sizeof(osw_hit));
    uint32_t *n_hits = (uint32_t *)calloc((size_t)nq, sizeof(uint32_t));
    int32_t *all = NULL;
    if (opt->dump_scores) all = (int32_t *)calloc((size_t)nq * (db.n_seqs ? db.n_seqs : 1), sizeof(int32_t));
    osw_timing tm;
    time_t current_time = time(NULL);
    rc = osw_search(ctx, a, a_disp, nq, matrix, opt->open_gap, opt->extend_gap, (int)top, hits, n_hits, all, &tm);
    if (rc != OSW_OK) {
        printf("OSWALD: search failed: %s (%s).\n", osw_strerror(rc), osw_last_error());
        osw_free(ctx);
        return 1;
    }
    PHASE("osw_search");
    if (opt->dump_scores) {
        FILE *f = fopen(opt->dump_scores, "wb");
        if (!f) { printf("OSWALD: cannot write %s.\n", opt->dump_scores); return 2; }
        fwrite(all, sizeof(int32_t), (size_t)nq * db.n_seqs, f);
        fclose(f);
    }
    /* titles of the printed hits only */
    size_t n_print = (size_t)nq * top;
    uint32_t *idx = (uint32_t *)malloc((n_print ? n_print : 1) * sizeof(uint32_t));
    char **titles = (char **)malloc((n_print ? n_print : 1) * sizeof(char *));
    for (size_t k = 0; k < n_print; ++k) idx[k] = hits[k].index;
    if ((rc = load_database_headers(opt->sequences_filename, idx, n_print, titles)) != 0) return rc;

    PHASE("load titles of the hits");
    for (int i = 0; i < nq; ++i) {                     /* report: HybridSearch.c:1213-1224 */
        printf("\nQuery no.\t\t\t%d\n", i + 1);
        printf("Query description: \t\t%s\n", qf.titles[qperm[i]]);
        printf("Query length:\t\t\t%d residues\n", (int)(a_disp[i + 1] - a_disp[i]));
        printf("\nScore\tSequence description\n");
        for (unsigned long j = 0; j < top; ++j)
            printf("%d\t%s", hits[(size_t)i * top + j].score, titles[(size_t)i * top + j] + 1);
    }
    double secs = tm.wall_ms / 1e3;
    printf("\nSearch date:\t\t\t%s", ctime(&current_time));
    printf("Search time:\t\t\t%lf seconds\n", secs);
    printf("Search speed:\t\t\t%.2lf GCUPS\n", secs > 0 ? (double)Q * (double)db.n_residues / (secs * 1e9) : 0.0);
    printf("GPU time:\t\t\t%lf seconds\n", tm.device_ms / 1e3);
    printf("GPU speed:\t\t\t%.2lf GCUPS\n", tm.device_ms > 0 ? (double)Q * (double)db.n_residues / (tm.device_ms * 1e6) : 0.0);
    printf("Number of GPUs:\t\t\t%u\n", opt->num_devices);
    printf("Kernel launches:\t\t%lu\n", (unsigned long)tm.launches);
    printf("Pairs re-scored at 32 bit:\t%lu\n", (unsigned long)tm.rescored_pairs);
    if (opt->max_chunk_size_given) printf("Max. chunk size on GPU:\t\t%lu bytes (database streamed through two windows)\n", opt->max_chunk_size);
    else printf("Max. chunk size on GPU:\t\tdatabase resident\n");

    for (size_t k = 0; k < n_print; ++k) free(titles[k]);
    free(titles); free(idx); free(hits); free(n_hits); free(all); free(a); free(a_disp); free(qperm);
    osw_free(ctx);
    free_database(&db);
    osw_fasta_free(&qf);
    return 0;
}
