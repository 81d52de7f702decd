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
    int ret = 0, rc;
    /* everything the function owns: released at `done` on every path */
    osw_fasta qf; memset(&qf, 0, sizeof qf);
    osw_database db; memset(&db, 0, sizeof db);
    osw_ctx *ctx = NULL;
    uint64_t *qperm = NULL; uint8_t *a = NULL; uint32_t *a_disp = NULL;
    osw_hit *hits = NULL; uint32_t *n_hits = NULL; int32_t *all = NULL;
    uint32_t *idx = NULL; char **titles = NULL; size_t n_print = 0;

    printf("\nOSWALD v%s \n\n", OSWALD_VERSION);
    printf("Database file:\t\t\t%s\n", opt->sequences_filename);

    /* queries: FASTA order -> stable ascending length (reference sequences.c:342) */
    rc = osw_fasta_read(opt->queries_filename, &qf);
    if (rc == -1) { printf("OSWALD: An error occurred while opening input sequence file.\n"); ret = 2; goto done; }
    if (rc) { printf("OSWALD: An error occurred while allocating memory.\n"); ret = 1; goto done; }
    if (qf.n == 0) { printf("OSWALD: the query file holds no sequence.\n"); ret = 2; goto done; }
    qperm = osw_length_order(&qf);
    const int nq = (int)qf.n;
    a = (uint8_t *)malloc(qf.n_residues ? qf.n_residues : 1);
    a_disp = (uint32_t *)malloc(((size_t)nq + 1) * sizeof(uint32_t));
    if (!qperm || !a || !a_disp) { printf("OSWALD: An error occurred while allocating memory.\n"); ret = 1; goto done; }
    uint64_t Q = 0;
    for (int k = 0; k < nq; ++k) {
        uint64_t i = qperm[k], len = qf.offsets[i + 1] - qf.offsets[i];
        if (len > OSW_MAX_QUERY_LEN) { printf("OSWALD: query %d is longer than %d residues.\n", k + 1, OSW_MAX_QUERY_LEN); ret = 2; goto done; }
        a_disp[k] = (uint32_t)Q;
        memcpy(a + Q, qf.residues + qf.offsets[i], len);
        Q += len;
    }
    a_disp[nq] = (uint32_t)Q;
    PHASE("load queries");

    /* the database: X.osw (device layout, written by -O preprocess) when it is there and matches
     * X.info, else the reference's X.seq, laid out now */
    uint64_t n_seqs = 0, n_residues = 0; uint32_t max_len = 0;
    char osw_name[4096];
    snprintf(osw_name, sizeof osw_name, "%s.osw", opt->sequences_filename);
    int use_osw = 0;
    {
        long ni = 0, di = 0; int mt = 0;
        char info_name[4096];
        snprintf(info_name, sizeof info_name, "%s.info", opt->sequences_filename);
        FILE *info = fopen(info_name, "r");
        uint64_t fn = 0, fd = 0; uint32_t fm = 0, ver = 0;
        if (info && fscanf(info, "%ld %ld %d", &ni, &di, &mt) == 3 && !getenv("OSW_NO_DBFILE") &&
            osw_db_file_info(osw_name, &fn, &fd, NULL, &fm, &ver) == OSW_OK && fn == (uint64_t)ni && fd == (uint64_t)di) {
            use_osw = 1; n_seqs = fn; n_residues = fd; max_len = fm;
        }
        if (info) fclose(info);
    }
    if (!use_osw) {
        if ((rc = load_database(opt->sequences_filename, &db)) != 0) { ret = rc; goto done; }
        n_seqs = db.n_seqs; n_residues = db.n_residues; max_len = db.max_len;
    }
    PHASE(use_osw ? "read X.osw header" : "map X.seq");
    printf("Database size:\t\t\t%ld sequences (%ld residues) \n", (long)n_seqs, (long)n_residues);
    printf("Longest database sequence: \t%d residues\n", (int)max_len);
    printf("Substitution matrix:\t\t%s\n", opt->submat_name);
    printf("Gap open penalty:\t\t%d\n", opt->open_gap);
    printf("Gap extend penalty:\t\t%d\n", opt->extend_gap);
    printf("Query filename:\t\t\t%s\n", opt->queries_filename);

    int8_t matrix[24 * 32];
    osw_matrix_by_name(opt->submat_arg, matrix);
    const unsigned long top = n_seqs < opt->top ? n_seqs : opt->top;     /* HybridSearch.c:64 */

    if ((rc = osw_init((int)opt->num_devices, NULL, &ctx)) != OSW_OK) {
        printf("OSWALD: cannot initialise %u GPU(s): %s (%s).\n", opt->num_devices, osw_strerror(rc), osw_last_error());
        ret = 1; goto done;
    }
    PHASE("osw_init");
    if (opt->max_chunk_size_given) osw_set_device_window(ctx, opt->max_chunk_size);      /* -k, arguments.c:113-117 */
    rc = use_osw ? osw_db_load_file(ctx, osw_name, 0, 1) : osw_db_load(ctx, db.residues, db.offsets, db.n_seqs, 0, 1, 0);
    if (rc != OSW_OK) {
        printf("OSWALD: cannot load the database on the GPU(s): %s (%s).\n", osw_strerror(rc), osw_last_error());
        ret = 1; goto done;
    }
    PHASE(use_osw ? "osw_db_load_file (copy + H2D)" : "osw_db_load (layout + H2D)");
    hits = (osw_hit *)malloc(((size_t)nq * (top ? top : 1)) * sizeof(osw_hit));
    n_hits = (uint32_t *)calloc((size_t)nq, sizeof(uint32_t));
    if (opt->dump_scores) all = (int32_t *)calloc((size_t)nq * (n_seqs ? n_seqs : 1), sizeof(int32_t));
    if (!hits || !n_hits || (opt->dump_scores && !all)) { printf("OSWALD: An error occurred while allocating memory.\n"); ret = 1; goto done; }
    osw_timing tm;
    time_t current_time = time(NULL);
    rc = osw_search(ctx, a, a_disp, nq, matrix, opt->open_gap, opt->extend_gap, (int)top, hits, n_hits, all, &tm);
    if (rc != OSW_OK) {
        printf("OSWALD: search failed: %s (%s).\n", osw_strerror(rc), osw_last_error());
        ret = 1; goto done;
    }
    PHASE("osw_search");
    if (opt->dump_scores) {
        FILE *f = fopen(opt->dump_scores, "wb");
        if (!f) { printf("OSWALD: cannot write %s.\n", opt->dump_scores); ret = 2; goto done; }
        fwrite(all, sizeof(int32_t), (size_t)nq * n_seqs, f);
        fclose(f);
    }
    /* titles of the printed hits only */
    n_print = (size_t)nq * top;
    idx = (uint32_t *)malloc((n_print ? n_print : 1) * sizeof(uint32_t));
    titles = (char **)calloc(n_print ? n_print : 1, sizeof(char *));
    if (!idx || !titles) { printf("OSWALD: An error occurred while allocating memory.\n"); ret = 1; goto done; }
    for (size_t k = 0; k < n_print; ++k) idx[k] = hits[k].index;
    if ((rc = load_database_headers(opt->sequences_filename, idx, n_print, titles)) != 0) { ret = rc; goto done; }
    PHASE("load titles of the hits");

    for (int i = 0; i < nq; ++i) {                     /* report: HybridSearch.c:1213-1224 */
        printf("\nQuery no.\t\t\t%d\n", i + 1);
        printf("Query description: \t\t%s\n", qf.titles[qperm[i]]);
        printf("Query length:\t\t\t%d residues\n", (int)(a_disp[i + 1] - a_disp[i]));
        printf("\nScore\tSequence description\n");
        for (unsigned long j = 0; j < top; ++j)
            printf("%d\t%s", hits[(size_t)i * top + j].score, titles[(size_t)i * top + j] + 1);
    }
    {
        const double secs = tm.wall_ms / 1e3;
        printf("\nSearch date:\t\t\t%s", ctime(&current_time));
        printf("Search time:\t\t\t%lf seconds\n", secs);
        printf("Search speed:\t\t\t%.2lf GCUPS\n", secs > 0 ? (double)Q * (double)n_residues / (secs * 1e9) : 0.0);
        printf("GPU time:\t\t\t%lf seconds\n", tm.device_ms / 1e3);
        printf("GPU speed:\t\t\t%.2lf GCUPS\n", tm.device_ms > 0 ? (double)Q * (double)n_residues / (tm.device_ms * 1e6) : 0.0);
        printf("Number of GPUs:\t\t\t%u\n", opt->num_devices);
        printf("Kernel launches:\t\t%lu\n", (unsigned long)tm.launches);
        printf("Pairs re-scored at 32 bit:\t%lu\n", (unsigned long)tm.rescored_pairs);
        printf("Database layout:\t\t%s\n", use_osw ? "X.osw (device layout from disk)" : "X.seq (laid out at start)");
        if (opt->max_chunk_size_given) printf("Max. chunk size on GPU:\t\t%lu bytes (database streamed through two windows)\n", opt->max_chunk_size);
        else printf("Max. chunk size on GPU:\t\tdatabase resident\n");
    }
done:
    if (titles) for (size_t k = 0; k < n_print; ++k) free(titles[k]);
    free(titles); free(idx); free(hits); free(n_hits); free(all); free(a); free(a_disp); free(qperm);
    if (ctx) osw_free(ctx);
    free_database(&db);
    osw_fasta_free(&qf);
    return ret;
}
