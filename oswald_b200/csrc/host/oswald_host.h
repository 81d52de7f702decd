/* oswald_host.h - host side of the command-line tool (plain C).
 *
 * Mirrors the reference's host interface for the hot path: same options (arguments.c:10-152),
 * same operations (main.c:35-67), same on-disk database triple X.info / X.seq / X.desc
 * (sequences.c:4-220) and the same report (HybridSearch.c:1213-1234).  The scoring itself is
 * behind the C ABI of include/oswald_cuda.h. */
#ifndef OSWALD_HOST_H
#define OSWALD_HOST_H
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#define OSWALD_VERSION "1.0-b200"

/* defaults of the reference: arguments.h:13-28, main.c:27-32 */
#define OPEN_GAP 10
#define EXTEND_GAP 2
#define TOP 10
#define MAX_CHUNK_SIZE 134217728
#define NUM_DEVICES 1
#define MAX_NUM_DEVICES 16
#define CPU_THREADS 4

typedef struct osw_options {
    const char *op;                     /* -O preprocess | search | info */
    const char *input_filename;         /* -i */
    const char *output_filename;        /* -o */
    const char *queries_filename;       /* -q */
    const char *sequences_filename;     /* -d */
    char        submat_arg[16];         /* -s as typed (lower case) */
    char        submat_name[16];        /* upper case, for the report */
    int         open_gap, extend_gap;   /* -g -e */
    unsigned long top;                  /* -r */
    unsigned long max_chunk_size;       /* -k */
    int         max_chunk_size_given;   /* -k was on the command line */
    unsigned    num_devices;            /* -f: number of GPUs (was: FPGAs) */
    int         cpu_threads;            /* -c: host threads for preprocessing */
    /* accepted for command-line compatibility, no effect on a GPU: */
    int         execution_mode;         /* -m */
    int         cpu_vector_length;      /* -v */
    int         cpu_block_size;         /* -b */
    double      test_db_percentage;     /* -p */
    const char *dump_scores;            /* --dump-scores FILE: raw int32 score rows (parity tests) */
} osw_options;

void program_arguments_processing(int argc, char **argv, osw_options *opt);

/* residue code of a FASTA letter, reference sequences.c:163-175.  The reference does not
 * validate its input (anything but A-Z is undefined behaviour there); here such bytes become the
 * dummy residue 23, like J/O/U. */
static inline uint8_t osw_encode_letter(unsigned char c) {
    if (c < 'A' || c > 'Z') return 23;
    unsigned x = (c == 'J' || c == 'O' || c == 'U') ? 'Z' + 1 : c;
    return (uint8_t)(x - 'A' - (x > 'J') - (x > 'O') - (x > 'U'));
}

/* ---- FASTA ------------------------------------------------------------------------------ */
typedef struct osw_fasta {
    uint64_t  n;             /* records */
    uint64_t  n_residues;
    char    **titles;        /* header lines without '>' and without the line end */
    uint64_t *offsets;       /* n+1: record i = residues[offsets[i] .. offsets[i+1]) (codes) */
    uint8_t  *residues;
    char     *title_pool;
} osw_fasta;
int  osw_fasta_read(const char *path, osw_fasta *out);      /* 0, or -1 (cannot open) / -2 (memory) */
int  osw_fasta_read_mt(const char *path, osw_fasta *out, int n_threads);
void osw_fasta_free(osw_fasta *f);
/* perm[k] = record at canonical position k: stable ascending length (sequences.c:1130-1225) */
uint64_t *osw_length_order(const osw_fasta *f);

/* ---- preprocessed database (reference triple) --------------------------------------------- */
int preprocess_db(const char *input_filename, const char *out_filename, int n_procs);

typedef struct osw_database {
    uint64_t  n_seqs, n_residues;
    int       max_title_length;
    uint32_t  max_len;
    uint64_t *offsets;       /* n_seqs+1 */
    uint8_t  *residues;      /* n_residues codes, canonical order */
    void     *map; size_t map_size;       /* mmap of X.seq */
} osw_database;
int  load_database(const char *prefix, osw_database *db);
void free_database(osw_database *db);
/* titles of the given canonical indices, read from X.desc (one pass; result[i] malloc'ed) */
int  load_database_headers(const char *prefix, const uint32_t *indices, size_t n, char **result);

/* ---- operations ------------------------------------------------------------------------- */
int gpu_search(const osw_options *opt);
int gpu_info(void);

#endif
