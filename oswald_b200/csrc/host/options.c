/* options.c - command line, same switches and defaults as the reference (arguments.c:10-152). */
#include "oswald_host.h"
#include "submat.h"
#include <argp.h>
#include <ctype.h>
#include <stdlib.h>
#include <string.h>

const char *argp_program_bug_address = "<oswald-b200 maintainers>";
static char doc[] =
    "\nOSWALD (B200 build): Smith-Waterman protein database search on NVIDIA B200 GPUs, "
    "command-line compatible with enzorucci/OSWALD";

enum { KEY_DUMP = 0x100 };

static struct argp_option options[] = {
    {0, 0, 0, 0, "OSWALD execution", 1},
    {0, 'O', "<string>", 0, "'preprocess' for database preprocessing, 'search' for database search, 'info' for GPU information [REQUIRED]", 1},
    {0, 0, 0, 0, "preprocess", 2},
    {"input", 'i', "<string>", 0, "Input sequence filename (must be in FASTA format). [REQUIRED]", 2},
    {"output", 'o', "<string>", 0, "Output filename. [REQUIRED]", 2},
    {0, 0, 0, 0, "search", 3},
    {"query", 'q', "<string>", 0, "Input query sequence filename (must be in FASTA format). [REQUIRED]", 3},
    {"db", 'd', "<string>", 0, "Preprocessed database output filename. [REQUIRED]", 3},
    {"sm", 's', "<string>", 0, "Substitution matrix. Supported values: blosum45, blosum50, blosum62, blosum80, blosum90, pam30, pam70, pam250 (default: blosum62).", 3},
    {"gap_open", 'g', "<integer>", 0, "Gap open penalty (default: 10).", 3},
    {"gap_extend", 'e', "<integer>", 0, "Gap extend penalty (default: 2).", 3},
    {"top", 'r', "<integer>", 0, "Number of scores to show (default: 10).", 3},
    {"max_chunk_size", 'k', "<integer>", 0, "Maximum size of the database on a GPU at a time (bytes): a larger database is streamed from host memory through two windows of this size while it is scored. Default: keep the database resident (the reference's default, 134217728, applies only when -k is given).", 3},
    {"num_fpgas", 'f', "<integer>", 0, "Number of GPUs (default: 1; the reference's number of FPGAs).", 3},
    {"cpu_threads", 'c', "<integer>", 0, "Number of host threads for preprocessing (default: 4).", 3},
    {"execution_mode", 'm', "<integer>", 0, "Accepted for compatibility (FPGA/hybrid mode); ignored.", 3},
    {"vector_length", 'v', "<integer>", 0, "Accepted for compatibility (host SIMD width); ignored.", 3},
    {"cpu_block_width", 'b', "<integer>", 0, "Accepted for compatibility (host block width); ignored.", 3},
    {"db_percentage", 'p', "<float>", 0, "Accepted for compatibility (FPGA/host calibration sample); ignored.", 3},
    {"dump-scores", KEY_DUMP, "<file>", 0, "Write every raw int32 score (one row per query, canonical order) to <file>.", 3},
    {0}};

static error_t parse_opt(int key, char *arg, struct argp_state *state) {
    osw_options *o = (osw_options *)state->input;
    switch (key) {
        case 'O':
            if (strcmp(arg, "preprocess") && strcmp(arg, "search") && strcmp(arg, "info"))
                argp_failure(state, 1, 0, "%s is not a valid option for execution.", arg);
            o->op = arg;
            break;
        case 'i': o->input_filename = arg; break;
        case 'o': o->output_filename = arg; break;
        case 'q': o->queries_filename = arg; break;
        case 'd': o->sequences_filename = arg; break;
        case 's': {
            int8_t probe[24 * 32];
            if (strlen(arg) >= sizeof o->submat_arg || osw_matrix_by_name(arg, probe) != 0)
                argp_failure(state, 1, 0, "%s is not a valid option for substitution matrix.", arg);
            strcpy(o->submat_arg, arg);
            for (size_t k = 0; k <= strlen(arg); ++k) o->submat_name[k] = (char)toupper((unsigned char)arg[k]);
            break;
        }
        case 'g':
            o->open_gap = atoi(arg);
            if (o->open_gap < 0 || o->open_gap > 255) argp_failure(state, 1, 0, "%d is not a valid option for gap open penalty.", o->open_gap);
            break;
        case 'e':
            o->extend_gap = atoi(arg);
            if (o->extend_gap < 0 || o->extend_gap > 127) argp_failure(state, 1, 0, "%d is not a valid option for gap extend penalty.", o->extend_gap);
            break;
        case 'm': o->execution_mode = atoi(arg); break;
        case 'c':
            o->cpu_threads = atoi(arg);
            if (o->cpu_threads <= 0) argp_failure(state, 1, 0, "The number of host threads must be greater than 0.");
            break;
        case 'b': o->cpu_block_size = atoi(arg); break;
        case 'v': o->cpu_vector_length = atoi(arg); break;
        case 'f': {
            int n = atoi(arg);
            if (n <= 0 || n > MAX_NUM_DEVICES) argp_failure(state, 1, 0, "The number of GPUs must be between 1 and %d.", MAX_NUM_DEVICES);
            o->num_devices = (unsigned)n;
            break;
        }
        case 'k': {
            long v = atol(arg);
            if (v <= 0) argp_failure(state, 1, 0, "The chunk size must be greater than 0.");
            o->max_chunk_size = (unsigned long)v;
            o->max_chunk_size_given = 1;
            break;
        }
        case 'p': o->test_db_percentage = atof(arg); break;
        case 'r': {
            long v = atol(arg);
            if (v < 0) argp_failure(state, 1, 0, "The number of scores to show must be greater than 0.");
            o->top = (unsigned long)v;
            break;
        }
        case KEY_DUMP: o->dump_scores = arg; break;
        case ARGP_KEY_END:
            if (state->argc <= 1) argp_failure(state, 1, 0, "Missing options");
            if (!o->op) argp_failure(state, 1, 0, "OSWALD execution option is required");
            else if (!strcmp(o->op, "preprocess")) {
                if (!o->input_filename) argp_failure(state, 1, 0, "Input sequence filename is required");
                if (!o->output_filename) argp_failure(state, 1, 0, "Output filename is required");
            } else if (!strcmp(o->op, "search")) {
                if (!o->sequences_filename) argp_failure(state, 1, 0, "Database filename is required");
                if (!o->queries_filename) argp_failure(state, 1, 0, "Query sequences filename is required");
            }
            break;
        default: return ARGP_ERR_UNKNOWN;
    }
    return 0;
}

void program_arguments_processing(int argc, char **argv, osw_options *o) {
    memset(o, 0, sizeof *o);
    strcpy(o->submat_arg, "blosum62");
    strcpy(o->submat_name, "BLOSUM62");
    o->open_gap = OPEN_GAP; o->extend_gap = EXTEND_GAP; o->top = TOP; o->max_chunk_size = MAX_CHUNK_SIZE;
    o->num_devices = NUM_DEVICES; o->cpu_threads = CPU_THREADS; o->execution_mode = 1; o->cpu_vector_length = 16;
    o->cpu_block_size = 256; o->test_db_percentage = 0.01;
    struct argp argp = {options, parse_opt, 0, doc};
    argp_parse(&argp, argc, argv, 0, 0, o);
}
