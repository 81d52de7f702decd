/* main.c - entry point, same dispatch as the reference (main.c:35-67). */
#include "oswald_host.h"
#include <string.h>

int main(int argc, char **argv) {
    osw_options opt;
    program_arguments_processing(argc, argv, &opt);
    if (strcmp(opt.op, "preprocess") == 0)
        return preprocess_db(opt.input_filename, opt.output_filename, opt.cpu_threads);
    if (strcmp(opt.op, "info") == 0)
        return gpu_info();
    return gpu_search(&opt);
}
