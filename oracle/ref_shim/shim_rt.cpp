// Inert OpenCL/AOCL runtime for the oracle build (TEST INFRASTRUCTURE, computes nothing).
//
// Every cl* call succeeds and does nothing; one fake device is reported.  The fake FPGA
// is made to look infinitely slow so that assemble_db_chunks gives it a 0 % share
// (reference sequences.c:842-847, 1026-1032) and the reference's own host AVX2/SSE team
// scores every database sequence.  Environment knobs:
//   OSWALD_ORACLE_DUMP=<file>    append each query's raw int32 score row (N values) there
//   OSWALD_ORACLE_TIMING=<file>  write "test_cpu_s work_s" measured around the CPU team
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include "AOCL_Utils.h"

static osw_shim_handle *token() { static long cell; return (osw_shim_handle *)&cell; }

namespace aocl_utils {
bool setCwdToExeDir() { return true; }               // keep relative paths relative to the caller
cl_platform_id findPlatform(const char *) { return token(); }
cl_device_id *getDevices(cl_platform_id, cl_device_type, cl_uint *count) {
    cl_device_id *d = new cl_device_id[16];
    for (int i = 0; i < 16; ++i) d[i] = token();
    if (count) *count = 1;
    return d;
}
std::string getBoardBinaryFile(const char *prefix, cl_device_id) { return std::string(prefix) + ".aocx"; }
cl_program createProgramFromBinary(cl_context, const char *, const cl_device_id *, unsigned) { return token(); }
void shim_check(int line, const char *file, cl_int status, const char *msg) {
    if (status != CL_SUCCESS) { fprintf(stderr, "shim: %s:%d: %s (%d)\n", file, line, msg, status); exit(1); }
}
}  // namespace aocl_utils

static cl_int ok(cl_int *st) { if (st) *st = CL_SUCCESS; return CL_SUCCESS; }

extern "C" {
cl_int clGetPlatformInfo(cl_platform_id, cl_platform_info, size_t n, void *v, size_t *) {
    if (v && n) snprintf((char *)v, n, "oracle-shim"); return CL_SUCCESS; }
cl_int clGetDeviceInfo(cl_device_id, cl_device_info what, size_t n, void *v, size_t *) {
    if (!v) return CL_SUCCESS;
    memset(v, 0, n);
    if (what == CL_DEVICE_GLOBAL_MEM_SIZE && n >= sizeof(cl_ulong)) {
        cl_ulong big = (cl_ulong)1 << 40;            // never the binding cap on -k (utils.c:160-168)
        memcpy(v, &big, sizeof big);
    } else if ((what == CL_DEVICE_NAME || what == CL_DEVICE_VENDOR || what == CL_DEVICE_VERSION ||
                what == CL_DRIVER_VERSION) && n) {
        snprintf((char *)v, n, "oracle-shim");
    }
    return CL_SUCCESS;
}
cl_context clCreateContext(const cl_context_properties *, cl_uint, const cl_device_id *,
                           void (*)(const char *, const void *, size_t, void *), void *, cl_int *st) { ok(st); return token(); }
cl_command_queue clCreateCommandQueue(cl_context, cl_device_id, cl_command_queue_properties, cl_int *st) { ok(st); return token(); }
cl_int clBuildProgram(cl_program, cl_uint, const cl_device_id *, const char *, void (*)(cl_program, void *), void *) { return CL_SUCCESS; }
cl_kernel clCreateKernel(cl_program, const char *, cl_int *st) { ok(st); return token(); }
cl_mem clCreateBuffer(cl_context, cl_mem_flags, size_t, void *, cl_int *st) { ok(st); return token(); }
cl_int clSetKernelArg(cl_kernel, cl_uint, size_t, const void *) { return CL_SUCCESS; }
cl_int clEnqueueWriteBuffer(cl_command_queue, cl_mem, cl_bool, size_t, size_t, const void *, cl_uint, const cl_event *, cl_event *) { return CL_SUCCESS; }
cl_int clEnqueueReadBuffer(cl_command_queue, cl_mem, cl_bool, size_t, size_t, void *, cl_uint, const cl_event *, cl_event *) { return CL_SUCCESS; }
cl_int clEnqueueNDRangeKernel(cl_command_queue, cl_kernel, cl_uint, const size_t *, const size_t *, const size_t *, cl_uint, const cl_event *, cl_event *ev) { if (ev) *ev = token(); return CL_SUCCESS; }
cl_int clFinish(cl_command_queue) { return CL_SUCCESS; }
cl_int clWaitForEvents(cl_uint, const cl_event *) { return CL_SUCCESS; }
cl_int clReleaseEvent(cl_event) { return CL_SUCCESS; }
cl_int clReleaseMemObject(cl_mem) { return CL_SUCCESS; }
cl_int clReleaseKernel(cl_kernel) { return CL_SUCCESS; }
cl_int clReleaseCommandQueue(cl_command_queue) { return CL_SUCCESS; }
cl_int clReleaseProgram(cl_program) { return CL_SUCCESS; }
cl_int clReleaseContext(cl_context) { return CL_SUCCESS; }
}

// ---- observation hooks (see hooks.h) --------------------------------------------------
static double now_s() { struct timeval tv; gettimeofday(&tv, 0); return tv.tv_sec + tv.tv_usec * 1e-6; }
static double t_cpu0, t_cpu1, t_work0, t_work1;

static void write_timing() {
    const char *path = getenv("OSWALD_ORACLE_TIMING");
    if (!path) return;
    FILE *f = fopen(path, "w");
    if (!f) return;
    fprintf(f, "%.6f %.6f\n", t_cpu1 - t_cpu0, t_work1 - t_work0);
    fclose(f);
}

double oracle_dwalltime(int line) {
    switch (line) {
        case 218: case 224: case 1494: case 1500: return INFINITY;   // fake FPGA: infinitely slow
        case 234: case 1511: t_cpu0 = now_s(); return t_cpu0;         // CPU team, calibration sample
        case 614: case 1882: t_cpu1 = now_s(); return t_cpu1;
        case 633: case 1901: t_work0 = now_s(); return t_work0;       // CPU team, rest of the database
        case 1181: case 2434: t_work1 = now_s(); write_timing(); return t_work1;
        default: return now_s();
    }
}

void sort_scores(int *scores, char **titles, unsigned long int size, int threads);   // reference utils.c:71
void oracle_sort_scores(int *scores, char **titles, unsigned long int size, int threads) {
    const char *path = getenv("OSWALD_ORACLE_DUMP");
    if (path) {
        FILE *f = fopen(path, "ab");
        if (f) { fwrite(scores, sizeof(int), size, f); fclose(f); }
    }
    sort_scores(scores, titles, size, threads);
}
