// api.cu - the C ABI (include/oswald_cuda.h): context, database upload, search orchestration.
//
// One DevState per GPU: its shard of the chunk streams resident in HBM, streams, events.
// osw_search enqueues, per GPU and without host synchronisation in between:
//   score clear -> first-stage launches (one per pass of the plan, plan.cu) -> flagged-pair scan,
// then (after reading each GPU's flagged count) 32-bit re-score -> top-r radix select,
// and only then waits for all GPUs, orders the r keys per query on the host and merges the
// GPUs' lists (the reference's sort_scores order, utils.c:3-86).
#include "osw_internal.h"
#include "sw_t16.h"
#include <algorithm>
#include <chrono>
#include <new>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <vector>

static thread_local char g_err[512];
static int cuda_fail(cudaError_t e, const char *what, int line) {
    snprintf(g_err, sizeof g_err, "%s failed at api.cu:%d: %s", what, line, cudaGetErrorString(e));
    return OSW_E_CUDA;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call, __LINE__); } while (0)

namespace {

constexpr uint32_t MAX_LAUNCH_SLOTS = 4096;
constexpr uint32_t FLAG_CAPACITY_MIN = 1u << 16;
constexpr size_t CONTROL_BYTES = MAX_LAUNCH_SLOTS * sizeof(unsigned long long) + 2 * sizeof(unsigned long long) + (1 + MAX_LAUNCH_SLOTS) * sizeof(uint32_t);

struct LaunchRecord { int G, R, has_in, has_out, pair_db; uint32_t first, end; uint64_t cols; uint32_t express; uint32_t slot; uint32_t x_chunks, x_ctas; int x_R; uint32_t pipe_chunks; int transposed; uint32_t x_slot; };

struct DevState {
    int dev = -1, n_sms = 0;
    std::vector<LaunchRecord> trace;           // first-stage launches of the last search
    cudaStream_t st = nullptr;
    cudaStream_t st_copy = nullptr;                // host -> window copies in streaming mode
    bool streaming = false;                        // the column stream is not resident: two windows
    uint8_t *d_win[2] = {nullptr, nullptr}; size_t win_bytes = 0;
    cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    uint8_t *d_stage = nullptr; size_t stage_cap = 0;         // sequences gathered for the 32-bit re-score (streaming mode)
    uint64_t *d_task_off = nullptr; size_t task_off_cap = 0;
    uint8_t *h_stage = nullptr; size_t h_stage_cap = 0;       // pinned staging of the same (kept until the next search)
    uint64_t *h_task_off = nullptr; size_t h_task_off_cap = 0;
    uint2 *h_pairs = nullptr; size_t h_pairs_cap = 0;
    cudaEvent_t ev[6] = {};
    // database shard
    osw_shard shard = {};
    uint8_t *h_stream = nullptr;               // pinned copy of shard.stream (source of re-uploads)
    uint8_t *h_pair = nullptr, *d_pair = nullptr;   // pair stream (pinned host copy, device)
    uint8_t *d_stream = nullptr; osw_chunk *d_chunks = nullptr, *d_pair_chunks = nullptr; uint32_t *d_canon = nullptr;
    uint64_t *d_seq_off = nullptr; uint32_t *d_seq_len = nullptr;
    // per-search buffers (grown on demand)
    int32_t *d_scores = nullptr; size_t scores_cap = 0;
    uint8_t *d_queries = nullptr; size_t queries_cap = 0;
    uint32_t *d_qoff = nullptr; size_t qoff_cap = 0;
    int8_t *d_matrix = nullptr;
    int2 *d_scratch = nullptr; size_t scratch_cap = 0;
    uint2 *d_bound = nullptr; size_t bound_cap = 0;   // bottom rows handed from pass to pass (in place), one segment of chunks at a time
    unsigned char *d_profile = nullptr;        // the current pass's profile table image
    unsigned char *d_profile_x = nullptr;      // ... of a long-chunk launch beside it
    uint32_t *d_first_table_x = nullptr;
    // pipelined passes over the longest chunks: one stream, profile image, deal table and progress row per pass
    std::vector<cudaStream_t> pipe_st;
    std::vector<cudaEvent_t> pipe_ev;
    unsigned char *d_profile_pipe = nullptr; uint32_t *d_first_table_pipe = nullptr; uint32_t *d_progress = nullptr;
    uint32_t *d_first_table = nullptr;         // the current pass's first-chunk deal (CTAs x warps)
    uint2 *d_pairs = nullptr; uint32_t pairs_cap = 0;
    uint32_t *d_counters = nullptr;            // [0] flagged count, [1..] chunk counters per launch
    unsigned long long *d_task_counter = nullptr;   // [0] i32 queue, [1] n_tasks mirror
    unsigned long long *d_cycles = nullptr;    // [MAX_LAUNCH_SLOTS]
    unsigned char *d_control = nullptr;        // the three above live in this one buffer (cleared by one memset per search)
    TopRWork topr = {}; int topr_nq = 0; uint32_t topr_r = 0;
    unsigned long long *h_keys = nullptr; size_t h_keys_cap = 0;      // pinned
    int32_t *h_scores = nullptr; size_t h_scores_cap = 0;             // pinned
    unsigned long long *h_cycles = nullptr;                            // pinned [MAX_LAUNCH_SLOTS]
    uint32_t *h_counts = nullptr;                                       // pinned [4]
    std::vector<uint32_t> t16_hist;            // the shard's pairs by length (sw_t16.h), made at the first search that looks at it
};

template <typename T>
int grow(T **ptr, size_t *cap, size_t need) {
    if (need <= *cap) return OSW_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr; *cap = 0;
    cudaError_t e = cudaMalloc((void **)ptr, need * sizeof(T));
    if (e != cudaSuccess) { cuda_fail(e, "cudaMalloc", __LINE__); return OSW_E_NOMEM; }
    *cap = need;
    return OSW_OK;
}
template <typename T>
int grow_pinned(T **ptr, size_t *cap, size_t need) {
    if (need <= *cap) return OSW_OK;
    if (*ptr) cudaFreeHost(*ptr);
    *ptr = nullptr; *cap = 0;
    cudaError_t e = cudaMallocHost((void **)ptr, need * sizeof(T));
    if (e != cudaSuccess) { cuda_fail(e, "cudaMallocHost", __LINE__); return OSW_E_NOMEM; }
    *cap = need;
    return OSW_OK;
}

void free_db(DevState &d) {
    cudaSetDevice(d.dev);
    cudaFree(d.d_pair); d.d_pair = nullptr;
    cudaFree(d.d_win[0]); cudaFree(d.d_win[1]); d.d_win[0] = d.d_win[1] = nullptr; d.win_bytes = 0; d.streaming = false;
    if (d.h_pair) cudaFreeHost(d.h_pair);
    d.h_pair = nullptr;
    cudaFree(d.d_stream); cudaFree(d.d_chunks); cudaFree(d.d_pair_chunks); cudaFree(d.d_canon); cudaFree(d.d_seq_off); cudaFree(d.d_seq_len);
    cudaFree(d.d_bound); d.bound_cap = 0;
    d.d_stream = nullptr; d.d_chunks = nullptr; d.d_pair_chunks = nullptr; d.d_canon = nullptr; d.d_seq_off = nullptr; d.d_seq_len = nullptr;
    d.d_bound = nullptr;
    d.t16_hist.clear();
    if (d.h_stream) cudaFreeHost(d.h_stream);
    d.h_stream = nullptr;
    osw_shard_free(&d.shard);
}

int upload_db(DevState &d) {
    const osw_shard &s = d.shard;
    CK(cudaSetDevice(d.dev));
    if (!d.streaming) {
        CK(cudaMemcpyAsync(d.d_stream, d.h_stream, s.stream_bytes, cudaMemcpyHostToDevice, d.st));
        if (d.h_pair && d.d_pair) CK(cudaMemcpyAsync(d.d_pair, d.h_pair, 2 * s.pair_cols, cudaMemcpyHostToDevice, d.st));
    }
    CK(cudaMemcpyAsync(d.d_chunks, s.chunks, s.n_chunks * sizeof(osw_chunk), cudaMemcpyHostToDevice, d.st));
    CK(cudaMemcpyAsync(d.d_pair_chunks, s.pair_chunks, s.n_pair_chunks * sizeof(osw_chunk), cudaMemcpyHostToDevice, d.st));
    CK(cudaMemcpyAsync(d.d_canon, s.canon, s.n_seqs * sizeof(uint32_t), cudaMemcpyHostToDevice, d.st));
    CK(cudaMemcpyAsync(d.d_seq_off, s.seq_off, s.n_seqs * sizeof(uint64_t), cudaMemcpyHostToDevice, d.st));
    CK(cudaMemcpyAsync(d.d_seq_len, s.seq_len, s.n_seqs * sizeof(uint32_t), cudaMemcpyHostToDevice, d.st));
    return OSW_OK;
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

// Experiment / test switches, read from the environment ONCE, in osw_init: the
// search path itself never calls getenv.  None of them changes a result.
struct Tunables {
    uint32_t chunk_cols = 0;            // OSW_CHUNK_COLS: residues per chunk (0 = by database size)
    size_t   bound_budget_cols = (size_t)2 << 30;   // OSW_BOUND_BUDGET_COLS: bottom-row buffer, columns (16 GiB)
    uint64_t score_budget = (uint64_t)8 << 30;      // OSW_SCORE_BUDGET_KB: score matrix per batch of queries
    uint32_t flag_cap = 0;              // OSW_FLAG_CAP: capacity of the flagged-pair list (0 = by problem size)
    bool     express = true;            // OSW_NO_EXPRESS
    double   express_ratio = 1.5;       // OSW_EXPRESS_RATIO
    int      force_g = 0;               // OSW_MIN_G: force a group width for single-pass plans
    int      rmax = 0;                  // OSW_RMAX: rows per lane of the full-height passes (0 = default)
    int      long_chunks = -1;          // OSW_LONG_CHUNKS: force the number of chunks of the long-chunk launch (0 = never)
    int      pipe_chunks = -1;          // OSW_PIPE_CHUNKS: force the number of chunks whose passes are pipelined (0 = never)
    int      transpose = -1;            // OSW_TRANSPOSE: 0 = never the transposed form (sw_t16.cu), 1 = whenever the queries fit it,
                                        // 2 = only for the long chunks of the passes
    double   t16_long_frac = 1.0;       // OSW_T16_LONG_FRAC: chunks whose contended walk exceeds this fraction of the launch's pipe time go to the transposed launch
    double   t16_gang = 1.2;            // OSW_T16_GANG: its gang threshold, as a fraction of the launch's estimated time
    bool     trace = false;             // OSW_TRACE: per-launch report on stderr
    void read() {
        if (const char *e = getenv("OSW_CHUNK_COLS")) { const int v = atoi(e); if (v >= 64) chunk_cols = (uint32_t)v; }
        if (const char *e = getenv("OSW_BOUND_BUDGET_COLS")) { const long long v = atoll(e); if (v >= 1024) bound_budget_cols = (size_t)v; }
        if (const char *e = getenv("OSW_SCORE_BUDGET_KB")) { const long long v = atoll(e); if (v >= 1) score_budget = (uint64_t)v << 10; }
        if (const char *e = getenv("OSW_FLAG_CAP")) { const long long v = atoll(e); if (v >= 1) flag_cap = (uint32_t)v; }
        express = getenv("OSW_NO_EXPRESS") == nullptr;
        if (const char *e = getenv("OSW_EXPRESS_RATIO")) express_ratio = atof(e);
        if (const char *e = getenv("OSW_MIN_G")) force_g = atoi(e);
        if (const char *e = getenv("OSW_RMAX")) rmax = atoi(e);
        if (const char *e = getenv("OSW_LONG_CHUNKS")) long_chunks = atoi(e);
        if (const char *e = getenv("OSW_PIPE_CHUNKS")) pipe_chunks = atoi(e);
        if (const char *e = getenv("OSW_TRANSPOSE")) transpose = atoi(e);
        if (const char *e = getenv("OSW_T16_GANG")) t16_gang = atof(e);
        if (const char *e = getenv("OSW_T16_LONG_FRAC")) t16_long_frac = atof(e);
        trace = getenv("OSW_TRACE") != nullptr;
    }
};

struct osw_ctx {
    Tunables tune;
    int n_dev = 0;
    DevState *devs = nullptr;
    int kernel_mask = OSW_K_DEFAULT;
    uint64_t window_bytes = 0;         // osw_set_device_window
    bool db_loaded = false;
    uint64_t n_seqs_canon = 0;        // size of the whole canonical database
    uint64_t n_seqs_local = 0, residues_local = 0, chunks_local = 0;
};

extern "C" const char *osw_strerror(int code) {
    switch (code) {
        case OSW_OK: return "ok";
        case OSW_E_ARG: return "invalid argument";
        case OSW_E_NODEV: return "no usable CUDA device";
        case OSW_E_CUDA: return "CUDA runtime error";
        case OSW_E_NOMEM: return "out of memory";
        case OSW_E_STATE: return "call order error (database not loaded?)";
        case OSW_E_ARCH: return "device is not sm_100 (B200)";
        case OSW_E_IO: return "file cannot be opened / written";
        case OSW_E_FORMAT: return "not a database file of this format version";
    }
    return "unknown error";
}
extern "C" const char *osw_last_error(void) { return g_err; }

extern "C" int osw_device_count(int *count) {
    if (!count) return OSW_E_ARG;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; cuda_fail(e, "cudaGetDeviceCount", __LINE__); return OSW_E_NODEV; }
    *count = n;
    return OSW_OK;
}

extern "C" int osw_device_info(int device, char *text, size_t n) {
    if (!text || !n) return OSW_E_ARG;
    cudaDeviceProp pr;
    cudaError_t e = cudaGetDeviceProperties(&pr, device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties", __LINE__);
    int clock_khz = 0, mem_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, device);
    cudaDeviceGetAttribute(&mem_khz, cudaDevAttrMemoryClockRate, device);
    snprintf(text, n,
             "Device %d: %s\n  compute capability %d.%d, %d SMs, %.0f MHz\n  global memory %.1f GiB, memory clock %.0f MHz, L2 %.0f MiB\n"
             "  shared memory per block (opt-in) %zu KiB, registers per SM %d\n",
             device, pr.name, pr.major, pr.minor, pr.multiProcessorCount, clock_khz / 1000.0,
             pr.totalGlobalMem / 1073741824.0, mem_khz / 1000.0, pr.l2CacheSize / 1048576.0,
             pr.sharedMemPerBlockOptin / 1024, pr.regsPerMultiprocessor);
    return OSW_OK;
}

extern "C" int osw_init(int n_devices, const int *devices, osw_ctx **out) {
    if (!out || n_devices < 1) return OSW_E_ARG;
    *out = nullptr;
    int avail = 0;
    int rc = osw_device_count(&avail);
    if (rc != OSW_OK || avail < 1) return OSW_E_NODEV;
    osw_ctx *c = new (std::nothrow) osw_ctx;
    if (!c) return OSW_E_NOMEM;
    c->devs = new (std::nothrow) DevState[n_devices];
    if (!c->devs) { delete c; return OSW_E_NOMEM; }
    c->n_dev = n_devices;
    c->tune.read();
    if (n_devices > 1) {
        // creating a device's primary context takes about a second: do it for all devices at once
        // (measured with eight GPUs: osw_init 7.9 s one after the other)
        std::vector<std::thread> warm;
        for (int i = 0; i < n_devices; ++i) {
            const int dev = devices ? devices[i] : i;
            if (dev >= 0 && dev < avail) warm.emplace_back([dev] { if (cudaSetDevice(dev) == cudaSuccess) cudaFree(nullptr); });
        }
        for (std::thread &t : warm) t.join();
    }
    for (int i = 0; i < n_devices; ++i) {
        DevState &d = c->devs[i];
        d.dev = devices ? devices[i] : i;
        if (d.dev < 0 || d.dev >= avail) { osw_free(c); return OSW_E_NODEV; }
        cudaDeviceProp pr;
        cudaError_t e = cudaGetDeviceProperties(&pr, d.dev);
        if (e != cudaSuccess) { osw_free(c); return cuda_fail(e, "cudaGetDeviceProperties", __LINE__); }
        if (pr.major != 10) {
            snprintf(g_err, sizeof g_err, "device %d (%s) is sm_%d%d; this library is built for sm_100a only",
                     d.dev, pr.name, pr.major, pr.minor);
            osw_free(c);
            return OSW_E_ARCH;
        }
        d.n_sms = pr.multiProcessorCount;
        if ((e = cudaSetDevice(d.dev)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&d.st, cudaStreamNonBlocking)) != cudaSuccess) {
            osw_free(c); return cuda_fail(e, "stream setup", __LINE__);
        }
        for (auto &ev : d.ev) if ((e = cudaEventCreate(&ev)) != cudaSuccess) { osw_free(c); return cuda_fail(e, "cudaEventCreate", __LINE__); }
        if ((e = cudaStreamCreateWithFlags(&d.st_copy, cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&d.ev_ready[0], cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&d.ev_ready[1], cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&d.ev_free[0], cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&d.ev_free[1], cudaEventDisableTiming)) != cudaSuccess) {
            osw_free(c); return cuda_fail(e, "stream setup", __LINE__);
        }
        if ((e = cudaMalloc(&d.d_matrix, 24 * 32)) != cudaSuccess ||
            (e = cudaMalloc(&d.d_profile, OSW_PROFILE_BYTES)) != cudaSuccess ||
            (e = cudaMalloc(&d.d_first_table, 1024 * 16 * sizeof(uint32_t))) != cudaSuccess ||
            (e = cudaMalloc(&d.d_profile_x, OSW_PROFILE_BYTES)) != cudaSuccess ||
            (e = cudaMalloc(&d.d_first_table_x, 1024 * 16 * sizeof(uint32_t))) != cudaSuccess ||
            (e = cudaMalloc(&d.d_control, CONTROL_BYTES)) != cudaSuccess ||
            (e = cudaMallocHost(&d.h_cycles, MAX_LAUNCH_SLOTS * sizeof(unsigned long long))) != cudaSuccess ||
            (e = cudaMallocHost(&d.h_counts, 4 * sizeof(uint32_t))) != cudaSuccess) {
            osw_free(c); return cuda_fail(e, "cudaMalloc", __LINE__);
        }
        d.d_cycles = reinterpret_cast<unsigned long long *>(d.d_control);
        d.d_task_counter = d.d_cycles + MAX_LAUNCH_SLOTS;
        d.d_counters = reinterpret_cast<uint32_t *>(d.d_task_counter + 2);
    }
    *out = c;
    return OSW_OK;
}

extern "C" void osw_free(osw_ctx *c) {
    if (!c) return;
    for (int i = 0; i < c->n_dev; ++i) {
        DevState &d = c->devs[i];
        if (d.dev < 0) continue;
        cudaSetDevice(d.dev);
        if (d.st) cudaStreamSynchronize(d.st);
        free_db(d);
        cudaFree(d.d_scores); cudaFree(d.d_queries); cudaFree(d.d_qoff); cudaFree(d.d_matrix); cudaFree(d.d_scratch);
        cudaFree(d.d_profile); cudaFree(d.d_first_table); cudaFree(d.d_profile_x); cudaFree(d.d_first_table_x);
        cudaFree(d.d_pairs); cudaFree(d.d_control);
        cudaFree(d.topr.hist); cudaFree(d.topr.prefix); cudaFree(d.topr.remaining); cudaFree(d.topr.out_count); cudaFree(d.topr.out_keys);
        if (d.h_keys) cudaFreeHost(d.h_keys);
        if (d.h_scores) cudaFreeHost(d.h_scores);
        if (d.h_cycles) cudaFreeHost(d.h_cycles);
        if (d.h_counts) cudaFreeHost(d.h_counts);
        for (auto &ev : d.ev) if (ev) cudaEventDestroy(ev);
        for (int k = 0; k < 2; ++k) { if (d.ev_ready[k]) cudaEventDestroy(d.ev_ready[k]); if (d.ev_free[k]) cudaEventDestroy(d.ev_free[k]); }
        if (d.st_copy) cudaStreamDestroy(d.st_copy);
        for (cudaStream_t ps : d.pipe_st) cudaStreamDestroy(ps);
        for (cudaEvent_t pe : d.pipe_ev) cudaEventDestroy(pe);
        cudaFree(d.d_profile_pipe); cudaFree(d.d_first_table_pipe); cudaFree(d.d_progress);
        cudaFree(d.d_stage); cudaFree(d.d_task_off);
        if (d.h_stage) cudaFreeHost(d.h_stage);
        if (d.h_task_off) cudaFreeHost(d.h_task_off);
        if (d.h_pairs) cudaFreeHost(d.h_pairs);
        if (d.st) cudaStreamDestroy(d.st);
    }
    delete[] c->devs;
    delete c;
}

extern "C" int osw_set_kernels(osw_ctx *c, int mask) {
    if (!c || !(mask & OSW_K_I32)) return OSW_E_ARG;   // the 32-bit stage is never optional
    if ((mask & OSW_K_TWO_TRACK) && (mask & OSW_K_PAIR_DB)) return OSW_E_ARG;
    c->kernel_mask = mask;
    return OSW_OK;
}

extern "C" int osw_set_device_window(osw_ctx *c, uint64_t bytes) {
    if (!c) return OSW_E_ARG;
    if (bytes && bytes < (1u << 20)) bytes = 1u << 20;          // a window holds at least a few of the longest chunks
    c->window_bytes = bytes;
    return OSW_OK;
}

namespace {

// Work-unit size: 8192 residues for large databases (0.8 % pipeline fill per chunk; measured best
// together with the quarter-size chunks at the end of the queue); smaller for small databases so
// that every group of lanes on every SM gets several chunks.
uint32_t pick_chunk_cols(const Tunables &tune, uint64_t n_residues, uint64_t n_devices_total, uint64_t max_chunk_residues) {
    uint32_t chunk_cols = OSW_CHUNK_COLS_DEFAULT;
    const uint64_t per_dev = n_residues / (n_devices_total ? n_devices_total : 1);
    const uint64_t fit = per_dev / 25000;            // ~ 148 SMs x 12 warps x 8 groups x 2
    if (n_residues && fit < chunk_cols) chunk_cols = (uint32_t)(fit < 256 ? 256 : fit);
    if (tune.chunk_cols) chunk_cols = tune.chunk_cols;             // experiments
    if (max_chunk_residues && max_chunk_residues < chunk_cols) chunk_cols = (uint32_t)max_chunk_residues;
    return chunk_cols;
}

// Builds every GPU's shard with `build(device index, global shard, n_shards, allocator, user, out)`
// - from the caller's canonical arrays or from a mapped X.osw file - straight into pinned host
// memory (the source of every upload), allocates the device copies and uploads them.
template <typename BuildShard>
int load_shards(osw_ctx *c, uint64_t n_seqs_canon, int shard_rank, int shard_count, BuildShard build) {
    const uint32_t n_shards = (uint32_t)shard_count * (uint32_t)c->n_dev;
    c->db_loaded = false;
    c->n_seqs_canon = n_seqs_canon; c->n_seqs_local = 0; c->residues_local = 0; c->chunks_local = 0;
    for (int i = 0; i < c->n_dev; ++i) {
        DevState &d = c->devs[i];
        free_db(d);
        struct Pinned { uint8_t *ptr[2]; int n; } pinned = {{nullptr, nullptr}, 0};
        auto pinned_alloc = [](size_t bytes, void *user) -> void * {
            Pinned *pn = (Pinned *)user;
            void *ptr = nullptr;
            if (pn->n >= 2 || cudaMallocHost(&ptr, bytes) != cudaSuccess) return nullptr;
            pn->ptr[pn->n++] = (uint8_t *)ptr;
            return ptr;
        };
        CK(cudaSetDevice(d.dev));
        const int brc = build((uint32_t)shard_rank * c->n_dev + i, n_shards, pinned_alloc, &pinned, &d.shard);
        if (brc != 0) {
            for (int k = 0; k < pinned.n; ++k) cudaFreeHost(pinned.ptr[k]);
            if (brc == -2) { snprintf(g_err, sizeof g_err, "the database holds a residue code outside 0..23"); return OSW_E_ARG; }
            return OSW_E_NOMEM;
        }
        d.h_stream = d.shard.stream; d.h_pair = d.shard.pair_stream;     // owned by DevState (freed in free_db)
        d.shard.stream = nullptr; d.shard.pair_stream = nullptr;
        const osw_shard &s = d.shard;
        d.streaming = c->window_bytes && s.stream_bytes > c->window_bytes;
        if (d.streaming) {
            d.win_bytes = (size_t)c->window_bytes;
            CK(cudaMalloc(&d.d_win[0], d.win_bytes));
            CK(cudaMalloc(&d.d_win[1], d.win_bytes));
        } else {
            CK(cudaMalloc(&d.d_stream, s.stream_bytes ? s.stream_bytes : 1));
        }
        CK(cudaMalloc(&d.d_chunks, (s.n_chunks ? s.n_chunks : 1) * sizeof(osw_chunk)));
        CK(cudaMalloc(&d.d_pair_chunks, (s.n_pair_chunks ? s.n_pair_chunks : 1) * sizeof(osw_chunk)));
        CK(cudaMalloc(&d.d_canon, (s.n_seqs ? s.n_seqs : 1) * sizeof(uint32_t)));
        CK(cudaMalloc(&d.d_seq_off, (s.n_seqs ? s.n_seqs : 1) * sizeof(uint64_t)));
        CK(cudaMalloc(&d.d_seq_len, (s.n_seqs ? s.n_seqs : 1) * sizeof(uint32_t)));
        int rc = upload_db(d);
        if (rc != OSW_OK) return rc;
        c->n_seqs_local += s.n_seqs; c->residues_local += s.n_residues; c->chunks_local += s.n_chunks;
    }
    for (int i = 0; i < c->n_dev; ++i) {
        DevState &d = c->devs[i];
        CK(cudaSetDevice(d.dev));
        CK(cudaStreamSynchronize(d.st));
    }
    c->db_loaded = true;
    return OSW_OK;
}

}  // namespace

extern "C" int osw_db_load(osw_ctx *c, const uint8_t *residues, const uint64_t *offsets, uint64_t n_seqs,
                           int shard_rank, int shard_count, uint64_t max_chunk_residues) {
    if (!c || shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count) return OSW_E_ARG;
    if (n_seqs && (!residues || !offsets)) return OSW_E_ARG;
    if (n_seqs > 0xffffffffull) return OSW_E_ARG;
    // The layout and the kernels rely on the canonical order (sequences.c:1130-1225): lengths
    // ascending (a chunk's last sequence is its longest; empty sequences lead), offsets monotonic.
    for (uint64_t i = 0, prev = 0; i < n_seqs; ++i) {
        if (offsets[i + 1] < offsets[i]) { snprintf(g_err, sizeof g_err, "database offsets decrease at sequence %llu", (unsigned long long)i); return OSW_E_ARG; }
        const uint64_t len = offsets[i + 1] - offsets[i];
        if (len < prev || len > 0x7fffffffull) {
            snprintf(g_err, sizeof g_err, "database sequence %llu (length %llu) breaks the ascending-length order",
                     (unsigned long long)i, (unsigned long long)len);
            return OSW_E_ARG;
        }
        prev = len;
    }
    const uint32_t chunk_cols = pick_chunk_cols(c->tune, n_seqs ? offsets[n_seqs] : 0, (uint64_t)shard_count * c->n_dev, max_chunk_residues);
    return load_shards(c, n_seqs, shard_rank, shard_count,
                       [&](uint32_t shard, uint32_t n_shards, osw_alloc_fn alloc, void *user, osw_shard *out) {
                           return osw_shard_build_ex(residues, offsets, n_seqs, shard, n_shards, chunk_cols,
                                                     0 /* the pair stream is made when a search first needs it */, alloc, user, out);
                       });
}

// ---- X.osw: the chunk streams on disk (dbformat.h) -------------------------------------------------
extern "C" int osw_db_write_file(const char *path, const uint8_t *residues, const uint64_t *offsets, uint64_t n_seqs,
                                 uint64_t max_chunk_residues) {
    if (!path || (n_seqs && (!residues || !offsets)) || n_seqs > 0xffffffffull) return OSW_E_ARG;
    for (uint64_t i = 0, prev = 0; i < n_seqs; ++i) {
        if (offsets[i + 1] < offsets[i] || offsets[i + 1] - offsets[i] < prev) {
            snprintf(g_err, sizeof g_err, "database sequence %llu breaks the ascending-length order", (unsigned long long)i);
            return OSW_E_ARG;
        }
        prev = offsets[i + 1] - offsets[i];
    }
    Tunables tune;
    tune.read();
    const int rc = osw_dbfile_write(path, residues, offsets, n_seqs, pick_chunk_cols(tune, n_seqs ? offsets[n_seqs] : 0, 1, max_chunk_residues));
    if (rc == -2) { snprintf(g_err, sizeof g_err, "the database holds a residue code outside 0..23"); return OSW_E_ARG; }
    if (rc == -1) { snprintf(g_err, sizeof g_err, "cannot write %s", path); return OSW_E_IO; }
    return rc ? OSW_E_NOMEM : OSW_OK;
}

static int dbfile_error(int rc, const char *path) {
    snprintf(g_err, sizeof g_err, rc == -1 ? "cannot open %s" : rc == -2 ? "%s is not an X.osw file of format version 1" : "%s is truncated or corrupt", path);
    return rc == -1 ? OSW_E_IO : OSW_E_FORMAT;
}

extern "C" int osw_db_file_info(const char *path, uint64_t *n_seqs, uint64_t *n_residues, uint64_t *n_chunks, uint32_t *max_len, uint32_t *version) {
    if (!path) return OSW_E_ARG;
    osw_dbfile f;
    const int rc = osw_dbfile_open(path, &f);
    if (rc) return dbfile_error(rc, path);
    if (n_seqs) *n_seqs = f.h.n_seqs;
    if (n_residues) *n_residues = f.h.n_residues;
    if (n_chunks) *n_chunks = f.h.n_chunks;
    if (max_len) *max_len = f.h.max_len;
    if (version) *version = f.h.version;
    osw_dbfile_close(&f);
    return OSW_OK;
}

extern "C" int osw_db_load_file(osw_ctx *c, const char *path, int shard_rank, int shard_count) {
    if (!c || !path || shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count) return OSW_E_ARG;
    osw_dbfile f;
    int rc = osw_dbfile_open(path, &f);
    if (rc) return dbfile_error(rc, path);
    rc = load_shards(c, f.h.n_seqs, shard_rank, shard_count,
                     [&](uint32_t shard, uint32_t n_shards, osw_alloc_fn alloc, void *user, osw_shard *out) {
                         return osw_shard_from_file(&f, shard, n_shards, alloc, user, out);
                     });
    osw_dbfile_close(&f);
    return rc;
}

extern "C" int osw_db_upload(osw_ctx *c, uint64_t *bytes) {
    if (!c || !c->db_loaded) return OSW_E_STATE;
    uint64_t total = 0;
    for (int i = 0; i < c->n_dev; ++i) {
        int rc = upload_db(c->devs[i]);
        if (rc != OSW_OK) return rc;
        const osw_shard &s = c->devs[i].shard;
        total += s.stream_bytes + (c->devs[i].h_pair ? 2 * s.pair_cols : 0) + s.n_chunks * sizeof(osw_chunk) + s.n_seqs * (sizeof(uint32_t) * 2 + sizeof(uint64_t));
    }
    for (int i = 0; i < c->n_dev; ++i) {
        CK(cudaSetDevice(c->devs[i].dev));
        CK(cudaStreamSynchronize(c->devs[i].st));
    }
    if (bytes) *bytes = total;
    return OSW_OK;
}

extern "C" int osw_db_stats(const osw_ctx *c, uint64_t *n_seqs_local, uint64_t *residues_local, uint64_t *n_chunks_local) {
    if (!c || !c->db_loaded) return OSW_E_STATE;
    if (n_seqs_local) *n_seqs_local = c->n_seqs_local;
    if (residues_local) *residues_local = c->residues_local;
    if (n_chunks_local) *n_chunks_local = c->chunks_local;
    return OSW_OK;
}

// Orders two hits like the reference's merge sort leaves them (utils.c:3-69).
static inline bool hit_before(const osw_hit &a, const osw_hit &b) {
    return a.score != b.score ? a.score > b.score : a.index > b.index;
}

extern "C" size_t osw_merge_hits(const osw_hit *const *lists, const uint32_t *counts, int n_lists,
                                 uint32_t top_r, osw_hit *out) {
    if (!lists || !counts || !out || n_lists < 1) return 0;
    std::vector<uint32_t> cur((size_t)n_lists, 0u);
    size_t n = 0;
    while (n < top_r) {
        int pick = -1;
        for (int s = 0; s < n_lists; ++s) {
            if (cur[s] >= counts[s]) continue;
            if (pick < 0 || hit_before(lists[s][cur[s]], lists[pick][cur[pick]])) pick = s;
        }
        if (pick < 0) break;
        out[n++] = lists[pick][cur[pick]++];
    }
    return n;
}

namespace {

constexpr int MAX_PASSES = 2048;

// Timing model of one first-stage launch, in SM cycles (measured on B200: tools/dpx_latency.cu, the
// OSW_TRACE reports of bench.py on small databases; profiles/r1_ncu_summary.md).  A warp-step (one
// column for each of the warp's 32/G groups) takes `issue` cycles of its scheduler's issue slots /
// DPX pipe, and at least `alone` cycles from start to end (mailbox and profile reads, then the
// dependent row chain: 14 cycles per row, two row segments in parallel).
struct LaunchModel {
    double issue, alone, contended, t_pipe;
    int groups, warps_per_scheduler;
    // n_chunks: chunks of the launch (0 = plenty).  On a database of few chunks not every warp of a
    // scheduler has work, and a chunk is walked faster than among three busy neighbours.
    LaunchModel(int G, int R, bool pd, double cols, int n_sms, double n_chunks = 0.0) {
        groups = 32 / G;
        warps_per_scheduler = R > 32 ? 3 : 4;                       // 384- / 512-thread CTAs
        issue = (pd ? 11.7 : 9.0) * R + 45.0;
        alone = 140.0 + 7.0 * R;
        double busy = warps_per_scheduler;
        if (n_chunks > 0.0) busy = std::min(busy, std::max(1.0, n_chunks / ((double)groups * std::max(n_sms, 1) * 4.0)));
        contended = std::max(alone, busy * issue);
        t_pipe = cols / ((double)groups * std::max(n_sms, 1) * 4.0) * issue;
    }
};

// Long chunks in the TRANSPOSED form (sw_t16.cu).  The chunks of a single-pass launch that could
// not be walked among busy neighbours within the launch's pipe time hold the shard's longest sequences; a
// transposed launch scores those pairs in a time that does not depend on their length (m + 31 steps per block
// of 128 rows, gangs of warps for the longest), on as many SMs as balance it - with a margin: its model is the
// less certain one, and an SM more costs the other launch under a per cent - against the launch that walks the
// rest.  (The same kernel with the widest array needed one dependent step per column: on 100 000 sequences x one
// short query its 15 SMs were still walking when the other 133 had been idle for a third of the search.)
struct LongT16 { uint32_t n_chunks = 0, n_ctas = 0; OswT16Plan plan; double est = 0; };
bool plan_long_t16(const osw_ctx *c, const DevState &d, const OswPass &ps, uint32_t first, uint32_t end, uint64_t cols,
                   const uint32_t *q_off, int nq, LongT16 *out) {
    const osw_shard &s = d.shard;
    const bool pd = ps.pair_db != 0;
    const osw_chunk *dir = pd ? s.pair_chunks : s.chunks;
    auto chunk_cols = [&](uint32_t k) -> double { return pd ? (double)dir[k].n_pair_cols : (double)dir[k].n_cols; };
    *out = LongT16();
    if (end <= first || !c->tune.express || d.streaming || c->tune.long_chunks == 0) return false;
    const LaunchModel m(ps.G, ps.R, pd, (double)cols, d.n_sms, (double)(end - first));
    if (!(chunk_cols(first) * m.contended > c->tune.express_ratio * m.t_pipe)) return false;
    const LaunchModel full(ps.G, ps.R, pd, (double)cols, d.n_sms);
    double x_cols = 0;
    uint32_t n = 0;
    while (first + n < end && n < 16384u && chunk_cols(first + n) * full.contended > c->tune.t16_long_frac * full.t_pipe) { x_cols += chunk_cols(first + n); ++n; }
    if (c->tune.long_chunks > 0) {          // experiments: force the number of chunks
        n = std::min<uint32_t>((uint32_t)c->tune.long_chunks, end - first);
        x_cols = 0;
        for (uint32_t k = 0; k < n; ++k) x_cols += chunk_cols(first + k);
    }
    if (!n) return false;
    // (chunks are in descending, sequences in ascending length order: the chunks' sequences are the last ones
    // of the shard; a pair that straddles the cut is scored by both launches, with the same result)
    const uint64_t N = s.n_seqs, s_cut = dir[first + n - 1].seq0 & ~1u;
    std::vector<uint32_t> hist_x(OSW_T16_MAX_ROWS32 + 1);
    osw_t16_histogram(s.seq_len + s_cut, N - s_cut, hist_x.data());
    for (int k : {2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64, 96}) {
        if (k > d.n_sms - 16) break;
        OswT16Plan pl;
        osw_t16_plan(hist_x.data(), N - s_cut, q_off, nq, k, c->tune.t16_gang, &pl);
        if (pl.warps != 16) break;
        const double t_rest = full.t_pipe * ((double)cols - x_cols) / (double)std::max<uint64_t>(cols, 1) * d.n_sms / (d.n_sms - k);
        const double t = std::max(t_rest, pl.est_cycles / 0.75);
        if (!out->n_chunks || t < out->est) { out->est = t; out->plan = pl; out->n_ctas = (uint32_t)k; out->n_chunks = n; }
    }
    if (!out->n_chunks) return false;
    out->plan.n_pairs = (uint32_t)((N + 1) / 2);
    return true;
}

int enqueue_search(osw_ctx *c, DevState &d, const uint8_t *queries, const uint32_t *q_off, int nq,
                   const int8_t *matrix, int go, int ge, uint32_t top_r, bool want_all,
                   const std::vector<OswPass> &passes, const OswPass *pass_x, const OswT16Plan *tplan, bool t16_long,
                   uint32_t *n_launch_slots, uint64_t *launches, uint64_t *padded_cells) {
    const osw_shard &s = d.shard;
    const uint64_t N = s.n_seqs;
    const size_t q_bytes = q_off[nq];
    CK(cudaSetDevice(d.dev));
    int rc;
    if ((rc = grow(&d.d_scores, &d.scores_cap, (size_t)nq * (N ? N : 1))) != OSW_OK) return rc;
    if ((rc = grow(&d.d_queries, &d.queries_cap, q_bytes ? q_bytes : 1)) != OSW_OK) return rc;
    if ((rc = grow(&d.d_qoff, &d.qoff_cap, (size_t)nq + 1)) != OSW_OK) return rc;
    const uint32_t r = (uint32_t)std::min<uint64_t>(top_r, N);
    if (nq > d.topr_nq || r > d.topr_r) {
        cudaFree(d.topr.hist); cudaFree(d.topr.prefix); cudaFree(d.topr.remaining); cudaFree(d.topr.out_count); cudaFree(d.topr.out_keys);
        d.topr = TopRWork(); d.topr_nq = 0; d.topr_r = 0;
        const int qn = std::max(nq, d.topr_nq); const uint32_t rn = std::max(r, std::max(d.topr_r, 1u));
        CK(cudaMalloc(&d.topr.hist, (size_t)qn * 8 * 256 * sizeof(uint32_t)));
        CK(cudaMalloc(&d.topr.prefix, (size_t)qn * 8 * sizeof(unsigned long long)));
        CK(cudaMalloc(&d.topr.remaining, (size_t)qn * 8 * sizeof(uint32_t)));
        CK(cudaMalloc(&d.topr.out_count, (size_t)qn * sizeof(uint32_t)));
        CK(cudaMalloc(&d.topr.out_keys, (size_t)qn * rn * sizeof(unsigned long long)));
        d.topr_nq = qn; d.topr_r = rn;
    }
    if ((rc = grow_pinned(&d.h_keys, &d.h_keys_cap, (size_t)nq * std::max(r, 1u))) != OSW_OK) return rc;
    if (want_all && (rc = grow_pinned(&d.h_scores, &d.h_scores_cap, (size_t)nq * (N ? N : 1))) != OSW_OK) return rc;

    const bool use_u16 = (c->kernel_mask & OSW_K_U16) != 0;
    d.h_counts[0] = 0;                 // an empty shard launches nothing: its flagged count must not be a previous search's
    uint32_t flag_cap = 0;
    if (use_u16) {
        uint64_t want = std::max<uint64_t>(FLAG_CAPACITY_MIN, (uint64_t)nq * N / 64);
        flag_cap = (uint32_t)std::min<uint64_t>(want, 1u << 26);
        if (c->tune.flag_cap) flag_cap = c->tune.flag_cap;             // tests: force several re-score rounds
        size_t cap = d.pairs_cap;
        if ((rc = grow(&d.d_pairs, &cap, flag_cap)) != OSW_OK) return rc;
        d.pairs_cap = (uint32_t)cap;
        flag_cap = d.pairs_cap;
        bool need_bound = false;
        for (const OswPass &p : passes) need_bound |= p.has_in || p.has_out;
        if (need_bound) {
            // The bottom rows take 8 bytes per column - 8 x the database itself.  They are kept for one
            // SEGMENT of consecutive chunks at a time (all passes run over a segment before the next
            // one starts), so the buffer is bounded whatever the database size.
            const size_t budget_cols = c->tune.bound_budget_cols;            // 16 GiB unless a test shrinks it
            const size_t all_cols = std::max<uint64_t>(std::max<uint64_t>(s.stream_bytes, s.pair_cols), 1);
            const size_t want = std::min(all_cols, budget_cols + 2 * 65536 + 1024);
            const size_t had = d.bound_cap;
            if ((rc = grow(&d.d_bound, &d.bound_cap, want)) != OSW_OK) return rc;
            if (d.bound_cap != had) CK(cudaMemsetAsync(d.d_bound, 0, d.bound_cap * sizeof(uint2), d.st));   // (padding columns are read)
        }
    }

    // ---- uploads ----
    CK(cudaMemcpyAsync(d.d_queries, queries, q_bytes, cudaMemcpyHostToDevice, d.st));
    CK(cudaMemcpyAsync(d.d_qoff, q_off, ((size_t)nq + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, d.st));
    CK(cudaMemcpyAsync(d.d_matrix, matrix, 24 * 32, cudaMemcpyHostToDevice, d.st));
    CK(cudaEventRecord(d.ev[0], d.st));
    CK(cudaMemsetAsync(d.d_scores, 0, (size_t)nq * N * sizeof(int32_t), d.st));
    CK(cudaMemsetAsync(d.d_control, 0, CONTROL_BYTES, d.st));          // counters, task queue, cycle sums: one clear

    uint32_t slot = 0;
    d.trace.clear();
    if (use_u16 && N && !passes.empty() && passes[0].pair_db && !d.h_pair) {
        // first pair-database search on this database: derive the pair stream from the plain one
        CK(cudaMallocHost(&d.h_pair, s.pair_cols ? 2 * s.pair_cols : 1));
        osw_shard_fill_pair(&s, d.h_stream, d.h_pair);
        if (!d.streaming) {
            CK(cudaMalloc(&d.d_pair, s.pair_cols ? 2 * s.pair_cols : 1));
            CK(cudaMemcpyAsync(d.d_pair, d.h_pair, 2 * s.pair_cols, cudaMemcpyHostToDevice, d.st));
        }
    }
    if (use_u16 && N && tplan && tplan->warps) {
        // The transposed form (sw_t16.cu): one launch scores every pair of sequences against every query.
        if (slot + OSW_T16_CLASSES > MAX_LAUNCH_SLOTS) { snprintf(g_err, sizeof g_err, "too many launches"); return OSW_E_ARG; }
        OswT16Params tp;
        tp.stream = d.d_stream; tp.seq_off = d.d_seq_off; tp.seq_len = d.d_seq_len;
        tp.queries = d.d_queries; tp.q_off = d.d_qoff; tp.nq = nq; tp.matrix = d.d_matrix;
        tp.scores = d.d_scores; tp.n_seqs = N;
        tp.counters = d.d_counters + 1 + slot; tp.cycle_acc = d.d_cycles + slot;
        tp.gap_open_extend = go + ge; tp.gap_extend = ge;
        if (osw_launch_t16(tp, *tplan, d.n_sms, d.st) != OSW_OK) { cuda_fail(cudaGetLastError(), "sw_t16 launch", __LINE__); return OSW_E_CUDA; }
        LaunchRecord lr = {};
        lr.end = tplan->n_pairs; lr.cols = (uint64_t)nq * s.n_residues; lr.slot = slot; lr.transposed = tplan->warps;
        lr.first = tplan->class_begin[1]; lr.x_chunks = tplan->class_begin[2]; lr.x_ctas = tplan->class_begin[3];
        d.trace.push_back(lr);
        slot += OSW_T16_CLASSES; *launches += 1;
        *padded_cells += tplan->padded_cells;
        CK(cudaEventRecord(d.ev[1], d.st));
    } else if (use_u16 && N) {
        uint64_t bound_col0 = 0, win_col0 = 0;
        const uint8_t *win_ptr = nullptr;             // streaming mode: the device window holding the current segment
        // pair-database passes walk the pair directory and the pair stream, the others the plain ones
        const bool pd = !passes.empty() && passes[0].pair_db;
        const osw_chunk *dir = pd ? s.pair_chunks : s.chunks;
        const uint32_t n_dir = pd ? s.n_pair_chunks : s.n_chunks;
        auto chunk_cols = [&](uint32_t k) -> uint64_t { return pd ? dir[k].n_pair_cols : dir[k].n_cols; };
        auto chunk_begin = [&](uint32_t k) -> uint64_t { return pd ? dir[k].pair_off : dir[k].stream_off; };
        auto chunk_end = [&](uint32_t k) -> uint64_t {
            const uint64_t al = pd ? 64 : OSW_CHUNK_ALIGN;
            return chunk_begin(k) + (chunk_cols(k) + al - 1) / al * al;
        };
        constexpr uint32_t PIPE_MAX_PASSES = 64, PIPE_MAX_CHUNKS = 16;
        uint32_t reserved_ctas = 0;           // SMs left to the pipelined launches of the longest chunks
        auto launch = [&](const OswPass &ps, uint32_t first, uint32_t end) -> int {
            if (slot >= MAX_LAUNCH_SLOTS) { snprintf(g_err, sizeof g_err, "too many launches"); return OSW_E_ARG; }
            U16Params up;
            up.stream = win_ptr ? win_ptr : d.d_stream; up.pair_stream = win_ptr ? win_ptr : d.d_pair; up.stream_col0 = win_col0;
            up.chunks = pd ? d.d_pair_chunks : d.d_chunks;
            up.chunk_first = first; up.chunk_end = end; up.static_first = 0;
            uint64_t cols = 0;
            if (first == 0 && end == n_dir) cols = pd ? s.pair_cols : s.n_residues;
            else for (uint32_t k = first; k < end; ++k) cols += chunk_cols(k);
            // When walking the longest chunk among three other busy warps would take clearly longer
            // than the whole launch needs for its cell updates, the longest chunks get express CTAs:
            // one warp per scheduler, so that they advance at the latency of a step rather than at a
            // quarter of the scheduler's issue rate.  (Threshold 1.5 measured: at a ratio of 1.15 the
            // express CTAs cost 4 %, at 1.9 they gain 6 %, at 3.5 and above 20-40 %.)
            up.express_ctas = 0; up.all_express = 0; up.progress_in = nullptr; up.progress_out = nullptr;
            uint32_t x_chunks = 0, x_ctas = 0;
            bool x_t16 = false;               // the long-chunk launch in the transposed form
            OswT16Plan xplan;
            uint32_t x_slot = 0;
            if (end > first && c->tune.express && !reserved_ctas) {
                const LaunchModel m(ps.G, ps.R, pd, (double)cols, d.n_sms, (double)(end - first));
                const double longest = (double)chunk_cols(first);
                if (longest * m.contended > c->tune.express_ratio * m.t_pipe) {
                    LongT16 lt;
                    if (t16_long && passes.size() == 1 && plan_long_t16(c, d, ps, first, end, cols, q_off, nq, &lt)) {
                        x_t16 = true; x_chunks = lt.n_chunks; x_ctas = lt.n_ctas; xplan = lt.plan;
                    }
                    if (!x_t16 && pass_x && !d.streaming && passes.size() == 1 && (pass_x->G != ps.G || pass_x->R < ps.R) && c->tune.long_chunks != 0) {
                        // Long-chunk launch: a step takes longer the more rows a lane holds (140 + 7 R cycles
                        // for a warp alone on its scheduler), so the chunks whose walk would outlast the
                        // launch go to a SECOND, concurrent launch of the same kernel with the widest array
                        // (32 lanes x 8..16 rows) on a few SMs of their own, every CTA an express CTA, while
                        // the geometry that suits the bulk (fewest padded rows) runs on the other SMs.
                        // It takes the chunks that could not be walked among busy neighbours within the
                        // launch's target time, on as many SMs as finish them in that time - unless that
                        // needs more than 32 (a tiny database, where nearly every chunk is "long": measured
                        // slower than the express CTAs).  Measured with it: 100 000 / 200 000 sequences x one
                        // 144-residue query + 12 % / + 6 %, two short queries x 10 000 sequences + 4 %.
                        const LaunchModel full(ps.G, ps.R, pd, (double)cols, d.n_sms);
                        const LaunchModel mx(32, pass_x->R, pd, 0.0, d.n_sms);
                        const double target = std::max(full.t_pipe, longest * mx.alone);
                        double x_cols = 0;
                        uint32_t n = 0;
                        while (first + n < end && n < 4096u && (double)chunk_cols(first + n) * full.contended > target) { x_cols += (double)chunk_cols(first + n); ++n; }
                        const double ctas = std::ceil(x_cols * mx.alone / target / 4.0);
                        if (n && ctas <= 32.0) { x_chunks = n; x_ctas = (uint32_t)std::max(1.0, std::min(ctas, std::ceil(n / 4.0))); }
                        if (c->tune.long_chunks > 0) {          // experiments: force the number of chunks of the long-chunk launch
                            x_chunks = std::min<uint32_t>((uint32_t)c->tune.long_chunks, end - first);
                            x_ctas = std::max<uint32_t>(1, std::min<uint32_t>(32, (x_chunks + 15) / 16));
                        }
                    }
                    if (!x_chunks) {
                        // Express CTAs inside the launch: the longest chunk groups get a scheduler each
                        // (threshold 1.5 measured: at a ratio of 1.15 they cost 4 %, at 1.9 they gain 6 %, at
                        // 3.5 and above 20-40 %).
                        const double cut = longest * m.alone / m.contended;      // shorter chunks finish in time anyway
                        uint32_t n = 0;
                        while (first + n < end && n < 64u * m.groups && (double)chunk_cols(first + n) > cut) ++n;
                        up.express_ctas = std::min<uint32_t>(16, std::max<uint32_t>(1, (n + 4 * m.groups - 1) / (4 * m.groups)));
                    }
                }
            }
            up.queries = d.d_queries; up.q_off = d.d_qoff; up.matrix = d.d_matrix;
            up.profile = d.d_profile; up.first_table = d.d_first_table; up.dyn_base = 0;
            up.scores = d.d_scores; up.n_seqs = N;
            up.bound = d.d_bound; up.bound_col0 = bound_col0;
            up.gap_open_extend = go + ge; up.gap_extend = ge;
            if (x_chunks) {
                if (slot + 1 >= MAX_LAUNCH_SLOTS) { snprintf(g_err, sizeof g_err, "too many launches"); return OSW_E_ARG; }
                // fork: the long-chunk launch starts when everything enqueued so far (uploads, clears) is done
                CK(cudaEventRecord(d.ev_ready[0], d.st));
                CK(cudaStreamWaitEvent(d.st_copy, d.ev_ready[0], 0));
                int rcx;
                if (x_t16) {
                    if (slot + OSW_T16_CLASSES >= MAX_LAUNCH_SLOTS) { snprintf(g_err, sizeof g_err, "too many launches"); return OSW_E_ARG; }
                    OswT16Params tp;
                    tp.stream = d.d_stream; tp.seq_off = d.d_seq_off; tp.seq_len = d.d_seq_len;
                    tp.queries = d.d_queries; tp.q_off = d.d_qoff; tp.nq = nq; tp.matrix = d.d_matrix;
                    tp.scores = d.d_scores; tp.n_seqs = N;
                    tp.counters = d.d_counters + 1 + slot; tp.cycle_acc = d.d_cycles + slot;
                    tp.gap_open_extend = go + ge; tp.gap_extend = ge;
                    x_slot = slot;
                    slot += OSW_T16_CLASSES;
                    rcx = osw_launch_t16(tp, xplan, (int)x_ctas, d.st_copy);
                    *launches += 1;
                    *padded_cells += xplan.padded_cells;
                } else {
                    U16Params xp = up;
                    xp.profile = d.d_profile_x; xp.first_table = d.d_first_table_x; xp.all_express = 1;
                    xp.chunk_first = first; xp.chunk_end = first + x_chunks;
                    xp.chunk_counter = d.d_counters + 1 + slot; xp.cycle_acc = nullptr;
                    ++slot;
                    rcx = osw_launch_u16(xp, *pass_x, (int)x_ctas, d.st_copy);
                    *launches += 2;
                }
                if (rcx != OSW_OK) { cuda_fail(cudaGetLastError(), "long-chunk launch", __LINE__); return rcx; }
                CK(cudaEventRecord(d.ev_free[0], d.st_copy));
                uint64_t xc = 0;
                for (uint32_t k = 0; k < x_chunks; ++k) xc += chunk_cols(first + k);
                if (!x_t16) *padded_cells += (uint64_t)32 * pass_x->R * 2 * xc;
                up.chunk_first = first + x_chunks;
                cols -= xc;
            }
            up.chunk_counter = d.d_counters + 1 + slot;
            up.cycle_acc = d.d_cycles + slot;
            // (beside a long-chunk launch the scoring launch leaves it its SMs: both are resident at once,
            // whichever the hardware places first - a full-width grid would make the other one wait)
            int rc2 = osw_launch_u16(up, ps, d.n_sms - (int)x_ctas - (int)reserved_ctas, d.st);
            if (rc2 != OSW_OK) { cuda_fail(cudaGetLastError(), "sw_u16 launch", __LINE__); return rc2; }
            if (x_chunks) CK(cudaStreamWaitEvent(d.st, d.ev_free[0], 0));          // join
            d.trace.push_back({ps.G, ps.R, ps.has_in, ps.has_out, ps.pair_db, first, end, cols, up.express_ctas, slot, x_chunks, x_ctas, x_chunks ? (x_t16 ? -1 : pass_x->R) : 0, 0u, 0, x_slot});
            ++slot; *launches += 2;               // profile_build_kernel + sw_u16_kernel
            *padded_cells += (uint64_t)ps.G * ps.R * 2 * cols;
            return OSW_OK;
        };
        {
            // One launch per pass, in order: a pass reads the bottom row the previous one parked (in
            // place: a warp writes a chunk's columns behind the ones it still has to read).  With
            // several passes the chunk list is cut into segments whose bottom rows fit the buffer;
            // chunks are stored back to back in ascending order, so a range of the (descending)
            // directory is one contiguous stretch of the stream.
            const uint64_t bpc = pd ? 2 : 1;                               // bytes per column of the stream in use
            const uint8_t *h_src = pd ? d.h_pair : d.h_stream;
            const uint64_t win_cols = d.streaming ? d.win_bytes / bpc : ~0ull;
            const bool bounded = d.d_bound && passes.size() > 1;
            uint32_t seg_first = 0, seg_no = 0;
            while (seg_first < n_dir) {
                uint32_t seg_end = n_dir;
                if (bounded || d.streaming) {
                    const uint64_t cap = std::min<uint64_t>(bounded ? d.bound_cap : ~0ull, win_cols);
                    seg_end = seg_first + 1;
                    while (seg_end < n_dir && chunk_end(seg_first) - chunk_begin(seg_end) <= cap) ++seg_end;
                    const uint64_t seg_col0 = chunk_begin(seg_end - 1);
                    if (chunk_end(seg_first) - seg_col0 > cap) {
                        snprintf(g_err, sizeof g_err, "a chunk is larger than the device window / bottom-row buffer");
                        return OSW_E_NOMEM;
                    }
                    bound_col0 = seg_col0;
                    if (d.streaming) {
                        // copy the segment into the free window while the previous segment is being scored
                        const int w = (int)(seg_no & 1);
                        CK(cudaStreamWaitEvent(d.st_copy, d.ev_free[w], 0));
                        CK(cudaMemcpyAsync(d.d_win[w], h_src + seg_col0 * bpc, (chunk_end(seg_first) - seg_col0) * bpc,
                                           cudaMemcpyHostToDevice, d.st_copy));
                        CK(cudaEventRecord(d.ev_ready[w], d.st_copy));
                        CK(cudaStreamWaitEvent(d.st, d.ev_ready[w], 0));
                        win_ptr = d.d_win[w]; win_col0 = seg_col0;
                    }
                }
                // Pipelined passes over the LONGEST chunks.  Every pass walks a chunk column by column, so a
                // titin-length sequence costs n_passes sequential walks (BASELINE.json config 5 on eight
                // GPUs: 17 x 14 ms of a 177 ms search were 337 ms).  But pass p + 1 needs, at column j, only
                // what pass p has produced up to column j: the walks can run at the same time, a few dozen
                // columns apart, each pass on an SM of its own with its own profile table, handing the
                // bottom row over through the same in-place buffer under per-chunk progress counters.
                uint32_t n_pipe = 0;
                if (seg_first == 0 && seg_end == n_dir && bounded && !d.streaming && c->tune.express && c->tune.pipe_chunks != 0 &&
                    passes.size() >= 2 && passes.size() <= PIPE_MAX_PASSES) {
                    double t_total = 0, step_sum = 0;
                    bool all32 = true;
                    for (const OswPass &ps : passes) {
                        const LaunchModel m(ps.G, ps.R, pd, (double)(pd ? s.pair_cols : s.n_residues), d.n_sms);
                        t_total += m.t_pipe; step_sum += m.contended; all32 &= ps.G == 32;
                    }
                    while (all32 && n_pipe < PIPE_MAX_CHUNKS && n_pipe < n_dir && (double)chunk_cols(n_pipe) * step_sum > 1.5 * t_total) ++n_pipe;
                    if (c->tune.pipe_chunks > 0 && all32) n_pipe = std::min<uint32_t>(std::min<uint32_t>((uint32_t)c->tune.pipe_chunks, PIPE_MAX_CHUNKS), n_dir);
                    const uint32_t ctas_per_pass = (n_pipe + 3) / 4;
                    if (ctas_per_pass * passes.size() > (size_t)d.n_sms / 2) n_pipe = 0;
                    if (n_pipe) {
                        const size_t np = passes.size();
                        if (!d.d_profile_pipe) {
                            CK(cudaMalloc(&d.d_profile_pipe, (size_t)PIPE_MAX_PASSES * OSW_PROFILE_BYTES));
                            CK(cudaMalloc(&d.d_first_table_pipe, (size_t)PIPE_MAX_PASSES * 1024 * sizeof(uint32_t)));
                            CK(cudaMalloc(&d.d_progress, (size_t)PIPE_MAX_PASSES * PIPE_MAX_CHUNKS * sizeof(uint32_t)));
                        }
                        while (d.pipe_st.size() < np) {
                            cudaStream_t st2; cudaEvent_t ev2;
                            CK(cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking));
                            d.pipe_st.push_back(st2);
                            CK(cudaEventCreateWithFlags(&ev2, cudaEventDisableTiming));
                            d.pipe_ev.push_back(ev2);
                        }
                        CK(cudaMemsetAsync(d.d_progress, 0, (size_t)PIPE_MAX_PASSES * PIPE_MAX_CHUNKS * sizeof(uint32_t), d.st));
                        CK(cudaEventRecord(d.ev_ready[0], d.st));                  // fork
                        uint64_t pipe_cols = 0;
                        for (uint32_t k = 0; k < n_pipe; ++k) pipe_cols += chunk_cols(k);
                        for (size_t pi = 0; pi < np; ++pi) {
                            if (slot >= MAX_LAUNCH_SLOTS) { snprintf(g_err, sizeof g_err, "too many launches"); return OSW_E_ARG; }
                            const OswPass &ps = passes[pi];
                            CK(cudaStreamWaitEvent(d.pipe_st[pi], d.ev_ready[0], 0));
                            U16Params xp;
                            xp.stream = d.d_stream; xp.pair_stream = d.d_pair; xp.stream_col0 = 0;
                            xp.chunks = pd ? d.d_pair_chunks : d.d_chunks;
                            xp.chunk_first = 0; xp.chunk_end = n_pipe; xp.static_first = 0;
                            xp.express_ctas = 0; xp.all_express = 1;
                            xp.queries = d.d_queries; xp.q_off = d.d_qoff; xp.matrix = d.d_matrix;
                            xp.profile = d.d_profile_pipe + pi * OSW_PROFILE_BYTES; xp.first_table = d.d_first_table_pipe + pi * 1024; xp.dyn_base = 0;
                            xp.scores = d.d_scores; xp.n_seqs = N;
                            xp.bound = d.d_bound; xp.bound_col0 = bound_col0;
                            xp.gap_open_extend = go + ge; xp.gap_extend = ge;
                            xp.chunk_counter = d.d_counters + 1 + slot; xp.cycle_acc = nullptr;
                            xp.progress_in = pi ? d.d_progress + (pi - 1) * PIPE_MAX_CHUNKS : nullptr;
                            xp.progress_out = pi + 1 < np ? d.d_progress + pi * PIPE_MAX_CHUNKS : nullptr;
                            ++slot;
                            int rcx = osw_launch_u16(xp, ps, (int)ctas_per_pass, d.pipe_st[pi]);
                            if (rcx != OSW_OK) { cuda_fail(cudaGetLastError(), "pipelined launch", __LINE__); return rcx; }
                            CK(cudaEventRecord(d.pipe_ev[pi], d.pipe_st[pi]));
                            *launches += 2;
                            *padded_cells += (uint64_t)ps.G * ps.R * 2 * pipe_cols;
                            d.trace.push_back({ps.G, ps.R, ps.has_in, ps.has_out, ps.pair_db, 0u, n_pipe, pipe_cols, ctas_per_pass, slot - 1, 0u, 0u, 0, n_pipe});
                        }
                        reserved_ctas = ctas_per_pass * (uint32_t)np;
                    }
                }
                for (const OswPass &ps : passes) {
                    if ((rc = launch(ps, seg_first + n_pipe, seg_end)) != OSW_OK) return rc;
                    // (only the first pass leaves the pipelined launches their SMs, so that they are resident
                    // at once; they are done after a pass or two, and the later passes run at full width - a
                    // CTA that finds no free SM yet simply starts when a pipelined one has left)
                    reserved_ctas = 0;
                }
                if (n_pipe) {
                    for (size_t pi = 0; pi < passes.size(); ++pi) CK(cudaStreamWaitEvent(d.st, d.pipe_ev[pi], 0));      // join
                    reserved_ctas = 0;
                }
                if (d.streaming) CK(cudaEventRecord(d.ev_free[seg_no & 1], d.st));
                seg_first = seg_end; ++seg_no;
            }
        }
        CK(cudaEventRecord(d.ev[1], d.st));
    } else {
        CK(cudaEventRecord(d.ev[1], d.st));
    }
    *n_launch_slots = slot;
    return OSW_OK;
}

// 32-bit kernel parameters for this GPU's shard.
I32Params i32_params(const DevState &d, int go, int ge) {
    const osw_shard &s = d.shard;
    I32Params ip;
    ip.stream = d.d_stream; ip.seq_off = d.d_seq_off; ip.seq_len = d.d_seq_len; ip.task_off = nullptr;
    ip.queries = d.d_queries; ip.q_off = d.d_qoff; ip.matrix = d.d_matrix;
    ip.pairs = nullptr; ip.n_tasks = 0; ip.n_tasks_dev = nullptr;
    ip.n_seqs = s.n_seqs; ip.scores = d.d_scores; ip.scratch = d.d_scratch; ip.max_len = s.max_len ? s.max_len : 1;
    ip.gap_open_extend = go + ge; ip.gap_extend = ge; ip.task_counter = d.d_task_counter;
    return ip;
}

// Per-warp scratch of the 32-bit kernel (a pass's bottom row) for a grid of `blocks` blocks.
int i32_scratch(DevState &d, int blocks) {
    const size_t warps = (size_t)blocks * (osw_i32_block_threads() / 32);
    return grow(&d.d_scratch, &d.scratch_cap, warps * (d.shard.max_len ? d.shard.max_len : 1));
}

// Top-r selection and the copies back to the host (keys, flagged count, cycle counters, all scores
// on request), enqueued behind whatever has been enqueued on the GPU's stream.
int enqueue_finish(osw_ctx *c, DevState &d, int nq, uint32_t top_r, bool want_all, bool list_flags, uint32_t slots,
                   cudaEvent_t ev_before, cudaEvent_t ev_after, uint64_t *launches) {
    const uint64_t N = d.shard.n_seqs;
    CK(cudaEventRecord(ev_before, d.st));
    const uint32_t r = (uint32_t)std::min<uint64_t>((uint64_t)top_r, N);
    // the first scan of the selection also lists the pairs the 16-bit stage flagged (d_counters[0] counts them)
    const FlagList fl = {d.d_pairs, d.d_counters, d.pairs_cap};
    if (N) *launches += osw_topr_select(d.d_scores, d.d_canon, N, std::max<uint64_t>(c->n_seqs_canon, 1), nq, r, d.topr, list_flags ? &fl : nullptr, d.st);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ev_after, d.st));
    if (list_flags) CK(cudaMemcpyAsync(d.h_counts, d.d_counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, d.st));
    if (r) CK(cudaMemcpyAsync(d.h_keys, d.topr.out_keys, (size_t)nq * r * sizeof(unsigned long long), cudaMemcpyDeviceToHost, d.st));
    if (want_all && N) CK(cudaMemcpyAsync(d.h_scores, d.d_scores, (size_t)nq * N * sizeof(int32_t), cudaMemcpyDeviceToHost, d.st));
    if (slots) CK(cudaMemcpyAsync(d.h_cycles, d.d_cycles, slots * sizeof(unsigned long long), cudaMemcpyDeviceToHost, d.st));
    return OSW_OK;
}

}  // namespace

static int search_batch(osw_ctx *c, const uint8_t *queries, const uint32_t *q_off, int nq,
                        const int8_t *matrix, int go, int ge, int top_r,
                        osw_hit *hits, uint32_t *n_hits, int32_t *all_scores, osw_timing *timing);

extern "C" int osw_search(osw_ctx *c, const uint8_t *queries, const uint32_t *q_off, int nq,
                          const int8_t *matrix, int go, int ge, int top_r,
                          osw_hit *hits, uint32_t *n_hits, int32_t *all_scores, osw_timing *timing) {
    if (!c || !queries || !q_off || nq < 1 || !matrix || go < 0 || ge < 0 || go > 255 || ge > 127 || top_r < 0)
        return OSW_E_ARG;
    if (top_r > 0 && !hits) return OSW_E_ARG;
    if (!c->db_loaded) return OSW_E_STATE;
    for (int q = 0; q < nq; ++q)
        if (q_off[q + 1] < q_off[q] || q_off[q + 1] - q_off[q] > OSW_MAX_QUERY_LEN) return OSW_E_ARG;
    for (uint32_t k = q_off[0]; k < q_off[nq]; ++k)
        if (queries[k] > 23) { snprintf(g_err, sizeof g_err, "query residue code %u is outside 0..23", queries[k]); return OSW_E_ARG; }
    // the 16-bit kernel's bias and wrap detection assume |score| <= 31 (the reference's tables span -17..17)
    for (int k = 0; k < 24 * 32; ++k)
        if (matrix[k] < -32 || matrix[k] > 31) { snprintf(g_err, sizeof g_err, "substitution score %d is outside -32..31", matrix[k]); return OSW_E_ARG; }
    // The score matrix (4 bytes per query and sequence) is the largest per-search buffer: many
    // queries go through in batches that keep it under 8 GiB per GPU (OSW_SCORE_BUDGET_KB
    // overrides).  A batch also stays within what one plan can hold: 16 384 queries (grid limit of
    // the per-query kernels) and a million query rows (a few hundred passes, far below MAX_PASSES
    // and, times the segments of a streamed database, MAX_LAUNCH_SLOTS).
    uint64_t n_max = 1;
    for (int i = 0; i < c->n_dev; ++i) n_max = std::max<uint64_t>(n_max, c->devs[i].shard.n_seqs);
    const uint64_t budget = c->tune.score_budget;
    const int batch = (int)std::max<uint64_t>(2, std::min<uint64_t>(std::min<uint64_t>((uint64_t)nq, 16384), budget / (4 * n_max)));
    const uint64_t max_rows = 1u << 20;
    osw_timing total;
    memset(&total, 0, sizeof total);
    for (int q0 = 0, nb = 0; q0 < nq; q0 += nb) {
        nb = 0;
        uint64_t rows = 0;
        while (q0 + nb < nq && nb < batch && (nb == 0 || rows + (q_off[q0 + nb + 1] - q_off[q0 + nb]) <= max_rows)) {
            rows += q_off[q0 + nb + 1] - q_off[q0 + nb];
            ++nb;
        }
        osw_timing tm;
        int rc = search_batch(c, queries, q_off + q0, nb, matrix, go, ge, top_r, hits ? hits + (size_t)q0 * top_r : nullptr,
                              n_hits ? n_hits + q0 : nullptr, all_scores ? all_scores + (size_t)q0 * c->n_seqs_canon : nullptr, &tm);
        if (rc != OSW_OK) return rc;
        total.device_ms += tm.device_ms; total.score_ms += tm.score_ms; total.rescore_ms += tm.rescore_ms; total.topr_ms += tm.topr_ms;
        total.h2d_ms += tm.h2d_ms; total.wall_ms += tm.wall_ms; total.cells += tm.cells; total.padded_cells += tm.padded_cells;
        total.rescored_pairs += tm.rescored_pairs; total.launches += tm.launches; total.sm_cycles += tm.sm_cycles;
        total.db_stream_bytes += tm.db_stream_bytes; total.bound_bytes += tm.bound_bytes; total.score_launches += tm.score_launches;
    }
    if (timing) *timing = total;
    return OSW_OK;
}

// One batch of queries: q_off points at the batch's first entry (offsets stay absolute).
static int search_batch(osw_ctx *c, const uint8_t *queries, const uint32_t *q_off, int nq,
                        const int8_t *matrix, int go, int ge, int top_r,
                        osw_hit *hits, uint32_t *n_hits, int32_t *all_scores, osw_timing *timing) {
    const double t_wall0 = now_ms();
    std::vector<OswPass> passes;
    OswPass plan_x;
    bool have_x = false;
    if (c->kernel_mask & OSW_K_U16) {
        std::vector<uint32_t> q_len((size_t)nq);
        for (int q = 0; q < nq; ++q) q_len[q] = q_off[q + 1] - q_off[q];
        const int mode = (c->kernel_mask & OSW_K_PAIR_DB) ? OSW_PLAN_PAIR_DB : (c->kernel_mask & OSW_K_TWO_TRACK) ? OSW_PLAN_TWO_TRACK : OSW_PLAN_AUTO;
        // (a pass descriptor is a kilobyte: room for 32 first - nearly every search - and for MAX_PASSES only when needed)
        passes.resize(32);
        int n_pass = osw_plan_passes_ex(q_len.data(), nq, passes.data(), 32, mode, 4, c->tune.rmax);
        if (n_pass < 0) {
            passes.resize(MAX_PASSES);
            n_pass = osw_plan_passes_ex(q_len.data(), nq, passes.data(), MAX_PASSES, mode, 4, c->tune.rmax);
        }
        if (n_pass < 0) { snprintf(g_err, sizeof g_err, "the queries need more than %d passes", MAX_PASSES); return OSW_E_ARG; }
        // A single-pass plan that also fits 32 lanes x <= 16 rows can hand its longest chunks to a
        // launch of that geometry (same halves, same directory; the launcher decides per launch).
        if (n_pass == 1) {
            OswPass alt[2];
            if (osw_plan_passes_ex(q_len.data(), nq, alt, 2, passes[0].pair_db ? OSW_PLAN_PAIR_DB : OSW_PLAN_TWO_TRACK, 32, 16) == 1 &&
                alt[0].R <= 16 && alt[0].pair_db == passes[0].pair_db) { plan_x = alt[0]; have_x = true; }
        }
        if (n_pass == 1 && passes[0].G < 32) {
            // The planner picked the geometry with the fewest padded rows.  That is the fastest one
            // when the launch is bound by the DPX pipe; on a small database the launch lasts as long
            // as its longest chunk (one dependent step per column, and a step takes longer the more
            // rows a lane holds), and a wider array with fewer rows per lane finishes sooner despite
            // its padding.  Estimate both terms for every group width and keep the best plan.
            const DevState &d0 = c->devs[0];
            const osw_shard &s0 = d0.shard;
            const bool pd = passes[0].pair_db != 0;
            const double cols = (double)(pd ? s0.pair_cols : s0.n_residues);
            const double longest = !(pd ? s0.n_pair_chunks : s0.n_chunks) ? 0.0 : (double)(pd ? s0.pair_chunks[0].n_pair_cols : s0.chunks[0].n_cols);
            auto estimate = [&](const OswPass &ps) {
                const LaunchModel m(ps.G, ps.R, pd, cols, d0.n_sms);
                const double t_chain = longest * m.alone;               // (the longest chunks get express CTAs)
                return std::max(m.t_pipe, t_chain) + 0.25 * std::min(m.t_pipe, t_chain);
            };
            double best = estimate(passes[0]);
            const int force_g = c->tune.force_g;                 // experiments
            std::vector<OswPass> alt(2);
            for (int g = passes[0].G * 2; g <= 32; g *= 2) {
                if (osw_plan_passes_ex(q_len.data(), nq, alt.data(), 2, pd ? OSW_PLAN_PAIR_DB : OSW_PLAN_TWO_TRACK, g, c->tune.rmax) != 1) continue;
                const double t = estimate(alt[0]);
                if (force_g ? g == force_g : t < best) { best = t; passes[0] = alt[0]; }
            }
        }
        passes.resize((size_t)n_pass);
    }
    const bool use_u16 = (c->kernel_mask & OSW_K_U16) != 0;
    // The transposed form (sw_t16.cu: database residues as rows, the query as the column stream) takes
    // short queries whenever its model says it is faster than the passes planned above - small
    // databases, whose launches last as long as the walk along their longest sequence, above all.
    std::vector<OswT16Plan> tplans;
    double t16_est = 0;
    bool t16_long = false;                // the passes' long-chunk launch may take the transposed form
    if (use_u16 && c->tune.transpose != 0 && !(c->kernel_mask & (OSW_K_TWO_TRACK | OSW_K_PAIR_DB))) {
        bool ok = true;
        for (int i = 0; i < c->n_dev; ++i) ok &= !c->devs[i].streaming && (c->devs[i].shard.n_seqs == 0 || c->devs[i].d_stream != nullptr);
        if (ok) {
            tplans.resize((size_t)c->n_dev);
            for (int i = 0; i < c->n_dev && ok; ++i) {
                DevState &d = c->devs[i];
                if (d.t16_hist.empty()) {
                    d.t16_hist.resize(OSW_T16_MAX_ROWS32 + 1);
                    osw_t16_histogram(d.shard.seq_len, d.shard.n_seqs, d.t16_hist.data());
                }
                osw_t16_plan(d.t16_hist.data(), d.shard.n_seqs, q_off, nq, d.n_sms, c->tune.t16_gang, &tplans[(size_t)i]);
                ok &= d.shard.n_seqs == 0 || tplans[(size_t)i].warps != 0;
            }
        }
        const bool forced = (c->kernel_mask & OSW_K_TRANSPOSED) || c->tune.transpose == 1;
        t16_long = ok && !forced && tplans[0].warps == 16 && passes.size() == 1;
        if (ok && !forced && c->tune.transpose == 2) ok = false;          // (experiments: the transposed form for long chunks only)
        if (ok && !forced) {
            // the passes' own estimate, on the first GPU's shard (all shards hold the same mix of lengths)
            const DevState &d0 = c->devs[0];
            const osw_shard &s0 = d0.shard;
            const bool pd = !passes.empty() && passes[0].pair_db != 0;
            const double cols = (double)(pd ? s0.pair_cols : s0.n_residues);
            const double longest = !(pd ? s0.n_pair_chunks : s0.n_chunks) ? 0.0 : (double)(pd ? s0.pair_chunks[0].n_pair_cols : s0.chunks[0].n_cols);
            double t_std = 0;
            for (const OswPass &ps : passes) {
                const LaunchModel m(ps.G, ps.R, pd, cols, d0.n_sms);
                const double t_chain = longest * m.alone;
                t_std += std::max(m.t_pipe, t_chain) + 0.25 * std::min(m.t_pipe, t_chain) + 6000.0;      // (+ a launch and its profile build)
            }
            // (the passes' model is optimistic on databases of this size: measured / model is 1.25-1.65, against
            // 0.95-1.1 for the transposed form's; 8-warp CTAs - queries of more than about 250 residues - have two
            // warps per scheduler and are slower than the passes everywhere: profiles/r2_transposed_form.md)
            ok = s0.n_seqs != 0 && tplans[0].warps == 16 && tplans[0].est_cycles < 1.4 * t_std;
            // ... and against the passes with their long chunks in the transposed form, when that applies (measured
            // / model there: 1.4 on 50 000-100 000 sequences, 1.8 on 10 000)
            LongT16 lt;
            if (ok && t16_long && plan_long_t16(c, d0, passes[0], 0, pd ? s0.n_pair_chunks : s0.n_chunks, pd ? s0.pair_cols : s0.n_residues, q_off, nq, &lt))
                ok = tplans[0].est_cycles < 1.3 * lt.est;
        }
        if (!ok) tplans.clear();
        else t16_est = tplans[0].est_cycles;
    }
    uint64_t launches = 0, padded = 0, rescored = 0;
    std::vector<uint32_t> slots((size_t)c->n_dev, 0u);

    // ---- phase 1: first stage on every GPU ------------------------------------------------
    for (int i = 0; i < c->n_dev; ++i) {
        int rc = enqueue_search(c, c->devs[i], queries, q_off, nq, matrix, go, ge, (uint32_t)top_r,
                                all_scores != nullptr, passes, have_x ? &plan_x : nullptr, tplans.empty() ? nullptr : &tplans[(size_t)i], t16_long && tplans.empty(),
                                &slots[i], &launches, &padded);
        if (rc != OSW_OK) return rc;
    }
    const double t_h2d = now_ms();
    // ---- phase 2: top-r, enqueued WITHOUT waiting for the first stage --------------------------
    // The first scan of the selection also lists the pairs whose 16-bit score may have wrapped.
    // Nearly every search has none, and then it was one uninterrupted sequence of launches per GPU
    // (a host round trip between scoring and top-r cost 15-40 ms per search under torchrun); when
    // there are some, the selection is void: phase 3 re-scores them at 32 bit and selects again.
    const int i32_blocks_rescore = 2, i32_blocks_all = 8;          // blocks per SM
    for (int i = 0; i < c->n_dev; ++i) {
        DevState &d = c->devs[i];
        const uint64_t N = d.shard.n_seqs;
        CK(cudaSetDevice(d.dev));
        if (!use_u16 && N) {
            if (d.streaming) { snprintf(g_err, sizeof g_err, "the 32-bit-only mode needs a resident database"); return OSW_E_STATE; }
            int rc2 = i32_scratch(d, d.n_sms * i32_blocks_all);
            if (rc2 != OSW_OK) return rc2;
            I32Params ip = i32_params(d, go, ge);
            ip.n_tasks = (uint64_t)nq * N;
            rescored += ip.n_tasks;
            osw_launch_i32(ip, d.n_sms * i32_blocks_all, d.st);
            ++launches;
        }
        CK(cudaGetLastError());
        int rc2 = enqueue_finish(c, d, nq, (uint32_t)top_r, all_scores != nullptr, use_u16, slots[i], d.ev[2], d.ev[3], &launches);
        if (rc2 != OSW_OK) return rc2;
    }
    // ---- phase 3: wait; rare slow paths; order, merge ------------------------------------------
    osw_timing tm;
    memset(&tm, 0, sizeof tm);
    std::vector<std::vector<osw_hit>> per_dev((size_t)c->n_dev);
    for (int i = 0; i < c->n_dev; ++i) {
        DevState &d = c->devs[i];
        const osw_shard &s = d.shard;
        const uint64_t N = s.n_seqs;
        CK(cudaSetDevice(d.dev));
        CK(cudaStreamSynchronize(d.st));
        cudaEvent_t ev_end = d.ev[3];
        uint32_t n_flag = use_u16 && N ? d.h_counts[0] : 0u;
        // The 32-bit stage (rare): re-score the listed pairs - which replaces their FLAGGED marker by
        // the exact score - and, when more pairs were flagged than the list holds, scan the score
        // matrix for the rest and repeat: like the reference, which simply recomputes whatever
        // saturated (HybridSearch.c:1032-1134), this never fails on data.  With a streamed database
        // the flagged sequences are gathered from the pinned host copy into a staging buffer.  Then
        // the selection is done again.
        if (n_flag) {
            bool listed = true;               // d_pairs holds min(n_flag, capacity) pairs still to be re-scored
            for (;;) {
                if (!listed) {
                    CK(cudaMemsetAsync(d.d_counters, 0, sizeof(uint32_t), d.st));
                    launches += osw_collect_flagged(d.d_scores, N, nq, d.d_pairs, d.d_counters, d.pairs_cap, d.st);
                    CK(cudaMemcpyAsync(d.h_counts, d.d_counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, d.st));
                    CK(cudaStreamSynchronize(d.st));
                    n_flag = d.h_counts[0];
                    if (!n_flag) break;
                }
                listed = false;
                const uint32_t n_now = std::min(n_flag, d.pairs_cap);
                rescored += n_now;
                int rc2 = i32_scratch(d, d.n_sms * i32_blocks_rescore);
                if (rc2 != OSW_OK) return rc2;
                I32Params ip = i32_params(d, go, ge);
                if (d.streaming) {
                    if ((rc2 = grow_pinned(&d.h_pairs, &d.h_pairs_cap, (size_t)n_now)) != OSW_OK) return rc2;
                    CK(cudaMemcpyAsync(d.h_pairs, d.d_pairs, (size_t)n_now * sizeof(uint2), cudaMemcpyDeviceToHost, d.st));
                    CK(cudaStreamSynchronize(d.st));
                    if ((rc2 = grow_pinned(&d.h_task_off, &d.h_task_off_cap, (size_t)n_now)) != OSW_OK) return rc2;
                    uint64_t total = 0;
                    for (uint32_t k = 0; k < n_now; ++k) { d.h_task_off[k] = total; total += s.seq_len[d.h_pairs[k].y]; }
                    if ((rc2 = grow_pinned(&d.h_stage, &d.h_stage_cap, (size_t)(total ? total : 1))) != OSW_OK) return rc2;
                    for (uint32_t k = 0; k < n_now; ++k)
                        memcpy(d.h_stage + d.h_task_off[k], d.h_stream + s.seq_off[d.h_pairs[k].y], s.seq_len[d.h_pairs[k].y]);
                    if ((rc2 = grow(&d.d_stage, &d.stage_cap, (size_t)(total ? total : 1))) != OSW_OK) return rc2;
                    if ((rc2 = grow(&d.d_task_off, &d.task_off_cap, (size_t)n_now)) != OSW_OK) return rc2;
                    CK(cudaMemcpyAsync(d.d_stage, d.h_stage, (size_t)total, cudaMemcpyHostToDevice, d.st));
                    CK(cudaMemcpyAsync(d.d_task_off, d.h_task_off, (size_t)n_now * sizeof(uint64_t), cudaMemcpyHostToDevice, d.st));
                    ip.stream = d.d_stage; ip.task_off = d.d_task_off;
                }
                ip.pairs = d.d_pairs; ip.n_tasks = n_now;
                CK(cudaMemsetAsync(d.d_task_counter, 0, 2 * sizeof(unsigned long long), d.st));
                osw_launch_i32(ip, (int)std::min<uint64_t>((uint64_t)d.n_sms * i32_blocks_rescore, ((uint64_t)n_now + 3) / 4), d.st);
                ++launches;
                CK(cudaGetLastError());
                if (n_flag <= d.pairs_cap) break;
            }
            int rc2 = enqueue_finish(c, d, nq, (uint32_t)top_r, all_scores != nullptr, false, slots[i], d.ev[4], d.ev[5], &launches);
            if (rc2 != OSW_OK) return rc2;
            CK(cudaStreamSynchronize(d.st));
            ev_end = d.ev[5];
        }
        float ms_total = 0, ms_score = 0, ms_resc = 0, ms_top = 0;
        CK(cudaEventElapsedTime(&ms_total, d.ev[0], ev_end));
        CK(cudaEventElapsedTime(&ms_score, d.ev[0], d.ev[1]));
        CK(cudaEventElapsedTime(&ms_resc, d.ev[1], d.ev[2]));
        CK(cudaEventElapsedTime(&ms_top, d.ev[2], d.ev[3]));
        if (ev_end != d.ev[3]) {                 // slow path: everything after the first top-r counts as re-score
            float ms_extra = 0;
            CK(cudaEventElapsedTime(&ms_extra, d.ev[3], ev_end));
            ms_resc += ms_extra;
        }
        tm.device_ms = std::max(tm.device_ms, (double)ms_total);
        tm.score_ms = std::max(tm.score_ms, (double)ms_score);
        tm.rescore_ms = std::max(tm.rescore_ms, (double)ms_resc);
        tm.topr_ms = std::max(tm.topr_ms, (double)ms_top);
        if (i == 0) for (uint32_t k = 0; k < slots[i]; ++k) tm.sm_cycles += d.h_cycles[k];
        if (i == 0 && c->tune.trace) {
            // per-launch report: geometry, elapsed SM cycles, padded cell updates per SM-cycle
            for (uint32_t k = 0; k < d.trace.size(); ++k) {
                const LaunchRecord &lr = d.trace[k];
                const unsigned long long cyc = std::max<unsigned long long>(d.h_cycles[lr.slot], 1);
                if (lr.transposed) {
                    fprintf(stderr, "osw trace: launch %u/%zu transposed, %d warps per CTA, %u pairs (gang classes up to rank %u / %u / %u)  %llu busy cycles/SM  %.2f padded cells/SM-clk  model %.0f cycles\n",
                            k + 1, d.trace.size(), lr.transposed, lr.end, lr.first, lr.x_chunks, lr.x_ctas, cyc / d.n_sms, (double)padded / (double)cyc, t16_est);
                    continue;
                }
                const double cells = 2.0 * lr.G * lr.R * (double)lr.cols;
                fprintf(stderr, "osw trace: launch %u/%zu G=%d R=%d in=%d out=%d pairdb=%d express=%u long[R=%d (-1: transposed) chunks=%u ctas=%u] pipelined=%u chunks [%u,%u)  %llu busy cycles/SM  %.2f padded cells/SM-clk",
                        k + 1, d.trace.size(), lr.G, lr.R, lr.has_in, lr.has_out, lr.pair_db, lr.express, lr.x_R, lr.x_chunks, lr.x_ctas, lr.pipe_chunks, lr.first, lr.end,
                        cyc / d.n_sms, cells / (double)cyc);
                if (lr.x_R < 0 && lr.x_ctas) fprintf(stderr, "  (%llu cycles per CTA here, %llu in the transposed launch)", cyc / (d.n_sms - lr.x_ctas), d.h_cycles[lr.x_slot] / lr.x_ctas);
                fprintf(stderr, "\n");
            }
        }
        const uint32_t r = (uint32_t)std::min<uint64_t>((uint64_t)top_r, N);
        per_dev[i].resize((size_t)nq * r);
        for (int q = 0; q < nq; ++q) {
            unsigned long long *k = d.h_keys + (size_t)q * r;
            std::sort(k, k + r, [](unsigned long long x, unsigned long long y) { return x > y; });
            for (uint32_t j = 0; j < r; ++j) {
                per_dev[i][(size_t)q * r + j].score = (int32_t)(k[j] >> 32);
                per_dev[i][(size_t)q * r + j].index = (uint32_t)k[j];
            }
        }
        if (all_scores)
            for (int q = 0; q < nq; ++q)
                for (uint64_t l = 0; l < N; ++l)
                    all_scores[(size_t)q * c->n_seqs_canon + d.shard.canon[l]] = d.h_scores[(size_t)q * N + l];
    }
    if (top_r > 0) {
        std::vector<const osw_hit *> lists((size_t)c->n_dev);
        std::vector<uint32_t> counts((size_t)c->n_dev);
        for (int q = 0; q < nq; ++q) {
            for (int i = 0; i < c->n_dev; ++i) {
                const uint32_t r = (uint32_t)std::min<uint64_t>((uint64_t)top_r, c->devs[i].shard.n_seqs);
                lists[i] = per_dev[i].data() + (size_t)q * r; counts[i] = r;
            }
            size_t n = osw_merge_hits(lists.data(), counts.data(), c->n_dev, (uint32_t)top_r, hits + (size_t)q * top_r);
            if (n_hits) n_hits[q] = (uint32_t)n;
        }
    }
    tm.h2d_ms = t_h2d - t_wall0;
    tm.wall_ms = now_ms() - t_wall0;
    tm.cells = (uint64_t)(q_off[nq] - q_off[0]) * c->residues_local;
    tm.padded_cells = padded;
    tm.rescored_pairs = rescored;
    tm.launches = launches;
    tm.db_stream_bytes = 0;
    for (int i = 0; i < c->n_dev; ++i)
        for (const LaunchRecord &lr : c->devs[i].trace) {
            tm.db_stream_bytes += (lr.pair_db ? 2 : 1) * lr.cols;
            tm.bound_bytes += 8 * lr.cols * (uint64_t)((lr.has_in ? 1 : 0) + (lr.has_out ? 1 : 0));
            ++tm.score_launches;
        }
    if (timing) *timing = tm;
    return OSW_OK;
}
