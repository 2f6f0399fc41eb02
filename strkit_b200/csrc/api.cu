// api.cu -- C ABI (include/strkit_b200.h), context and batch management, launch orchestration.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/strkit_b200.h"
#include "dp_general.cuh"
#include "dp_packed.cuh"
#include "plan.cuh"
#include "dedupe.cuh"
#include "int_peak.cuh"
#include "replay.cuh"
#include "ref_path.cuh"
#include "strk_common.cuh"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int set_err(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return set_err(STRK_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,            \
                           cudaGetErrorString(e_));                                                       \
    } while (0)

extern "C" const char *strk_last_error(void) { return g_err; }
extern "C" const char *strk_version(void) { return "strkit_b200 0.1 (sm_100a)"; }

extern "C" int strk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ------------------------------------------------------------------------------------------------
// device buffer that only grows
// ------------------------------------------------------------------------------------------------
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 16;
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct strk_ctx {
    int device = 0;
    int n_sm = 0;
    size_t l2_persist_max = 0, l2_window_max = 0;  // L2 set-aside for the capture scratch (see launch_packed_r)
    const void *l2_window_ptr = nullptr;
    size_t l2_window_bytes = 0;
    cudaStream_t l2_window_stream = nullptr;
    // small passes: one side stream per rows-per-lane class, so the class launches (each under one wave) overlap
    cudaStream_t side[STRK_PK_NBIN] = {nullptr};
    cudaEvent_t side_ev[STRK_PK_NBIN] = {nullptr};
    cudaEvent_t fork_ev = nullptr;
    cudaStream_t stream = nullptr;
    // Uploads and planning kernels of strk_batch_fill run on a LOW-priority stream, the DP streams are created with the
    // greatest priority: in the streamed path the fill of block i + 1 (two contexts) shares the GPU with the DP kernels
    // of block i, and its large grids of short blocks (one warp per read) would otherwise take the slots every
    // finishing DP class frees before the next class' CTAs can.
    cudaStream_t fill_stream = nullptr;
    int prio_hi = 0;
    ScoreConsts h_consts;
    ScoreConsts *d_consts = nullptr;
    int gap = 5, end_flags = STRK_MODE_SG, tie_flags = 0;
    DevBuf<int> scratch;
    DevBuf<FamDesc> fams;
    DevBuf<int> table;
    DevBuf<long long> table64;
    DevBuf<int> list_a, list_b, list_d;  // widening-pass lists
    DevBuf<long long> list_c;
    DevBuf<int> fallback;            // reads the packed kernel handed to the general kernel
    DevBuf<uint4> pk_scratch;        // captured DP columns of the packed kernel (per resident warp)
    DevBuf<double> al_rep[3];        // allele calling: replicate means / weights / stdevs of one chunk of loci
    DevBuf<unsigned char> al_peaks;  //                 replicate peak counts
    // reference path (strk_ref_counts): per-locus arrays, pending lists, and the one-read-per-locus batch of phase 2
    DevBuf<unsigned char> ref_arena;
    DevBuf<unsigned long long> ref_u64[2];
    DevBuf<int> ref_i[13];
    struct strk_batch *ref_batch = nullptr;
    std::vector<int> ref_flat, ref_hwd;  // host sources of asynchronous copies of the reference path's fast path
    DevBuf<int> ref_out;                 // its results, assembled on the device
    DevBuf<int> al_i[8];             //                 per-read / per-locus integer arrays (recycled across calls)
    DevBuf<double> al_d[3];
    DevBuf<long long> al_rb;
    unsigned int *d_queue = nullptr;  // [0] work queue, [1] miss counter
    double *d_acc = nullptr;          // [0] ref cells, [1] executed cells
    PlanStats *d_plan = nullptr;      // device-side planning counters
    unsigned int *d_bin_off = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    struct strk_batch *reuse = nullptr;  // device buffers recycled by strk_count_reads
};

struct strk_batch {
    long long n_reads = 0, n_loci = 0;
    DevBuf<unsigned char> arena, status;
    DevBuf<uint4> arena4;  // nibble-packed copy as uploaded (STRK_ARENA_NIBBLE), expanded into `arena` on the device
    DevBuf<unsigned long long> seq_off, motif_off;
    DevBuf<int> lens, est, motif_len, read_locus, order, out;
    DevBuf<long long> read_begin;
    unsigned char *d_arena = nullptr;
    unsigned long long *d_seq_off = nullptr, *d_motif_off = nullptr;
    int *d_lens = nullptr, *d_est = nullptr, *d_motif_len = nullptr, *d_read_locus = nullptr, *d_order = nullptr;
    long long *d_read_begin = nullptr;
    int *d_out = nullptr;
    unsigned char *d_status = nullptr;
    // host mirror used by the widening passes (per locus, small); per-read planning happens on the device
    std::vector<long long> h_read_begin;
    DevBuf<unsigned char> bin;
    // identical reads of a locus share one table (dedupe.cuh): rep[r] = the read whose row r uses
    DevBuf<unsigned long long> hash;
    DevBuf<int> rep;
    DevBuf<double> hint;  // per locus {smallest, largest} carried-offset fraction seen at a window miss (strk_slot_window)
    int *d_rep = nullptr;
    long long n_dup = 0;
    int max_n1 = 0, mb_cols_base = 0, mb_m = 0;
    // first-window policy, carried from one block of loci to the next one filled into this object: 1 = short motifs
    // get the wider first window of strk_read_wd (set when a block sent > 2.5 % of its loci to a second pass,
    // cleared when < 1 % of a block's loci used the margin).  Speed only: results never depend on the window.
    int wide_short = 0;
    // work plan over h_order: [0, n_general) general-kernel-only reads, then one segment per packed R
    long long n_general = 0;
    long long bin_off[STRK_PK_NBIN] = {0}, bin_cnt[STRK_PK_NBIN] = {0};  // index R (0 = general kernel)
    int bin_mmax[STRK_PK_NBIN] = {0}, bin_flank[STRK_PK_NBIN] = {0};
    void release() {
        arena.release(), status.release(), seq_off.release(), motif_off.release(), lens.release(), est.release();
        motif_len.release(), read_locus.release(), order.release(), out.release(), read_begin.release();
        bin.release(), hash.release(), rep.release(), arena4.release(), hint.release();
    }
};

// ------------------------------------------------------------------------------------------------
// init / destroy
// ------------------------------------------------------------------------------------------------
static void build_lut(unsigned char lut[256]) {
    // parasail matrix_create mapper: alphabet "ACGTRYSWKMBDHVNX" (align_matrix.py:25), case-insensitive,
    // everything else -> the wildcard column (index 16)
    const char *alpha = "ACGTRYSWKMBDHVNX";
    for (int i = 0; i < 256; ++i) lut[i] = 16;
    for (int i = 0; i < 16; ++i) {
        lut[(unsigned char)alpha[i]] = (unsigned char)i;
        lut[(unsigned char)(alpha[i] - 'A' + 'a')] = (unsigned char)i;
    }
}

extern "C" int strk_destroy(strk_ctx *ctx);

extern "C" int strk_init(int device, const int8_t matrix[STRK_NSYM * STRK_NSYM], int gap_open, int gap_extend,
                         int end_flags, int tie_flags, strk_ctx **out) {
    if (!out || !matrix) return set_err(STRK_ERR_ARG, "strk_init: null argument");
    *out = nullptr;
    if (gap_open != gap_extend)
        return set_err(STRK_ERR_UNSUPPORTED,
                       "gap_open (%d) != gap_extend (%d): only the linear-gap case the reference uses "
                       "(indel_penalty for both, repeats.py:33) is implemented",
                       gap_open, gap_extend);
    if (gap_open < 1 || gap_open > 60) return set_err(STRK_ERR_ARG, "gap penalty %d out of range [1, 60]", gap_open);
    if (end_flags < 0 || end_flags > 15) return set_err(STRK_ERR_ARG, "end_flags %d out of range", end_flags);
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return set_err(STRK_ERR_CUDA, "no CUDA device available (%s); strkit_b200 has no CPU fallback",
                       e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n_dev) return set_err(STRK_ERR_ARG, "device %d out of range (%d devices)", device, n_dev);
    CU(cudaSetDevice(device));
    // STRK_BLOCKING_SYNC=1: host threads sleep in stream synchronisation instead of spinning (for boxes with fewer
    // host cores than ranks x host threads; the streamed path keeps three host threads per GPU busy otherwise)
    if (getenv("STRK_BLOCKING_SYNC")) {
        if (cudaSetDeviceFlags(cudaDeviceScheduleBlockingSync) != cudaSuccess) cudaGetLastError();
    }
    strk_ctx *ctx = new (std::nothrow) strk_ctx();
    if (!ctx) return set_err(STRK_ERR_NOMEM, "out of host memory");
    // any failure below releases what has been created so far (strk_destroy copes with a half-built context)
    struct Guard {
        strk_ctx *c;
        ~Guard() {
            if (c) strk_destroy(c);
        }
    } guard{ctx};
    ctx->device = device;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    ctx->n_sm = prop.multiProcessorCount;
    ctx->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
    ctx->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
    // L2 set-aside for persisting lines (the packed kernel's capture scratch): a device-wide limit, set once here
    if (ctx->l2_persist_max && getenv("STRK_L2_WINDOW")) {
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, ctx->l2_persist_max) != cudaSuccess) {
            cudaGetLastError();
            ctx->l2_persist_max = 0;
        }
    }
    ctx->gap = gap_open;
    ctx->end_flags = end_flags;
    ctx->tie_flags = tie_flags;
    build_lut(ctx->h_consts.lut);
    for (int a = 0; a < STRK_NSYM; ++a)
        for (int b = 0; b < STRK_NSYM; ++b) {
            int v = matrix[a * STRK_NSYM + b];
            if (v < -60 || v > 60) return set_err(STRK_ERR_ARG, "matrix entry %d out of range [-60, 60]", v);
            ctx->h_consts.smat[a * STRK_NSYM + b] = (signed char)v;
        }
    for (int b = 0; b < STRK_NSYM; ++b) {
        ctx->h_consts.smat[STRK_PAD_FREE * STRK_NSYM + b] = 0;
        ctx->h_consts.smat[STRK_PAD_PEN * STRK_NSYM + b] = (signed char)(-2 * gap_open);
    }
    ctx->h_consts.gap = gap_open;
    ctx->h_consts.end_flags = end_flags;
    {
        // PRMT tables of the packed kernel: byte c of t8[b] = score(row class c, column symbol b) + 2g
        static const int class_code[7] = {0, 1, 2, 3, 14, 15, 16};  // A C G T N X other
        int ok = 1;
        for (int b = 0; b < STRK_NSYM; ++b) {
            unsigned long long tf = 0, tb = 0;
            for (int c = 0; c < 7; ++c) {
                int v = matrix[class_code[c] * STRK_NSYM + b] + 2 * gap_open;
                if (v < 0 || v > 127) ok = 0;
                tf |= (unsigned long long)(v & 0xff) << (8 * c);
            }
            tb = tf;  // byte 7 (pad rows) stays 0: pad rows copy the row above them (dp_packed.cuh, borders)
            ctx->h_consts.t8f[b] = tf;
            ctx->h_consts.t8b[b] = tb;
        }
        for (int a = 0; a < STRK_NSYM; ++a)
            for (int b = 0; b < STRK_NSYM; ++b) {
                int v = matrix[a * STRK_NSYM + b] + 2 * gap_open;
                if (v < 0 || v > 127) ok = 0;
            }
        for (int k = 0; k <= STRK_SMAT_ROWS; ++k) ctx->h_consts.cls_of[k] = 0x80;
        for (int c = 0; c < 7; ++c) ctx->h_consts.cls_of[class_code[c]] = (unsigned char)c;
        ctx->h_consts.cls_of[STRK_PAD_FREE] = 7;
        ctx->h_consts.cls_of[STRK_PAD_PEN] = 7;
        for (int k = 0; k <= STRK_SMAT_ROWS; ++k) {
            const unsigned c = ctx->h_consts.cls_of[k];
            unsigned w = 0x800u | 0x88u;  // not representable
            if (!(c & 0x80)) {
                const unsigned cls = c & 7u;
                const unsigned add = cls < 4 || cls == 7 ? 0u : (unsigned)(ctx->h_consts.t8f[0] >> (8 * cls)) & 0xffu;
                w = (cls < 4 ? cls : 8u) | ((cls < 4 ? 4u + cls : 8u) << 4) | (cls << 8) |
                    ((cls < 4 || cls == 7) ? 0u : 0x1000u) | (add << 16);
            }
            ctx->h_consts.rowinfo[k] = w;
        }
        ctx->h_consts.packed_ok = ok;
        int one = 1;  // one-table flank path: a non-ACGT row symbol must score the same against A, C, G and T
        for (int c = 4; c < 7; ++c)
            for (int b = 1; b < 4; ++b)
                if (matrix[class_code[c] * STRK_NSYM + b] != matrix[class_code[c] * STRK_NSYM]) one = 0;
        ctx->h_consts.one_table_ok = one;
        for (int k = 0; k < 32; ++k) ctx->h_consts.one_v[k] = 1u;
    }
    {
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // lo = least, hi = greatest (numerically smaller)
        if (getenv("STRK_NO_PRIORITY")) hi = lo;          // (measurement only: every stream at the default priority)
        ctx->prio_hi = hi;
        CU(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, hi));
        CU(cudaStreamCreateWithPriority(&ctx->fill_stream, cudaStreamNonBlocking, lo));
    }
    CU(cudaMalloc((void **)&ctx->d_consts, sizeof(ScoreConsts)));
    CU(cudaMemcpy(ctx->d_consts, &ctx->h_consts, sizeof(ScoreConsts), cudaMemcpyHostToDevice));
    CU(cudaMalloc((void **)&ctx->d_queue, (8 + STRK_PK_NBIN + 1) * sizeof(unsigned int)));  // [8 + k]: work queue of packed class k
    CU(cudaMalloc((void **)&ctx->d_acc, 4 * sizeof(double)));
    CU(cudaMalloc((void **)&ctx->d_plan, sizeof(PlanStats)));
    CU(cudaMalloc((void **)&ctx->d_bin_off, STRK_PK_NBIN * sizeof(unsigned int)));
    for (int k = 0; k < 6; ++k) CU(cudaEventCreate(&ctx->ev[k]));
    guard.c = nullptr;
    *out = ctx;
    return STRK_OK;
}

extern "C" int strk_batch_free(strk_ctx *ctx, strk_batch *b);

extern "C" int strk_sync(strk_ctx *ctx) {
    if (!ctx) return set_err(STRK_ERR_ARG, "strk_sync: null context");
    CU(cudaSetDevice(ctx->device));
    for (int k = 0; k < STRK_PK_NBIN; ++k)
        if (ctx->side[k]) CU(cudaStreamSynchronize(ctx->side[k]));
    CU(cudaStreamSynchronize(ctx->fill_stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return STRK_OK;
}

extern "C" int strk_destroy(strk_ctx *ctx) {
    if (!ctx) return STRK_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->reuse) strk_batch_free(ctx, ctx->reuse);
    ctx->reuse = nullptr;
    if (ctx->ref_batch) strk_batch_free(ctx, ctx->ref_batch);
    ctx->ref_batch = nullptr;
    ctx->ref_arena.release();
    ctx->ref_out.release();
    for (int k = 0; k < 2; ++k) ctx->ref_u64[k].release();
    for (int k = 0; k < 13; ++k) ctx->ref_i[k].release();
    ctx->scratch.release();
    ctx->fams.release();
    ctx->table.release();
    ctx->table64.release();
    ctx->list_a.release();
    ctx->list_b.release();
    ctx->list_d.release();
    ctx->list_c.release();
    ctx->fallback.release();
    ctx->pk_scratch.release();
    for (int k = 0; k < 3; ++k) ctx->al_rep[k].release();
    ctx->al_peaks.release();
    for (int k = 0; k < 8; ++k) ctx->al_i[k].release();
    for (int k = 0; k < 3; ++k) ctx->al_d[k].release();
    ctx->al_rb.release();
    if (ctx->d_consts) cudaFree(ctx->d_consts);
    if (ctx->d_queue) cudaFree(ctx->d_queue);
    if (ctx->d_acc) cudaFree(ctx->d_acc);
    if (ctx->d_plan) cudaFree(ctx->d_plan);
    if (ctx->d_bin_off) cudaFree(ctx->d_bin_off);
    for (int k = 0; k < 6; ++k)
        if (ctx->ev[k]) cudaEventDestroy(ctx->ev[k]);
    for (int k = 0; k < STRK_PK_NBIN; ++k) {
        if (ctx->side[k]) cudaStreamDestroy(ctx->side[k]);
        if (ctx->side_ev[k]) cudaEventDestroy(ctx->side_ev[k]);
    }
    if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
    if (ctx->fill_stream) cudaStreamDestroy(ctx->fill_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return STRK_OK;
}

extern "C" int strk_host_register(void *ptr, uint64_t bytes) {
    if (!ptr || !bytes) return set_err(STRK_ERR_ARG, "strk_host_register: null/empty buffer");
    CU(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return STRK_OK;
}
extern "C" int strk_host_unregister(void *ptr) {
    if (!ptr) return set_err(STRK_ERR_ARG, "strk_host_unregister: null buffer");
    CU(cudaHostUnregister(ptr));
    return STRK_OK;
}

extern "C" int strk_get_stats(strk_ctx *ctx, double stats[8]) {
    if (!ctx || !stats) return set_err(STRK_ERR_ARG, "strk_get_stats: null argument");
    memcpy(stats, ctx->stats, sizeof(ctx->stats));
    return STRK_OK;
}

// ------------------------------------------------------------------------------------------------
// validation shared by every entry point that takes sequences
// ------------------------------------------------------------------------------------------------
static int validate_reads(const char *who, uint64_t arena_bytes, const uint64_t *seq_off, const int32_t *lens,
                          int64_t n_reads) {
    for (int64_t r = 0; r < n_reads; ++r) {
        const int a = lens[3 * r], b = lens[3 * r + 1], c = lens[3 * r + 2];
        if (a < 0 || b < 0 || c < 0) return set_err(STRK_ERR_ARG, "%s: read %lld has a negative length", who, (long long)r);
        const int64_t n1 = (int64_t)a + b + c;
        if (n1 <= 0) return set_err(STRK_ERR_ARG, "%s: read %lld is empty (fl + tr + fr has no bases)", who, (long long)r);
        if (n1 > (1 << 24)) return set_err(STRK_ERR_ARG, "%s: read %lld is too long (%lld)", who, (long long)r, (long long)n1);
        if (seq_off[r] > arena_bytes || (uint64_t)n1 > arena_bytes - seq_off[r])
            return set_err(STRK_ERR_ARG, "%s: read %lld runs past the end of the arena", who, (long long)r);
    }
    return STRK_OK;
}

static int validate_motifs(const char *who, uint64_t arena_bytes, const uint64_t *motif_off, const int32_t *motif_len,
                           int64_t n) {
    for (int64_t l = 0; l < n; ++l) {
        if (motif_len[l] <= 0) return set_err(STRK_ERR_ARG, "%s: motif %lld is empty", who, (long long)l);
        if (motif_off[l] > arena_bytes || (uint64_t)motif_len[l] > arena_bytes - motif_off[l])
            return set_err(STRK_ERR_ARG, "%s: motif %lld runs past the end of the arena", who, (long long)l);
    }
    return STRK_OK;
}

// ------------------------------------------------------------------------------------------------
// general DP launch
// ------------------------------------------------------------------------------------------------
// One family per CTA (dp_general.cuh): the strips of a long read run as a pipeline over the warps of the CTA.
// Throughput shape: GEN_WARPS warps per CTA, 16 / GEN_WARPS CTAs per SM (128 registers per thread).
// Latency shape: a launch that holds long (multi-strip) reads but no more families than fit the SMs at 8 warps each -- a
// locus or a few per call, the widening passes of a block of expansions -- is bound by the time of ONE read; there a
// CTA runs 8 warps, so that twice as many strips of a read are in flight.  Measured on 21 reads of 6 147 rows
// (profiles/r2_general_latency_shape.txt): 2.96 ms at 4 warps, 2.36 ms at 8, 2.52 ms at 16 -- a warp alone on its
// scheduler issues one instruction per 3.7 cycles (dependent issue), and beyond 2 warps per scheduler the strips of one
// read contend for the same issue slots and ALU pipe, so the gain stops at 8.
static int general_latency_warps() {  // (STRK_GEN_LATENCY_WARPS: measurement switch)
    static const int w = getenv("STRK_GEN_LATENCY_WARPS") ? atoi(getenv("STRK_GEN_LATENCY_WARPS")) : 8;
    return std::min(std::max(w, GEN_WARPS), GEN_WARPS_MAX);
}

struct GenShape {
    int grid, warps;
    size_t scratch;  // ints: per CTA [b_len][warps + 1 boundary rows of rowlen]
};

static GenShape general_shape(strk_ctx *ctx, long long n_fams, int b_len, int rowlen, bool latency) {
    GenShape g;
    g.warps = latency ? general_latency_warps() : GEN_WARPS;
    long long blocks = n_fams;
    const long long cap = (long long)ctx->n_sm * (16 / g.warps);  // persistent
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    g.grid = (int)blocks;
    g.scratch = ((size_t)b_len + (size_t)(g.warps + 1) * (size_t)rowlen) * (size_t)g.grid;
    return g;
}

// most families a launch may hold and still take the latency shape (STRK_GEN_LATENCY_MAX: measurement switch)
static long long lat_max(strk_ctx *ctx) {
    static const long long env = getenv("STRK_GEN_LATENCY_MAX") ? atoll(getenv("STRK_GEN_LATENCY_MAX")) : -1;
    return env >= 0 ? env : (long long)ctx->n_sm * (16 / general_latency_warps());  // one wave of latency-shape CTAs
}

static bool general_latency_mode(strk_ctx *ctx, long long n_fams, int rowlen, bool dev_count) {
    static const bool off = getenv("STRK_GEN_LATENCY") && atoi(getenv("STRK_GEN_LATENCY")) == 0;  // (measurement switch)
    return !off && !dev_count && rowlen > 2 && n_fams <= lat_max(ctx);
}

// scratch that any general launch of up to n_fams families may need (either shape)
static size_t general_scratch_max(strk_ctx *ctx, long long n_fams, int b_len, int rowlen) {
    size_t a = general_shape(ctx, n_fams, b_len, rowlen, false).scratch;
    if (rowlen > 2) a = std::max(a, general_shape(ctx, std::min(n_fams, lat_max(ctx)), b_len, rowlen, true).scratch);
    return a;
}

static int launch_general(strk_ctx *ctx, bool ref, const FamDesc *d_fams, const int *d_order, long long n_fams,
                          const unsigned char *d_arena, void *d_table, int b_len, int rowlen, cudaStream_t st,
                          const unsigned int *d_count = nullptr) {
    // d_count != nullptr: the list length lives on the device; n_fams is only its upper bound
    if (n_fams <= 0) return STRK_OK;
    if (d_count && n_fams > (long long)ctx->n_sm * 4) n_fams = (long long)ctx->n_sm * 4;
    if (n_fams > 0x7fffffffLL) return set_err(STRK_ERR_ARG, "too many families in one launch");
    const GenShape gs = general_shape(ctx, n_fams, b_len, rowlen, general_latency_mode(ctx, n_fams, rowlen, d_count != nullptr));
    const int grid = gs.grid, GEN_THREADS = gs.warps * 32;
    const size_t total = gs.scratch;
    if (ctx->scratch.reserve(total) != cudaSuccess) {
        cudaGetLastError();
        return set_err(STRK_ERR_NOMEM, "cannot allocate %zu bytes of DP scratch", total * sizeof(int));
    }
    CU(cudaMemsetAsync(ctx->d_queue, 0, sizeof(unsigned int), st));
    if (ref)
        dp_general_kernel<true><<<grid, GEN_THREADS, 0, st>>>(d_fams, d_order, (int)n_fams, d_arena, ctx->d_consts,
                                                              d_table, ctx->scratch.p, rowlen, b_len, ctx->d_queue,
                                                              d_count);
    else
        dp_general_kernel<false><<<grid, GEN_THREADS, 0, st>>>(d_fams, d_order, (int)n_fams, d_arena, ctx->d_consts,
                                                               d_table, ctx->scratch.p, rowlen, b_len, ctx->d_queue,
                                                               d_count);
    CU(cudaGetLastError());
    ctx->stats[2] += 1;
    return STRK_OK;
}

// ------------------------------------------------------------------------------------------------
// packed DP launch (one template instantiation per R; persistent warps, per-warp L2-resident scratch)
// ------------------------------------------------------------------------------------------------
template <int R, int L>
static int launch_packed_r(strk_ctx *ctx, const FamDesc *fams, const int *list, int n, const unsigned char *arena,
                           int *table, PackedDims dims, cudaStream_t st, int ref_mode, size_t *plan_scratch,
                           long long scratch_off) {
    // plan_scratch != nullptr: only report the capture scratch (uint4 units) this launch needs.
    // scratch_off >= 0: the launch uses pk_scratch + scratch_off, reserved by the caller (concurrent class launches).
    constexpr int HALVES = 32 / L;
    size_t smem = pk_smem_bytes(R, dims, L);
    if (const char *env = getenv("STRK_PK_EXTRA_SMEM")) smem += (size_t)atoi(env);  // occupancy experiments only
    // the opt-in shared-memory size is a per-device attribute of the function, shared by every context of the
    // process on that device (the streamed path keeps two): only ever raised, under a lock
    {
        static std::mutex mu;
        static size_t configured[64] = {0};
        std::lock_guard<std::mutex> lock(mu);
        size_t &cur = configured[ctx->device & 63];
        if (smem > cur) {
            CU(cudaFuncSetAttribute(dp_packed_kernel<R, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cur = smem;
        }
    }
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dp_packed_kernel<R, L>, pk_warps(L) * 32, smem));
    if (per_sm < 1) return set_err(STRK_ERR_CUDA, "packed kernel R=%d L=%d does not fit an SM (%zu B shared)", R, L, smem);
    long long grid = (long long)per_sm * ctx->n_sm;
    const long long need = ((long long)n + pk_warps(L) * HALVES - 1) / (pk_warps(L) * HALVES);
    if (grid > need) grid = need;
    const size_t words = pk_scratch_words_per_unit(R, dims.w_max, L) * (size_t)grid * pk_warps(L) * HALVES;
    if (plan_scratch) {
        *plan_scratch = (words + 3) / 4;
        return STRK_OK;
    }
    if (scratch_off < 0 && ctx->pk_scratch.reserve((words + 3) / 4) != cudaSuccess) {
        cudaGetLastError();
        return set_err(STRK_ERR_NOMEM, "cannot allocate %zu bytes of capture scratch", words * 4);
    }
    // The capture scratch is written and read back within one read's lifetime and then overwritten by the next
    // read of the same warp.  STRK_L2_WINDOW=1 keeps it resident in L2 (persisting access window) so that it is not
    // written back to HBM on eviction: measured 1176 -> 89 MB of DRAM traffic per launch of the R = 10 class, no
    // change in kernel time (the kernel is integer-issue bound, the write-backs use ~10 % of HBM bandwidth), -2 % on
    // the streamed path where two contexts' windows compete for the set-aside.  Off by default for that reason.
    {
        const size_t bytes = ctx->pk_scratch.cap * sizeof(uint4);  // the whole buffer: set again only when it moves
        static const bool off = getenv("STRK_L2_WINDOW") == nullptr;
        if (!off && scratch_off < 0 && ctx->l2_persist_max && ctx->l2_window_max &&
            (ctx->l2_window_ptr != ctx->pk_scratch.p || ctx->l2_window_bytes != bytes || ctx->l2_window_stream != st)) {
            const size_t win = bytes < ctx->l2_window_max ? bytes : ctx->l2_window_max;
            const size_t carve = win < ctx->l2_persist_max ? win : ctx->l2_persist_max;
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof(attr));
            attr.accessPolicyWindow.base_ptr = (void *)ctx->pk_scratch.p;
            attr.accessPolicyWindow.num_bytes = win;
            attr.accessPolicyWindow.hitRatio = win ? (float)((double)carve / (double)win) : 0.f;
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
            ctx->l2_window_ptr = ctx->pk_scratch.p;
            ctx->l2_window_bytes = bytes;
            ctx->l2_window_stream = st;
        }
    }
    unsigned int *work_counter = ctx->d_queue + 8 + (L == 16 ? R / 2 : R);  // one queue per rows-per-lane class
    CU(cudaMemsetAsync(work_counter, 0, sizeof(unsigned int), st));
    dp_packed_kernel<R, L><<<(unsigned)grid, pk_warps(L) * 32, smem, st>>>(fams, list, n, arena, ctx->d_consts, table, dims,
                                                                        ctx->pk_scratch.p + (scratch_off > 0 ? scratch_off : 0),
                                                                        ctx->fallback.p, ctx->d_queue + 2, ref_mode, work_counter);
    CU(cudaGetLastError());
    ctx->stats[2] += 1;
    return STRK_OK;
}

// Lanes per read of a rows-per-lane class: reads of the classes up to PK_PAIR_RMAX (db < 32 * PK_PAIR_RMAX bases) are
// swept two per warp, 16 lanes x 2R rows each (half the skew, half the per-step overhead per read).
// Larger classes were measured slower paired (R = 9: -1.2 %, 10: -1.8 %, 12: -3.2 %: 2R rows per lane need 127+
// registers).  STRK_PK_PAIRS=<largest paired class, 0 = none, at most 8> overrides the default (measurement only).
#define PK_PAIR_RMAX 8
static int pk_lanes_for_class(int R) {
    static const int rmax = [] {
        const char *e = getenv("STRK_PK_PAIRS");
        const int v = e ? atoi(e) : PK_PAIR_RMAX;
        return v < 0 ? 0 : (v > PK_PAIR_RMAX ? PK_PAIR_RMAX : v);
    }();
    return R <= rmax ? 16 : 32;
}
// sizes of the class' shared-memory tables for the lane count it runs with
static PackedDims pk_dims_for_class(int R, int max_flank, int max_m, int w_max) {
    const int L = pk_lanes_for_class(R);
    PackedDims d;
    d.colt_entries = max_flank + 2 * L;  // columns -(L - 1) .. flank + (L - 1), one spare
    d.prof_words = L == 16 ? pk_prof_words(2 * R, max_m, 16) : pk_prof_words(R, max_m, 32);
    d.w_max = w_max;
    return d;
}
static size_t pk_smem_for_class(int R, const PackedDims &d) {
    return pk_lanes_for_class(R) == 16 ? pk_smem_bytes(2 * R, d, 16) : pk_smem_bytes(R, d, 32);
}

// ref_mode: the families are reference windows and `table` holds the 64-bit boundary keys (score_ref_boundaries)
static int launch_packed(strk_ctx *ctx, int R, const FamDesc *fams, const int *list, int n, const unsigned char *arena,
                         int *table, PackedDims dims, cudaStream_t st, int ref_mode = 0, size_t *plan_scratch = nullptr,
                         long long scratch_off = -1) {
    if (pk_lanes_for_class(R) == 16) {
        switch (R) {
#define PK_CASE(N) \
    case N: return launch_packed_r<2 * N, 16>(ctx, fams, list, n, arena, table, dims, st, ref_mode, plan_scratch, scratch_off);
            PK_CASE(2) PK_CASE(3) PK_CASE(4) PK_CASE(5) PK_CASE(6) PK_CASE(7) PK_CASE(8)
#undef PK_CASE
            default: break;
        }
    }
    switch (R) {
#define PK_CASE(N) \
    case N: return launch_packed_r<N, 32>(ctx, fams, list, n, arena, table, dims, st, ref_mode, plan_scratch, scratch_off);
        PK_CASE(2) PK_CASE(3) PK_CASE(4) PK_CASE(5) PK_CASE(6) PK_CASE(7) PK_CASE(8) PK_CASE(9) PK_CASE(10) PK_CASE(11)
        PK_CASE(12) PK_CASE(13) PK_CASE(14) PK_CASE(15) PK_CASE(16)
#undef PK_CASE
        default: return set_err(STRK_ERR_ARG, "packed kernel: no instantiation for %d rows per lane", R);
    }
}

// One launch per rows-per-lane class k >= 1 of a pass (list[k], cnt[k]; tables sized by flank[k], mmax[k], w_max[k]);
// a class whose tables do not fit shared memory goes to the general kernel.  Large passes: back to back on the
// run's stream (spreading the classes over several streams was measured slower there, -2.8 %: concurrent grids with
// different footprints fragment the SMs).  Small passes, where a class is less than a wave of CTAs and each launch
// lasts one read's sweep whatever its grid: every class on its own side stream with its own slice of the capture
// scratch, so the launches overlap (blocks of a few hundred loci, per-locus calls, second passes, the reference path).
static int launch_packed_classes(strk_ctx *ctx, const int *const *list, const long long *cnt, const int *flank,
                                 const int *mmax, const int *w_max, long long n_slots, const FamDesc *fams,
                                 const unsigned char *arena, void *table, int b_len, int rowlen, cudaStream_t st,
                                 int ref_mode, long long *n_packed) {
    static const long long fan_max = [] {
        const char *e = getenv("STRK_PK_FANOUT_MAX");  // reads per pass up to which the classes overlap
        return e ? atoll(e) : 131072LL;
    }();
    int rc = STRK_OK;
    long long fan_off[STRK_PK_NBIN];
    bool fan_out = false;
    if (n_slots <= fan_max) {
        size_t total = 0;
        int n_classes = 0;
        for (int k = STRK_PK_RMAX; k >= 1; --k) {
            fan_off[k] = -1;
            if (!cnt[k]) continue;
            const PackedDims dims = pk_dims_for_class(k, flank[k], mmax[k], w_max[k]);
            if (pk_smem_for_class(k, dims) > 200 * 1024) continue;
            size_t need = 0;
            rc = launch_packed(ctx, k, fams, list[k], (int)cnt[k], arena, (int *)table, dims, st, ref_mode, &need);
            if (rc) return rc;
            fan_off[k] = (long long)total;
            total += need;
            ++n_classes;
        }
        if (n_classes >= 2) {
            if (ctx->pk_scratch.reserve(total) != cudaSuccess) {
                cudaGetLastError();
                return set_err(STRK_ERR_NOMEM, "cannot allocate %zu bytes of capture scratch", total * sizeof(uint4));
            }
            if (!ctx->fork_ev) CU(cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
            CU(cudaEventRecord(ctx->fork_ev, st));
            fan_out = true;
        }
    }
    for (int k = STRK_PK_RMAX; k >= 1; --k) {
        if (!cnt[k]) continue;
        const PackedDims dims = pk_dims_for_class(k, flank[k], mmax[k], w_max[k]);
        if (fan_out && fan_off[k] >= 0) {
            if (!ctx->side[k]) CU(cudaStreamCreateWithPriority(&ctx->side[k], cudaStreamNonBlocking, ctx->prio_hi));
            if (!ctx->side_ev[k]) CU(cudaEventCreateWithFlags(&ctx->side_ev[k], cudaEventDisableTiming));
            CU(cudaStreamWaitEvent(ctx->side[k], ctx->fork_ev, 0));
            rc = launch_packed(ctx, k, fams, list[k], (int)cnt[k], arena, (int *)table, dims, ctx->side[k], ref_mode, nullptr,
                               fan_off[k]);
            if (rc) return rc;
            CU(cudaEventRecord(ctx->side_ev[k], ctx->side[k]));
            CU(cudaStreamWaitEvent(st, ctx->side_ev[k], 0));
            *n_packed += cnt[k];
            continue;
        }
        if (pk_smem_for_class(k, dims) > 200 * 1024) {
            // shared memory would not fit: hand the whole segment to the general kernel
            rc = launch_general(ctx, ref_mode != 0, fams, list[k], cnt[k], arena, table, b_len, rowlen, st);
            if (rc) return rc;
            continue;
        }
        rc = launch_packed(ctx, k, fams, list[k], (int)cnt[k], arena, (int *)table, dims, st, ref_mode);
        if (rc) return rc;
        *n_packed += cnt[k];
    }
    return STRK_OK;
}

// One DP pass over a work plan: list[0] = families only the general kernel takes, list[k >= 1] = the packed classes,
// then the packed kernel's fallbacks (a device-side list) through the general kernel.  The general launch of list[0]
// is a handful of families (reads / reference windows of 512+ bases, IUPAC reads) whose duration is the LATENCY of one
// family's dependency chain (0.3 - 0.6 ms) whatever the grid: it starts first, on its own side stream, and runs under
// the packed launches instead of after them (it used to be 3 % of a 32 768-locus read step and a quarter of the
// reference path of the same block).
static int launch_pass(strk_ctx *ctx, const int *const *list, const long long *cnt, const int *flank, const int *mmax,
                       const int *w_max, long long n_slots, const FamDesc *fams, const unsigned char *arena, void *table,
                       int b_len, int rowlen, cudaStream_t st, int ref_mode, long long *n_packed) {
    bool overlap = cnt[0] > 0 && getenv("STRK_GENERAL_SERIAL") == nullptr;  // (the switch: measurement only)
    long long packed_total = 0;
    for (int k = 1; k < STRK_PK_NBIN; ++k) {
        if (!cnt[k]) continue;
        packed_total += cnt[k];
        const PackedDims dims = pk_dims_for_class(k, flank[k], mmax[k], w_max[k]);
        if (pk_smem_for_class(k, dims) > 200 * 1024) overlap = false;  // that class runs on the general kernel itself
    }
    {
        // scratch of the largest general launch of this pass, reserved before anything is in flight
        long long fams_max = std::max(cnt[0], std::min(packed_total, (long long)ctx->n_sm * 4));
        for (int k = 1; k < STRK_PK_NBIN; ++k) fams_max = std::max(fams_max, cnt[k]);
        const size_t total = general_scratch_max(ctx, fams_max, b_len, rowlen);
        if (ctx->scratch.reserve(total) != cudaSuccess) {
            cudaGetLastError();
            return set_err(STRK_ERR_NOMEM, "cannot allocate %zu bytes of DP scratch", total * sizeof(int));
        }
    }
    int rc = STRK_OK;
    if (overlap) {
        if (!ctx->side[0]) CU(cudaStreamCreateWithPriority(&ctx->side[0], cudaStreamNonBlocking, ctx->prio_hi));
        if (!ctx->side_ev[0]) CU(cudaEventCreateWithFlags(&ctx->side_ev[0], cudaEventDisableTiming));
        if (!ctx->fork_ev) CU(cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
        CU(cudaEventRecord(ctx->fork_ev, st));
        CU(cudaStreamWaitEvent(ctx->side[0], ctx->fork_ev, 0));
        rc = launch_general(ctx, ref_mode != 0, fams, list[0], cnt[0], arena, table, b_len, rowlen, ctx->side[0]);
        if (rc) return rc;
        CU(cudaEventRecord(ctx->side_ev[0], ctx->side[0]));
    }
    rc = launch_packed_classes(ctx, list, cnt, flank, mmax, w_max, n_slots, fams, arena, table, b_len, rowlen, st, ref_mode,
                               n_packed);
    if (rc) return rc;
    if (overlap) {
        CU(cudaStreamWaitEvent(st, ctx->side_ev[0], 0));
    } else {
        rc = launch_general(ctx, ref_mode != 0, fams, list[0], cnt[0], arena, table, b_len, rowlen, st);
        if (rc) return rc;
    }
    if (*n_packed) {
        rc = launch_general(ctx, ref_mode != 0, fams, ctx->fallback.p, *n_packed, arena, table, b_len, rowlen, st,
                            ctx->d_queue + 2);
        if (rc) return rc;
    }
    return STRK_OK;
}

// ------------------------------------------------------------------------------------------------
// batches
// ------------------------------------------------------------------------------------------------
template <typename T>
static cudaError_t upload(DevBuf<T> &buf, T **dst, const T *src, size_t n, cudaStream_t st) {
    cudaError_t e = buf.reserve(n ? n : 1);
    if (e != cudaSuccess) return e;
    *dst = buf.p;
    if (n) e = cudaMemcpyAsync(buf.p, src, n * sizeof(T), cudaMemcpyHostToDevice, st);
    return e;
}

extern "C" int strk_batch_free(strk_ctx *ctx, strk_batch *b) {
    if (!b) return STRK_OK;
    if (ctx) cudaSetDevice(ctx->device);
    b->release();
    delete b;
    return STRK_OK;
}

// Validation + work planning of a batch whose arrays are already on the device (b->d_*, n_reads, n_loci set).
static int batch_plan(strk_ctx *ctx, strk_batch *b, uint64_t arena_bytes) {
    const long long n_reads = b->n_reads, n_loci = b->n_loci;
    cudaStream_t st = ctx->fill_stream;  // (callers hand over host-synchronised data; every path below ends synchronised)
    b->d_rep = nullptr;
    b->n_dup = 0;
    if (n_reads == 0) {
        CU(cudaStreamSynchronize(st));
        return STRK_OK;
    }
    // ---- plan on the device
    PlanStats hp;
    CU(cudaMemsetAsync(ctx->d_plan, 0, sizeof(PlanStats), st));
    CU(cudaMemsetAsync(&ctx->d_plan->first_error, 0xff, sizeof(unsigned long long), st));
    const int T = 256;
    plan_loci_kernel<<<(unsigned)((n_loci + T - 1) / T), T, 0, st>>>(b->d_read_begin, n_loci, n_reads, b->d_motif_off,
                                                                     b->d_motif_len, arena_bytes, b->d_read_locus,
                                                                     ctx->d_plan);
    CU(cudaGetLastError());
    // the read scan dereferences read_locus / motif_len: stop here if the locus table itself is broken
    CU(cudaMemcpyAsync(&hp, ctx->d_plan, sizeof(PlanStats), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (hp.first_error == ~0ull) {
        plan_reads_scan_kernel<<<(unsigned)((n_reads + T - 1) / T), T, 0, st>>>(
            b->d_seq_off, b->d_lens, b->d_est, b->d_read_locus, b->d_motif_len, n_reads, arena_bytes,
            ctx->h_consts.packed_ok, b->bin.p, ctx->d_plan);
        CU(cudaGetLastError());
        // identical reads of a locus share one table (STRK_DEDUPE=0 switches it off: measurement only)
        const char *dd = getenv("STRK_DEDUPE");
        if ((!dd || atoi(dd) != 0) && n_reads > n_loci) {
            if (b->hash.reserve((size_t)n_reads) != cudaSuccess || b->rep.reserve((size_t)n_reads) != cudaSuccess) {
                cudaGetLastError();
                return set_err(STRK_ERR_NOMEM, "batch: cannot allocate the duplicate-read map");
            }
            hash_reads_kernel<<<(unsigned)((n_reads * 32 + T - 1) / T), T, 0, st>>>(b->d_arena, b->d_seq_off, b->d_lens, b->d_est,
                                                                                 n_reads, ctx->d_plan, b->hash.p);
            CU(cudaGetLastError());
            dedupe_loci_kernel<<<(unsigned)((n_loci * 32 + T - 1) / T), T, 0, st>>>(b->d_arena, b->d_seq_off, b->d_lens, b->d_est,
                                                                                  b->d_read_begin, n_loci, b->hash.p, b->bin.p,
                                                                                  ctx->d_plan, b->rep.p);
            CU(cudaGetLastError());
            b->d_rep = b->rep.p;
        }
        CU(cudaMemcpyAsync(&hp, ctx->d_plan, sizeof(PlanStats), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    if (hp.first_error != ~0ull) {
        const long long idx = (long long)(hp.first_error >> 8);
        static const char *what[] = {"", "has a negative length", "is empty (fl + tr + fr has no bases)", "is too long",
                                     "runs past the end of the arena", "has an out-of-range est_cn", "has an empty motif",
                                     "has a motif that runs past the end of the arena", "has a non-monotone read_begin"};
        const int kind = (int)(hp.first_error & 0xff);
        return set_err(STRK_ERR_ARG, "batch: %s %lld %s", kind >= PLAN_ERR_MOTIF_EMPTY ? "locus" : "read", idx,
                       what[kind <= 8 ? kind : 0]);
    }
    // segment offsets: general first, then packed classes from R = 16 down to 2 (class 1 stays empty)
    unsigned off[STRK_PK_NBIN];
    unsigned acc = hp.bin_cnt[0];
    off[0] = 0;
    b->n_general = hp.bin_cnt[0];
    for (int k = STRK_PK_RMAX; k >= 1; --k) {
        off[k] = acc;
        b->bin_off[k] = acc;
        b->bin_cnt[k] = hp.bin_cnt[k];
        b->bin_mmax[k] = hp.bin_mmax[k];
        b->bin_flank[k] = hp.bin_flank[k];
        acc += hp.bin_cnt[k];
    }
    b->max_n1 = hp.max_n1;
    b->mb_cols_base = hp.mb_cols_base;
    b->mb_m = hp.mb_m;
    b->n_dup = b->d_rep ? (long long)hp.n_dup : 0;
    CU(cudaMemcpyAsync(ctx->d_bin_off, off, sizeof(off), cudaMemcpyHostToDevice, st));
    plan_reads_scatter_kernel<<<(unsigned)((n_reads + T - 1) / T), T, 0, st>>>(b->bin.p, n_reads, ctx->d_bin_off,
                                                                              ctx->d_plan, b->d_order, b->d_rep);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));  // `off` is on this stack frame; the caller's buffers may go away after return
    return STRK_OK;
}

// The same validation + planning on the host, for small batches (the per-call wrappers of the drop-in API send one
// read at a time): no planning kernels, no read-backs, no stream synchronisation.  Mirrors plan.cuh rule for rule.
#define STRK_SMALL_BATCH 512
static int batch_plan_host(strk_ctx *ctx, strk_batch *b, uint64_t arena_bytes, const uint64_t *seq_off, const int32_t *lens,
                           const int32_t *est_cn, const int64_t *read_begin, const uint64_t *motif_off,
                           const int32_t *motif_len) {
    const long long n_reads = b->n_reads, n_loci = b->n_loci;
    cudaStream_t st = ctx->fill_stream;
    b->d_rep = nullptr;  // small batches: every read gets its own table
    b->n_dup = 0;
    if (n_reads == 0) {
        CU(cudaStreamSynchronize(st));
        return STRK_OK;
    }
    static const char *what[] = {"", "has a negative length", "is empty (fl + tr + fr has no bases)", "is too long",
                                 "runs past the end of the arena", "has an out-of-range est_cn", "has an empty motif",
                                 "has a motif that runs past the end of the arena", "has a non-monotone read_begin"};
    std::vector<int> read_locus((size_t)n_reads, 0), order((size_t)n_reads);
    std::vector<unsigned char> bin((size_t)n_reads, 0);
    unsigned long long first_error = ~0ull;
    auto report = [&](long long index, int kind) {
        const unsigned long long key = ((unsigned long long)index << 8) | (unsigned long long)kind;
        if (key < first_error) first_error = key;
    };
    for (long long l = 0; l < n_loci; ++l) {
        const long long r0 = read_begin[l], r1 = read_begin[l + 1];
        if (r1 < r0 || r0 < 0 || r1 > n_reads) {
            report(l, PLAN_ERR_READ_BEGIN);
            continue;
        }
        if (motif_len[l] <= 0)
            report(l, PLAN_ERR_MOTIF_EMPTY);
        else if (motif_off[l] > arena_bytes || (unsigned long long)motif_len[l] > arena_bytes - motif_off[l])
            report(l, PLAN_ERR_MOTIF_PAST_ARENA);
        for (long long r = r0; r < r1; ++r) read_locus[(size_t)r] = (int)l;
    }
    long long cnt[STRK_PK_NBIN] = {0};
    if (first_error == ~0ull) {
        for (long long r = 0; r < n_reads; ++r) {
            const int fl = lens[3 * r], tr = lens[3 * r + 1], fr = lens[3 * r + 2];
            const long long n1 = (long long)fl + tr + fr;
            const int est = est_cn[r];
            int bb = 0;
            if (fl < 0 || tr < 0 || fr < 0)
                report(r, PLAN_ERR_NEG_LEN);
            else if (n1 <= 0)
                report(r, PLAN_ERR_EMPTY);
            else if (n1 > (1 << 24))
                report(r, PLAN_ERR_TOO_LONG);
            else if (seq_off[r] > arena_bytes || (unsigned long long)n1 > arena_bytes - seq_off[r])
                report(r, PLAN_ERR_PAST_ARENA);
            else if (est < 0 || est > (1 << 22) ||
                     (long long)motif_len[read_locus[(size_t)r]] * (long long)est > (1ll << 24))
                report(r, PLAN_ERR_EST);
            else {
                const int m = motif_len[read_locus[(size_t)r]];
                int R = ctx->h_consts.packed_ok ? strk_pick_rows_packed((int)n1 + 1) : 0;
                if (fl < 1 || fr < 1 || fl > PK_FLANK_MAX || fr > PK_FLANK_MAX || m * R > 128 || m <= 0) R = 0;
                bb = R;
                b->max_n1 = std::max(b->max_n1, (int)n1);
                if (bb) {
                    b->bin_mmax[bb] = std::max(b->bin_mmax[bb], m);
                    b->bin_flank[bb] = std::max(b->bin_flank[bb], std::max(fl, fr));
                }
                if (n1 > 32 * 16 && m > 0) {
                    b->mb_cols_base = std::max(b->mb_cols_base, std::max(fl, fr) + m * est);
                    b->mb_m = std::max(b->mb_m, m);
                }
            }
            bin[(size_t)r] = (unsigned char)bb;
            ++cnt[bb];
        }
    }
    if (first_error != ~0ull) {
        const long long idx = (long long)(first_error >> 8);
        const int kind = (int)(first_error & 0xff);
        return set_err(STRK_ERR_ARG, "batch: %s %lld %s", kind >= PLAN_ERR_MOTIF_EMPTY ? "locus" : "read", idx,
                       what[kind <= 8 ? kind : 0]);
    }
    // segment offsets: general first, then packed classes from R = 16 down to 2 (as batch_plan)
    long long off[STRK_PK_NBIN], acc = cnt[0];
    off[0] = 0;
    b->n_general = cnt[0];
    for (int k = STRK_PK_RMAX; k >= 1; --k) {
        off[k] = acc;
        b->bin_off[k] = acc;
        b->bin_cnt[k] = cnt[k];
        acc += cnt[k];
    }
    for (long long r = 0; r < n_reads; ++r) order[(size_t)off[bin[(size_t)r]]++] = (int)r;
    CU(cudaMemcpyAsync(b->d_read_locus, read_locus.data(), (size_t)n_reads * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b->d_order, order.data(), (size_t)n_reads * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b->bin.p, bin.data(), (size_t)n_reads, cudaMemcpyHostToDevice, st));
    // The caller's arrays (possibly pinned: truly asynchronous DMA) and the vectors above are still being read by the
    // copies queued on `st`; the header promises that no host pointer is kept after a call returns, and strk_batch_run
    // may run on a caller stream that is not ordered after `st`.
    CU(cudaStreamSynchronize(st));
    return STRK_OK;
}

// H2D into (possibly recycled) device buffers, then validation + work planning on the device (plan.cuh)
static int batch_fill(strk_ctx *ctx, strk_batch *b, const uint8_t *arena, uint64_t arena_bytes, const uint64_t *seq_off,
                      const int32_t *lens, const int32_t *est_cn, int64_t n_reads, const int64_t *read_begin,
                      const uint64_t *motif_off, const int32_t *motif_len, int64_t n_loci,
                      int arena_format = STRK_ARENA_ASCII) {
    if (arena_format != STRK_ARENA_ASCII && arena_format != STRK_ARENA_NIBBLE)
        return set_err(STRK_ERR_ARG, "batch: unknown arena format %d", arena_format);
    if (n_reads < 0 || n_loci < 0 || n_reads > 0x7ffffff0LL || n_loci > 0x7ffffff0LL)
        return set_err(STRK_ERR_ARG, "batch: bad counts (%lld reads, %lld loci)", (long long)n_reads, (long long)n_loci);
    if ((n_reads && (!arena || !seq_off || !lens || !est_cn)) || !read_begin || (n_loci && (!motif_off || !motif_len)))
        return set_err(STRK_ERR_ARG, "batch: null array");
    if (read_begin[0] != 0 || read_begin[n_loci] != n_reads)
        return set_err(STRK_ERR_ARG, "batch: read_begin must run from 0 to n_reads");
    CU(cudaSetDevice(ctx->device));
    b->n_reads = n_reads;
    b->n_loci = n_loci;
    b->h_read_begin.assign(read_begin, read_begin + n_loci + 1);
    b->n_general = 0;
    for (int k = 0; k < STRK_PK_NBIN; ++k) b->bin_off[k] = b->bin_cnt[k] = 0, b->bin_mmax[k] = b->bin_flank[k] = 0;
    b->max_n1 = b->mb_cols_base = b->mb_m = 0;

    cudaStream_t st = ctx->fill_stream;
    cudaError_t e = cudaSuccess;
#define UP(buf, dst, src, n) \
    if (e == cudaSuccess) e = upload(b->buf, &b->dst, src, (size_t)(n), st)
    if (arena_format == STRK_ARENA_NIBBLE) {
        // packed bytes in, byte-per-symbol arena out (offsets and lengths are in symbols either way)
        const size_t quads = (size_t)((arena_bytes + 15) / 16);
        e = b->arena4.reserve(quads ? quads : 1);
        if (e == cudaSuccess) e = b->arena.reserve((size_t)(arena_bytes ? 2 * arena_bytes : 1));
        b->d_arena = b->arena.p;
        if (e == cudaSuccess && arena_bytes) {
            e = cudaMemcpyAsync(b->arena4.p, arena, (size_t)arena_bytes, cudaMemcpyHostToDevice, st);
            const unsigned long long threads = (arena_bytes >> 4) + 1;
            if (e == cudaSuccess) {
                expand_nibbles_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(b->arena4.p, arena_bytes, b->arena.p);
                e = cudaGetLastError();
            }
        }
        arena_bytes *= 2;  // symbols
    } else {
        UP(arena, d_arena, (const unsigned char *)arena, arena_bytes);
    }
    UP(seq_off, d_seq_off, (const unsigned long long *)seq_off, n_reads);
    UP(lens, d_lens, (const int *)lens, 3 * n_reads);
    UP(est, d_est, (const int *)est_cn, n_reads);
    UP(read_begin, d_read_begin, (const long long *)read_begin, n_loci + 1);
    UP(motif_off, d_motif_off, (const unsigned long long *)motif_off, n_loci);
    UP(motif_len, d_motif_len, (const int *)motif_len, n_loci);
#undef UP
    const size_t nr = (size_t)(n_reads ? n_reads : 1);
    if (e == cudaSuccess) e = b->read_locus.reserve(nr);
    if (e == cudaSuccess) e = b->order.reserve(nr);
    if (e == cudaSuccess) e = b->bin.reserve(nr);
    if (e == cudaSuccess) e = b->out.reserve(nr * 4);
    if (e == cudaSuccess) e = b->status.reserve((size_t)(n_loci ? n_loci : 1));
    b->d_read_locus = b->read_locus.p;
    b->d_order = b->order.p;
    b->d_out = b->out.p;
    b->d_status = b->status.p;
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_err(e == cudaErrorMemoryAllocation ? STRK_ERR_NOMEM : STRK_ERR_CUDA, "batch upload: %s",
                       cudaGetErrorString(e));
    }
    if (n_reads <= STRK_SMALL_BATCH) return batch_plan_host(ctx, b, arena_bytes, seq_off, lens, est_cn, read_begin, motif_off, motif_len);
    return batch_plan(ctx, b, arena_bytes);
}

extern "C" int strk_batch_create(strk_ctx *ctx, strk_batch **out) {
    if (!ctx || !out) return set_err(STRK_ERR_ARG, "strk_batch_create: null context/output");
    *out = new (std::nothrow) strk_batch();
    if (!*out) return set_err(STRK_ERR_NOMEM, "out of host memory");
    return STRK_OK;
}

extern "C" int strk_batch_fill(strk_ctx *ctx, strk_batch *b, const uint8_t *arena, uint64_t arena_bytes,
                               const uint64_t *seq_off, const int32_t *lens, const int32_t *est_cn, int64_t n_reads,
                               const int64_t *read_begin, const uint64_t *motif_off, const int32_t *motif_len,
                               int64_t n_loci) {
    if (!ctx || !b) return set_err(STRK_ERR_ARG, "strk_batch_fill: null context/batch");
    return batch_fill(ctx, b, arena, arena_bytes, seq_off, lens, est_cn, n_reads, read_begin, motif_off, motif_len, n_loci);
}

extern "C" int strk_batch_fill_fmt(strk_ctx *ctx, strk_batch *b, int arena_format, const uint8_t *arena,
                                   uint64_t arena_bytes, const uint64_t *seq_off, const int32_t *lens,
                                   const int32_t *est_cn, int64_t n_reads, const int64_t *read_begin,
                                   const uint64_t *motif_off, const int32_t *motif_len, int64_t n_loci) {
    if (!ctx || !b) return set_err(STRK_ERR_ARG, "strk_batch_fill_fmt: null context/batch");
    return batch_fill(ctx, b, arena, arena_bytes, seq_off, lens, est_cn, n_reads, read_begin, motif_off, motif_len, n_loci,
                      arena_format);
}

extern "C" int strk_batch_upload(strk_ctx *ctx, const uint8_t *arena, uint64_t arena_bytes, const uint64_t *seq_off,
                                 const int32_t *lens, const int32_t *est_cn, int64_t n_reads, const int64_t *read_begin,
                                 const uint64_t *motif_off, const int32_t *motif_len, int64_t n_loci,
                                 strk_batch **out) {
    if (!ctx || !out) return set_err(STRK_ERR_ARG, "strk_batch_upload: null context/output");
    *out = nullptr;
    strk_batch *b = new (std::nothrow) strk_batch();
    if (!b) return set_err(STRK_ERR_NOMEM, "out of host memory");
    int rc = batch_fill(ctx, b, arena, arena_bytes, seq_off, lens, est_cn, n_reads, read_begin, motif_off, motif_len,
                        n_loci);
    if (rc) {
        strk_batch_free(ctx, b);
        return rc;
    }
    *out = b;
    return STRK_OK;
}

// scratch sizes of the general kernel: backward column (longest db + 2) and, for multi-pass reads (db > 512),
// the boundary row (longest candidate prefix + 2)
static void scratch_dims(const strk_batch *b, int wd, int *b_len, int *rowlen) {
    *b_len = b->max_n1 + 2;
    *rowlen = b->max_n1 > 32 * 16 ? b->mb_cols_base + b->mb_m * wd + 2 : 2;
}

extern "C" int strk_batch_run(strk_ctx *ctx, strk_batch *b, int max_iters, int local_search_range, int step_size,
                              int kernel, void *stream) {
    if (!ctx || !b) return set_err(STRK_ERR_ARG, "strk_batch_run: null argument");
    if (max_iters < 0 || local_search_range < 0 || step_size < 0 || local_search_range > 1000 || step_size > 1000)
        return set_err(STRK_ERR_ARG, "strk_batch_run: bad search parameters");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    for (int k = 0; k < 8; ++k) ctx->stats[k] = 0;
    if (b->n_reads == 0) {
        return STRK_OK;
    }
    CU(cudaMemsetAsync(ctx->d_acc, 0, 4 * sizeof(double), st));

    // first-pass window half-width around the per-read estimate: the replay of a well-started search visits
    // start +- (range + step); the margin absorbs the carried start offset.  STRK_WD overrides (tuning only).
    int wd = local_search_range + step_size + 2;
    if (wd < 6) wd = 6;
    if (const char *env = getenv("STRK_WD")) {
        int v = atoi(env);
        if (v >= 1 && v <= 64) wd = v;
    }
    static const bool trace = getenv("STRK_RUN_TRACE") != nullptr;
    int ws = b->wide_short;
    if (const char *env = getenv("STRK_WIDE_SHORT")) ws = atoi(env) != 0;  // tuning only
    const int wd_cap = (STRK_MAX_WINDOW - 1) / 2 - 2;
    long long n_slots = b->n_reads;
    long long n_list = b->n_loci;
    const int *d_read_ids = nullptr, *d_locus_ids = nullptr;
    const long long *d_slot_begin = nullptr;
    std::vector<int> h_read_ids, h_locus_ids;
    std::vector<long long> h_slot_begin;
    std::vector<unsigned char> h_status, h_bin;
    std::vector<int> h_class_lists;
    float ms_dp = 0.f, ms_replay = 0.f;
    double acc[2] = {0, 0};  // [0] reference-equivalent cells, [1] executed cells

    // Where a widening pass has to look: per locus the carried-offset fractions its misses happened at (strk_slot_window)
    if (b->hint.reserve(2 * (size_t)(b->n_loci ? b->n_loci : 1)) != cudaSuccess) {
        cudaGetLastError();
        return set_err(STRK_ERR_NOMEM, "cannot allocate the window hints");
    }
    CU(cudaMemsetAsync(b->hint.p, 0, 2 * (size_t)b->n_loci * sizeof(double), st));

    for (int pass = 0;; ++pass) {
        if (wd > wd_cap) wd = wd_cap;
        const int threads = 256;
        const double *d_hint = pass ? b->hint.p : nullptr;
        int W = 2 * strk_read_wd(wd, 1, ws) + 1;  // table stride: the widest per-read window
        if (pass) {
            // windows of a widening pass are stretched towards the starts the locus' offset has produced: measure the widest
            CU(cudaMemsetAsync(ctx->d_queue + 6, 0, sizeof(unsigned int), st));
            plan_reads_kernel<<<(unsigned)((n_slots + threads - 1) / threads), threads, 0, st>>>(
                d_read_ids, n_slots, b->d_seq_off, b->d_lens, b->d_est, b->d_read_locus, b->d_motif_off, b->d_motif_len, wd, ws,
                W, ctx->fams.p, ctx->d_acc + 1, nullptr, d_hint, ctx->d_queue + 6);
            CU(cudaGetLastError());
            unsigned int w_needed = 0;
            CU(cudaMemcpyAsync(&w_needed, ctx->d_queue + 6, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            if ((int)w_needed > W) W = (int)w_needed;
            if (W > STRK_MAX_WINDOW)
                return set_err(STRK_ERR_SEARCH, "search left the widest supported window (%d sizes)", STRK_MAX_WINDOW);
        }
        if (ctx->fams.reserve((size_t)n_slots) != cudaSuccess || ctx->table.reserve((size_t)n_slots * (size_t)W) != cudaSuccess) {
            cudaGetLastError();
            return set_err(STRK_ERR_NOMEM, "cannot allocate score tables for %lld reads x %d sizes", n_slots, W);
        }
        int b_len, rowlen;
        scratch_dims(b, pass ? W : strk_read_wd(wd, 1, ws), &b_len, &rowlen);  // (n_hi - est <= W in a stretched window)
        plan_reads_kernel<<<(unsigned)((n_slots + threads - 1) / threads), threads, 0, st>>>(
            d_read_ids, n_slots, b->d_seq_off, b->d_lens, b->d_est, b->d_read_locus, b->d_motif_off, b->d_motif_len, wd,
            ws, W, ctx->fams.p, ctx->d_acc + 1, pass == 0 ? b->d_rep : nullptr, d_hint);
        CU(cudaGetLastError());
        ctx->stats[2] += 1;
        CU(cudaEventRecord(ctx->ev[0], st));
        int rc = STRK_OK;
        // packed kernel: every pass whose window it can hold (the first widening pass included: 8x the first window)
        const bool use_packed = kernel != STRK_KERNEL_GENERAL && ctx->h_consts.packed_ok && W <= PK_WINDOW_MAX;
        long long n_packed = 0;
        if (!use_packed) {
            // (first pass: the work order lists the representatives only when identical reads share tables)
            rc = launch_general(ctx, false, ctx->fams.p, pass ? nullptr : b->d_order, n_slots - (pass == 0 ? b->n_dup : 0),
                                b->d_arena, ctx->table.p, b_len, rowlen, st);
            if (rc) return rc;
        } else {
            // Packed kernel per rows-per-lane class; what it cannot take goes to the general kernel.
            // Pass 0: slot == read, the class segments of the batch's work order.  Widening passes: the slots of
            // the loci being redone, binned on the host by the class planned for their reads at upload time.
            if (ctx->fallback.reserve((size_t)b->n_reads) != cudaSuccess) {
                cudaGetLastError();
                return set_err(STRK_ERR_NOMEM, "cannot allocate the fallback list");
            }
            const int *seg_list[STRK_PK_NBIN];
            long long seg_cnt[STRK_PK_NBIN];
            int seg_flank[STRK_PK_NBIN], seg_mmax[STRK_PK_NBIN];
            for (int k = 0; k < STRK_PK_NBIN; ++k) seg_flank[k] = b->bin_flank[k], seg_mmax[k] = b->bin_mmax[k];
            if (pass == 0) {
                for (int k = 0; k < STRK_PK_NBIN; ++k) seg_list[k] = b->d_order + b->bin_off[k], seg_cnt[k] = b->bin_cnt[k];
                seg_cnt[0] = b->n_general;
                seg_list[0] = b->d_order;
            } else {
                if (h_bin.empty()) {
                    h_bin.resize((size_t)b->n_reads);
                    CU(cudaMemcpyAsync(h_bin.data(), b->bin.p, (size_t)b->n_reads, cudaMemcpyDeviceToHost, st));
                    CU(cudaStreamSynchronize(st));
                }
                std::vector<int> by_class[STRK_PK_NBIN];
                // A small pass is launch-latency bound (one read's sweep is ~100 us whatever the grid): its reads run
                // in the largest class of their lane group (pad rows cost nothing there), two launches instead of 15.
                int merged[STRK_PK_NBIN];
                for (int k = 0; k < STRK_PK_NBIN; ++k) merged[k] = k;
                if (h_read_ids.size() <= 8192) {
                    bool used[STRK_PK_NBIN] = {false};
                    for (int r : h_read_ids) used[h_bin[(size_t)r]] = true;
                    for (int L = 16; L <= 32; L += 16) {
                        int top = 0;
                        for (int k = 1; k < STRK_PK_NBIN; ++k)
                            if (used[k] && pk_lanes_for_class(k) == L) top = k;
                        for (int k = 1; k < top; ++k)
                            if (used[k] && pk_lanes_for_class(k) == L) {
                                merged[k] = top;
                                seg_flank[top] = seg_flank[top] > seg_flank[k] ? seg_flank[top] : seg_flank[k];
                                seg_mmax[top] = seg_mmax[top] > seg_mmax[k] ? seg_mmax[top] : seg_mmax[k];
                            }
                    }
                }
                for (size_t sl = 0; sl < h_read_ids.size(); ++sl)
                    by_class[merged[h_bin[(size_t)h_read_ids[sl]]]].push_back((int)sl);
                h_class_lists.clear();
                size_t at[STRK_PK_NBIN];
                for (int k = 0; k < STRK_PK_NBIN; ++k) {
                    at[k] = h_class_lists.size();
                    h_class_lists.insert(h_class_lists.end(), by_class[k].begin(), by_class[k].end());
                }
                if (ctx->list_d.reserve(h_class_lists.size()) != cudaSuccess) {
                    cudaGetLastError();
                    return set_err(STRK_ERR_NOMEM, "cannot allocate widening lists");
                }
                CU(cudaMemcpyAsync(ctx->list_d.p, h_class_lists.data(), h_class_lists.size() * sizeof(int),
                                   cudaMemcpyHostToDevice, st));
                CU(cudaStreamSynchronize(st));  // h_class_lists is reused by the next pass
                for (int k = 0; k < STRK_PK_NBIN; ++k) seg_list[k] = ctx->list_d.p + at[k], seg_cnt[k] = (long long)by_class[k].size();
            }
            CU(cudaMemsetAsync(ctx->d_queue + 2, 0, sizeof(unsigned int), st));
            int seg_w[STRK_PK_NBIN];
            for (int k = 0; k < STRK_PK_NBIN; ++k) seg_w[k] = (W + 3) / 4 * 4;
            rc = launch_pass(ctx, seg_list, seg_cnt, seg_flank, seg_mmax, seg_w, n_slots, ctx->fams.p, b->d_arena, ctx->table.p,
                             b_len, rowlen, st, 0, &n_packed);
            if (rc) return rc;
        }
        CU(cudaEventRecord(ctx->ev[1], st));
        CU(cudaMemsetAsync(ctx->d_queue + 1, 0, sizeof(unsigned int), st));
        CU(cudaMemsetAsync(ctx->d_queue + 3, 0, sizeof(unsigned int), st));
        if (W <= REPLAY_WMAX)
            replay_reads_small_kernel<<<(unsigned)((n_list + REPLAY_THREADS - 1) / REPLAY_THREADS), REPLAY_THREADS, 0, st>>>(
                ctx->table.p, W, wd, ws, d_locus_ids, d_slot_begin, (int)n_list, b->d_read_begin, b->d_est, b->d_lens,
                b->d_motif_len, max_iters, local_search_range, step_size, ctx->tie_flags, b->d_out, b->d_status,
                ctx->d_queue + 1, ctx->d_acc, pass == 0 ? b->d_rep : nullptr, b->hint.p);
        else
            replay_reads_kernel<<<(unsigned)((n_list + 127) / 128), 128, 0, st>>>(
                ctx->table.p, W, wd, ws, d_locus_ids, d_slot_begin, (int)n_list, b->d_read_begin, b->d_est, b->d_lens,
                b->d_motif_len, max_iters, local_search_range, step_size, ctx->tie_flags, b->d_out, b->d_status,
                ctx->d_queue + 1, ctx->d_acc, pass == 0 ? b->d_rep : nullptr, b->hint.p);
        CU(cudaGetLastError());
        ctx->stats[2] += 1;
        CU(cudaEventRecord(ctx->ev[2], st));
        // [0] loci whose search left the window, [1] packed-kernel fallbacks, [2] loci saved by the wide_short margin
        unsigned int miss_fb[3] = {0, 0, 0};
        CU(cudaMemcpyAsync(miss_fb, ctx->d_queue + 1, 3 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(acc, ctx->d_acc, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));  // cumulative over passes
        CU(cudaStreamSynchronize(st));
        const unsigned int miss = miss_fb[0];
        if (use_packed) {
            ctx->stats[6] += (double)(n_packed - (long long)miss_fb[1]);
            ctx->stats[7] -= (double)(n_packed - (long long)miss_fb[1]);
        }
        float t0 = 0.f, t1 = 0.f;
        CU(cudaEventElapsedTime(&t0, ctx->ev[0], ctx->ev[1]));
        CU(cudaEventElapsedTime(&t1, ctx->ev[1], ctx->ev[2]));
        ms_dp += t0;
        ms_replay += t1;
        ctx->stats[7] += (double)(n_slots - (pass == 0 ? b->n_dup : 0));  // duplicates run no DP of their own
        if (trace)
            fprintf(stderr, "[strk_batch_run %p] pass %d: wd %d (wide_short %d, %d sizes), %lld reads of %lld loci, %s kernel, "
                    "dp %.3f ms, replay %.3f ms, %u loci left the window, %u used the margin, %u fallbacks\n", (void *)ctx, pass, wd, ws, W,
                    n_slots, n_list, use_packed ? "packed" : "general", t0, t1, miss, miss_fb[2], miss_fb[1]);
        if (pass == 0 && b->n_loci >= 256) {  // policy for the next block filled into this batch object
            if (!ws && (long long)miss * 40 > b->n_loci) b->wide_short = 1;
            if (ws && (long long)(miss + miss_fb[2]) * 100 < b->n_loci) b->wide_short = 0;
        }
        if (miss == 0) break;

        // widening pass: redo the loci whose search left the window (status 1); status 2 is an error
        ctx->stats[5] += 1;
        if (wd >= wd_cap)
            return set_err(STRK_ERR_SEARCH, "search left the widest supported window (%d sizes)", STRK_MAX_WINDOW);
        h_status.resize((size_t)b->n_loci);
        CU(cudaMemcpy(h_status.data(), b->d_status, (size_t)b->n_loci, cudaMemcpyDeviceToHost));
        std::vector<int> loci;
        if (pass == 0) {
            for (long long l = 0; l < b->n_loci; ++l)
                if (h_status[(size_t)l]) loci.push_back((int)l);
        } else {
            for (int l : h_locus_ids)
                if (h_status[(size_t)l]) loci.push_back(l);
        }
        h_locus_ids.swap(loci);
        h_read_ids.clear();
        h_slot_begin.clear();
        for (int l : h_locus_ids) {
            if (h_status[(size_t)l] == 2)
                return set_err(STRK_ERR_SEARCH, "locus %d: the search scored no size (max_iters = %d)", l, max_iters);
            h_slot_begin.push_back((long long)h_read_ids.size());
            for (long long r = b->h_read_begin[(size_t)l]; r < b->h_read_begin[(size_t)l + 1]; ++r)
                h_read_ids.push_back((int)r);
        }
        n_slots = (long long)h_read_ids.size();
        n_list = (long long)h_locus_ids.size();
        if (ctx->list_a.reserve(h_read_ids.size()) != cudaSuccess || ctx->list_b.reserve(h_locus_ids.size()) != cudaSuccess ||
            ctx->list_c.reserve(h_slot_begin.size()) != cudaSuccess) {
            cudaGetLastError();
            return set_err(STRK_ERR_NOMEM, "cannot allocate widening lists");
        }
        CU(cudaMemcpyAsync(ctx->list_a.p, h_read_ids.data(), h_read_ids.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->list_b.p, h_locus_ids.data(), h_locus_ids.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->list_c.p, h_slot_begin.data(), h_slot_begin.size() * sizeof(long long),
                           cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
        d_read_ids = ctx->list_a.p;
        d_locus_ids = ctx->list_b.p;
        d_slot_begin = ctx->list_c.p;
        // The next pass looks where the misses happened (hints) with twice the margin; once that has not been enough twice
        // it widens 8x per pass like the blind scheme did (a search that goes nowhere near its start or its estimate).
        int widen = pass < 2 ? 2 : 8;
        if (const char *env = getenv("STRK_WIDEN1")) widen = pass == 0 && atoi(env) >= 2 ? atoi(env) : widen;  // tuning only
        wd *= widen;
    }
    ctx->stats[0] = acc[1];
    ctx->stats[1] = acc[0];
    ctx->stats[3] = ms_dp;
    ctx->stats[4] = ms_replay;
    return STRK_OK;
}

extern "C" int strk_batch_download(strk_ctx *ctx, strk_batch *b, int32_t *out) {
    if (!ctx || !b || (!out && b->n_reads)) return set_err(STRK_ERR_ARG, "strk_batch_download: null argument");
    CU(cudaSetDevice(ctx->device));
    if (b->n_reads) CU(cudaMemcpy(out, b->d_out, (size_t)b->n_reads * 4 * sizeof(int), cudaMemcpyDeviceToHost));
    return STRK_OK;
}

extern "C" int strk_count_reads_fmt(strk_ctx *ctx, int arena_format, const uint8_t *arena, uint64_t arena_bytes,
                                    const uint64_t *seq_off, const int32_t *lens, const int32_t *est_cn, int64_t n_reads,
                                    const int64_t *read_begin, const uint64_t *motif_off, const int32_t *motif_len,
                                    int64_t n_loci, int max_iters, int local_search_range, int step_size, int kernel,
                                    int32_t *out) {
    if (!ctx) return set_err(STRK_ERR_ARG, "strk_count_reads: null context");
    if (!ctx->reuse) ctx->reuse = new (std::nothrow) strk_batch();
    if (!ctx->reuse) return set_err(STRK_ERR_NOMEM, "out of host memory");
    strk_batch *b = ctx->reuse;  // device buffers are recycled across calls (no cudaMalloc per block of loci)
    int rc = batch_fill(ctx, b, arena, arena_bytes, seq_off, lens, est_cn, n_reads, read_begin, motif_off, motif_len,
                        n_loci, arena_format);
    if (rc) return rc;
    rc = strk_batch_run(ctx, b, max_iters, local_search_range, step_size, kernel, nullptr);
    if (!rc) rc = strk_batch_download(ctx, b, out);
    return rc;
}

extern "C" int strk_count_reads(strk_ctx *ctx, const uint8_t *arena, uint64_t arena_bytes, const uint64_t *seq_off,
                                const int32_t *lens, const int32_t *est_cn, int64_t n_reads, const int64_t *read_begin,
                                const uint64_t *motif_off, const int32_t *motif_len, int64_t n_loci, int max_iters,
                                int local_search_range, int step_size, int kernel, int32_t *out) {
    return strk_count_reads_fmt(ctx, STRK_ARENA_ASCII, arena, arena_bytes, seq_off, lens, est_cn, n_reads, read_begin,
                                motif_off, motif_len, n_loci, max_iters, local_search_range, step_size, kernel, out);
}

// The PyO3 function this library replaces, argument for argument (repeats.py:58-68): one read, strings in, four
// ints out.  A one-read batch through the same machinery (host-side planning, no device planning kernels).
extern "C" int strk_get_repeat_count(strk_ctx *ctx, int start_count, const char *tr_seq, int n_tr,
                                     const char *flank_left_seq, int n_fl, const char *flank_right_seq, int n_fr,
                                     const char *motif, int m, int max_iters, int local_search_range, int step_size,
                                     int use_shortcuts, int32_t out4[4]) {
    if (!ctx || !out4 || (n_tr && !tr_seq) || (n_fl && !flank_left_seq) || (n_fr && !flank_right_seq) || !motif)
        return set_err(STRK_ERR_ARG, "strk_get_repeat_count: null argument");
    if (n_tr < 0 || n_fl < 0 || n_fr < 0 || m <= 0) return set_err(STRK_ERR_ARG, "strk_get_repeat_count: bad length");
    if (use_shortcuts)
        return set_err(STRK_ERR_UNSUPPORTED, "use_shortcuts=True: the reference never passes it (repeats.py:67) and its "
                                             "semantics are not in the reference tree");
    std::vector<unsigned char> arena((size_t)n_fl + n_tr + n_fr + m);
    if (n_fl) memcpy(arena.data(), flank_left_seq, (size_t)n_fl);
    if (n_tr) memcpy(arena.data() + n_fl, tr_seq, (size_t)n_tr);
    if (n_fr) memcpy(arena.data() + n_fl + n_tr, flank_right_seq, (size_t)n_fr);
    memcpy(arena.data() + n_fl + n_tr + n_fr, motif, (size_t)m);
    const uint64_t seq_off = 0, motif_off = (uint64_t)n_fl + n_tr + n_fr;
    const int32_t lens[3] = {n_fl, n_tr, n_fr}, est = start_count, mlen = m;
    const int64_t read_begin[2] = {0, 1};
    int32_t row[4] = {0, 0, 0, 0};
    int rc = strk_count_reads_fmt(ctx, STRK_ARENA_ASCII, arena.data(), arena.size(), &seq_off, lens, &est, 1, read_begin,
                                  &motif_off, &mlen, 1, max_iters, local_search_range, step_size, STRK_KERNEL_AUTO, row);
    if (rc) return rc;
    out4[0] = row[0], out4[1] = row[1], out4[2] = row[2], out4[3] = row[0] - start_count;  // repeats.py:55-56
    return STRK_OK;
}

// ------------------------------------------------------------------------------------------------
// raw tables (parity checks of the DP itself) and the reference-boundary path
// ------------------------------------------------------------------------------------------------
struct TmpDev {
    std::vector<void *> ptrs;
    ~TmpDev() {
        for (void *p : ptrs) cudaFree(p);
    }
    template <typename T>
    cudaError_t up(T **dst, const T *src, size_t n) {
        cudaError_t e = cudaMalloc((void **)dst, (n ? n : 1) * sizeof(T));
        if (e != cudaSuccess) return e;
        ptrs.push_back(*dst);
        if (n && src) e = cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice);
        return e;
    }
};

static int tables_common(strk_ctx *ctx, bool ref, const uint8_t *arena, uint64_t arena_bytes, const uint64_t *seq_off,
                         const int32_t *lens, const int32_t *motif_idx, const int32_t *n_lo, const int32_t *n_hi,
                         int64_t n_reads, const uint64_t *motif_off, const int32_t *motif_len, int64_t n_motifs,
                         const uint64_t *out_off, int kernel, void *out_host) {
    const char *who = ref ? "strk_ref_boundary_tables" : "strk_score_tables";
    if (!ctx || !arena || !seq_off || !lens || !n_lo || !n_hi || !motif_off || !motif_len || !out_off || !out_host)
        return set_err(STRK_ERR_ARG, "%s: null argument", who);
    if (n_reads <= 0 || n_reads > 0x7ffffff0LL) return set_err(STRK_ERR_ARG, "%s: bad family count", who);
    int rc = validate_reads(who, arena_bytes, seq_off, lens, n_reads);
    if (rc) return rc;
    rc = validate_motifs(who, arena_bytes, motif_off, motif_len, n_motifs);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    std::vector<FamDesc> fams((size_t)n_reads);
    uint64_t total = 0;
    int b_len = 2, rowlen = 2;
    for (int64_t r = 0; r < n_reads; ++r) {
        const int64_t mi = motif_idx ? motif_idx[r] : r;
        if (mi < 0 || mi >= n_motifs) return set_err(STRK_ERR_ARG, "%s: motif index out of range", who);
        if (n_lo[r] < 0 || n_hi[r] < n_lo[r] || n_hi[r] > (1 << 22))
            return set_err(STRK_ERR_ARG, "%s: bad size window [%d, %d] for family %lld", who, n_lo[r], n_hi[r], (long long)r);
        FamDesc &f = fams[(size_t)r];
        f.db_off = seq_off[r];
        f.motif_off = motif_off[mi];
        f.out_off = out_off[r];
        f.n_fl = lens[3 * r];
        f.n_tr = lens[3 * r + 1];
        f.n_fr = lens[3 * r + 2];
        f.m = motif_len[mi];
        f.n_lo = n_lo[r];
        f.n_hi = n_hi[r];
        // an empty candidate has no alignment (parasail rejects empty sequences)
        if (ref && f.n_lo == 0 && (f.n_fl == 0 || f.n_fr == 0))
            return set_err(STRK_ERR_ARG, "%s: family %lld scores an empty candidate (n = 0 with an empty flank)", who,
                           (long long)r);
        if (!ref && f.n_lo == 0 && f.n_fl + f.n_fr == 0)
            return set_err(STRK_ERR_ARG, "%s: family %lld scores an empty candidate (n = 0 without flanks)", who,
                           (long long)r);
        const int n1 = f.n_fl + f.n_tr + f.n_fr;
        const int64_t cols = (int64_t)std::max(f.n_fl, f.n_fr) + (int64_t)f.m * f.n_hi;
        if (cols > (1 << 26)) return set_err(STRK_ERR_ARG, "%s: candidate too long", who);
        b_len = std::max(b_len, n1 + 2);
        if (n1 > 32 * 16) rowlen = std::max(rowlen, (int)cols + 2);
        total = std::max<uint64_t>(total, out_off[r] + (uint64_t)(f.n_hi - f.n_lo + 1));
    }
    TmpDev tmp;
    unsigned char *d_arena = nullptr;
    FamDesc *d_fams = nullptr;
    cudaError_t e = tmp.up(&d_arena, arena, (size_t)arena_bytes);
    if (e == cudaSuccess) e = tmp.up(&d_fams, fams.data(), fams.size());
    void *d_table = nullptr;
    const size_t elem = ref ? 2 * sizeof(long long) : sizeof(int);
    if (e == cudaSuccess) {
        e = cudaMalloc(&d_table, (size_t)total * elem);
        if (e == cudaSuccess) tmp.ptrs.push_back(d_table);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_err(STRK_ERR_NOMEM, "%s: %s", who, cudaGetErrorString(e));
    }
    for (int k = 0; k < 8; ++k) ctx->stats[k] = 0;
    if (ref || kernel == STRK_KERNEL_GENERAL || !ctx->h_consts.packed_ok) {
        rc = launch_general(ctx, ref, d_fams, nullptr, n_reads, d_arena, d_table, b_len, rowlen, ctx->stream);
        if (rc) return rc;
    } else {
        // same routing as strk_batch_run: packed kernel per R, the rest (and its fallbacks) to the general kernel
        std::vector<int> lists[STRK_PK_NBIN];
        int mmax[STRK_PK_NBIN] = {0}, flank[STRK_PK_NBIN] = {0}, wmax[STRK_PK_NBIN] = {0};
        for (int64_t r = 0; r < n_reads; ++r) {
            const FamDesc &f = fams[(size_t)r];
            int R = strk_pick_rows_packed(f.n_fl + f.n_tr + f.n_fr + 1);
            if (f.n_fl < 1 || f.n_fr < 1 || f.n_fl > PK_FLANK_MAX || f.n_fr > PK_FLANK_MAX || f.m * R > 128 ||
                f.n_hi - f.n_lo + 1 > 64)
                R = 0;
            lists[R].push_back((int)r);
            mmax[R] = std::max(mmax[R], f.m);
            flank[R] = std::max(flank[R], std::max(f.n_fl, f.n_fr));
            wmax[R] = std::max(wmax[R], f.n_hi - f.n_lo + 1);
        }
        std::vector<int> flat;
        size_t offs[STRK_PK_NBIN];
        for (int k = 0; k < STRK_PK_NBIN; ++k) {
            offs[k] = flat.size();
            flat.insert(flat.end(), lists[k].begin(), lists[k].end());
        }
        int *d_lists = nullptr;
        e = tmp.up(&d_lists, flat.data(), flat.size());
        if (e == cudaSuccess) e = ctx->fallback.reserve((size_t)n_reads);
        int *d_fb = ctx->fallback.p;
        if (e != cudaSuccess) return set_err(STRK_ERR_NOMEM, "%s: %s", who, cudaGetErrorString(e));
        CU(cudaMemsetAsync(ctx->d_queue + 2, 0, sizeof(unsigned int), ctx->stream));
        long long n_packed = 0;
        for (int k = STRK_PK_RMAX; k >= 1; --k) {
            if (lists[k].empty()) continue;
            const PackedDims dims = pk_dims_for_class(k, flank[k], mmax[k], (wmax[k] + 3) / 4 * 4);
            if (pk_smem_for_class(k, dims) > 200 * 1024) {
                rc = launch_general(ctx, false, d_fams, d_lists + offs[k], (long long)lists[k].size(), d_arena, d_table,
                                    b_len, rowlen, ctx->stream);
                if (rc) return rc;
                continue;
            }
            rc = launch_packed(ctx, k, d_fams, d_lists + offs[k], (int)lists[k].size(), d_arena, (int *)d_table, dims,
                               ctx->stream);
            if (rc) return rc;
            n_packed += (long long)lists[k].size();
        }
        rc = launch_general(ctx, false, d_fams, d_lists + offs[0], (long long)lists[0].size(), d_arena, d_table, b_len,
                            rowlen, ctx->stream);
        if (rc) return rc;
        if (n_packed) {
            rc = launch_general(ctx, false, d_fams, d_fb, n_packed, d_arena, d_table, b_len, rowlen, ctx->stream,
                                ctx->d_queue + 2);
            if (rc) return rc;
            unsigned int fb = 0;
            CU(cudaMemcpyAsync(&fb, ctx->d_queue + 2, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            ctx->stats[6] = (double)(n_packed - (long long)fb);
        }
        ctx->stats[7] = (double)n_reads - ctx->stats[6];
    }
    CU(cudaStreamSynchronize(ctx->stream));
    if (!ref) {
        CU(cudaMemcpy(out_host, d_table, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost));
    } else {
        // unpack (score, end_query) pairs: the device layout per family is [fwd keys W][rev keys W]
        std::vector<long long> keys((size_t)total * 2);
        CU(cudaMemcpy(keys.data(), d_table, keys.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        int32_t *o = (int32_t *)out_host;
        for (int64_t r = 0; r < n_reads; ++r) {
            const FamDesc &f = fams[(size_t)r];
            const int W = f.n_hi - f.n_lo + 1;
            const long long *k0 = keys.data() + 2 * f.out_off;
            for (int k = 0; k < W; ++k) {
                int fs, fe, rs, re;
                ref_unpack(k0[k], fs, fe);
                ref_unpack(k0[W + k], rs, re);
                int32_t *dst = o + 4 * (f.out_off + (uint64_t)k);
                dst[0] = fs, dst[1] = fe, dst[2] = rs, dst[3] = re;
            }
        }
    }
    return STRK_OK;
}

extern "C" int strk_score_tables(strk_ctx *ctx, const uint8_t *arena, uint64_t arena_bytes, const uint64_t *seq_off,
                                 const int32_t *lens, const int32_t *motif_idx, const int32_t *n_lo, const int32_t *n_hi,
                                 int64_t n_reads, const uint64_t *motif_off, const int32_t *motif_len, int64_t n_motifs,
                                 const uint64_t *out_off, int kernel, int32_t *scores) {
    return tables_common(ctx, false, arena, arena_bytes, seq_off, lens, motif_idx, n_lo, n_hi, n_reads, motif_off,
                         motif_len, n_motifs, out_off, kernel, scores);
}

extern "C" int strk_ref_boundary_tables(strk_ctx *ctx, const uint8_t *arena, uint64_t arena_bytes,
                                        const uint64_t *seq_off, const int32_t *lens, const int32_t *n_lo,
                                        const int32_t *n_hi, int64_t n_loci, const uint64_t *motif_off,
                                        const int32_t *motif_len, const uint64_t *out_off, int32_t *out) {
    return tables_common(ctx, true, arena, arena_bytes, seq_off, lens, nullptr, n_lo, n_hi, n_loci, motif_off, motif_len,
                         n_loci, out_off, STRK_KERNEL_GENERAL, out);
}

// One DP pass over `fams` (general kernel) and download of the raw table (int scores, or packed argmax keys).
template <typename T>
static int dp_tables_to_host(strk_ctx *ctx, bool ref, const std::vector<FamDesc> &fams, const unsigned char *d_arena,
                             size_t total_elems, std::vector<T> &host_table) {
    int b_len = 2, rowlen = 2;
    for (const FamDesc &f : fams) {
        const int n1 = f.n_fl + f.n_tr + f.n_fr;
        b_len = std::max(b_len, n1 + 2);
        if (n1 > 32 * 16) rowlen = std::max(rowlen, std::max(f.n_fl, f.n_fr) + f.m * f.n_hi + 2);
    }
    if (ctx->fams.reserve(fams.size()) != cudaSuccess) {
        cudaGetLastError();
        return set_err(STRK_ERR_NOMEM, "cannot allocate family descriptors");
    }
    void *d_table = nullptr;
    if (ref) {
        if (ctx->table64.reserve(total_elems) != cudaSuccess) {
            cudaGetLastError();
            return set_err(STRK_ERR_NOMEM, "cannot allocate the boundary table");
        }
        d_table = ctx->table64.p;
    } else {
        if (ctx->table.reserve(total_elems) != cudaSuccess) {
            cudaGetLastError();
            return set_err(STRK_ERR_NOMEM, "cannot allocate the score table");
        }
        d_table = ctx->table.p;
    }
    CU(cudaMemcpyAsync(ctx->fams.p, fams.data(), fams.size() * sizeof(FamDesc), cudaMemcpyHostToDevice, ctx->stream));
    int rc = launch_general(ctx, ref, ctx->fams.p, nullptr, (long long)fams.size(), d_arena, d_table, b_len, rowlen,
                            ctx->stream);
    if (rc) return rc;
    host_table.resize(total_elems);
    CU(cudaMemcpyAsync(host_table.data(), d_table, total_elems * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return STRK_OK;
}

// get_ref_repeat_count (repeats.py:73-192) for a batch of loci, device-resident from upload to results.
// Phase 1: boundary tables (two sg_qe sweeps per locus, every candidate size of a window at once, general kernel)
// + the dual-score search replayed one thread per locus -> l_offset / r_offset; loci whose search leaves the
// window are redone 4x wider.  Phase 2: the final get_repeat_count on the adjusted flanks = the read-path batch
// machinery (packed kernel, device replay, widening) with one "read" per locus, one run per search-parameter tier.
static int ref_counts_fast(strk_ctx *ctx, const uint8_t *arena, uint64_t arena_bytes, const uint64_t *seq_off,
                           const int32_t *lens, const int32_t *start_count, const int32_t *ref_size,
                           const int32_t *rc_params, int64_t n_loci, const uint64_t *motif_off, const int32_t *motif_len,
                           int vcf_anchor_size, int32_t *out, std::vector<int> &redo, bool *taken);
static int ref_counts_slow(strk_ctx *ctx, const uint8_t *arena, uint64_t arena_bytes, const uint64_t *seq_off,
                           const int32_t *lens, const int32_t *start_count, const int32_t *ref_size,
                           const int32_t *rc_params, int64_t n_loci, const uint64_t *motif_off,
                           const int32_t *motif_len, int vcf_anchor_size, int respect_coords, int32_t *out);

// Fast path first (one host synchronisation for the whole call: every decision between the kernels is taken on the
// device); the loci it could not finish -- a search that left its first window -- and the calls it does not take
// (several search-parameter tiers, respect_coords) go through the general path below, pass by pass.
extern "C" int strk_ref_counts(strk_ctx *ctx, const uint8_t *arena, uint64_t arena_bytes, const uint64_t *seq_off,
                               const int32_t *lens, const int32_t *start_count, const int32_t *ref_size,
                               const int32_t *rc_params, int64_t n_loci, const uint64_t *motif_off,
                               const int32_t *motif_len, int vcf_anchor_size, int respect_coords, int32_t *out) {
    if (!ctx || !arena || !seq_off || !lens || !start_count || !ref_size || !rc_params || !motif_off || !motif_len || !out)
        return set_err(STRK_ERR_ARG, "strk_ref_counts: null argument");
    if (n_loci <= 0 || n_loci > 0x7ffffff0LL) return set_err(STRK_ERR_ARG, "strk_ref_counts: bad locus count");
    static const bool no_fast = getenv("STRK_REF_SLOW") != nullptr;  // (measurement only)
    bool taken = false;
    std::vector<int> redo;
    if (!respect_coords && !no_fast) {
        int rc = ref_counts_fast(ctx, arena, arena_bytes, seq_off, lens, start_count, ref_size, rc_params, n_loci, motif_off,
                                 motif_len, vcf_anchor_size, out, redo, &taken);
        if (rc) return rc;
    }
    if (!taken)
        return ref_counts_slow(ctx, arena, arena_bytes, seq_off, lens, start_count, ref_size, rc_params, n_loci, motif_off,
                               motif_len, vcf_anchor_size, respect_coords, out);
    if (redo.empty()) return STRK_OK;
    // the few loci left over: their own little call through the general path (compact copies of their sequences)
    double keep[8];
    memcpy(keep, ctx->stats, sizeof(keep));
    const size_t nr = redo.size();
    std::vector<unsigned char> sub_arena;
    std::vector<uint64_t> so(nr), mo(nr);
    std::vector<int32_t> ln(3 * nr), sc(nr), rs(nr), rcp(3 * nr), ml(nr), sub_out(8 * nr);
    for (size_t q = 0; q < nr; ++q) {
        const int l = redo[q];
        const int n1 = lens[3 * l] + lens[3 * l + 1] + lens[3 * l + 2];
        so[q] = sub_arena.size();
        sub_arena.insert(sub_arena.end(), arena + seq_off[l], arena + seq_off[l] + n1);
        mo[q] = sub_arena.size();
        sub_arena.insert(sub_arena.end(), arena + motif_off[l], arena + motif_off[l] + motif_len[l]);
        for (int k = 0; k < 3; ++k) ln[3 * q + k] = lens[3 * l + k], rcp[3 * q + k] = rc_params[3 * l + k];
        sc[q] = start_count[l], rs[q] = ref_size[l], ml[q] = motif_len[l];
    }
    int rc = ref_counts_slow(ctx, sub_arena.data(), sub_arena.size(), so.data(), ln.data(), sc.data(), rs.data(), rcp.data(),
                             (int64_t)nr, mo.data(), ml.data(), vcf_anchor_size, 0, sub_out.data());
    if (rc) return rc;
    for (size_t q = 0; q < nr; ++q) memcpy(out + 8 * (size_t)redo[q], sub_out.data() + 8 * q, 8 * sizeof(int32_t));
    for (int k = 0; k < 8; ++k) ctx->stats[k] += keep[k];
    ctx->stats[5] = keep[5] + (double)nr;  // loci redone
    return STRK_OK;
}

static int ref_counts_slow(strk_ctx *ctx, const uint8_t *arena, uint64_t arena_bytes, const uint64_t *seq_off,
                           const int32_t *lens, const int32_t *start_count, const int32_t *ref_size,
                           const int32_t *rc_params, int64_t n_loci, const uint64_t *motif_off,
                           const int32_t *motif_len, int vcf_anchor_size, int respect_coords, int32_t *out) {
    if (n_loci <= 0 || n_loci > 0x7ffffff0LL) return set_err(STRK_ERR_ARG, "strk_ref_counts: bad locus count");
    int rc = validate_reads("strk_ref_counts", arena_bytes, seq_off, lens, n_loci);
    if (rc) return rc;
    rc = validate_motifs("strk_ref_counts", arena_bytes, motif_off, motif_len, n_loci);
    if (rc) return rc;
    const int WD_MAX = (STRK_MAX_WINDOW - 1) / 2;
    std::vector<int> h_wd((size_t)n_loci);
    int max_wd = 6, max_n1 = 0, mb_cols = 0, mb_m = 0;
    for (int64_t l = 0; l < n_loci; ++l) {
        if (start_count[l] < 0 || start_count[l] > (1 << 22) ||
            (long long)motif_len[l] * (long long)start_count[l] > (1ll << 24) || rc_params[3 * l] < 0 || rc_params[3 * l + 1] < 0 ||
            rc_params[3 * l + 2] < 0 || rc_params[3 * l + 1] > 1000 || rc_params[3 * l + 2] > 1000)
            return set_err(STRK_ERR_ARG, "strk_ref_counts: bad start count / search parameters for locus %lld", (long long)l);
        // first window of the boundary search: start +- max(6, range + step + 2) sizes, like the read path's (a search
        // that starts where it ends touches start +- (range + step)); one that leaves it is redone 4x wider.  Measured
        // on 32 768 HiFi-sized windows: +- 9 sizes 3.89 ms, +- 7 3.79 ms, +- 6 3.33 ms per block.  STRK_REF_WD=<n>
        // raises the floor of 6 (tuning only).
        static const int wd_env = getenv("STRK_REF_WD") ? atoi(getenv("STRK_REF_WD")) : 6;
        h_wd[(size_t)l] = std::max(wd_env > 0 ? wd_env : 6, rc_params[3 * l + 1] + rc_params[3 * l + 2] + 2);
        max_wd = std::max(max_wd, h_wd[(size_t)l]);
        const int n1 = lens[3 * l] + lens[3 * l + 1] + lens[3 * l + 2];
        max_n1 = std::max(max_n1, n1);
        if (n1 > 32 * 16) {  // multi-pass reads of the general kernel keep a boundary row per candidate column
            mb_cols = std::max(mb_cols, std::max(lens[3 * l], lens[3 * l + 2]) + motif_len[l] * start_count[l]);
            mb_m = std::max(mb_m, motif_len[l]);
        }
    }
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    for (int k = 0; k < 8; ++k) ctx->stats[k] = 0;
    // counters of this call (strk_get_stats afterwards): the boundary sweeps of phase 1 + the batch runs of phase 2
    double acc_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const size_t n = (size_t)n_loci;
    const bool timing = getenv("STRK_REF_TIMING") != nullptr;  // coarse phase timers on stderr (tuning only)
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    // ---- device-resident inputs and per-locus state (context-owned, recycled across calls)
    cudaError_t e = ctx->ref_arena.reserve((size_t)arena_bytes);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) e = ctx->ref_u64[k].reserve(n);
    const size_t isz[13] = {3 * n, n, n, 3 * n, n, n, n, n, n, n, n, 4, n};
    for (int k = 0; k < 13 && e == cudaSuccess; ++k) e = ctx->ref_i[k].reserve(isz[k]);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_err(STRK_ERR_NOMEM, "strk_ref_counts: %s", cudaGetErrorString(e));
    }
    unsigned char *d_arena = ctx->ref_arena.p;
    unsigned long long *d_seq_off = ctx->ref_u64[0].p, *d_motif_off = ctx->ref_u64[1].p;
    int *d_lens = ctx->ref_i[0].p, *d_start = ctx->ref_i[1].p, *d_ref_size = ctx->ref_i[2].p, *d_rc = ctx->ref_i[3].p;
    int *d_motif_len = ctx->ref_i[4].p, *d_wd = ctx->ref_i[5].p, *d_l_off = ctx->ref_i[6].p, *d_r_off = ctx->ref_i[7].p;
    int *d_n_off = ctx->ref_i[8].p, *d_ids_a = ctx->ref_i[9].p, *d_ids_b = ctx->ref_i[10].p;
    unsigned int *d_cnt = (unsigned int *)ctx->ref_i[11].p;
    CU(cudaMemcpyAsync(d_arena, arena, (size_t)arena_bytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_seq_off, seq_off, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_motif_off, motif_off, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_lens, lens, 3 * n * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_start, start_count, n * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_ref_size, ref_size, n * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_rc, rc_params, 3 * n * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_motif_len, motif_len, n * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_wd, h_wd.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(d_l_off, 0, n * sizeof(int), st));
    CU(cudaMemsetAsync(d_r_off, 0, n * sizeof(int), st));
    CU(cudaMemsetAsync(d_n_off, 0, n * sizeof(int), st));
    const int T = 128;

    // ---- phase 1: boundary extension (skipped with respect_coords, repeats.py:99)
    if (!respect_coords) {
        long long n_pending = n_loci;
        const int *d_ids = nullptr;
        int *d_again = d_ids_a;
        int cur_wd = max_wd;
        while (n_pending > 0) {
            const int stride_w = 2 * cur_wd + 1;
            if (ctx->fams.reserve((size_t)n_pending) != cudaSuccess ||
                ctx->table64.reserve((size_t)n_pending * (size_t)stride_w * 2) != cudaSuccess) {
                cudaGetLastError();
                return set_err(STRK_ERR_NOMEM, "strk_ref_counts: cannot allocate the boundary tables");
            }
            ref_plan1_kernel<<<(unsigned)((n_pending + T - 1) / T), T, 0, st>>>(d_ids, (int)n_pending, d_seq_off, d_lens,
                                                                                d_start, d_wd, d_motif_off, d_motif_len,
                                                                                stride_w, ctx->fams.p);
            CU(cudaGetLastError());
            const int b_len = max_n1 + 2;
            const int rowlen = max_n1 > 32 * 16 ? mb_cols + mb_m * cur_wd + 2 : 2;
            CU(cudaEventRecord(ctx->ev[0], st));
            if (d_ids == nullptr) {  // executed cells of the first pass: two sweeps of db x (flank + motif * n_hi)
                for (int64_t l = 0; l < n_loci; ++l) {
                    const double n1 = (double)lens[3 * l] + lens[3 * l + 1] + lens[3 * l + 2];
                    const double hi = (double)start_count[l] + h_wd[(size_t)l];
                    acc_stats[0] += n1 * ((double)lens[3 * l] + lens[3 * l + 2] + 2.0 * motif_len[l] * hi);
                }
            }
            if (d_ids == nullptr && ctx->h_consts.packed_ok && 2 * stride_w <= PK_WINDOW_MAX) {
                // first pass (locus q = q): packed kernel in reference mode per rows-per-lane class, the forward and
                // the reverse sg_qe alignment of a locus in the two halves of its lane words; the rest -> general
                std::vector<int> lists[STRK_PK_NBIN];
                int mmax[STRK_PK_NBIN] = {0}, flank[STRK_PK_NBIN] = {0};
                for (int64_t l = 0; l < n_loci; ++l) {
                    const int fl = lens[3 * l], tr = lens[3 * l + 1], fr = lens[3 * l + 2], m = motif_len[l];
                    int R = strk_pick_rows_packed(fl + tr + fr + 1);
                    if (fl < 1 || fr < 1 || fl > PK_FLANK_MAX || fr > PK_FLANK_MAX || m * R > 128) R = 0;
                    lists[R].push_back((int)l);
                    mmax[R] = std::max(mmax[R], m);
                    flank[R] = std::max(flank[R], std::max(fl, fr));
                }
                std::vector<int> flat;
                size_t at[STRK_PK_NBIN];
                for (int k = 0; k < STRK_PK_NBIN; ++k) {
                    at[k] = flat.size();
                    flat.insert(flat.end(), lists[k].begin(), lists[k].end());
                }
                int *d_lists = ctx->ref_i[12].p;
                if (ctx->fallback.reserve(n) != cudaSuccess) {
                    cudaGetLastError();
                    return set_err(STRK_ERR_NOMEM, "strk_ref_counts: cannot allocate the fallback list");
                }
                CU(cudaMemcpyAsync(d_lists, flat.data(), flat.size() * sizeof(int), cudaMemcpyHostToDevice, st));
                CU(cudaMemsetAsync(ctx->d_queue + 2, 0, sizeof(unsigned int), st));
                long long n_packed = 0;
                const int *seg_list[STRK_PK_NBIN];
                long long seg_cnt[STRK_PK_NBIN];
                int seg_w[STRK_PK_NBIN];
                for (int k = 0; k < STRK_PK_NBIN; ++k) {
                    seg_list[k] = d_lists + at[k];
                    seg_cnt[k] = (long long)lists[k].size();
                    seg_w[k] = 2 * stride_w - 1;  // forward + reverse columns of every size of the window
                }
                rc = launch_pass(ctx, seg_list, seg_cnt, flank, mmax, seg_w, n_loci, ctx->fams.p, d_arena, ctx->table64.p, b_len,
                                 rowlen, st, 1, &n_packed);
                if (rc) return rc;
                CU(cudaStreamSynchronize(st));  // `flat` is read by the copy above
            } else {
                rc = launch_general(ctx, true, ctx->fams.p, nullptr, n_pending, d_arena, ctx->table64.p, b_len, rowlen, st);
                if (rc) return rc;
            }
            CU(cudaEventRecord(ctx->ev[1], st));
            CU(cudaMemsetAsync(d_cnt, 0, 4 * sizeof(unsigned int), st));
            ref_replay1_kernel<<<(unsigned)((n_pending + T - 1) / T), T, 0, st>>>(
                d_ids, (int)n_pending, ctx->table64.p, ctx->fams.p, d_start, d_rc, d_ref_size, vcf_anchor_size, WD_MAX, d_wd,
                d_l_off, d_r_off, d_n_off, d_again, d_cnt);
            CU(cudaGetLastError());
            CU(cudaEventRecord(ctx->ev[2], st));
            ctx->stats[2] += 2;
            unsigned int h_cnt[4] = {0, 0, 0, 0};
            CU(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            {
                float t0 = 0.f, t1 = 0.f;
                CU(cudaEventElapsedTime(&t0, ctx->ev[0], ctx->ev[1]));
                CU(cudaEventElapsedTime(&t1, ctx->ev[1], ctx->ev[2]));
                acc_stats[3] += t0;
                acc_stats[4] += t1;
            }
            if (h_cnt[2])
                return set_err(STRK_ERR_SEARCH, "strk_ref_counts: locus %lld left the widest window",
                               (long long)(0x7fffffffu - h_cnt[2]));
            if (h_cnt[1]) {
                const long long l = (long long)(0x7fffffffu - h_cnt[1]);
                return set_err(STRK_ERR_SEARCH, "strk_ref_counts: locus %lld scored no size (max_iters = %d)", l,
                               rc_params[3 * l]);
            }
            n_pending = h_cnt[0];
            acc_stats[5] += (double)n_pending;
            d_ids = d_again;
            d_again = d_again == d_ids_a ? d_ids_b : d_ids_a;
            cur_wd = std::min(WD_MAX, cur_wd * 4);
        }
    }
    const double t_phase1 = now();
    acc_stats[2] = ctx->stats[2];
    std::vector<int> l_off(n), r_off(n), n_off(n);
    CU(cudaMemcpyAsync(l_off.data(), d_l_off, n * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(r_off.data(), d_r_off, n * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(n_off.data(), d_n_off, n * sizeof(int), cudaMemcpyDeviceToHost, st));

    // ---- phase 2: final count on the adjusted flanks (repeats.py:171-188), one batch run per parameter tier
    std::vector<std::vector<int>> tiers;
    std::vector<int> tier_key;  // 3 ints per tier
    for (int64_t l = 0; l < n_loci; ++l) {
        size_t t = 0;
        for (; t < tiers.size(); ++t)
            if (tier_key[3 * t] == rc_params[3 * l] && tier_key[3 * t + 1] == rc_params[3 * l + 1] &&
                tier_key[3 * t + 2] == rc_params[3 * l + 2])
                break;
        if (t == tiers.size()) {
            tiers.emplace_back();
            tier_key.insert(tier_key.end(), rc_params + 3 * l, rc_params + 3 * l + 3);
        }
        tiers[t].push_back((int)l);
    }
    if (!ctx->ref_batch) ctx->ref_batch = new (std::nothrow) strk_batch();
    if (!ctx->ref_batch) return set_err(STRK_ERR_NOMEM, "out of host memory");
    strk_batch *b = ctx->ref_batch;
    std::vector<int> res;
    for (size_t t = 0; t < tiers.size(); ++t) {
        const std::vector<int> &ids = tiers[t];
        const size_t nt = ids.size();
        e = b->seq_off.reserve(nt);
        if (e == cudaSuccess) e = b->lens.reserve(3 * nt);
        if (e == cudaSuccess) e = b->est.reserve(nt);
        if (e == cudaSuccess) e = b->motif_off.reserve(nt);
        if (e == cudaSuccess) e = b->motif_len.reserve(nt);
        if (e == cudaSuccess) e = b->read_begin.reserve(nt + 1);
        if (e == cudaSuccess) e = b->read_locus.reserve(nt);
        if (e == cudaSuccess) e = b->order.reserve(nt);
        if (e == cudaSuccess) e = b->bin.reserve(nt);
        if (e == cudaSuccess) e = b->out.reserve(nt * 4);
        if (e == cudaSuccess) e = b->status.reserve(nt);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return set_err(STRK_ERR_NOMEM, "strk_ref_counts: %s", cudaGetErrorString(e));
        }
        b->n_reads = b->n_loci = (long long)nt;
        b->d_arena = d_arena;
        b->d_seq_off = b->seq_off.p, b->d_lens = b->lens.p, b->d_est = b->est.p, b->d_motif_off = b->motif_off.p;
        b->d_motif_len = b->motif_len.p, b->d_read_begin = b->read_begin.p, b->d_read_locus = b->read_locus.p;
        b->d_order = b->order.p, b->d_out = b->out.p, b->d_status = b->status.p;
        b->h_read_begin.resize(nt + 1);
        for (size_t q = 0; q <= nt; ++q) b->h_read_begin[q] = (long long)q;
        b->n_general = 0;
        for (int k = 0; k < STRK_PK_NBIN; ++k) b->bin_off[k] = b->bin_cnt[k] = 0, b->bin_mmax[k] = b->bin_flank[k] = 0;
        b->max_n1 = b->mb_cols_base = b->mb_m = 0;
        CU(cudaMemcpyAsync(d_ids_a, ids.data(), nt * sizeof(int), cudaMemcpyHostToDevice, st));
        ref_plan2_kernel<<<(unsigned)((nt + T - 1) / T), T, 0, st>>>(d_ids_a, (int)nt, d_seq_off, d_lens, d_start, d_motif_off,
                                                                     d_motif_len, d_l_off, d_r_off, b->d_seq_off, b->d_lens,
                                                                     b->d_est, b->d_motif_off, b->d_motif_len,
                                                                     b->d_read_begin);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(st));  // `ids` is read by the copy above
        rc = batch_plan(ctx, b, arena_bytes);
        if (rc) return rc;
        rc = strk_batch_run(ctx, b, tier_key[3 * t], tier_key[3 * t + 1], tier_key[3 * t + 2], STRK_KERNEL_AUTO, nullptr);
        if (rc) return rc;
        for (int k = 0; k < 8; ++k) acc_stats[k] += ctx->stats[k];  // strk_batch_run reports its own run only
        res.resize(nt * 4);
        CU(cudaMemcpy(res.data(), b->d_out, nt * 4 * sizeof(int), cudaMemcpyDeviceToHost));
        for (size_t q = 0; q < nt; ++q) {
            const int l = ids[q];
            int32_t *o = out + 8 * (size_t)l;
            o[0] = res[4 * q], o[1] = res[4 * q + 1], o[2] = l_off[(size_t)l], o[3] = r_off[(size_t)l];
            o[4] = n_off[(size_t)l], o[5] = res[4 * q + 2];
            o[6] = lens[3 * l] - std::max(0, l_off[(size_t)l]);
            o[7] = lens[3 * l + 2] - std::max(0, r_off[(size_t)l]);
        }
    }
    for (int k = 0; k < 8; ++k) ctx->stats[k] = acc_stats[k];
    if (timing)
        fprintf(stderr, "[strk_ref_counts] %lld loci: upload + phase 1 %.2f ms, phase 2 %.2f ms (host clock); DP kernels %.2f ms, "
                "replay kernels %.2f ms (device clock), %d kernels\n", (long long)n_loci, t_phase1 - t_begin, now() - t_phase1,
                acc_stats[3], acc_stats[4], (int)acc_stats[2]);
    return STRK_OK;
}

// The fast path of strk_ref_counts: one search-parameter tier, first windows that fit the packed kernel.  Everything
// from the upload to the assembled results is queued on the context's stream without a host round trip in between:
// phase 1 (boundary tables + dual-score replay), phase 2 (the final count: the same loci as one "read" each through
// the read-path kernels -- their rows-per-lane class depends on the window's length only, which moving bases between
// flank and tract does not change, so the class lists built on the host for phase 1 serve both phases and no device-side
// planning pass is needed), and a kernel that assembles the 8 result ints per locus.  ONE synchronisation at the end
// brings back the results and the flags of the loci whose search left a window; those are returned in `redo`.
// (The general path takes 8 host synchronisations per call: harmless alone, but under another context's read kernels
// every one of them waits for SM slots that the persistent read CTAs free only at the end of a class launch.)
static int ref_counts_fast(strk_ctx *ctx, const uint8_t *arena, uint64_t arena_bytes, const uint64_t *seq_off,
                           const int32_t *lens, const int32_t *start_count, const int32_t *ref_size,
                           const int32_t *rc_params, int64_t n_loci, const uint64_t *motif_off, const int32_t *motif_len,
                           int vcf_anchor_size, int32_t *out, std::vector<int> &redo, bool *taken) {
    *taken = false;
    if (!ctx->h_consts.packed_ok) return STRK_OK;
    for (int64_t l = 1; l < n_loci; ++l)
        if (rc_params[3 * l] != rc_params[0] || rc_params[3 * l + 1] != rc_params[1] || rc_params[3 * l + 2] != rc_params[2])
            return STRK_OK;  // several tiers: the general path groups them
    const int max_iters = rc_params[0], range = rc_params[1], step = rc_params[2];
    int rc = validate_reads("strk_ref_counts", arena_bytes, seq_off, lens, n_loci);
    if (rc) return rc;
    rc = validate_motifs("strk_ref_counts", arena_bytes, motif_off, motif_len, n_loci);
    if (rc) return rc;
    if (max_iters < 0 || range < 0 || step < 0 || range > 1000 || step > 1000)
        return set_err(STRK_ERR_ARG, "strk_ref_counts: bad search parameters");
    static const int wd_env = getenv("STRK_REF_WD") ? atoi(getenv("STRK_REF_WD")) : 6;
    const int wd1 = std::max(wd_env > 0 ? wd_env : 6, range + step + 2);
    const int stride_w = 2 * wd1 + 1;
    int wd2 = range + step + 2;  // phase 2: the read path's first window
    if (wd2 < 6) wd2 = 6;
    const int W2 = 2 * wd2 + 1;
    if (2 * stride_w > PK_WINDOW_MAX || W2 > PK_WINDOW_MAX) return STRK_OK;
    const size_t n = (size_t)n_loci;
    int max_n1 = 0, mb_cols = 0, mb_m = 0;
    std::vector<int> lists[STRK_PK_NBIN];
    int mmax[STRK_PK_NBIN] = {0}, flank[STRK_PK_NBIN] = {0};
    double cells1 = 0.0;
    for (int64_t l = 0; l < n_loci; ++l) {
        if (start_count[l] < 0 || start_count[l] > (1 << 22) || (long long)motif_len[l] * (long long)start_count[l] > (1ll << 24))
            return set_err(STRK_ERR_ARG, "strk_ref_counts: bad start count / search parameters for locus %lld", (long long)l);
        const int fl = lens[3 * l], tr = lens[3 * l + 1], fr = lens[3 * l + 2], m = motif_len[l];
        const int n1 = fl + tr + fr;
        max_n1 = std::max(max_n1, n1);
        if (n1 > 32 * 16) {
            // boundary row of multi-strip sweeps: the longest candidate prefix of either phase (phase 2 moves at most
            // both flanks into the tract: (fl + fr) / m + 1 more copies)
            mb_cols = std::max(mb_cols, std::max(fl, fr) + m * (start_count[l] + (fl + fr) / m + 1));
            mb_m = std::max(mb_m, m);
        }
        int R = strk_pick_rows_packed(n1 + 1);
        if (fl < 1 || fr < 1 || fl > PK_FLANK_MAX || fr > PK_FLANK_MAX || m * R > 128) R = 0;
        lists[R].push_back((int)l);
        mmax[R] = std::max(mmax[R], m);
        flank[R] = std::max(flank[R], std::max(fl, fr));
        cells1 += (double)n1 * ((double)fl + fr + 2.0 * m * ((double)start_count[l] + wd1));
    }
    *taken = true;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    for (int k = 0; k < 8; ++k) ctx->stats[k] = 0;
    // ---- device buffers (context-owned, recycled)
    cudaError_t e = ctx->ref_arena.reserve((size_t)arena_bytes);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) e = ctx->ref_u64[k].reserve(n);
    const size_t isz[13] = {3 * n, n, n, 3 * n, n, n, n, n, n, n, n, 8, n};
    for (int k = 0; k < 13 && e == cudaSuccess; ++k) e = ctx->ref_i[k].reserve(isz[k]);
    if (e == cudaSuccess) e = ctx->ref_out.reserve(8 * n + 8);
    if (!ctx->ref_batch) ctx->ref_batch = new (std::nothrow) strk_batch();
    if (!ctx->ref_batch) return set_err(STRK_ERR_NOMEM, "out of host memory");
    strk_batch *b = ctx->ref_batch;
    if (e == cudaSuccess) e = b->seq_off.reserve(n);
    if (e == cudaSuccess) e = b->lens.reserve(3 * n);
    if (e == cudaSuccess) e = b->est.reserve(n);
    if (e == cudaSuccess) e = b->motif_off.reserve(n);
    if (e == cudaSuccess) e = b->motif_len.reserve(n);
    if (e == cudaSuccess) e = b->read_begin.reserve(n + 1);
    if (e == cudaSuccess) e = b->read_locus.reserve(n);
    if (e == cudaSuccess) e = b->out.reserve(4 * n);
    if (e == cudaSuccess) e = b->status.reserve(n);
    // second look at phase 1, still on the device: up to CAP2 loci whose search left the first window, 4x wider
    const int CAP2 = 4096;
    const int WD_MAX = (STRK_MAX_WINDOW - 1) / 2;
    const int wd1b = std::min(WD_MAX, 4 * wd1), stride_w2 = 2 * wd1b + 1;
    if (e == cudaSuccess) e = ctx->fams.reserve(std::max(n, (size_t)CAP2));
    if (e == cudaSuccess) e = ctx->table64.reserve(std::max(n * (size_t)stride_w * 2, (size_t)CAP2 * (size_t)stride_w2 * 2));
    if (e == cudaSuccess) e = ctx->table.reserve(n * (size_t)W2);
    if (e == cudaSuccess) e = ctx->fallback.reserve(n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_err(STRK_ERR_NOMEM, "strk_ref_counts: %s", cudaGetErrorString(e));
    }
    unsigned char *d_arena = ctx->ref_arena.p;
    unsigned long long *d_seq_off = ctx->ref_u64[0].p, *d_motif_off = ctx->ref_u64[1].p;
    int *d_lens = ctx->ref_i[0].p, *d_start = ctx->ref_i[1].p, *d_ref_size = ctx->ref_i[2].p, *d_rc = ctx->ref_i[3].p;
    int *d_motif_len = ctx->ref_i[4].p, *d_wd = ctx->ref_i[5].p, *d_l_off = ctx->ref_i[6].p, *d_r_off = ctx->ref_i[7].p;
    int *d_n_off = ctx->ref_i[8].p, *d_again = ctx->ref_i[9].p, *d_lists = ctx->ref_i[12].p;
    unsigned int *d_cnt = (unsigned int *)ctx->ref_i[11].p;
    ctx->ref_hwd.assign(n, wd1);
    ctx->ref_flat.clear();
    size_t at[STRK_PK_NBIN];
    for (int k = 0; k < STRK_PK_NBIN; ++k) {
        at[k] = ctx->ref_flat.size();
        ctx->ref_flat.insert(ctx->ref_flat.end(), lists[k].begin(), lists[k].end());
    }
    CU(cudaMemcpyAsync(d_arena, arena, (size_t)arena_bytes, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_seq_off, seq_off, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_motif_off, motif_off, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_lens, lens, 3 * n * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_start, start_count, n * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_ref_size, ref_size, n * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_rc, rc_params, 3 * n * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_motif_len, motif_len, n * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_wd, ctx->ref_hwd.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_lists, ctx->ref_flat.data(), ctx->ref_flat.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(d_l_off, 0, n * sizeof(int), st));
    CU(cudaMemsetAsync(d_r_off, 0, n * sizeof(int), st));
    CU(cudaMemsetAsync(d_n_off, 0, n * sizeof(int), st));
    CU(cudaMemsetAsync(d_cnt, 0, 8 * sizeof(unsigned int), st));
    CU(cudaMemsetAsync(ctx->d_acc, 0, 4 * sizeof(double), st));
    CU(cudaMemsetAsync(ctx->d_queue + 1, 0, 3 * sizeof(unsigned int), st));
    const int T = 128;
    const int b_len = max_n1 + 2;
    const int rowlen = max_n1 > 32 * 16 ? mb_cols + mb_m * std::max(std::max(wd1, wd2), wd1b) + 2 : 2;
    const int *seg_list[STRK_PK_NBIN];
    long long seg_cnt[STRK_PK_NBIN];
    int seg_w1[STRK_PK_NBIN], seg_w2[STRK_PK_NBIN];
    for (int k = 0; k < STRK_PK_NBIN; ++k) {
        seg_list[k] = d_lists + at[k];
        seg_cnt[k] = (long long)lists[k].size();
        seg_w1[k] = 2 * stride_w - 1;
        seg_w2[k] = (W2 + 3) / 4 * 4;
    }
    // ---- phase 1
    ref_plan1_kernel<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(nullptr, (int)n, d_seq_off, d_lens, d_start, d_wd, d_motif_off,
                                                               d_motif_len, stride_w, ctx->fams.p);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev[0], st));
    long long n_packed = 0;
    rc = launch_pass(ctx, seg_list, seg_cnt, flank, mmax, seg_w1, n_loci, ctx->fams.p, d_arena, ctx->table64.p, b_len, rowlen, st,
                     1, &n_packed);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev[1], st));
    ref_replay1_kernel<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(nullptr, (int)n, ctx->table64.p, ctx->fams.p, d_start, d_rc,
                                                                 d_ref_size, vcf_anchor_size, WD_MAX, d_wd, d_l_off, d_r_off,
                                                                 d_n_off, d_again, d_cnt);
    CU(cudaGetLastError());
    {
        // the loci whose search left the first window (a percent of a block at +-6 sizes): 4x wider through the general
        // kernel, list and count on the device (d_again, d_cnt[0]); what still does not settle is left to the host
        int *d_again2 = ctx->ref_i[10].p;
        ref_clamp_count_kernel<<<1, 32, 0, st>>>(d_cnt, (unsigned)CAP2);  // (loci beyond the cap keep their flag: host redo)
        ref_plan1_kernel<<<(unsigned)((CAP2 + T - 1) / T), T, 0, st>>>(d_again, CAP2, d_seq_off, d_lens, d_start, d_wd, d_motif_off,
                                                                      d_motif_len, stride_w2, ctx->fams.p, d_cnt);
        CU(cudaGetLastError());
        rc = launch_general(ctx, true, ctx->fams.p, nullptr, CAP2, d_arena, ctx->table64.p, b_len, rowlen, st, d_cnt);
        if (rc) return rc;
        ref_replay1_kernel<<<(unsigned)((CAP2 + T - 1) / T), T, 0, st>>>(d_again, CAP2, ctx->table64.p, ctx->fams.p, d_start, d_rc,
                                                                        d_ref_size, vcf_anchor_size, WD_MAX, d_wd, d_l_off,
                                                                        d_r_off, d_n_off, d_again2, d_cnt + 4, d_cnt);
        CU(cudaGetLastError());
    }
    // ---- phase 2 (loci whose phase 1 is not final yet run with their unadjusted flanks; their rows are redone)
    b->n_reads = b->n_loci = (long long)n;
    b->d_arena = d_arena;
    b->d_seq_off = b->seq_off.p, b->d_lens = b->lens.p, b->d_est = b->est.p, b->d_motif_off = b->motif_off.p;
    b->d_motif_len = b->motif_len.p, b->d_read_begin = b->read_begin.p, b->d_read_locus = b->read_locus.p;
    b->d_out = b->out.p, b->d_status = b->status.p;
    ref_plan2_kernel<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(nullptr, (int)n, d_seq_off, d_lens, d_start, d_motif_off,
                                                               d_motif_len, d_l_off, d_r_off, b->d_seq_off, b->d_lens, b->d_est,
                                                               b->d_motif_off, b->d_motif_len, b->d_read_begin, b->d_read_locus);
    CU(cudaGetLastError());
    plan_reads_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(nullptr, (long long)n, b->d_seq_off, b->d_lens, b->d_est,
                                                                  b->d_read_locus, b->d_motif_off, b->d_motif_len, wd2, 0, W2,
                                                                  ctx->fams.p, ctx->d_acc + 1, nullptr);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev[2], st));
    CU(cudaMemsetAsync(ctx->d_queue + 2, 0, sizeof(unsigned int), st));
    long long n_packed2 = 0;
    rc = launch_pass(ctx, seg_list, seg_cnt, flank, mmax, seg_w2, n_loci, ctx->fams.p, d_arena, ctx->table.p, b_len, rowlen, st, 0,
                     &n_packed2);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev[3], st));
    if (W2 <= REPLAY_WMAX)
        replay_reads_small_kernel<<<(unsigned)((n + REPLAY_THREADS - 1) / REPLAY_THREADS), REPLAY_THREADS, 0, st>>>(
            ctx->table.p, W2, wd2, 0, nullptr, nullptr, (int)n, b->d_read_begin, b->d_est, b->d_lens, b->d_motif_len, max_iters,
            range, step, ctx->tie_flags, b->d_out, b->d_status, ctx->d_queue + 1, ctx->d_acc, nullptr);
    else
        replay_reads_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(
            ctx->table.p, W2, wd2, 0, nullptr, nullptr, (int)n, b->d_read_begin, b->d_est, b->d_lens, b->d_motif_len, max_iters,
            range, step, ctx->tie_flags, b->d_out, b->d_status, ctx->d_queue + 1, ctx->d_acc, nullptr);
    CU(cudaGetLastError());
    ref_assemble_kernel<<<(unsigned)((n + T - 1) / T), T, 0, st>>>((int)n, b->d_out, b->d_status, d_lens, d_l_off, d_r_off, d_n_off,
                                                                  ctx->ref_out.p);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev[4], st));
    // ---- the one synchronisation: results, flags, counters
    std::vector<unsigned char> status(n);
    unsigned int h_cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double acc[2] = {0, 0};
    CU(cudaMemcpyAsync(out, ctx->ref_out.p, 8 * n * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(status.data(), b->d_status, n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(acc, ctx->d_acc, sizeof(acc), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (h_cnt[1] || h_cnt[5]) {
        const long long l = (long long)(0x7fffffffu - (h_cnt[1] ? h_cnt[1] : h_cnt[5]));
        return set_err(STRK_ERR_SEARCH, "strk_ref_counts: locus %lld scored no size (max_iters = %d)", l, max_iters);
    }
    // loci to redo through the general path: phase 1 left its window (flagged in the assembled row), or phase 2 did
    for (size_t l = 0; l < n; ++l)
        if (status[l] || out[8 * l + 4] < 0) redo.push_back((int)l);
    float t1 = 0.f, t2 = 0.f, t3 = 0.f, t4 = 0.f;
    CU(cudaEventElapsedTime(&t1, ctx->ev[0], ctx->ev[1]));
    CU(cudaEventElapsedTime(&t2, ctx->ev[1], ctx->ev[2]));
    CU(cudaEventElapsedTime(&t3, ctx->ev[2], ctx->ev[3]));
    CU(cudaEventElapsedTime(&t4, ctx->ev[3], ctx->ev[4]));
    ctx->stats[0] = cells1 + acc[1];
    ctx->stats[1] = acc[0];
    ctx->stats[2] += 7;  // (launch_* count their own)
    ctx->stats[3] = t1 + t3;
    ctx->stats[4] = t2 + t4;
    ctx->stats[5] = 0;
    if (getenv("STRK_REF_TIMING"))
        fprintf(stderr, "[strk_ref_counts] fast path, %lld loci: DP kernels %.2f + %.2f ms, replay / second look / planning kernels "
                "%.2f + %.2f ms (device clock), %u loci took the second look, %zu left to redo\n", (long long)n_loci, t1, t3, t2, t4,
                h_cnt[0], redo.size());
    return STRK_OK;
}

// ------------------------------------------------------------------------------------------------
// integer issue-rate micro-benchmark (roofline denominator)
// ------------------------------------------------------------------------------------------------
extern "C" int strk_measure_int_peak(strk_ctx *ctx, double out_tiops[3]) {
    if (!ctx || !out_tiops) return set_err(STRK_ERR_ARG, "strk_measure_int_peak: null argument");
    CU(cudaSetDevice(ctx->device));
    int *d_out = (int *)(ctx->d_queue + 7);  // context-owned scratch word and events: nothing to release on an error path
    const int grid = ctx->n_sm * 8, threads = 256, iters = 4096;
    const double instr = (double)grid * threads * (double)iters * 64.0;  // lane-level integer instructions
    cudaEvent_t e0 = ctx->ev[0], e1 = ctx->ev[1];
    for (int mode = 0; mode < 3; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            CU(cudaEventRecord(e0, ctx->stream));
            if (mode == 0) int_peak_kernel<0><<<grid, threads, 0, ctx->stream>>>(iters, rep + 3, d_out);
            if (mode == 1) int_peak_kernel<1><<<grid, threads, 0, ctx->stream>>>(iters, rep + 3, d_out);
            if (mode == 2) int_peak_kernel<2><<<grid, threads, 0, ctx->stream>>>(iters, rep + 3, d_out);
            CU(cudaGetLastError());
            CU(cudaEventRecord(e1, ctx->stream));
            CU(cudaEventSynchronize(e1));
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        out_tiops[mode] = instr / ((double)best * 1e-3) / 1e12;
    }
    return STRK_OK;
}

// ------------------------------------------------------------------------------------------------
// bootstrap / GMM allele calls (the consumer of the per-read counts)
// ------------------------------------------------------------------------------------------------
#include "alleles_api.cuh"
#include "realign_api.cuh"
