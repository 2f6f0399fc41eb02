// alleles_api.cuh -- C entry points of the bootstrap / GMM allele caller (included at the end of api.cu, which
// defines strk_ctx, set_err and CU).  Declarations and the reference lines they replace: include/strkit_b200.h.
#pragma once
#include "alleles.cuh"

static int alleles_params(AlleleParams *P, int n_alleles, int num_bootstrap, int min_reads, int min_allele_reads,
                          int force_gm_filter, double expansion_ratio, int filter_factor, int n_init, uint64_t seed) {
    if (n_alleles < 1 || n_alleles > 2)
        return set_err(STRK_ERR_UNSUPPORTED, "call_alleles: n_alleles = %d (1 and 2 are implemented)", n_alleles);
    if (num_bootstrap < 2 || num_bootstrap > ALL_MAX_BOOT)  // the reference itself fails for 1 (allele.py:163-167,258)
        return set_err(STRK_ERR_ARG, "call_alleles: num_bootstrap must be in 2..%d", ALL_MAX_BOOT);
    if (n_init < 1 || n_init > ALL_N_INIT_MAX || filter_factor < 1)
        return set_err(STRK_ERR_ARG, "call_alleles: bad GMM parameters (n_init %d, filter_factor %d)", n_init, filter_factor);
    P->n_alleles = n_alleles;
    P->num_bootstrap = num_bootstrap;
    P->min_reads = min_reads;
    P->n_init = n_init;
    P->max_iter = 100;
    P->force_gm_filter = force_gm_filter;
    P->tol = 1e-3;
    P->reg_covar = 1e-6;
    P->allele_filter = ((double)min_allele_reads - 0.1) / (double)num_bootstrap;
    P->expansion_ratio = expansion_ratio;
    P->filter_weight = 1.0 / ((double)filter_factor * 2.0);
    P->small_allele_min = 8.0;
    P->seed = seed;
    return STRK_OK;
}

struct AllDev {
    std::vector<void *> ptrs;
    ~AllDev() {
        for (void *p : ptrs) cudaFree(p);
    }
    template <typename T>
    cudaError_t get(T **dst, size_t n) {
        cudaError_t e = cudaMalloc((void **)dst, (n ? n : 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(*dst);
        return e;
    }
    template <typename T>
    cudaError_t up(T **dst, const T *src, size_t n, cudaStream_t st) {
        cudaError_t e = get(dst, n);
        if (e == cudaSuccess && n) e = cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, st);
        return e;
    }
};

template <int KMAX>
static void launch_fit(const int *vals, const double *cdf, const int *cnt, const int *K, const int *n, const int *status,
                       int n_loci, int kcap, const AlleleParams &P, double *rm, double *rw, double *rs,
                       unsigned char *rp, cudaStream_t st) {
    const long long threads = (long long)n_loci * P.num_bootstrap;
    alleles_fit_kernel<KMAX><<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(vals, cdf, cnt, K, n, status, n_loci, kcap,
                                                                                P, rm, rw, rs, rp);
}

extern "C" int strk_call_alleles(strk_ctx *ctx, const int32_t *cn, const double *weights, const int64_t *read_begin,
                                 int64_t n_loci, int n_alleles, int num_bootstrap, int min_reads, int min_allele_reads,
                                 int force_gm_filter, double expansion_ratio, int filter_factor, int n_init,
                                 uint64_t seed, int32_t *out_i, double *out_d, int32_t *out_status, double *ms_out) {
    if (!ctx || !read_begin || !out_i || !out_d || !out_status)
        return set_err(STRK_ERR_ARG, "strk_call_alleles: null argument");
    if (n_loci < 0 || n_loci > 0x7fffffffLL / 4096) return set_err(STRK_ERR_ARG, "strk_call_alleles: bad locus count");
    AlleleParams P;
    int rc = alleles_params(&P, n_alleles, num_bootstrap, min_reads, min_allele_reads, force_gm_filter, expansion_ratio,
                            filter_factor, n_init, seed);
    if (rc) return rc;
    if (n_loci == 0) return STRK_OK;
    const int64_t n_reads = read_begin[n_loci];
    if (read_begin[0] != 0 || n_reads < 0 || (n_reads && (!cn || !weights)))
        return set_err(STRK_ERR_ARG, "strk_call_alleles: read_begin must run from 0 to the number of reads");
    int max_n = 1;
    for (int64_t l = 0; l < n_loci; ++l) {
        const int64_t d = read_begin[l + 1] - read_begin[l];
        if (d < 0) return set_err(STRK_ERR_ARG, "strk_call_alleles: locus %lld has a non-monotone read_begin", (long long)l);
        if (d > 256)
            return set_err(STRK_ERR_UNSUPPORTED, "strk_call_alleles: locus %lld has %lld reads (limit 256; the reference "
                                                 "caps at max_reads = 250)", (long long)l, (long long)d);
        if (d > max_n) max_n = (int)d;
    }
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int kcap = max_n, A = n_alleles, B = num_bootstrap;
    // context-owned buffers, recycled across calls (no cudaMalloc / cudaFree in the steady state)
    cudaError_t e = cudaSuccess;
    auto take_i = [&](int slot, size_t n) -> int * {
        if (e == cudaSuccess) e = ctx->al_i[slot].reserve(n ? n : 1);
        return ctx->al_i[slot].p;
    };
    auto take_d = [&](int slot, size_t n) -> double * {
        if (e == cudaSuccess) e = ctx->al_d[slot].reserve(n ? n : 1);
        return ctx->al_d[slot].p;
    };
    int *d_cn = take_i(0, (size_t)n_reads), *d_vals = take_i(1, (size_t)n_loci * kcap);
    int *d_cnt = take_i(2, (size_t)n_loci * kcap), *d_K = take_i(3, (size_t)n_loci), *d_n = take_i(4, (size_t)n_loci);
    int *d_status = take_i(5, (size_t)n_loci), *d_kmax = take_i(6, 1), *d_oi = take_i(7, (size_t)n_loci * (1 + 5 * A));
    double *d_w = take_d(0, (size_t)n_reads), *d_cdf = take_d(1, (size_t)n_loci * kcap);
    double *d_od = take_d(2, (size_t)n_loci * 3 * A);
    if (e == cudaSuccess) e = ctx->al_rb.reserve((size_t)n_loci + 1);
    long long *d_rb = ctx->al_rb.p;
    if (e == cudaSuccess && n_reads) e = cudaMemcpyAsync(d_cn, cn, (size_t)n_reads * sizeof(int), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && n_reads) e = cudaMemcpyAsync(d_w, weights, (size_t)n_reads * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_rb, read_begin, ((size_t)n_loci + 1) * sizeof(long long), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_err(STRK_ERR_NOMEM, "strk_call_alleles: %s", cudaGetErrorString(e));
    }
    cudaEvent_t e0, e1;
    e0 = ctx->ev[0], e1 = ctx->ev[1];  // context-owned: nothing to release on an error path
    CU(cudaEventRecord(e0, st));
    CU(cudaMemsetAsync(d_kmax, 0, sizeof(int), st));
    alleles_prepare_kernel<<<(unsigned)((n_loci + 127) / 128), 128, 0, st>>>(d_cn, d_w, d_rb, (int)n_loci, min_reads, kcap,
                                                                             d_vals, d_cdf, d_cnt, d_K, d_n, d_status,
                                                                             d_kmax);
    CU(cudaGetLastError());
    int kmax = 0;
    CU(cudaMemcpyAsync(&kmax, d_kmax, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    // replicate estimates, in chunks of loci that keep the scratch near 256 MB (context-owned, recycled across calls)
    const size_t per_locus = (size_t)A * B * 3 * sizeof(double) + (size_t)B;
    int64_t chunk = (int64_t)((size_t)1 << 28) / (int64_t)per_locus;
    if (chunk < 1) chunk = 1;
    if (chunk > n_loci) chunk = n_loci;
    for (int k = 0; k < 3 && e == cudaSuccess; ++k) e = ctx->al_rep[k].reserve((size_t)chunk * A * B);
    if (e == cudaSuccess) e = ctx->al_peaks.reserve((size_t)chunk * B);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_err(STRK_ERR_NOMEM, "strk_call_alleles: %s", cudaGetErrorString(e));
    }
    double *rm = ctx->al_rep[0].p, *rw = ctx->al_rep[1].p, *rs = ctx->al_rep[2].p;
    unsigned char *rp = ctx->al_peaks.p;
    int np2 = 1;
    while (np2 < B) np2 <<= 1;
    const size_t agg_smem = (size_t)np2 * (sizeof(double) + sizeof(int));
    if (agg_smem > 48 * 1024)
        CU(cudaFuncSetAttribute(alleles_aggregate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)agg_smem));
    for (int64_t l0 = 0; l0 < n_loci; l0 += chunk) {
        const int nl = (int)std::min<int64_t>(chunk, n_loci - l0);
        const int *v = d_vals + (size_t)l0 * kcap, *c = d_cnt + (size_t)l0 * kcap;
        const double *p = d_cdf + (size_t)l0 * kcap;
        AlleleParams Pc = P;
        Pc.seed = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(l0 / chunk);  // distinct streams per chunk
        if (kmax <= 8)
            launch_fit<8>(v, p, c, d_K + l0, d_n + l0, d_status + l0, nl, kcap, Pc, rm, rw, rs, rp, st);
        else if (kmax <= 32)
            launch_fit<32>(v, p, c, d_K + l0, d_n + l0, d_status + l0, nl, kcap, Pc, rm, rw, rs, rp, st);
        else
            launch_fit<256>(v, p, c, d_K + l0, d_n + l0, d_status + l0, nl, kcap, Pc, rm, rw, rs, rp, st);
        CU(cudaGetLastError());
        alleles_aggregate_kernel<<<(unsigned)nl, 256, agg_smem, st>>>(rm, rw, rs, rp, d_status + l0, v, kcap, nl, A, B,
                                                                      d_oi + (size_t)l0 * (1 + 5 * A),
                                                                      d_od + (size_t)l0 * 3 * A);
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(e1, st));
    CU(cudaMemcpyAsync(out_i, d_oi, (size_t)n_loci * (1 + 5 * A) * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out_d, d_od, (size_t)n_loci * 3 * A * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out_status, d_status, (size_t)n_loci * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (ms_out) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        *ms_out = (double)ms;
    }
    return STRK_OK;
}

extern "C" int strk_gmm_fit_counts(strk_ctx *ctx, const double *x, const int32_t *counts, const int32_t *K,
                                   const int32_t *init, int64_t n_problems, int kcap, int n_alleles, int num_bootstrap,
                                   int min_allele_reads, int force_gm_filter, double expansion_ratio, int filter_factor,
                                   int n_init, double *out) {
    if (!ctx || !x || !counts || !K || !init || !out) return set_err(STRK_ERR_ARG, "strk_gmm_fit_counts: null argument");
    if (n_problems < 0 || n_problems > 0x7fffffff || kcap < 1 || kcap > 256)
        return set_err(STRK_ERR_ARG, "strk_gmm_fit_counts: bad sizes");
    AlleleParams P;
    int rc = alleles_params(&P, n_alleles, num_bootstrap, 0, min_allele_reads, force_gm_filter, expansion_ratio,
                            filter_factor, n_init, 0);
    if (rc) return rc;
    if (n_problems == 0) return STRK_OK;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    AllDev dev;
    double *d_x = nullptr, *d_out = nullptr;
    int *d_c = nullptr, *d_K = nullptr, *d_init = nullptr;
    cudaError_t e = dev.up(&d_x, x, (size_t)n_problems * kcap, st);
    if (e == cudaSuccess) e = dev.up(&d_c, (const int *)counts, (size_t)n_problems * kcap, st);
    if (e == cudaSuccess) e = dev.up(&d_K, (const int *)K, (size_t)n_problems, st);
    if (e == cudaSuccess) e = dev.up(&d_init, (const int *)init, (size_t)n_problems * 2 * n_init, st);
    if (e == cudaSuccess) e = dev.get(&d_out, (size_t)n_problems * 7);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_err(STRK_ERR_NOMEM, "strk_gmm_fit_counts: %s", cudaGetErrorString(e));
    }
    for (int64_t q = 0; q < n_problems; ++q)
        if (K[q] < 1 || K[q] > kcap) return set_err(STRK_ERR_ARG, "strk_gmm_fit_counts: problem %lld has K = %d", (long long)q, K[q]);
    const unsigned grid = (unsigned)((n_problems + 127) / 128);
    if (kcap <= 8)
        gmm_fit_counts_kernel<8><<<grid, 128, 0, st>>>(d_x, d_c, d_K, d_init, (int)n_problems, kcap, P, d_out);
    else if (kcap <= 32)
        gmm_fit_counts_kernel<32><<<grid, 128, 0, st>>>(d_x, d_c, d_K, d_init, (int)n_problems, kcap, P, d_out);
    else
        gmm_fit_counts_kernel<256><<<grid, 128, 0, st>>>(d_x, d_c, d_K, d_init, (int)n_problems, kcap, P, d_out);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, d_out, (size_t)n_problems * 7 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return STRK_OK;
}

extern "C" int strk_alleles_aggregate(strk_ctx *ctx, const double *rep_means, const double *rep_weights,
                                      const double *rep_stdevs, const uint8_t *rep_peaks, int64_t n_loci, int n_alleles,
                                      int num_bootstrap, int32_t *out_i, double *out_d) {
    if (!ctx || !rep_means || !rep_weights || !rep_stdevs || !rep_peaks || !out_i || !out_d)
        return set_err(STRK_ERR_ARG, "strk_alleles_aggregate: null argument");
    if (n_alleles < 1 || n_alleles > 2 || num_bootstrap < 1 || num_bootstrap > ALL_MAX_BOOT || n_loci < 0 ||
        n_loci > 0x7fffffff)
        return set_err(STRK_ERR_ARG, "strk_alleles_aggregate: bad sizes");
    if (n_loci == 0) return STRK_OK;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int A = n_alleles, B = num_bootstrap;
    AllDev dev;
    double *rm = nullptr, *rw = nullptr, *rs = nullptr, *d_od = nullptr;
    unsigned char *rp = nullptr;
    int *d_status = nullptr, *d_oi = nullptr;
    const size_t nrep = (size_t)n_loci * A * B;
    cudaError_t e = dev.up(&rm, rep_means, nrep, st);
    if (e == cudaSuccess) e = dev.up(&rw, rep_weights, nrep, st);
    if (e == cudaSuccess) e = dev.up(&rs, rep_stdevs, nrep, st);
    if (e == cudaSuccess) e = dev.up(&rp, (const unsigned char *)rep_peaks, (size_t)n_loci * B, st);
    if (e == cudaSuccess) e = dev.get(&d_status, (size_t)n_loci);
    if (e == cudaSuccess) e = dev.get(&d_oi, (size_t)n_loci * (1 + 5 * A));
    if (e == cudaSuccess) e = dev.get(&d_od, (size_t)n_loci * 3 * A);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_err(STRK_ERR_NOMEM, "strk_alleles_aggregate: %s", cudaGetErrorString(e));
    }
    CU(cudaMemsetAsync(d_status, 0, (size_t)n_loci * sizeof(int), st));
    int np2 = 1;
    while (np2 < B) np2 <<= 1;
    const size_t agg_smem = (size_t)np2 * (sizeof(double) + sizeof(int));
    if (agg_smem > 48 * 1024)
        CU(cudaFuncSetAttribute(alleles_aggregate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)agg_smem));
    alleles_aggregate_kernel<<<(unsigned)n_loci, 256, agg_smem, st>>>(rm, rw, rs, rp, d_status, nullptr, 1, (int)n_loci, A,
                                                                      B, d_oi, d_od);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_i, d_oi, (size_t)n_loci * (1 + 5 * A) * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out_d, d_od, (size_t)n_loci * 3 * A * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return STRK_OK;
}
