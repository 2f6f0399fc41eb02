// alleles.cuh -- bootstrap resampling + Gaussian-mixture allele calling, batched over loci.
//
// Follows the reference's call_alleles (strkit/call/allele.py:176-336) with its helpers
// get_resampled_bootstrapped_reads (:126-173, the separate_strands=False branch both call sites use,
// call_locus.py:201-214,255-268), fit_gmm (:56-123), make_single_gaussian / GMMParams.make_fitted_gmm
// (strkit/call/gmm.py:59-80) and, underneath, scikit-learn 1.9.0's GaussianMixture (spherical, k-means++
// init, n_init restarts, tol 1e-3, max_iter 100, reg_covar 1e-6) restated in float64:
//
//   alleles_prepare_kernel   one thread per locus: distinct copy numbers, their resampling probabilities
//   alleles_fit_kernel       one thread per (locus, bootstrap replicate): multinomial resample (counter-based
//                            Philox RNG), k-means++ seeding, EM, the reference's peak filters
//   alleles_aggregate_kernel one CTA per locus: per-allele stable sort of the replicate estimates,
//                            interpolated-inverted-CDF percentiles, median, modal peak count
//
// A bootstrap replicate of integer copy numbers is a multiset over the locus' K distinct values, so it is
// carried as K counts; EM on (value, count) pairs is the same arithmetic as EM on the expanded sample
// (identical points have identical responsibilities).  Random streams differ from numpy's / sklearn's, so
// results agree with the reference statistically, not bit for bit; the deterministic parts (EM given the
// seeds, the aggregation) are tested exactly.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define ALL_MAX_BOOT 4096   // replicates per locus the aggregation kernel sorts in shared memory
#define ALL_N_INIT_MAX 8

struct AlleleParams {
    int n_alleles;          // 1 or 2
    int num_bootstrap;
    int min_reads;
    int n_init;             // GMMParams.n_init (params.py:166-172: 3)
    int max_iter;           // sklearn default 100
    int force_gm_filter;
    double tol;             // sklearn default 1e-3
    double reg_covar;       // sklearn default 1e-6
    double allele_filter;   // (min_allele_reads - 0.1) / num_bootstrap  (allele.py:243: concat_samples.shape[0])
    double expansion_ratio; // params.gm_filter_expansion_ratio
    double filter_weight;   // 1 / (filter_factor * 2)
    double small_allele_min;  // allele.py:47
    unsigned long long seed;
};

// ---------------------------------------------------------------------------------------------- RNG
struct Philox {
    uint32_t key0, key1, c0, c1, c2, c3;
    uint32_t out[4];
    int have;
    __device__ __forceinline__ void init(unsigned long long seed, uint32_t a, uint32_t b) {
        key0 = (uint32_t)seed;
        key1 = (uint32_t)(seed >> 32);
        c0 = 0;
        c1 = 0;
        c2 = a;
        c3 = b;
        have = 0;
    }
    __device__ __forceinline__ void round(uint32_t &x0, uint32_t &x1, uint32_t &x2, uint32_t &x3, uint32_t k0,
                                          uint32_t k1) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
        const uint32_t y0 = hi1 ^ x1 ^ k0, y1 = lo1, y2 = hi0 ^ x3 ^ k1, y3 = lo0;
        x0 = y0, x1 = y1, x2 = y2, x3 = y3;
    }
    __device__ __forceinline__ void refill() {
        uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3, k0 = key0, k1 = key1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            round(x0, x1, x2, x3, k0, k1);
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        out[0] = x0, out[1] = x1, out[2] = x2, out[3] = x3;
        if (++c0 == 0) ++c1;
        have = 4;
    }
    // uniform double in [0, 1), 53 bits
    __device__ __forceinline__ double uniform() {
        if (have < 2) refill();
        const uint32_t a = out[have - 1], b = out[have - 2];
        have -= 2;
        const unsigned long long u = ((unsigned long long)a << 32) | b;
        return (double)(u >> 11) * (1.0 / 9007199254740992.0);
    }
};

// ---------------------------------------------------------------------------------------------- prepare
// status: 0 = bootstrap + GMM, 1 = fewer than min_reads reads (reference returns None, allele.py:192-193),
//         2 = a single distinct value (no bootstrap, allele.py:196-214)
__global__ void alleles_prepare_kernel(const int *__restrict__ cn, const double *__restrict__ w,
                                       const long long *__restrict__ read_begin, int n_loci, int min_reads, int kcap,
                                       int *__restrict__ vals, double *__restrict__ cdf, int *__restrict__ cnt,
                                       int *__restrict__ K_out,
                                       int *__restrict__ n_out, int *__restrict__ status, int *__restrict__ k_max) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_loci) return;
    const long long r0 = read_begin[l], r1 = read_begin[l + 1];
    const int n = (int)(r1 - r0);
    n_out[l] = n;
    int *v = vals + (size_t)l * kcap;
    double *p = cdf + (size_t)l * kcap;
    int *q = cnt + (size_t)l * kcap;
    int K = 0;
    // insertion into a sorted list of distinct values, probabilities and read counts accumulated per value
    // (kcap >= the longest locus of the batch, so the list cannot overflow)
    for (long long r = r0; r < r1; ++r) {
        const int x = cn[r];
        const double wx = w[r];
        int lo = 0;
        while (lo < K && v[lo] < x) ++lo;
        if (lo < K && v[lo] == x) {
            p[lo] += wx;
            q[lo] += 1;
        } else if (K < kcap) {
            for (int j = K; j > lo; --j) v[j] = v[j - 1], p[j] = p[j - 1], q[j] = q[j - 1];
            v[lo] = x;
            p[lo] = wx;
            q[lo] = 1;
            ++K;
        }
    }
    // cumulative, normalised by the total like numpy's Generator.choice (cdf /= cdf[-1])
    double acc = 0.0;
    for (int j = 0; j < K; ++j) {
        acc += p[j];
        p[j] = acc;
    }
    for (int j = 0; j < K; ++j) p[j] /= acc;
    K_out[l] = K;
    status[l] = n < min_reads ? 1 : (K == 1 ? 2 : 0);
    atomicMax(k_max, K);
}

// ---------------------------------------------------------------------------------------------- GMM
struct Gmm2 {
    double mean[2], cov[2], weight[2];
    int n_comp;
};

// sklearn _estimate_log_gaussian_prob (spherical, one feature) + log weights, for one x
__device__ __forceinline__ void gmm_wlp(double x, const double mean[2], const double pchol[2], const double logw[2],
                                        double out[2]) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        // separately rounded products and sums, in sklearn's order: the three terms cancel almost completely for
        // a point near the mean (precision up to 1e6), so a fused multiply-add here would change the low bits
        const double prec = __dmul_rn(pchol[k], pchol[k]);
        const double t1 = __dmul_rn(__dmul_rn(mean[k], mean[k]), prec);
        const double t2 = __dmul_rn(2.0, __dmul_rn(__dmul_rn(x, mean[k]), prec));
        const double t3 = __dmul_rn(__dmul_rn(x, x), prec);
        const double lp = __dadd_rn(__dsub_rn(t1, t2), t3);
        out[k] = __dadd_rn(__dadd_rn(__dmul_rn(-0.5, __dadd_rn(1.8378770664093453, lp)), log(pchol[k])), logw[k]);
    }
}

// EM from the k-means++ seeds (value indices i0, i1): GaussianMixture._initialize with one-hot responsibilities on
// the two seed points, then e-step / m-step until |change of the lower bound| < tol (sklearn BaseMixture.fit_predict).
template <int KMAX>
__device__ double gmm_em(const double *x, const int *c, int K, int n, int i0, int i1, const AlleleParams &P, Gmm2 &g,
                         int *n_iter_out) {
    const double eps10 = 10.0 * 2.220446049250313e-16;
    double mean[2], cov[2], weight[2], pchol[2], logw[2];
    {
        const int idx[2] = {i0, i1};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const double nk = 1.0 + eps10, xv = x[idx[k]];
            mean[k] = xv / nk;
            cov[k] = __dadd_rn(__dsub_rn(__ddiv_rn(__dmul_rn(xv, xv), nk), __dmul_rn(mean[k], mean[k])), P.reg_covar);
            weight[k] = nk / (double)n;
            pchol[k] = 1.0 / sqrt(cov[k]);
            logw[k] = log(weight[k]);
        }
    }
    double lower = -INFINITY;
    int it = 1;
    for (; it <= P.max_iter; ++it) {
        const double prev = lower;
        double nk[2] = {0.0, 0.0}, sx[2] = {0.0, 0.0}, sxx[2] = {0.0, 0.0}, ll = 0.0;
        for (int v = 0; v < K; ++v) {
            if (c[v] == 0) continue;
            double wl[2];
            gmm_wlp(x[v], mean, pchol, logw, wl);
            const double mx = fmax(wl[0], wl[1]);
            const double lse = mx + log(exp(wl[0] - mx) + exp(wl[1] - mx));
            const double cv = (double)c[v];
            ll += cv * lse;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const double r = exp(wl[k] - lse) * cv;
                nk[k] += r;
                sx[k] += r * x[v];
                sxx[k] += r * x[v] * x[v];
            }
        }
        lower = ll / (double)n;
        double wsum = 0.0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            nk[k] += eps10;
            mean[k] = sx[k] / nk[k];
            cov[k] = __dadd_rn(__dsub_rn(__ddiv_rn(sxx[k], nk[k]), __dmul_rn(mean[k], mean[k])), P.reg_covar);
            weight[k] = nk[k] / (double)n;
            wsum += weight[k];
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            weight[k] /= wsum;
            // sklearn raises on cov <= 0 (ill-defined covariance); it cannot happen for integer data of this
            // range with reg_covar = 1e-6, the clamp only keeps the arithmetic finite
            pchol[k] = 1.0 / sqrt(cov[k] > 0.0 ? cov[k] : P.reg_covar);
            logw[k] = log(weight[k]);
        }
        if (fabs(lower - prev) < P.tol) break;
    }
    if (it > P.max_iter) it = P.max_iter;
#pragma unroll
    for (int k = 0; k < 2; ++k) g.mean[k] = mean[k], g.cov[k] = cov[k], g.weight[k] = weight[k];
    g.n_comp = 2;
    if (n_iter_out) *n_iter_out = it;
    return lower;
}

// sklearn.cluster.kmeans_plusplus for two centres on the (sorted) replicate: first centre uniform over the n points,
// second = the better of 2 + int(log 2) = 2 candidates drawn proportionally to the squared distance.
template <int KMAX>
__device__ void kmeanspp2(const double *x, const int *c, int K, int n, Philox &rng, int &i0, int &i1) {
    // first centre: point index floor(u * n) -> value index  (Generator-free restatement of RandomState.choice)
    {
        const double u = rng.uniform();
        long long pt = (long long)(u * (double)n);
        if (pt >= n) pt = n - 1;
        int v = 0;
        long long acc = 0;
        for (; v < K; ++v) {
            acc += c[v];
            if (pt < acc) break;
        }
        i0 = v < K ? v : K - 1;
    }
    double pot = 0.0;
    for (int v = 0; v < K; ++v) {
        const double d = x[v] - x[i0];
        pot += (double)c[v] * d * d;
    }
    double best_pot = INFINITY;
    int best = i0;
#pragma unroll 1
    for (int t = 0; t < 2; ++t) {
        const double r = rng.uniform() * pot;
        int v = 0;
        double acc = 0.0;
        int last_nonzero = 0;
        for (; v < K; ++v) {
            if (c[v]) last_nonzero = v;
            const double d = x[v] - x[i0];
            acc += (double)c[v] * d * d;
            if (c[v] && acc >= r) break;
        }
        const int cand = v < K ? v : last_nonzero;
        double cp = 0.0;
        for (int q = 0; q < K; ++q) {
            const double d0 = x[q] - x[i0], d1 = x[q] - x[cand];
            cp += (double)c[q] * fmin(d0 * d0, d1 * d1);
        }
        if (cp < best_pot) best_pot = cp, best = cand;
    }
    i1 = best;
}

// make_single_gaussian (gmm.py:72-80): mean and population variance of the replicate
__device__ __forceinline__ void single_gaussian(const double *x, const int *c, int K, int n, double &mean, double &var) {
    double s = 0.0;
    for (int v = 0; v < K; ++v) s += (double)c[v] * x[v];
    mean = s / (double)n;
    double q = 0.0;
    for (int v = 0; v < K; ++v) {
        const double d = x[v] - mean;
        q += (double)c[v] * d * d;
    }
    var = q / (double)n;
}

// fit_gmm's peak filters (allele.py:88-121) on a fitted 2-component model: number of useless components
__device__ __forceinline__ int gmm_useless(const Gmm2 &g, const AlleleParams &P) {
    const double lo = fmin(g.mean[0], g.mean[1]), hi = fmax(g.mean[0], g.mean[1]);
    const bool f2_strict = P.force_gm_filter || hi < P.expansion_ratio * fmax(lo, P.small_allele_min);
    const double thr2 = f2_strict ? P.filter_weight : 1.1920928955078125e-07;  // np.finfo(np.float32).eps
    int useless = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k)
        if (!(g.weight[k] > P.allele_filter && g.weight[k] > thr2)) ++useless;
    return useless;
}

// One replicate -> the allele estimates the reference appends per bootstrap iteration (allele.py:249-293):
// out_m / out_w / out_s [n_alleles] sorted by mean, *n_peaks = g.means_.shape[0].
template <int KMAX>
__device__ void fit_replicate(const double *x, const int *c, int K, int n, const AlleleParams &P, Philox &rng,
                              const int *forced_init, double *out_m, double *out_w, double *out_s, int *n_peaks) {
    int distinct = 0;
    for (int v = 0; v < K; ++v) distinct += c[v] != 0;
    bool single = distinct == 1 || P.n_alleles == 1;  // fit_gmm: len(mc) == 1, or n_components == 1
    Gmm2 best;
    if (!single) {
        double best_lb = -INFINITY;
        for (int t = 0; t < P.n_init; ++t) {
            int i0, i1;
            if (forced_init) {
                i0 = forced_init[2 * t];
                i1 = forced_init[2 * t + 1];
            } else {
                kmeanspp2<KMAX>(x, c, K, n, rng, i0, i1);
            }
            Gmm2 g;
            const double lb = gmm_em<KMAX>(x, c, K, n, i0, i1, P, g, nullptr);
            if (lb > best_lb || best_lb == -INFINITY) best_lb = lb, best = g;
        }
        // while n_components > 0: 2 -> (2 - n_useless); 1 -> single Gaussian; 0 -> the loop ends and g is returned
        const int useless = gmm_useless(best, P);
        if (useless == 1) single = true;
    }
    if (single) {
        double m, var;
        single_gaussian(x, c, K, n, m, var);
        const double s = sqrt(var);
        for (int a = 0; a < P.n_alleles; ++a) out_m[a] = m, out_w[a] = 1.0, out_s[a] = s;
        *n_peaks = 1;
    } else {
        const int first = best.mean[1] < best.mean[0] ? 1 : 0;  // stable argsort of two means
        out_m[0] = best.mean[first], out_w[0] = best.weight[first], out_s[0] = sqrt(best.cov[first]);
        out_m[1] = best.mean[1 - first], out_w[1] = best.weight[1 - first], out_s[1] = sqrt(best.cov[1 - first]);
        *n_peaks = 2;
    }
}

// one thread per (locus, replicate); replicate arrays laid out [locus][allele][replicate]
template <int KMAX>
__global__ void __launch_bounds__(128)
alleles_fit_kernel(const int *__restrict__ vals, const double *__restrict__ cdf, const int *__restrict__ cnt,
                   const int *__restrict__ K_arr,
                   const int *__restrict__ n_arr, const int *__restrict__ status, int n_loci, int kcap, AlleleParams P,
                   double *__restrict__ rep_m, double *__restrict__ rep_w, double *__restrict__ rep_s,
                   unsigned char *__restrict__ rep_peaks) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int B = P.num_bootstrap;
    const int l = (int)(tid / B), b = (int)(tid % B);
    if (l >= n_loci || status[l] != 0) return;
    const int K = K_arr[l], n = n_arr[l];
    if (K > KMAX) return;  // the host picks KMAX >= the batch maximum
    double x[KMAX];
    int c[KMAX];
    const int *v = vals + (size_t)l * kcap;
    const double *p = cdf + (size_t)l * kcap;
    for (int j = 0; j < K; ++j) x[j] = (double)v[j], c[j] = 0;
    Philox rng;
    rng.init(P.seed, (uint32_t)l, (uint32_t)b);
    if (B > 1) {
        // Generator.choice(replace=True, p): index = cdf.searchsorted(uniform, side="right")
        for (int i = 0; i < n; ++i) {
            const double u = rng.uniform();
            int j = 0;
            while (j < K - 1 && p[j] <= u) ++j;
            ++c[j];
        }
    } else {
        // num_bootstrap == 1: the replicate is the sample itself (allele.py:163-167)
        for (int j = 0; j < K; ++j) c[j] = cnt[(size_t)l * kcap + j];
    }
    double m[2], w[2], s[2];
    int peaks;
    fit_replicate<KMAX>(x, c, K, n, P, rng, nullptr, m, w, s, &peaks);
    for (int a = 0; a < P.n_alleles; ++a) {
        const size_t o = ((size_t)l * P.n_alleles + a) * B + b;
        rep_m[o] = m[a], rep_w[o] = w[a], rep_s[o] = s[a];
    }
    rep_peaks[(size_t)l * B + b] = (unsigned char)peaks;
}

// deterministic building block (tests, and callers that bring their own replicates): problem q has K[q] values
// x[q*kcap ..], counts c[q*kcap ..], and n_init forced seed pairs init[q*2*n_init ..]
template <int KMAX>
__global__ void gmm_fit_counts_kernel(const double *__restrict__ xs, const int *__restrict__ cs, const int *__restrict__ Ks,
                                      const int *__restrict__ init, int n_problems, int kcap, AlleleParams P,
                                      double *__restrict__ out /* [q][7]: m0 w0 s0 m1 w1 s1 peaks */) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_problems) return;
    const int K = Ks[q];
    if (K > KMAX) return;
    double x[KMAX];
    int c[KMAX], n = 0;
    for (int j = 0; j < K; ++j) x[j] = xs[(size_t)q * kcap + j], c[j] = cs[(size_t)q * kcap + j], n += c[j];
    Philox rng;
    rng.init(P.seed, (uint32_t)q, 0u);
    double m[2] = {0, 0}, w[2] = {0, 0}, s[2] = {0, 0};
    int peaks = 0;
    fit_replicate<KMAX>(x, c, K, n, P, rng, init + (size_t)q * 2 * P.n_init, m, w, s, &peaks);
    double *o = out + (size_t)q * 7;
    o[0] = m[0], o[1] = w[0], o[2] = s[0], o[3] = m[1], o[4] = w[1], o[5] = s[1], o[6] = (double)peaks;
}

// ---------------------------------------------------------------------------------------------- aggregate
// np.percentile(..., method="interpolated_inverted_cdf") on an ascending array a[0..n): virtual index n*q - 1
__device__ __forceinline__ double pct_iicdf(const double *a, int n, double q) {
    double vi = (double)n * q - 1.0;
    if (vi < 0.0) vi = 0.0;
    if (vi > (double)(n - 1)) vi = (double)(n - 1);
    const int lo = (int)floor(vi);
    const int hi = lo + 1 < n ? lo + 1 : n - 1;
    const double g = vi - (double)lo;
    // numpy _lerp: a + (b - a) * t, switched to b - (b - a) * (1 - t) for t >= 0.5
    const double d = a[hi] - a[lo];
    return g >= 0.5 ? a[hi] - d * (1.0 - g) : a[lo] + d * g;
}

// out_i [locus][1 + 2*A + 4*A]: modal_n, call[A], ci95[A][2], ci99[A][2]      out_d [locus][3*A]: means, weights, stdevs
__global__ void __launch_bounds__(256)
alleles_aggregate_kernel(const double *__restrict__ rep_m, const double *__restrict__ rep_w, const double *__restrict__ rep_s,
                         const unsigned char *__restrict__ rep_peaks, const int *__restrict__ status,
                         const int *__restrict__ vals, int kcap, int n_loci, int A, int B, int *__restrict__ out_i,
                         double *__restrict__ out_d) {
    extern __shared__ unsigned char smem_raw_all[];
    const int l = blockIdx.x;
    if (l >= n_loci) return;
    int np2 = 1;
    while (np2 < B) np2 <<= 1;
    double *key = (double *)smem_raw_all;  // [np2] replicate means of one allele
    int *idx = (int *)(key + np2);         // [np2] replicate index (tie-break = stable order)
    int *oi = out_i + (size_t)l * (1 + 5 * A);
    double *od = out_d + (size_t)l * (3 * A);
    const int st = status[l];
    if (st != 0) {
        if (threadIdx.x == 0) {
            const int cnv = st == 2 ? vals[(size_t)l * kcap] : 0;
            oi[0] = st == 2 ? 1 : 0;
            for (int a = 0; a < A; ++a) {
                oi[1 + a] = cnv;
                oi[1 + A + 2 * a] = oi[1 + A + 2 * a + 1] = cnv;
                oi[1 + 3 * A + 2 * a] = oi[1 + 3 * A + 2 * a + 1] = cnv;
                od[a] = (double)cnv;
                od[A + a] = st == 2 ? 1.0 / (double)A : 0.0;
                od[2 * A + a] = 0.0;
            }
        }
        return;
    }
    __shared__ int s_cnt[3];
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x) atomicAdd(&s_cnt[rep_peaks[(size_t)l * B + b] == 2 ? 2 : 1], 1);
    __syncthreads();
    double wsel[2] = {0.0, 0.0};
    for (int a = 0; a < A; ++a) {
        const double *m = rep_m + ((size_t)l * A + a) * B;
        for (int b = threadIdx.x; b < np2; b += blockDim.x) {
            key[b] = b < B ? m[b] : INFINITY;
            idx[b] = b;
        }
        __syncthreads();
        // bitonic sort of (key, original index): lexicographic order = numpy's stable argsort
        for (int k = 2; k <= np2; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = threadIdx.x; t < np2; t += blockDim.x) {
                    const int u = t ^ j;
                    if (u > t) {
                        const bool up = (t & k) == 0;
                        const double ka = key[t], kb = key[u];
                        const int ia = idx[t], ib = idx[u];
                        const bool gt = ka > kb || (ka == kb && ia > ib);
                        if (gt == up) key[t] = kb, key[u] = ka, idx[t] = ib, idx[u] = ia;
                    }
                }
                __syncthreads();
            }
        if (threadIdx.x == 0) {
            const int med = B / 2;
            const size_t o = ((size_t)l * A + a) * B + idx[med];
            od[a] = key[med];
            wsel[a] = rep_w[o];
            od[2 * A + a] = rep_s[o];
            oi[1 + a] = (int)rint(key[med]);
            oi[1 + A + 2 * a] = (int)rint(pct_iicdf(key, B, 0.025));
            oi[1 + A + 2 * a + 1] = (int)rint(pct_iicdf(key, B, 0.975));
            oi[1 + 3 * A + 2 * a] = (int)rint(pct_iicdf(key, B, 0.005));
            oi[1 + 3 * A + 2 * a + 1] = (int)rint(pct_iicdf(key, B, 0.995));
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double ws = 0.0;
        for (int a = 0; a < A; ++a) ws += wsel[a];
        for (int a = 0; a < A; ++a) od[A + a] = wsel[a] / ws;
        // statistics.mode of the sorted peak counts: the smallest of the most common
        oi[0] = s_cnt[1] >= s_cnt[2] ? 1 : 2;
    }
}
