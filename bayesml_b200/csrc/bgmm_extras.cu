// Widening rows of SURVEY.md §8(f): per-row work either side of the VB loop that the reference does on the host.
//
//   bgmm_pred_logdensity  f3: ln p(x_new | x^n) of every row under the predictive mixture of Student-t distributions
//                         (formula: /root/reference/bayesml/gaussianmixture/__init__.py:86-97; the reference evaluates the
//                         same density through scipy.stats.multivariate_t at _gaussianmixture.py:1086-1099, one point at a
//                         time).  The quadratic forms (x - mu_k)^T Lambda_k (x - mu_k) come from bgmm_pass (ln rho output of
//                         a parameter set whose Lambda is p_lambda_mats); this kernel is the Student-t / log-sum-exp epilogue.
//   bgmm_dirichlet1       f2: r_n ~ Dirichlet(1_K) per row on the device (`_init_random_responsibility` :734-736 draws them
//                         with numpy's PCG64 + ziggurat stream, 3.2e8 variates on the host at BASELINE config C2).  Counter-
//                         based Philox4x32-10 keyed by (seed, global row): the draw of a row does not depend on how the rows
//                         are sharded.  NOT the reference's random stream — an opt-in (`device_init=True`); same distribution.
#include "bgmm_common.cuh"
#include <math.h>

namespace bgmm {

__global__ void __launch_bounds__(256) pred_logdensity_kernel(const double* __restrict__ lnrho, const int64_t n, const int K,
                                                              const double* __restrict__ acst, const double* __restrict__ ck,
                                                              const double* __restrict__ hk, const double* __restrict__ nuk,
                                                              double* __restrict__ out) {
    // one warp per row: lane k (k += 32) evaluates component k, then a log-sum-exp over the warp
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    double mx = -INFINITY, t[2] = {-INFINITY, -INFINITY};       // K <= 64
    for (int k = lane, s = 0; k < K; k += 32, ++s) {
        // ln rho = a_k - 0.5 * Delta^2  ->  Delta^2 >= 0 up to rounding
        const double d2 = fmax(0.0, -2.0 * (lnrho[row * K + k] - acst[k]));
        t[s] = ck[k] - hk[k] * log1p(d2 / nuk[k]);
        mx = fmax(mx, t[s]);
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    double sum = 0.0;
    for (int s = 0; s < 2; ++s)
        if (t[s] > -INFINITY) sum += exp(t[s] - mx);
    sum = warp_sum(sum);
    if (lane == 0) out[row] = mx + log(sum);
}

// Philox4x32-10 (Salmon et al. 2011): counter (c0..c3), key (k0, k1) -> 4 x 32 random bits
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ key.x, lo1, hi0 ^ c.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return c;
}

__device__ __forceinline__ double uniform_open(uint32_t hi, uint32_t lo) {       // (0, 1): 53 random bits, never 0
    const unsigned long long u = ((unsigned long long)hi << 32) | lo;
    return ((double)(u >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

__global__ void __launch_bounds__(256) dirichlet1_kernel(double* __restrict__ r, const int64_t n, const int K,
                                                         const unsigned long long seed, const int64_t row_offset) {
    // one thread per row: K standard exponentials -ln(u), normalised (Dirichlet(1,..,1) = normalised Gamma(1) draws)
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const unsigned long long g = (unsigned long long)(row_offset + i);
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    double* out = r + i * K;
    double sum = 0.0;
    for (int j = 0; j < K; j += 2) {
        const uint4 v = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)(j >> 1), 0x42474d4du), key);
        const double e0 = -log(uniform_open(v.x, v.y));
        out[j] = e0;
        sum += e0;
        if (j + 1 < K) {
            const double e1 = -log(uniform_open(v.z, v.w));
            out[j + 1] = e1;
            sum += e1;
        }
    }
    const double inv = 1.0 / sum;
    for (int j = 0; j < K; ++j) out[j] *= inv;
}

// GenModel.gen_sample on the device (/root/reference/bayesml/gaussianmixture/_gaussianmixture.py:241-264: a Python loop,
// one `rng.choice` + one `rng.multivariate_normal` per sample): one thread per sample, class from the inverse CDF of pi,
// x = mu_z + L_z eps with L_z = chol(Lambda_z^-1) and Box-Muller normals from Philox keyed by (seed, global row).
__global__ void __launch_bounds__(128) gen_sample_kernel(double* __restrict__ x, int32_t* __restrict__ z, const int64_t n,
                                                         const int K, const int D, const double* __restrict__ cdf,
                                                         const double* __restrict__ mu, const double* __restrict__ chol,
                                                         const unsigned long long seed, const int64_t row_offset) {
    extern __shared__ double eps_s[];                           // [128][D + 1]: this thread's standard normals
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    const unsigned long long g = (unsigned long long)(row_offset + i);
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    double* eps = eps_s + threadIdx.x * (D + 1);
    const uint4 v0 = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), 0u, 0x47454e53u), key);
    const double u = uniform_open(v0.x, v0.y);
    int k = 0;
    while (k < K - 1 && u > cdf[k]) ++k;
    for (int j = 0; j < D; j += 2) {
        const uint4 v = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)(1 + (j >> 1)), 0x47454e53u), key);
        const double rad = sqrt(-2.0 * log(uniform_open(v.x, v.y))), ang = 6.283185307179586476925286766559 * uniform_open(v.z, v.w);
        double sn, cs;
        sincos(ang, &sn, &cs);
        eps[j] = rad * cs;
        if (j + 1 < D) eps[j + 1] = rad * sn;
    }
    const double* L = chol + (int64_t)k * D * D;
    const double* m = mu + (int64_t)k * D;
    for (int a = 0; a < D; ++a) {
        double acc = m[a];
        for (int b = 0; b <= a; ++b) acc = fma(L[a * D + b], eps[b], acc);
        x[i * D + a] = acc;
    }
    z[i] = k;
}

}  // namespace bgmm

using namespace bgmm;

extern "C" int bgmm_gen_sample(double* x_out, int32_t* z_out, int64_t n, int K, int D, const double* cdf, const double* mu,
                               const double* chol, uint64_t seed, int64_t row_offset, void* stream) {
    if (n < 0 || K <= 0 || D <= 0 || D > 256 || row_offset < 0 || cdf == nullptr || mu == nullptr || chol == nullptr ||
        ((x_out == nullptr || z_out == nullptr) && n > 0)) {
        set_error("bgmm_gen_sample: bad argument (n=%lld K=%d D=%d; D <= 256)", (long long)n, K, D);
        return BGMM_EINVAL;
    }
    if (n == 0) return BGMM_OK;
    const size_t smem = sizeof(double) * 128 * (size_t)(D + 1);
    if (smem > 48 * 1024) {
        int rc = check_cuda(cudaFuncSetAttribute(gen_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(gen_sample_kernel)");
        if (rc) return rc;
    }
    gen_sample_kernel<<<(unsigned)((n + 127) / 128), 128, smem, (cudaStream_t)stream>>>(x_out, z_out, n, K, D, cdf, mu, chol,
                                                                                        (unsigned long long)seed, row_offset);
    return check_cuda(cudaGetLastError(), "gen_sample_kernel launch");
}

extern "C" int bgmm_pred_logdensity(const double* lnrho, int64_t n, int K, const double* acst, const double* ck,
                                    const double* hk, const double* nuk, double* out, void* stream) {
    if (n < 0 || K <= 0 || K > 64 || ((lnrho == nullptr || out == nullptr) && n > 0) || acst == nullptr || ck == nullptr ||
        hk == nullptr || nuk == nullptr) {
        set_error("bgmm_pred_logdensity: bad argument (n=%lld K=%d; K <= 64)", (long long)n, K);
        return BGMM_EINVAL;
    }
    if (n == 0) return BGMM_OK;
    pred_logdensity_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(lnrho, n, K, acst, ck, hk, nuk, out);
    return check_cuda(cudaGetLastError(), "pred_logdensity_kernel launch");
}

extern "C" int bgmm_dirichlet1(double* r_out, int64_t n, int K, uint64_t seed, int64_t row_offset, void* stream) {
    if (n < 0 || K <= 0 || row_offset < 0 || (r_out == nullptr && n > 0)) {
        set_error("bgmm_dirichlet1: bad argument (n=%lld K=%d)", (long long)n, K);
        return BGMM_EINVAL;
    }
    if (n == 0) return BGMM_OK;
    dirichlet1_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(r_out, n, K, (unsigned long long)seed,
                                                                                     row_offset);
    return check_cuda(cudaGetLastError(), "dirichlet1_kernel launch");
}
