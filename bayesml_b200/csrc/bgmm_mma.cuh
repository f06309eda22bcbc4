// FP64 tensor-pipe (DMMA) helpers shared by the pass kernels: the mma wrapper, the shared-memory tile layout
// (k-step-pair column permutation + 16-byte-chunk XOR swizzle) and the softmax exponential.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bgmm {

__device__ __forceinline__ void dmma(double& c0, double& c1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// element (row, physical column c) of a swizzled tile with row pitch `pitch` doubles (pitch % 16 == 0)
__device__ __forceinline__ int swz(int row, int c, int pitch) {
    return row * pitch + (c ^ (((row & 1) << 3) | ((row & 2) << 1)));
}
// logical feature -> physical column
__device__ __forceinline__ int phys_col(int p) { return (p & ~7) | ((p & 3) << 1) | ((p >> 2) & 1); }

__device__ __forceinline__ int fsw(int row) { return ((row & 1) << 3) | ((row & 2) << 1); }
__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }
// exp(z) for finite z <= 0, branch-free: z = n ln2 + f, |f| <= ln2/2, degree-12 Taylor (truncation 1.7e-16 relative),
// scaled by adding n to the exponent field.  Below z = -700 the result is flushed to 0 (true value < 1e-304).
__device__ __forceinline__ double exp_nonpos(double z) {
    const double magic = 6755399441055744.0;                    // 1.5 * 2^52: rint() by addition
    const double t = fma(z, 1.4426950408889634074, magic);
    const int n = __double2loint(t);
    const double nd = t - magic;
    double f = fma(nd, -6.93147180369123816490e-01, z);         // ln2 split (Cody-Waite)
    f = fma(nd, -1.90821492927058770002e-10, f);
    double p = 1.0 / 479001600.0;
    p = fma(p, f, 1.0 / 39916800.0);
    p = fma(p, f, 1.0 / 3628800.0);
    p = fma(p, f, 1.0 / 362880.0);
    p = fma(p, f, 1.0 / 40320.0);
    p = fma(p, f, 1.0 / 5040.0);
    p = fma(p, f, 1.0 / 720.0);
    p = fma(p, f, 1.0 / 120.0);
    p = fma(p, f, 1.0 / 24.0);
    p = fma(p, f, 1.0 / 6.0);
    p = fma(p, f, 0.5);
    p = fma(p, f, 1.0);
    p = fma(p, f, 1.0);
    const double r = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
    return z < -700.0 ? 0.0 : r;
}

// ---- mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wg_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

}  // namespace bgmm
