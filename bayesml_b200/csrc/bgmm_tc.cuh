// tcgen05 / TMEM primitives for sm_100a (inline PTX) used by the fp32-mode tensor-core kernels (bgmm_pass_tf32.cu).
// Conventions follow the CUTLASS SM100 descriptors (cute/arch/mma_sm100_desc.hpp): shared-memory matrix descriptors in the
// no-swizzle ("interleave") canonical layout made of 8 x 16-byte core matrices, instruction descriptor for kind::tf32 with
// fp32 accumulators in tensor memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bgmm {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- tensor memory ----
// one full warp allocates `ncols` (power of two >= 32) columns; the base address lands in *slot (shared memory)
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes (st.shared) -> visible to the async proxy (the tensor core reads operands through it)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- descriptors ----
// Shared-memory matrix descriptor, no swizzle.  A core matrix is 8 rows x 16 bytes, stored contiguously (128 bytes).
//   K-major operand  (rows = M or N index, 16-byte chunk = 4 consecutive k):   LBO = byte distance between the two k-chunks
//                     of one MMA (K = 8 tf32), SBO = byte distance between consecutive 8-row groups;
//   MN-major operand (rows = k index, 16-byte chunk = 4 consecutive m or n):   SBO = byte distance between consecutive
//                     4-element chunks along M / N, LBO = byte distance between consecutive 8-row k groups.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell)
    return d;                                 // base offset 0, lbo mode 0, layout type 0 (no swizzle)
}
// K-major operand in the 128-byte-swizzle layout: a row (M or N index) holds 32 consecutive k (128 bytes), 8 rows form a
// 1024-byte atom (1024-byte aligned) whose 16-byte chunks are XOR-ed with the row index (chunk ^ (row & 7)); SBO = byte
// distance between 8-row groups; one MMA (8 k) starts 32 bytes further along the row.  The tensor core fetches whole
// 128-byte rows: measured twice the operand rate of the no-swizzle layout (whose core matrices it fetches one by one).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                   // LBO: unused for a swizzled K-major operand
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                   // layout type: SWIZZLE_128B
    return d;
}
// float offset of element (row, k) of a [rows][32 k] swizzled atom column (atoms of 8 rows stacked every 256 floats)
__device__ __forceinline__ int sw128_off(int row, int k) {
    return (row >> 3) * 256 + (row & 7) * 32 + ((((k >> 2) & 7) ^ (row & 7)) << 2) + (k & 3);
}
// Instruction descriptor for kind::tf32: fp32 accumulate, A and B tf32, M x N tile, K-major (0) or MN-major (1) operands.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem];  issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// the same with the A operand in tensor memory (lane = row m, one 32-bit column per k): no shared-memory read for A
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// the same with the descriptor handed over as two 32-bit words: an issue loop that precomputes the low words (the address
// field) and keeps the high word constant needs one integer add per MMA instead of rebuilding the 64-bit descriptor — the
// issuing thread is a single warp at ~10 cycles per dependent instruction, so the 16 instructions per MMA of the naive form
// (2430 cycles for the 16 MMAs of a statistics sub-tile, measured with clock64) were the critical path of the kernel
__device__ __forceinline__ void mma_tf32_ts_w(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 bd, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_tf32_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 ad, bd;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 ad, {%1, %2};\n\tmov.b64 bd, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
// registers -> tensor memory: this warp's 32 lanes, 8 consecutive columns per lane
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                   "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                   "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// mbarrier arrive when every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
}

// ---- tensor memory -> registers: lane = row of the accumulator tile, this warp's 32 lanes, 32 consecutive columns ----
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// asynchronous: the registers are valid only after tmem_wait_ld() — issue several loads, wait once (a load + wait pair
// costs its full latency, ~500 cycles measured through the E kernel's first version)
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    tmem_ld16_async(taddr, r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- 3xTF32 split: v = hi + lo with hi = round-to-nearest tf32 (unbiased), lo = the exact remainder (the MMA reads its
//      top 19 bits); hi*hi + hi*lo + lo*hi keeps ~21 bits of every product ----
__device__ __forceinline__ float tf32_hi(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// the cheap split of the streaming kernels: hi = the top 19 bits (one LOP3 instead of the four instructions cvt.rna
// compiles to), lo = v - hi exact with <= 13 significant bits of which the MMA reads 11: 2^-21 |v| per term, one bit worse
// than the rounded split
__device__ __forceinline__ float tf32_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

// canonical no-swizzle offsets (in floats) of element (row, col) of a tile whose row groups / chunks are laid out as stated
//   K-major tile [rows][kcols]: core (row / 8, col / 4) at (row / 8) * sbo_f + (col / 4) * lbo_f
__device__ __forceinline__ int kmajor_off(int row, int col, int lbo_f, int sbo_f) {
    return (row >> 3) * sbo_f + (col >> 2) * lbo_f + (row & 7) * 4 + (col & 3);
}
//   MN-major tile [krows][mncols]: core (k / 8, mn / 4) at (mn / 4) * sbo_f + (k / 8) * lbo_f
__device__ __forceinline__ int mnmajor_off(int k, int mn, int lbo_f, int sbo_f) {
    return (mn >> 2) * sbo_f + (k >> 3) * lbo_f + (k & 7) * 4 + (mn & 3);
}

}  // namespace tc
}  // namespace bgmm
