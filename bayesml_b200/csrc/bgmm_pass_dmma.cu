// bgmm_pass, fp64 tensor-pipe variant (BGMM_PASS_DMMA) for sm_100a.
//
// One fused sweep over the centred rows of X per VB iteration.  Per 64-sample tile a persistent CTA
//   1. receives the tile by TMA bulk copy (cp.async.bulk + mbarrier, two-stage ring, issued one tile ahead),
//   2. expands it into the feature tile  Phi[64][SP]  (phi = [1, x, x_i x_j (i>=j)], zero padded) in shared memory,
//   3. E-step GEMM on the FP64 tensor pipe:  ln rho[64][K] = Phi . coef^T      (mma.sync.m8n8k4.f64 = DMMA.8x8x4),
//   4. softmax over k in the accumulator fragments (quad shuffles), entropy term, optional r / ln rho / argmax stores,
//      r tile -> shared memory,
//   5. M-step GEMM on the same pipe:  raw[K][SP] += R^T . Phi   with the accumulators resident in registers for the
//      whole sweep (r never touches HBM; HBM traffic = X once),
// then writes its partial statistics once and the last CTA reduces all partials in CTA order (deterministic).
//
// Replaces `_update_q_z` :772-784 (incl. the K-loop of N x D temporaries), `_calc_n_x_bar_s` :725-732 and the
// `xlogy` term :704 of /root/reference/bayesml/gaussianmixture/_gaussianmixture.py.
//
// Shared-memory layout: rows of Phi / coef / R are stored with a "k-step pair" column permutation
// (logical feature p = 8u + 4h + q  ->  physical column 8u + 2q + h) so that one LDS.128 yields the A (or B)
// fragments of two consecutive k-steps, and with a 16-byte-chunk XOR swizzle  f(row) = ((row&1)<<2)|(row&2)
// that makes every fragment load (LDS.128 by quarter-warp, LDS.64 by half-warp) bank-conflict free.
#include "bgmm_common.cuh"
#include <math.h>

namespace bgmm {

constexpr int DM_THREADS = 256;
constexpr int DM_WARPS = 8;
constexpr int DM_TILE = 64;          // samples per tile = 8 m-blocks

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

constexpr int DM_MAXTQ = 5;         // quadratic features per lane per row: D(D+1)/2 <= 160

__device__ __forceinline__ int fsw(int row) { return ((row & 1) << 3) | ((row & 2) << 1); }
__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }

template <int KB, int SP>
__global__ void __launch_bounds__(DM_THREADS, 1)
pass_dmma_kernel(const PassArgs a, const Layout L) {
    constexpr int RP = (8 * KB < 16) ? 16 : 8 * KB;          // pitch of the r tile (multiple of 16)
    constexpr int NBT = SP / 8;                               // n-blocks of the M-GEMM
    constexpr int MAXNB = ((NBT + 3) / 4 + 1) / 2;            // n-blocks owned by one warp
    constexpr int EW = SP / 32;                               // E-GEMM: 16-column groups per k half
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int K = L.K, D = L.D, P = L.P;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
    volatile int* ctrl = reinterpret_cast<volatile int*>(a.state + L.ctrl);
    if (!a.force && ctrl[BGMM_CTRL_DONE]) return;
    const double* __restrict__ coef_g = a.state + L.params[ctrl[BGMM_CTRL_CUR]] + L.p_coef;
    const double* __restrict__ x = static_cast<const double*>(a.x);

    // ---- carve shared memory ----
    double* phiS = reinterpret_cast<double*>(smem_raw);                  // [64][SP]
    double* coefS = phiS + DM_TILE * SP;                                 // [8*KB][SP]
    double* rS = coefS + 8 * KB * SP;                                    // [64][RP]
    double* xS = rS + DM_TILE * RP;                                      // [2][64*D]
    double* exch = xS + 2 * DM_TILE * D;                                 // [8 warps][32 lanes][2*KB]
    double* red = exch + DM_WARPS * 32 * 2 * KB;                         // [40]
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + 40);              // [2]

    const int64_t ntiles = (a.n + DM_TILE - 1) / DM_TILE;
    const uint32_t tile_bytes = (uint32_t)(DM_TILE * D * sizeof(double));

    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    // full tiles go through TMA; a ragged last tile is loaded with plain loads (size may not be a multiple of 16 B)
    auto issue_tile = [&](int64_t t, int stage) {
        const int64_t row0 = t * DM_TILE;
        if (row0 + DM_TILE <= a.n) {
            mbar_expect_tx(&bars[stage], tile_bytes);
            tma_load_1d(xS + stage * DM_TILE * D, x + row0 * D, tile_bytes, &bars[stage]);
        }
    };
    if (tid == 0) {
        if ((int64_t)blockIdx.x < ntiles) issue_tile(blockIdx.x, 0);
        if ((int64_t)blockIdx.x + gridDim.x < ntiles) issue_tile((int64_t)blockIdx.x + gridDim.x, 1);
    }

    // ---- one-time: swizzled coefficient matrix; the tile-invariant columns of Phi (constant 1 and zero padding) ----
    for (int e = tid; e < 8 * KB * SP; e += DM_THREADS) {
        const int k = e / SP, p = e - k * SP;
        double v = 0.0;
        if (k < K) { if (p < P) v = coef_g[(int64_t)k * L.pitch + p]; }
        else if (p == 0) v = -1.0e300;                                    // padded components: r == 0 exactly
        coefS[swz(k, phys_col(p), SP)] = v;
    }
    for (int e = tid; e < DM_TILE * (SP - P + 1); e += DM_THREADS) {
        const int r = e / (SP - P + 1), c = e - r * (SP - P + 1);
        if (c == 0) phiS[swz(r, phys_col(0), SP)] = 1.0;
        else phiS[swz(r, phys_col(P + c - 1), SP)] = 0.0;
    }
    // per-lane feature table of the Phi expansion (this lane produces the same features for every row)
    const int NQ = D * (D + 1) / 2;
    int qi[DM_MAXTQ], qj[DM_MAXTQ], qc[DM_MAXTQ];
#pragma unroll
    for (int t = 0; t < DM_MAXTQ; ++t) {
        const int qq = lane + 32 * t;
        qi[t] = 0; qj[t] = 0; qc[t] = -1;
        if (qq < NQ) {
            int i = (int)((sqrtf(8.0f * qq + 1.0f) - 1.0f) * 0.5f);
            while (i * (i + 1) / 2 > qq) --i;
            while ((i + 1) * (i + 2) / 2 <= qq) ++i;
            qi[t] = i; qj[t] = qq - i * (i + 1) / 2; qc[t] = phys_col(1 + D + qq);
        }
    }
    const int lc = (lane < D) ? phys_col(1 + lane) : -1;                  // linear feature of this lane

    // ---- persistent accumulators of the M-GEMM: this warp owns n-blocks b = (warp%4) + 4*(2*l + warp/4) ----
    double macc[MAXNB][KB][2];
#pragma unroll
    for (int l = 0; l < MAXNB; ++l)
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) { macc[l][kb][0] = 0.0; macc[l][kb][1] = 0.0; }
    double ent = 0.0;

    // ---- loop-invariant fragment pointers (all swizzles resolved here; the loops below use immediates only) ----
    const int mp = warp & 3, kh = warp >> 2;                // E-GEMM: m-block pair, k half
    const int fg = fsw(g), fq = fsw(q);
    const double* eA = phiS + (16 * mp + g) * SP + kh * (16 * EW);   // rows 16mp+g (+8): same swizzle as row g
    const double* eB = coefS + g * SP + kh * (16 * EW);
    const int eo0 = (2 * q) ^ fg, eo1 = (8 + 2 * q) ^ fg;
    const int lrow = (2 * mp + kh) * 8 + g;                 // the row this thread finalises in the softmax
    double* rst = rS + lrow * RP + ((4 * q) ^ fg);
    const double* mR = rS + q * RP + ((2 * g) ^ fq);
    const double* mP[MAXNB];
    bool mOn[MAXNB];
#pragma unroll
    for (int l = 0; l < MAXNB; ++l) {
        const int b = (warp & 3) + 4 * (2 * l + (warp >> 2));
        mOn[l] = b < NBT;
        mP[l] = phiS + q * SP + ((8 * (mOn[l] ? b : 0) + g) ^ fq);
    }
    uint32_t phase[2] = {0u, 0u};
    __syncthreads();

    int it = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int stage = it & 1;
        const int64_t row0 = t * DM_TILE;
        const int rows = (int)min((int64_t)DM_TILE, a.n - row0);
        double* xt = xS + stage * DM_TILE * D;
        if (rows == DM_TILE) {
            mbar_wait(&bars[stage], phase[stage]);
            phase[stage] ^= 1u;
        } else {
            for (int e = tid; e < DM_TILE * D; e += DM_THREADS) xt[e] = (e < rows * D) ? x[row0 * D + e] : 0.0;
            __syncthreads();
        }

        // ---- 2. feature tile: warp w expands rows 8w..8w+7; a lane owns the same features in every row ----
        {
            const double* xr = xt + (8 * warp) * D;
            double* pr = phiS + (8 * warp) * SP;
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
                const int f = fsw(rr);                       // 8*warp does not change the low row bits
                if (lc >= 0) pr[lc ^ f] = xr[lane];
#pragma unroll
                for (int tq = 0; tq < DM_MAXTQ; ++tq)
                    if (qc[tq] >= 0) pr[qc[tq] ^ f] = xr[qi[tq]] * xr[qj[tq]];
                xr += D;
                pr += SP;
            }
        }
        __syncthreads();
        // x tile consumed: refill this stage two tiles ahead
        if (tid == 0 && t + 2 * (int64_t)gridDim.x < ntiles) issue_tile(t + 2 * (int64_t)gridDim.x, stage);

        // ---- 3. E-GEMM: this warp -> m-blocks {2mp, 2mp+1}, the k half kh (EW groups of 16 physical columns) ----
        double acc[2][KB][2];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) { acc[m][kb][0] = 0.0; acc[m][kb][1] = 0.0; }
#pragma unroll
        for (int w = 0; w < EW; ++w) {
            double2 af[2][2], bf[KB][2];
            af[0][0] = lds2(eA + 16 * w + eo0);
            af[0][1] = lds2(eA + 16 * w + eo1);
            af[1][0] = lds2(eA + 8 * SP + 16 * w + eo0);
            af[1][1] = lds2(eA + 8 * SP + 16 * w + eo1);
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
                bf[kb][0] = lds2(eB + kb * 8 * SP + 16 * w + eo0);
                bf[kb][1] = lds2(eB + kb * 8 * SP + 16 * w + eo1);
            }
#pragma unroll
            for (int par = 0; par < 2; ++par) {
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) {
                    dmma(acc[0][kb][0], acc[0][kb][1], af[0][par].x, bf[kb][par].x);
                    dmma(acc[1][kb][0], acc[1][kb][1], af[1][par].x, bf[kb][par].x);
                }
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) {
                    dmma(acc[0][kb][0], acc[0][kb][1], af[0][par].y, bf[kb][par].y);
                    dmma(acc[1][kb][0], acc[1][kb][1], af[1][par].y, bf[kb][par].y);
                }
            }
        }
        // exchange the k halves: this warp finalises m-block 2mp + kh
        {
            double* mine = exch + (warp * 32 + lane) * 2 * KB;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb)                   // kh is warp-uniform: selects, not dynamic register indexing
                *reinterpret_cast<double2*>(mine + 2 * kb) =
                    kh ? make_double2(acc[0][kb][0], acc[0][kb][1]) : make_double2(acc[1][kb][0], acc[1][kb][1]);
        }
        __syncthreads();
        double lr[KB][2];
        {
            const double* theirs = exch + (((warp ^ 4) * 32) + lane) * 2 * KB;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
                const double2 o = lds2(theirs + 2 * kb);
                lr[kb][0] = (kh ? acc[1][kb][0] : acc[0][kb][0]) + o.x;
                lr[kb][1] = (kh ? acc[1][kb][1] : acc[0][kb][1]) + o.y;
            }
        }

        // ---- 4. softmax over k for row lrow; this thread holds components 8kb + 2q + {0,1} ----
        const int64_t grow = row0 + lrow;
        const bool valid = lrow < rows;
        double mx = -INFINITY;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) mx = fmax(mx, fmax(lr[kb][0], lr[kb][1]));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        if (a.lnrho_out != nullptr && valid) {
#pragma unroll
            for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = 8 * kb + 2 * q + e;
                    if (k < K) a.lnrho_out[grow * K + k] = lr[kb][e];
                }
        }
        double sum = 0.0, dot = 0.0;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double z = lr[kb][e] - mx;
                const double ex = exp(z);
                lr[kb][e] = ex;
                sum += ex;
                dot = fma(ex, z, dot);                  // z is finite (padding uses -1e300), so 0 * z == 0
            }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        dot += __shfl_xor_sync(0xffffffffu, dot, 2);
        const double inv = valid ? 1.0 / sum : 0.0;     // rows past the end contribute r = 0
        if (valid && q == 0) ent += dot * inv - log(sum);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) { lr[kb][0] *= inv; lr[kb][1] *= inv; }
        // r tile: physical column 16(kb>>1) + 4q + 2e + (kb&1), swizzle folded into `rst`
        if constexpr (KB >= 2) {
#pragma unroll
            for (int v = 0; v < KB / 2; ++v)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    *reinterpret_cast<double2*>(rst + 16 * v + 2 * e) = make_double2(lr[2 * v][e], lr[2 * v + 1][e]);
        } else {
            rst[0] = lr[0][0];
            rst[2] = lr[0][1];
        }
        if (a.r_out != nullptr && valid) {
#pragma unroll
            for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = 8 * kb + 2 * q + e;
                    if (k < K) a.r_out[grow * K + k] = lr[kb][e];
                }
        }
        if (a.argmax_out != nullptr) {
            int best = 0x7fffffff;
            double bestv = -1.0;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (lr[kb][e] > bestv) { bestv = lr[kb][e]; best = 8 * kb + 2 * q + e; }   // ascending k: first wins
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bestv, o);
                const int ok = __shfl_xor_sync(0xffffffffu, best, o);
                if (ov > bestv || (ov == bestv && ok < best)) { bestv = ov; best = ok; }
            }
            if (valid && q == 0) a.argmax_out[grow] = best;
        }
        __syncthreads();

        // ---- 5. M-GEMM: raw[k][p] += sum_n r[n][k] phi[n][p]; A = R^T (8 comps x 4 samples), B = Phi (4 x 8) ----
#pragma unroll
        for (int ks = 0; ks < DM_TILE / 4; ++ks) {
            double ra[KB];
            if constexpr (KB >= 2) {
#pragma unroll
                for (int v = 0; v < KB / 2; ++v) {
                    const double2 r2 = lds2(mR + ks * 4 * RP + 16 * v);
                    ra[2 * v] = r2.x;
                    ra[2 * v + 1] = r2.y;
                }
            } else {
                ra[0] = mR[ks * 4 * RP];
            }
#pragma unroll
            for (int l = 0; l < MAXNB; ++l) {
                if (mOn[l]) {
                    const double bfr = mP[l][ks * 4 * SP];
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb) dmma(macc[l][kb][0], macc[l][kb][1], ra[kb], bfr);
                }
            }
        }
        __syncthreads();   // phiS / rS free for the next tile
    }

    // ---- write this CTA's partial statistics (logical layout [K][pitch]) ----
    const int64_t len = L.stats_len;
    double* part = a.workspace + (int64_t)blockIdx.x * len;
#pragma unroll
    for (int l = 0; l < MAXNB; ++l) {
        const int b = (warp & 3) + 4 * (2 * l + (warp >> 2));
        if (b < NBT) {
#pragma unroll
            for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = 8 * kb + g;              // C fragment: row g = component, cols 2q+e = physical 8b+2q+e
                    const int p = 8 * b + 4 * e + q;       // physical 8b + 2q + e  <->  logical 8b + 4e + q
                    if (k < K && p < L.pitch) part[(int64_t)k * L.pitch + p] = macc[l][kb][e];
                }
        }
    }
    ent = block_sum(ent, red);
    if (tid == 0) {
        part[(int64_t)K * L.pitch] = ent;
        for (int o = 1; o < 8; ++o) part[(int64_t)K * L.pitch + o] = 0.0;
    }

    // ---- last CTA reduces the partials in CTA order ----
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int tk = atomicAdd(const_cast<int*>(&ctrl[BGMM_CTRL_PASS_TICKET]), 1);
        is_last = (tk == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double* out = a.state + L.stats;
    const double* ws = a.workspace;
    for (int64_t o = tid; o < len; o += DM_THREADS) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int bidx = 0;
        for (; bidx + 3 < (int)gridDim.x; bidx += 4) {
            s0 += __ldcg(&ws[(int64_t)(bidx + 0) * len + o]);
            s1 += __ldcg(&ws[(int64_t)(bidx + 1) * len + o]);
            s2 += __ldcg(&ws[(int64_t)(bidx + 2) * len + o]);
            s3 += __ldcg(&ws[(int64_t)(bidx + 3) * len + o]);
        }
        for (; bidx < (int)gridDim.x; ++bidx) s0 += __ldcg(&ws[(int64_t)bidx * len + o]);
        double acc = (s0 + s1) + (s2 + s3);
        if (o == (int64_t)K * L.pitch + 1) acc = (double)a.n;
        out[o] = a.accumulate ? out[o] + acc : acc;
    }
    if (tid == 0) ctrl[BGMM_CTRL_PASS_TICKET] = 0;
}

// ---- host side ----
struct DmmaPlan {
    bool ok;
    int KB, SP, RP;
    size_t smem;
};

static DmmaPlan plan_dmma(int K, int D) {
    DmmaPlan p{};
    const int P = feat_count(D);
    p.SP = (P + 31) & ~31;
    p.KB = K <= 8 ? 1 : (((K + 15) & ~15) / 8);
    p.RP = 8 * p.KB < 16 ? 16 : 8 * p.KB;
    p.smem = sizeof(double) * ((size_t)DM_TILE * p.SP + (size_t)8 * p.KB * p.SP + (size_t)DM_TILE * p.RP +
                               (size_t)2 * DM_TILE * D + (size_t)DM_WARPS * 32 * 2 * p.KB + 40) +
             2 * sizeof(uint64_t) + 128;
    p.ok = D * (D + 1) / 2 <= 32 * DM_MAXTQ && D <= 32 && (p.KB == 1 || p.KB == 2 || p.KB == 4) && p.SP <= 192 &&
           p.smem <= 227 * 1024 - 512 && (D * sizeof(double) * DM_TILE) % 16 == 0;
    return p;
}

bool dmma_supported(int K, int D, int dtype) { return dtype == BGMM_F64 && plan_dmma(K, D).ok; }

static int dmma_grid(int64_t n) {
    const int64_t ntiles = (n + DM_TILE - 1) / DM_TILE;
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return (int)(ntiles < 1 ? 1 : (ntiles < sms ? ntiles : sms));
}

int64_t dmma_workspace_doubles(int K, int D) {
    if (!plan_dmma(K, D).ok) return 0;
    return (int64_t)160 * ((int64_t)K * feat_pitch(D) + 8);   // >= SM count of any sm_100 part
}

template <int KB, int SP>
static int launch_cfg(const PassArgs& a, const Layout& L, const DmmaPlan& p, cudaStream_t stream) {
    auto kern = pass_dmma_kernel<KB, SP>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 512);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(pass_dmma)");
        configured = true;
    }
    kern<<<dmma_grid(a.n), DM_THREADS, p.smem, stream>>>(a, L);
    return check_cuda(cudaGetLastError(), "pass_dmma_kernel launch");
}

int launch_pass_dmma(const PassArgs& a, int K, int D, int dtype, cudaStream_t stream) {
    const DmmaPlan p = plan_dmma(K, D);
    if (dtype != BGMM_F64 || !p.ok) {
        set_error("bgmm_pass(dmma): unsupported shape K=%d D=%d dtype=%d", K, D, dtype);
        return BGMM_ENOSUP;
    }
    if ((reinterpret_cast<uintptr_t>(a.x) & 15) != 0) {
        set_error("bgmm_pass(dmma): x must be 16-byte aligned");
        return BGMM_EINVAL;
    }
    const Layout L = make_layout(K, D, 1);
#define BGMM_DM_CASE(kb, sp) if (p.KB == kb && p.SP == sp) return launch_cfg<kb, sp>(a, L, p, stream);
#define BGMM_DM_ROW(kb) BGMM_DM_CASE(kb, 32) BGMM_DM_CASE(kb, 64) BGMM_DM_CASE(kb, 96) BGMM_DM_CASE(kb, 128) \
                        BGMM_DM_CASE(kb, 160) BGMM_DM_CASE(kb, 192)
    BGMM_DM_ROW(1) BGMM_DM_ROW(2) BGMM_DM_ROW(4)
#undef BGMM_DM_ROW
#undef BGMM_DM_CASE
    set_error("bgmm_pass(dmma): no instantiation for KB=%d SP=%d", p.KB, p.SP);
    return BGMM_ENOSUP;
}

}  // namespace bgmm
