// Microbenchmarks for the compute-side roofline denominators on B200 (sm_100a).
// MEASURED_PEAKS.json (driver-written) only has HBM copy and bf16 GEMM; the VB-GMM
// hot path is FP64 (and FP32 in fp32 mode), so these numbers are measured here:
//   dfma      : scalar fp64 FMA issue rate
//   dmma884   : mma.sync.m8n8k4.f64 rate
//   dmma16816 : mma.sync.m16n8k16.f64 rate
//   ffma      : scalar fp32 FMA rate
//   ffma2     : packed fma.rn.f32x2 rate (Blackwell)
//   dexp      : fp64 exp() calls per second
//   hbm_read  : read-only streaming bandwidth (the pass kernel only reads X)
// Output: one JSON object on stdout.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b) {
    double r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = fma(r[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_ffma(float* out, float a, float b) {
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = fmaf(r[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_ffma2(float* out, float a, float b) {
    unsigned long long r[8];
    unsigned long long av, bv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float v = threadIdx.x * 1e-3f + i;
        asm("mov.b64 %0, {%1, %1};" : "=l"(r[i]) : "f"(v));
    }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(av), "l"(bv));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r[i]));
        s += lo + hi;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dmma884(double* out, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dmma16816(double* out, double a, double b) {
    double c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
                         "{%4,%4,%4,%4,%4,%4,%4,%4}, {%5,%5,%5,%5}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dmma1684(double* out, double a, double b) {
    double c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
                         "{%4,%4}, {%5}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dexp(double* out, double a) {
    double r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = -1.0 - 1e-3 * (threadIdx.x + i);
    for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = exp(r[i]) * a - 2.0;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r[0] + r[1] + r[2] + r[3];
}

__global__ void __launch_bounds__(256) k_read(const double2* __restrict__ x, size_t n2, double* out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (; i + 3 * stride < n2; i += 4 * stride) {
        double2 a = x[i], b = x[i + stride], c = x[i + 2 * stride], d = x[i + 3 * stride];
        s0 += a.x + a.y; s1 += b.x + b.y; s2 += c.x + c.y; s3 += d.x + d.y;
    }
    for (; i < n2; i += stride) { double2 a = x[i]; s0 += a.x + a.y; }
    double s = s0 + s1 + s2 + s3;
    if (s == 12345.678) out[0] = s;
}

template <class F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    int grid = sms * 8, block = 256;
    double* dout; CK(cudaMalloc(&dout, sizeof(double) * grid * block));
    float* fout = reinterpret_cast<float*>(dout);
    size_t nthreads = (size_t)grid * block;

    double t;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);

    t = time_ms([&] { k_dfma<<<grid, block>>>(dout, 0.999, 0.001); }, 5);
    printf(", \"dfma_tflops\": %.3f", 2.0 * nthreads * ITERS * 8 / (t * 1e-3) / 1e12);
    // sustained: back-to-back for ~2 s
    {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        int reps = (int)(2000.0 / t) + 1;
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i) k_dfma<<<grid, block>>>(dout, 0.999, 0.001);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf(", \"dfma_tflops_sustained\": %.3f", 2.0 * nthreads * ITERS * 8 * reps / (ms * 1e-3) / 1e12);
    }
    t = time_ms([&] { k_dmma884<<<grid, block>>>(dout, 0.999, 0.001); }, 5);
    printf(", \"dmma_m8n8k4_tflops\": %.3f", 2.0 * (nthreads / 32) * ITERS * 8 * 256 / (t * 1e-3) / 1e12);
    {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        int reps = (int)(2000.0 / t) + 1;
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; ++i) k_dmma884<<<grid, block>>>(dout, 0.999, 0.001);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf(", \"dmma_m8n8k4_tflops_sustained\": %.3f", 2.0 * (nthreads / 32) * ITERS * 8 * 256 * reps / (ms * 1e-3) / 1e12);
    }
    t = time_ms([&] { k_dmma1684<<<grid, block>>>(dout, 0.999, 0.001); }, 5);
    printf(", \"dmma_m16n8k4_tflops\": %.3f", 2.0 * (nthreads / 32) * ITERS * 4 * 512 / (t * 1e-3) / 1e12);
    t = time_ms([&] { k_dmma16816<<<grid, block>>>(dout, 0.999, 0.001); }, 5);
    printf(", \"dmma_m16n8k16_tflops\": %.3f", 2.0 * (nthreads / 32) * ITERS * 4 * 2048 / (t * 1e-3) / 1e12);
    t = time_ms([&] { k_ffma<<<grid, block>>>(fout, 0.999f, 0.001f); }, 5);
    printf(", \"ffma_tflops\": %.3f", 2.0 * nthreads * ITERS * 8 / (t * 1e-3) / 1e12);
    t = time_ms([&] { k_ffma2<<<grid, block>>>(fout, 0.999f, 0.001f); }, 5);
    printf(", \"ffma2_tflops\": %.3f", 4.0 * nthreads * ITERS * 8 / (t * 1e-3) / 1e12);
    t = time_ms([&] { k_dexp<<<grid, block>>>(dout, 0.5); }, 5);
    printf(", \"dexp_gcalls_per_s\": %.3f", (double)nthreads * (ITERS / 8) * 4 / (t * 1e-3) / 1e9);

    size_t bytes = (size_t)4 << 30;
    double2* big; CK(cudaMalloc(&big, bytes)); CK(cudaMemset(big, 0, bytes));
    t = time_ms([&] { k_read<<<sms * 16, 256>>>(big, bytes / sizeof(double2), dout); }, 10);
    printf(", \"hbm_read_gbs\": %.1f", bytes / (t * 1e-3) / 1e9);
    printf("}\n");
    return 0;
}
