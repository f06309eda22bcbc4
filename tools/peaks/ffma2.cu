// FFMA vs FFMA2 (fma.rn.f32x2) issue-rate microbenchmark on sm_100a: 8 independent chains per thread, 1024 threads/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{ .reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n mov.b64 rc, {%6, %7};\n"
        " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

template <bool PACKED>
__global__ void __launch_bounds__(256) kern(float* out, int iters, float s) {
    float2 acc[8];
    const float2 m = make_float2(s, s * 0.5f);
    for (int j = 0; j < 8; ++j) acc[j] = make_float2(threadIdx.x * 1e-3f + j, j * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (PACKED) acc[j] = ffma2(acc[j], m, m);
            else { acc[j].x = fmaf(acc[j].x, m.x, m.y); acc[j].y = fmaf(acc[j].y, m.y, m.x); }
        }
    }
    float r = 0.f;
    for (int j = 0; j < 8; ++j) r += acc[j].x + acc[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    const int iters = 20000;
    for (int packed = 0; packed < 2; ++packed) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (packed) kern<true><<<sms * 8, 256>>>(out, iters, 0.999f);
            else kern<false><<<sms * 8, 256>>>(out, iters, 0.999f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fma = (double)sms * 8 * 256 * iters * 16;     // scalar fused multiply-adds executed
        printf("%s: %.3f ms  %.2f TFLOP/s fp32 (2 flops per FMA)\n", packed ? "FFMA2 (f32x2)" : "FFMA scalar ", ms,
               2.0 * fma / ms / 1e9);
    }
    return 0;
}
