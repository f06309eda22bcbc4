// DMMA.8x8x4 issue rate as a function of resident warps per SM sub-partition and independent accumulator chains.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 8192;
template <int CH>
__global__ void k(double* out, double a, double b) {
    double c[CH][2];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
void run(double* d, int warps_per_sm) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int block = warps_per_sm * 32, grid = 148;
    k<CH><<<grid, block>>>(d, 0.999, 0.001); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<CH><<<grid, block>>>(d, 0.999, 0.001); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double dmma_per_smsp = (double)ITERS * CH * warps_per_sm / 4.0;
    double cyc = ms * 1e-3 * 1.965e9;
    printf("chains=%2d warps/SMSP=%.2f  cycles/DMMA/SMSP=%.2f  TFLOP/s=%.2f\n", CH, warps_per_sm / 4.0, cyc / dmma_per_smsp,
           2.0 * 256 * ITERS * CH * warps_per_sm * grid / (ms * 1e-3) / 1e12);
}
int main() {
    double* d; cudaMalloc(&d, 8 * 148 * 1024);
    for (int w : {4, 8, 12, 16}) { run<1>(d, w); run<4>(d, w); run<8>(d, w); run<20>(d, w); }
    return 0;
}
