// FP64 peak micro-benchmark for the roofline of the contraction-bound configurations (C3: P3 with per-point coefficients,
// C4: P2^3 elasticity).  SURVEY 8(d): "FP64 peak is not in MEASURED_PEAKS -- measure".  Two kernels, both register-only:
//   dfma : independent DFMA chains per thread (vector FP64 pipe)
//   dmma : mma.sync.aligned.m8n8k4.row.col.f64 with independent accumulator tiles (FP64 tensor path)
// Prints one JSON line {"dfma_tflops": ..., "dmma_tflops": ..., "sm_count": ..., "clock_mhz": ...}.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b, int iters) {
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3 + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 12345.678) out[0] = s;   // never true: keeps the chains alive
}

template <int TILES>
__global__ void __launch_bounds__(256) k_dmma(double* out, double a, double b, int iters) {
    double c0[TILES], c1[TILES];
#pragma unroll
    for (int t = 0; t < TILES; ++t) { c0[t] = threadIdx.x * 1e-3 + t; c1[t] = t; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int t = 0; t < TILES; ++t)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[t]), "+d"(c1[t]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int t = 0; t < TILES; ++t) s += c0[t] + c1[t];
    if (s == 12345.678) out[0] = s;
}

template <typename K>
static double time_ms(K launch, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main() {
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, 0) != cudaSuccess) { printf("{\"error\": \"no CUDA device\"}\n"); return 1; }
    double* out;
    cudaMalloc(&out, 64);
    const int sms = pr.multiProcessorCount, iters = 4096, blocks = sms * 8;
    constexpr int CH = 8, TL = 8;
    const double ms_f = time_ms([&] { k_dfma<CH><<<blocks, 256>>>(out, 1.0000001, 1e-9, iters); }, 5);
    const double ms_m = time_ms([&] { k_dmma<TL><<<blocks, 256>>>(out, 1.0000001, 1e-9, iters); }, 5);
    const double flop_f = 2.0 * CH * iters * 256.0 * blocks;
    const double flop_m = 2.0 * 256.0 * TL * iters * 8.0 * blocks;   // 8 warps per block, 8x8x4 FMAs per mma
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, \"sm_count\": %d, \"clock_mhz\": %.0f, \"dfma_ms\": %.3f, \"dmma_ms\": %.3f}\n",
           flop_f / ms_f * 1e-9, flop_m / ms_m * 1e-9, sms, clk * 1e-3, ms_f, ms_m);
    return cudaGetLastError() != cudaSuccess;
}
