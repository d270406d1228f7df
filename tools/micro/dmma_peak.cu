// Microbenchmark: FP64 tensor-core (mma.sync.m8n8k4.f64 -> DMMA.8x8x4) throughput vs DFMA on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak dmma_peak.cu && ./dmma_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a, double b) {
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    double c4[2] = {0, 0}, c5[2] = {0, 0}, c6[2] = {0, 0}, c7[2] = {0, 0};
    const double av = a + threadIdx.x * 1e-9, bv = b;
#define MMA(c) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(av), "d"(bv))
    for (int it = 0; it < iters; ++it) {
        MMA(c0); MMA(c1); MMA(c2); MMA(c3); MMA(c4); MMA(c5); MMA(c6); MMA(c7);
    }
    const double v = c0[0] + c1[0] + c2[0] + c3[0] + c4[1] + c5[1] + c6[1] + c7[1];
    if (v == 123.456) out[0] = v;
}

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double v = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (v == 123.456) out[0] = v;
}

// 4 DMMA (= 32 warp-DFMA equivalents) + 32 DFMA per iteration, independent chains: if the two instruction classes ran
// on separate pipes the loop would take max(t_dmma, t_dfma); on a shared FP64 datapath it takes the sum.
__global__ void __launch_bounds__(256) mixed_kernel(double* out, int iters, double a, double b) {
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    const double av = a + threadIdx.x * 1e-9, bv = b;
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int it = 0; it < iters; ++it) {
        MMA(c0);
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        MMA(c1);
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        MMA(c2);
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        MMA(c3);
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double v = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7)) + c0[0] + c1[0] + c2[1] + c3[1];
    if (v == 123.456) out[0] = v;
}

// FP64 adds/multiplies next to FP32 FMAs and shared-memory loads: which instruction classes co-issue with DFMA?
__global__ void __launch_bounds__(256) dfma_ffma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    float y0 = threadIdx.x, y1 = y0 + 1, y2 = y0 + 2, y3 = y0 + 3;
    const float fa = (float)a, fb = (float)b;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); y0 = fmaf(y0, fa, fb); x1 = fma(x1, a, b); y1 = fmaf(y1, fa, fb);
            x2 = fma(x2, a, b); y2 = fmaf(y2, fa, fb); x3 = fma(x3, a, b); y3 = fmaf(y3, fa, fb);
        }
    }
    const double v = (x0 + x1) + (x2 + x3) + (double)((y0 + y1) + (y2 + y3));
    if (v == 123.456) out[0] = v;
}

// Does a DMMA keep the sub-partition's issue port busy for its 16 cycles?  4 DMMA (64 datapath cycles) next to 64
// independent FFMAs (64 issue cycles) per iteration: ~64 cycles per iteration if they overlap, ~128 if they do not.
__global__ void __launch_bounds__(256) dmma_ffma_kernel(double* out, int iters, double a, double b) {
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    const double av = a + threadIdx.x * 1e-9, bv = b;
    float y[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) y[k] = threadIdx.x + k;
    const float fa = (float)a, fb = (float)b;
    for (int it = 0; it < iters; ++it) {
        MMA(c0);
#pragma unroll
        for (int k = 0; k < 16; ++k) y[k] = fmaf(y[k], fa, fb);
        MMA(c1);
#pragma unroll
        for (int k = 0; k < 16; ++k) y[k] = fmaf(y[k], fa, fb);
        MMA(c2);
#pragma unroll
        for (int k = 0; k < 16; ++k) y[k] = fmaf(y[k], fa, fb);
        MMA(c3);
#pragma unroll
        for (int k = 0; k < 16; ++k) y[k] = fmaf(y[k], fa, fb);
    }
    float ys = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) ys += y[k];
    const double v = c0[0] + c1[0] + c2[1] + c3[1] + (double)ys;
    if (v == 123.456) out[0] = v;
}

__global__ void __launch_bounds__(256) ffma_only_kernel(double* out, int iters, double a, double b) {
    float y[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) y[k] = threadIdx.x + k;
    const float fa = (float)a, fb = (float)b;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 16; ++k) y[k] = fmaf(y[k], fa, fb);
    }
    float ys = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) ys += y[k];
    if ((double)ys == 123.456) out[0] = ys;
}

int main() {
    double* d;
    cudaMalloc(&d, 64);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int grid = p.multiProcessorCount * 8, iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int pass = 0; pass < 2; ++pass) {
        for (int w = 0; w < 3; ++w) { if (pass) dmma_kernel<<<grid, 256>>>(d, iters, 1.0000001, 1e-9); else dfma_kernel<<<grid, 256>>>(d, iters, 1.0000001, 1e-9); }
        cudaDeviceSynchronize();
        float best = 1e30f;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(e0);
            if (pass) dmma_kernel<<<grid, 256>>>(d, iters, 1.0000001, 1e-9); else dfma_kernel<<<grid, 256>>>(d, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            best = ms < best ? ms : best;
        }
        // dfma: 32 FMA/thread/iter; dmma: 8 MMA/warp/iter, each 8*8*4 FMA
        const double flop = pass ? 2.0 * 8 * 256.0 * iters * grid * 8 : 2.0 * 32 * iters * (double)grid * 256;
        printf("%s: %.3f ms  %.2f TFLOP/s\n", pass ? "DMMA m8n8k4" : "DFMA", best, flop / (best * 1e-3) / 1e12);
    }
    for (int pass = 0; pass < 2; ++pass) {
        for (int w = 0; w < 3; ++w) { if (pass) dfma_ffma_kernel<<<grid, 256>>>(d, iters, 1.0000001, 1e-9); else mixed_kernel<<<grid, 256>>>(d, iters, 1.0000001, 1e-9); }
        cudaDeviceSynchronize();
        float best = 1e30f;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(e0);
            if (pass) dfma_ffma_kernel<<<grid, 256>>>(d, iters, 1.0000001, 1e-9); else mixed_kernel<<<grid, 256>>>(d, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            best = ms < best ? ms : best;
        }
        // mixed: 4 DMMA + 32 DFMA per warp-iteration; dfma_ffma: 32 DFMA (+ 32 FFMA) per thread-iteration
        const double flop = pass ? 2.0 * 32 * iters * (double)grid * 256 : 2.0 * (4 * 256.0 + 32 * 32.0) * iters * grid * 8;
        printf("%s: %.3f ms  %.2f FP64 TFLOP/s\n", pass ? "DFMA + FFMA 1:1 (FP64 flop only)" : "DMMA + DFMA mixed 1:1", best, flop / (best * 1e-3) / 1e12);
    }
    for (int pass = 0; pass < 2; ++pass) {
        for (int w = 0; w < 3; ++w) { if (pass) ffma_only_kernel<<<grid, 256>>>(d, iters, 1.0000001, 1e-9); else dmma_ffma_kernel<<<grid, 256>>>(d, iters, 1.0000001, 1e-9); }
        cudaDeviceSynchronize();
        float best = 1e30f;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(e0);
            if (pass) ffma_only_kernel<<<grid, 256>>>(d, iters, 1.0000001, 1e-9); else dmma_ffma_kernel<<<grid, 256>>>(d, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            best = ms < best ? ms : best;
        }
        printf("%s: %.3f ms\n", pass ? "64 FFMA per iteration alone" : "4 DMMA + 64 FFMA per iteration", best);
    }
    printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
