// dmma_lds.cu -- how busy can the FP64 tensor pipe be kept when the DMMA operands come out of shared memory?
// Both production kernels that use DMMA (rho_lattice_mma_kernel, isf_corr_mma_kernel) show the FP64 datapath ~2/3 busy
// in ncu although the warps are stalled on it ("math pipe throttle"), while DMMA chains on register operands reach 99 %
// (tools/micro/dmma_peak.cu).  This probe runs the tau-correlation's inner loop pattern -- per step 6 LDS.64 feeding 4
// DMMA.8x8x4 on four independent accumulators -- in several variants and at 1..8 warps per SM sub-partition:
//   0  operands in registers, no LDS                                  (ceiling)
//   1  6 LDS per 4 DMMA, operands consumed right after the loads      (what ptxas makes of the production loop)
//   2  the same, software pipelined: the loads of step h + 1 are issued before the DMMAs of step h (two register sets)
//   3  6 LDS per 4 DMMA, but the DMMAs use register operands          (shared-memory traffic without the dependence)
//   4  3 LDS per 4 DMMA, consumed right after the loads               (half the operand traffic)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_lds dmma_lds.cu && ./dmma_lds [sm_clock_mhz]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ double lds(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

constexpr int kRegion = 768;       // doubles per warp (6 KB), walked with a 96-byte stride like the production loop

template <int MODE>
__global__ void __launch_bounds__(128) k(double* out, int iters) {
    extern __shared__ __align__(16) double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* reg = sm + warp * kRegion;
    for (int i = lane; i < kRegion; i += 32) reg[i] = 1.0 + 1e-9 * i;
    __syncwarp();
    const int fi = lane >> 2, fk = lane & 3;
    // fragment bases as in isf_corr_mma_kernel: rows 12 doubles apart (A), consecutive elements (B)
    const unsigned base = static_cast<unsigned>(__cvta_generic_to_shared(reg));
    const unsigned pa = base + 8u * (fk + 12 * (7 - fi));
    const unsigned pb = base + 8u * (fk + fi + 4 * ((fk + fi) >> 3));
    double acc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    double a0 = 1.0 + lane, a1 = 2.0, a2 = 3.0, a3 = 0.5, b0 = 1e-3, b1 = 2e-3;
    const unsigned mask = 2047;                              // byte offsets 0..2047: every address stays inside the warp's region
    if constexpr (MODE == 0) {
        for (int h = 0; h < iters; ++h) {
            dmma(acc[0], a0, b0); dmma(acc[1], a1, b0); dmma(acc[2], a2, b1); dmma(acc[3], a3, b1);
        }
    } else if constexpr (MODE == 1) {
#pragma unroll 2
        for (int h = 0; h < iters; ++h) {
            const unsigned o = (96u * h) & mask;
            b0 = lds(pb + o); b1 = lds(pb + o + 32);
            a0 = lds(pa + o); a1 = lds(pa + o + 768); a2 = lds(pa + o + 32); a3 = lds(pa + o + 800);
            dmma(acc[0], a0, b0); dmma(acc[1], a1, b0); dmma(acc[2], a2, b1); dmma(acc[3], a3, b1);
        }
    } else if constexpr (MODE == 2) {
        double na0, na1, na2, na3, nb0, nb1;
        b0 = lds(pb); b1 = lds(pb + 32); a0 = lds(pa); a1 = lds(pa + 768); a2 = lds(pa + 32); a3 = lds(pa + 800);
#pragma unroll 2
        for (int h = 0; h < iters; ++h) {
            const unsigned o = (96u * (h + 1)) & mask;
            nb0 = lds(pb + o); nb1 = lds(pb + o + 32);
            na0 = lds(pa + o); na1 = lds(pa + o + 768); na2 = lds(pa + o + 32); na3 = lds(pa + o + 800);
            dmma(acc[0], a0, b0); dmma(acc[1], a1, b0); dmma(acc[2], a2, b1); dmma(acc[3], a3, b1);
            a0 = na0; a1 = na1; a2 = na2; a3 = na3; b0 = nb0; b1 = nb1;
        }
    } else if constexpr (MODE == 3) {
        double s = 0.0;
#pragma unroll 2
        for (int h = 0; h < iters; ++h) {
            const unsigned o = (96u * h) & mask;
            const double x0 = lds(pb + o), x1 = lds(pb + o + 32), x2 = lds(pa + o), x3 = lds(pa + o + 768), x4 = lds(pa + o + 32), x5 = lds(pa + o + 800);
            dmma(acc[0], a0, b0); dmma(acc[1], a1, b0); dmma(acc[2], a2, b1); dmma(acc[3], a3, b1);
            s = __hiloint2double(__double2hiint(s) ^ __double2hiint(x0) ^ __double2hiint(x1) ^ __double2hiint(x2) ^ __double2hiint(x3) ^
                                 __double2hiint(x4) ^ __double2hiint(x5), __double2loint(s));
        }
        acc[0][0] += s;
    } else {
#pragma unroll 2
        for (int h = 0; h < iters; ++h) {
            const unsigned o = (96u * h) & mask;
            b0 = lds(pb + o);
            a0 = lds(pa + o); a1 = lds(pa + o + 768);
            dmma(acc[0], a0, b0); dmma(acc[1], a1, b0); dmma(acc[2], a0, b0); dmma(acc[3], a1, b0);
        }
    }
    double t = 0.0;
    for (int i = 0; i < 4; ++i) t += acc[i][0] + acc[i][1];
    if (t == 1.2345e300) out[0] = t;
}

template <int MODE>
void run(double* out, double mhz, int sms) {
    const int iters = 20000;
    std::printf("mode %d:", MODE);
    for (int w = 1; w <= 8; ++w) {
        // w CTAs of 4 warps per SM = w warps per sub-partition; pad the shared-memory request so that exactly w fit
        size_t smem = 4 * kRegion * sizeof(double);
        smem = std::max(smem, static_cast<size_t>(226 * 1024) / w - 1024);
        cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<MODE>, 128, smem);
        const int grid = sms * occ;
        k<MODE><<<grid, 128, smem>>>(out, 16);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<MODE><<<grid, 128, smem>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        const double dmma_per_smsp = static_cast<double>(occ) * iters * 4.0;      // one warp of each CTA per sub-partition
        const double cycles = ms * 1e-3 * mhz * 1e6;
        std::printf("  w=%d(occ %d) %.3f", w, occ, dmma_per_smsp * 16.0 / cycles);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    std::printf("\n");
}

int main(int argc, char** argv) {
    const double mhz = argc > 1 ? std::atof(argv[1]) : 1965.0;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, 64);
    std::printf("fraction of the FP64 tensor pipe kept busy (16 cycles per DMMA.8x8x4 and sub-partition, %.0f MHz assumed), by warps per sub-partition\n", mhz);
    run<0>(out, mhz, sms); run<1>(out, mhz, sms); run<2>(out, mhz, sms); run<3>(out, mhz, sms); run<4>(out, mhz, sms);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { std::printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
