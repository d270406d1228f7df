// gather_peak.cu -- how many random table reads per second does one B200 deliver, as a function of the table's footprint
// and of the read width (8 bytes = what TabulatedPotential::direct reads; 32 bytes = one whole sector, LDG.E.256)?
// This is the roofline of the pair-potential kernels: every pair of beads costs one read at an index that is
// uncorrelated with its neighbours' (dr = 1e-6 rm, a 32-byte sector spans 1.2e-5 Angstrom of separation).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_peak gather_peak.cu && ./gather_peak
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ unsigned mix(unsigned x) {   // xorshift-multiply hash; cheap next to the load
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int W, int U>   // W = bytes per read (8 or 32), U = reads in flight per thread
__global__ void __launch_bounds__(256) gather_kernel(const unsigned long long* __restrict__ tab, unsigned nelem, int iters,
                                                     unsigned long long* __restrict__ out) {
    unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    unsigned long long acc = 0;
    for (int it = 0; it < iters; ++it) {
        unsigned idx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { s = mix(s + 0x9e3779b9u); idx[u] = static_cast<unsigned>((static_cast<unsigned long long>(s) * nelem) >> 32); }
        if constexpr (W == 8) {
            unsigned long long v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = __ldg(tab + idx[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) acc += v[u];
        } else {
            unsigned long long a[U], b[U], c[U], d[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a[u]), "=l"(b[u]), "=l"(c[u]), "=l"(d[u])
                             : "l"(tab + 4 * static_cast<size_t>(idx[u])));
#pragma unroll
            for (int u = 0; u < U; ++u) acc += (a[u] ^ b[u]) + (c[u] ^ d[u]);
        }
    }
    if (acc == 0x1234567ull) out[0] = acc;
}

template <int W, int U>
double run(const unsigned long long* tab, size_t bytes, unsigned long long* out, int ctas_per_sm) {
    const unsigned nelem = static_cast<unsigned>(bytes / W);
    const int grid = 148 * ctas_per_sm, iters = 2048 / U;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    gather_kernel<W, U><<<grid, 256>>>(tab, nelem, iters, out);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 3; ++r) gather_kernel<W, U><<<grid, 256>>>(tab, nelem, iters, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return 3.0 * grid * 256.0 * iters * U / (ms * 1e-3) / 1e9;   // G reads / s
}

int main() {
    const size_t maxb = 256u << 20;
    unsigned long long *tab, *out;
    cudaMalloc(&tab, maxb);
    cudaMalloc(&out, 64);
    cudaMemset(tab, 1, maxb);
    const double mb[] = {2, 8, 13.3, 26.5, 40, 53, 80, 106, 160, 256};
    printf("G reads/s (x W bytes = useful GB/s; x 32 = sector GB/s)\n");
    printf("%8s | %10s %10s %10s | %10s %10s %10s | occupancy 8 CTAs/SM x 256 thr\n", "MB", "W8 U2", "W8 U4", "W8 U8", "W32 U2", "W32 U4", "W32 U8");
    for (double m : mb) {
        const size_t bytes = static_cast<size_t>(m * (1 << 20)) & ~size_t(255);
        printf("%8.1f | %10.1f %10.1f %10.1f | %10.1f %10.1f %10.1f\n", m, run<8, 2>(tab, bytes, out, 8), run<8, 4>(tab, bytes, out, 8),
               run<8, 8>(tab, bytes, out, 8), run<32, 2>(tab, bytes, out, 8), run<32, 4>(tab, bytes, out, 8), run<32, 8>(tab, bytes, out, 8));
    }
    printf("half occupancy (4 CTAs/SM):\n");
    for (double m : {26.5, 53.0, 106.0}) {
        const size_t bytes = static_cast<size_t>(m * (1 << 20)) & ~size_t(255);
        printf("%8.1f | %10.1f %10.1f %10.1f | %10.1f %10.1f %10.1f\n", m, run<8, 2>(tab, bytes, out, 4), run<8, 4>(tab, bytes, out, 4),
               run<8, 8>(tab, bytes, out, 4), run<32, 2>(tab, bytes, out, 4), run<32, 4>(tab, bytes, out, 4), run<32, 8>(tab, bytes, out, 4));
    }
    return 0;
}
