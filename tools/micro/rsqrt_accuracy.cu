// rsqrt_accuracy.cu -- how good is the MUFU.RSQ64H seed (rsqrt.approx.ftz.f64), and r = sqrt(x), 1/r after one and two
// Goldschmidt steps, against the correctly rounded sqrt?  Decides the safety margin of the division-free table index in
// kernels_pair.cuh.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rsqrt_accuracy rsqrt_accuracy.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k(double lo, double hi, int n, double* maxerr) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    double e_seed = 0, e_r1 = 0, e_r2 = 0, e_i1 = 0, e_i2 = 0;
    for (int i = tid; i < n; i += nt) {
        const double x = lo + (hi - lo) * (static_cast<double>(i) + 0.37) / n;
        const double exact = __dsqrt_rn(x);
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        e_seed = fmax(e_seed, fabs(y * exact - 1.0));
        double g = x * y, h = 0.5 * y;
        double e = fma(-h, g, 0.5);
        const double g1 = fma(g, e, g), h1 = fma(h, e, h);
        e_r1 = fmax(e_r1, fabs(g1 - exact) / exact);
        e_i1 = fmax(e_i1, fabs(2.0 * h1 * exact - 1.0));
        e = fma(-h1, g1, 0.5);
        const double g2 = fma(g1, e, g1), h2 = fma(h1, e, h1);
        e_r2 = fmax(e_r2, fabs(g2 - exact) / exact);
        e_i2 = fmax(e_i2, fabs(2.0 * h2 * exact - 1.0));
    }
    // crude max over threads: atomicMax on the bit pattern of non-negative doubles
    atomicMax(reinterpret_cast<unsigned long long*>(maxerr + 0), __double_as_longlong(e_seed));
    atomicMax(reinterpret_cast<unsigned long long*>(maxerr + 1), __double_as_longlong(e_r1));
    atomicMax(reinterpret_cast<unsigned long long*>(maxerr + 2), __double_as_longlong(e_i1));
    atomicMax(reinterpret_cast<unsigned long long*>(maxerr + 3), __double_as_longlong(e_r2));
    atomicMax(reinterpret_cast<unsigned long long*>(maxerr + 4), __double_as_longlong(e_i2));
}

int main() {
    double* d; cudaMalloc(&d, 5 * sizeof(double));
    const double ranges[][2] = {{1.0, 4.0}, {4.0, 400.0}, {1e-6, 1e-3}, {1e3, 1e9}};
    for (auto& r : ranges) {
        cudaMemset(d, 0, 5 * sizeof(double));
        k<<<148 * 8, 256>>>(r[0], r[1], 400000000, d);
        double h[5]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        printf("x in [%g, %g]: max rel err  seed %.3g (2^%.1f) | 1 step: r %.3g (2^%.1f), 1/r %.3g | 2 steps: r %.3g (%.2f ulp), 1/r %.3g\n", r[0], r[1],
               h[0], log2(h[0]), h[1], log2(h[1]), h[2], h[3], h[3] / 1.11e-16, h[4]);
    }
    return 0;
}
