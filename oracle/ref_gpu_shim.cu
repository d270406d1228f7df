// ref_gpu_shim.cu -- test infrastructure (never linked by the product): C entry points around the REFERENCE'S OWN GPU
// estimator kernels, compiled unmodified from /root/reference/src/estimator_gpu.cu (oracle/Makefile, target `ref`).
// Only the launcher prototypes are repeated here (include/estimator_gpu.cuh:6-16); the host-side call pattern follows
// StaticStructureFactorGPUEstimator::accumulate (src/estimator.cpp:3822-3861: inorm = 1/N, one launch for all q) and
// IntermediateScatteringFunctionEstimatorGpu::accumulate (src/estimator.cpp:4063-4100: inorm = 1/(N M), one launch per
// q, tau = 0..M/2).  Host buffers in, host buffers out; the full padded AoS beads array is copied as upstream does.
#include <cuda_runtime.h>

void gpu_isf_launcher(double* isf, double* qvecs, double* beads, double inorm, int M, int N, int N_extent);
void gpu_ssf_launcher(double* ssf, double* qvecs, double* beads, double inorm, int M, int N, int N_extent, int n_qvecs);
void gpu_es_launcher(double* isf, double* qvecs, double* beads, double inorm, int M, int N, int N_extent);

#define REFCU(x) do { if ((x) != cudaSuccess) { rc = -1; goto done; } } while (0)

extern "C" int ref_ndim(void) { return NDIM; }

// out[nq]: exactly what the reference adds to `estimator` per measurement (its norm is 0.5/M, src/estimator.cpp:3789)
extern "C" int ref_gpu_ssf(const double* beads, int M, int N, int Next, const double* q, int nq, double* out) {
    int rc = 0;
    double *d_b = nullptr, *d_q = nullptr, *d_o = nullptr;
    const size_t nb = sizeof(double) * static_cast<size_t>(M) * Next * NDIM;
    REFCU(cudaMalloc(&d_b, nb));
    REFCU(cudaMalloc(&d_q, sizeof(double) * nq * NDIM));
    REFCU(cudaMalloc(&d_o, sizeof(double) * nq));
    REFCU(cudaMemcpy(d_b, beads, nb, cudaMemcpyHostToDevice));
    REFCU(cudaMemcpy(d_q, q, sizeof(double) * nq * NDIM, cudaMemcpyHostToDevice));
    gpu_ssf_launcher(d_o, d_q, d_b, 1.0 / N, M, N, Next, nq);
    REFCU(cudaGetLastError());
    REFCU(cudaDeviceSynchronize());
    REFCU(cudaMemcpy(out, d_o, sizeof(double) * nq, cudaMemcpyDeviceToHost));
done:
    cudaFree(d_b); cudaFree(d_q); cudaFree(d_o);
    return rc;
}

// out[nq][M/2 + 1]: the reference GPU estimator's per-measurement increment (norm 0.5, src/estimator.cpp:4033-4034)
extern "C" int ref_gpu_isf(const double* beads, int M, int N, int Next, const double* q, int nq, double* out) {
    int rc = 0;
    double *d_b = nullptr, *d_q = nullptr, *d_o = nullptr;
    const int nt = M / 2 + 1;
    const size_t nb = sizeof(double) * static_cast<size_t>(M) * Next * NDIM;
    REFCU(cudaMalloc(&d_b, nb));
    REFCU(cudaMalloc(&d_q, sizeof(double) * nq * NDIM));
    REFCU(cudaMalloc(&d_o, sizeof(double) * nq * nt));
    REFCU(cudaMemcpy(d_b, beads, nb, cudaMemcpyHostToDevice));
    REFCU(cudaMemcpy(d_q, q, sizeof(double) * nq * NDIM, cudaMemcpyHostToDevice));
    for (int k = 0; k < nq; ++k)
        gpu_isf_launcher(d_o + static_cast<size_t>(nt) * k, d_q + NDIM * k, d_b, 1.0 / (static_cast<double>(N) * M), M, N, Next);
    REFCU(cudaGetLastError());
    REFCU(cudaDeviceSynchronize());
    REFCU(cudaMemcpy(out, d_o, sizeof(double) * nq * nt, cudaMemcpyDeviceToHost));
done:
    cudaFree(d_b); cudaFree(d_q); cudaFree(d_o);
    return rc;
}

// out[nq]: the upstream "elastic scattering gpu" estimator's per-measurement increment (src/estimator.cpp:4197-4235:
// memset, then one gpu_es_launcher per q into d_es + nq with inorm = 1/(N M); its norm is 0.5, :4161)
extern "C" int ref_gpu_es(const double* beads, int M, int N, int Next, const double* q, int nq, double* out) {
    int rc = 0;
    double *d_b = nullptr, *d_q = nullptr, *d_o = nullptr;
    const size_t nb = sizeof(double) * static_cast<size_t>(M) * Next * NDIM;
    REFCU(cudaMalloc(&d_b, nb));
    REFCU(cudaMalloc(&d_q, sizeof(double) * nq * NDIM));
    REFCU(cudaMalloc(&d_o, sizeof(double) * nq));
    REFCU(cudaMemcpy(d_b, beads, nb, cudaMemcpyHostToDevice));
    REFCU(cudaMemcpy(d_q, q, sizeof(double) * nq * NDIM, cudaMemcpyHostToDevice));
    REFCU(cudaMemset(d_o, 0, sizeof(double) * nq));
    for (int k = 0; k < nq; ++k)
        gpu_es_launcher(d_o + k, d_q + NDIM * k, d_b, 1.0 / (static_cast<double>(N) * M), M, N, Next);
    REFCU(cudaGetLastError());
    REFCU(cudaDeviceSynchronize());
    REFCU(cudaMemcpy(out, d_o, sizeof(double) * nq, cudaMemcpyDeviceToHost));
done:
    cudaFree(d_b); cudaFree(d_q); cudaFree(d_o);
    return rc;
}
