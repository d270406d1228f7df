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

// Timing of the upstream kernels for the A/B leg of bench.py (roofline.upstream_gpu_ab; SURVEY.md section 2 sets "beats
// src/estimator_gpu.cu compiled for sm_100 on the same box" as the bar).  One evaluation = what the two upstream GPU
// estimators do per measurement: one gpu_ssf_launcher for all q + one gpu_isf_launcher per q (tau = 0..M/2).
//   ms[0] = kernels only, beads resident on the device (CUDA events, average over `reps` evaluations)
//   ms[1] = the upstream accumulate() pattern: each estimator copies the padded beads array to the device, launches,
//           synchronises and copies its result back (src/estimator.cpp:3833-3861, 4074-4100), wall clock
extern "C" int ref_gpu_bench(const double* beads, int M, int N, int Next, const double* q, int nq, int reps, double* ms) {
    int rc = 0;
    double *d_b = nullptr, *d_q = nullptr, *d_s = nullptr, *d_o = nullptr, *h_s = nullptr, *h_o = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const int nt = M / 2 + 1;
    const size_t nb = sizeof(double) * static_cast<size_t>(M) * Next * NDIM;
    float t = 0.f;
    REFCU(cudaMalloc(&d_b, nb));
    REFCU(cudaMalloc(&d_q, sizeof(double) * nq * NDIM));
    REFCU(cudaMalloc(&d_s, sizeof(double) * nq));
    REFCU(cudaMalloc(&d_o, sizeof(double) * nq * nt));
    REFCU(cudaMallocHost(&h_s, sizeof(double) * nq));
    REFCU(cudaMallocHost(&h_o, sizeof(double) * nq * nt));
    REFCU(cudaEventCreate(&e0));
    REFCU(cudaEventCreate(&e1));
    REFCU(cudaMemcpy(d_b, beads, nb, cudaMemcpyHostToDevice));
    REFCU(cudaMemcpy(d_q, q, sizeof(double) * nq * NDIM, cudaMemcpyHostToDevice));
    for (int pass = 0; pass < 2; ++pass) {                     // pass 0 = warm-up
        const int n = pass ? reps : 1;
        REFCU(cudaEventRecord(e0));
        for (int r = 0; r < n; ++r) {
            gpu_ssf_launcher(d_s, d_q, d_b, 1.0 / N, M, N, Next, nq);
            for (int k = 0; k < nq; ++k)
                gpu_isf_launcher(d_o + static_cast<size_t>(nt) * k, d_q + NDIM * k, d_b, 1.0 / (static_cast<double>(N) * M), M, N, Next);
        }
        REFCU(cudaEventRecord(e1));
        REFCU(cudaEventSynchronize(e1));
        REFCU(cudaEventElapsedTime(&t, e0, e1));
        ms[0] = t / n;
    }
    REFCU(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) {
        REFCU(cudaMemcpy(d_b, beads, nb, cudaMemcpyHostToDevice));
        gpu_ssf_launcher(d_s, d_q, d_b, 1.0 / N, M, N, Next, nq);
        REFCU(cudaDeviceSynchronize());
        REFCU(cudaMemcpy(h_s, d_s, sizeof(double) * nq, cudaMemcpyDeviceToHost));
        REFCU(cudaMemcpy(d_b, beads, nb, cudaMemcpyHostToDevice));
        for (int k = 0; k < nq; ++k)
            gpu_isf_launcher(d_o + static_cast<size_t>(nt) * k, d_q + NDIM * k, d_b, 1.0 / (static_cast<double>(N) * M), M, N, Next);
        REFCU(cudaDeviceSynchronize());
        REFCU(cudaMemcpy(h_o, d_o, sizeof(double) * nq * nt, cudaMemcpyDeviceToHost));
    }
    REFCU(cudaEventRecord(e1));
    REFCU(cudaEventSynchronize(e1));
    REFCU(cudaEventElapsedTime(&t, e0, e1));
    ms[1] = t / reps;
    REFCU(cudaGetLastError());
done:
    cudaFree(d_b); cudaFree(d_q); cudaFree(d_s); cudaFree(d_o);
    if (h_s) cudaFreeHost(h_s);
    if (h_o) cudaFreeHost(h_o);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
}

extern "C" int ref_gpu_block_size(void) {
#ifdef GPU_BLOCK_SIZE
    return GPU_BLOCK_SIZE;
#else
    return 256;            // include/common_gpu.h:9-11 (the header's default; upstream's CMake passes 1024)
#endif
}
