// kernels_ext.cuh -- the scattering-family variants and the virial slice sums.
// Same conventions as kernels.cuh: FP64 throughout, one CTA per (configuration, slice), SoA slice rows
// pos[sl][d][Npad], fixed-order reductions (results do not depend on scheduling).
#pragma once

#include "kernels.cuh"

namespace pimcb {

// ---------------------------------------------------------------------------------------------
// Elastic scattering: the upstream "elastic scattering gpu" estimator launches gpu_isf<true> once per q with
// M/2 + 1 blocks that all atomicAdd 2*inorm*sum into es[q], inorm = 1/(N M)
// (src/estimator_gpu.cu:66-164, 408-413; src/estimator.cpp:4197-4235):
//     es[q] = 2/(N M) sum_{tau=0}^{M/2} sum_t sum_{i,j} cos(q.(r_j(t+tau) - r_i(t))) = (2/M) sum_{tau<=M/2} cfg_isf[q][tau]
// with cfg_isf = isf/N as left in cfg[b][nq + q*M + tau] by the tau-correlation.  One warp per (configuration, q);
// lanes stride tau, fixed butterfly.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) elastic_kernel(const double* __restrict__ cfg, double* __restrict__ out, int B, int nq, int M) {
    const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (pair >= B * nq) return;
    const int b = pair / nq, iq = pair - b * nq;
    const size_t stride = static_cast<size_t>(nq) * (1 + M);
    const double* f = cfg + b * stride + nq + static_cast<size_t>(iq) * M;
    double acc = 0.0;
    for (int tau = lane; tau <= M / 2; tau += 32) acc += f[tau];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[pair] = acc * (2.0 / M);
}

// ---------------------------------------------------------------------------------------------
// Cylinder static structure factor (CylinderStaticStructureFactorEstimator::accumulate, src/estimator.cpp:5415-5456;
// include(r, maxR) = r[0]^2 + r[1]^2 < maxR^2, :4309-4311):
//     partial[sl][q] = sum_{i in} [ 1 + 2 sum_{j > i, j in} cos(q . minimage(r_i - r_j)) ]
// q commensurate with the periodic box (the estimator's own "line" q-set: multiples of 2 pi / L_z along z):
// |sum_{i in} exp(i q.r_i)|^2, one warp per q with lanes over particles; other q: the masked pair loop.
// inside[b] = number of slice-0 beads inside the radius (num1DParticles, :4318-4327).
// ---------------------------------------------------------------------------------------------
template <int ND>
__global__ void __launch_bounds__(256) ssf_cyl_kernel(const double* __restrict__ pos, const double* __restrict__ qsoa,
                                                       const unsigned char* __restrict__ comm, int nq, double maxR,
                                                       double* __restrict__ partial, int* __restrict__ inside, int nslices, int M,
                                                       int N, int Npad, BoxDev box) {
    extern __shared__ __align__(16) double sm[];
    double* xs = sm;                                          // [ND][Npad]
    unsigned char* in = reinterpret_cast<unsigned char*>(sm + ND * Npad);   // [Npad]
    __shared__ int s_count;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const double maxR2 = maxR * maxR;
    for (int sl = blockIdx.x; sl < nslices; sl += gridDim.x) {
        load_slice(xs, pos + static_cast<size_t>(sl) * ND * Npad, ND * Npad);
        if (threadIdx.x == 0) s_count = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < Npad; i += blockDim.x) {
            const double x = xs[i], y = xs[Npad + i];
            const bool ok = i < N && __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)) < maxR2;
            in[i] = ok ? 1 : 0;
            if (ok) atomicAdd(&s_count, 1);
        }
        __syncthreads();
        const int n_in = s_count;
        if (threadIdx.x == 0 && (sl % M) == 0) inside[sl / M] = n_in;
        for (int iq = warp; iq < nq; iq += nwarps) {
            double qv[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) qv[d] = __ldg(qsoa + d * nq + iq);
            double val;
            if (comm[iq]) {
                double cs = 0.0, sn = 0.0;
                for (int i = lane; i < N; i += 32) {
                    if (!in[i]) continue;
                    double ph = qv[0] * xs[i];
#pragma unroll
                    for (int d = 1; d < ND; ++d) ph = fma(qv[d], xs[d * Npad + i], ph);
                    double s, c;
                    sincos_fast(ph, s, c);
                    cs += c;
                    sn += s;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    cs += __shfl_xor_sync(0xffffffffu, cs, o);
                    sn += __shfl_xor_sync(0xffffffffu, sn, o);
                }
                val = fma(cs, cs, sn * sn);
            } else {
                double acc = 0.0;
                for (int i = 0; i < N; ++i) {
                    if (!in[i]) continue;
                    for (int j = i + 1 + lane; j < N; j += 32) {
                        if (!in[j]) continue;
                        double ph = 0.0;
#pragma unroll
                        for (int d = 0; d < ND; ++d) {
                            const double s = xs[d * Npad + i] - xs[d * Npad + j];
                            ph = fma(qv[d], s - box.pSide[d] * floor(fma(s, box.sideInv[d], 0.5)), ph);
                        }
                        double s, c;
                        sincos_fast(ph, s, c);
                        acc += c;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                val = fma(2.0, acc, static_cast<double>(n_in));
            }
            if (lane == 0) partial[static_cast<size_t>(sl) * nq + iq] = val;
        }
        __syncthreads();
    }
}

// out[b][q] = sum_t partial[b*M + t][q]   (t ascending)
__global__ void ssf_cyl_finalize_kernel(const double* __restrict__ partial, double* __restrict__ out, int B, int M, int nq) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * nq) return;
    const int b = idx / nq, iq = idx - b * nq;
    double acc = 0.0;
    for (int t = 0; t < M; ++t) acc += partial[(static_cast<size_t>(b) * M + t) * nq + iq];
    out[idx] = acc;
}

// ---------------------------------------------------------------------------------------------
// Virial slice sums (VirialEnergyEstimator::accumulate, src/estimator.cpp:1086-1250) -- the O(N^2) per-slice parts of
//   LocalAction::rDOTgradUterm1 / term2        (src/action.cpp:1446-1575)   w_i = r_i            (raw position)
//   LocalAction::deltadotgradUterm1 / term2    (src/action.cpp:1588-1784)   w_i = delta_i        (bead - centroid of its
//                                                                            world-line window, from the host)
// For every particle i of the slice, over all partners j != i with sep = minimage(r_i - r_j), r = |sep|,
// k = int(r/dr) (bit-identical to the CPU, see pair_kernel):
//   gV_i  = sum_j (dVdr[k]/r) sep                                        (AzizPotential::gradV, potential.h:997-1003)
//   T_i   = sum_j [ sep sep^T (d2V[k]/r^2 - dV/r^3) + 1 dV/r ],  dV = |(dVdr[k]/r) sep|   (action.cpp:1547-1554, 1722-1729)
// and out[sl] = { sum_i gV_i.r_i, sum_i (T_i gV_i).r_i, sum_i gV_i.delta_i, sum_i (T_i gV_i).delta_i }.
// The prefactors (VFactor tau, 2 gradVFactor tau^3 lambda) are applied by the caller.  t2_parity selects the slices
// whose gradVFactor is finite (-1 all, 0 even, 1 odd, -2 none): the T-matrix terms of the others are returned as 0,
// as upstream skips them (action.cpp:1505, 1680).
// EXT (non-free external potential; both-ends kernel only): gext[sl][d][Npad] = externalPtr->gradV(r_i), g2ext[sl][Npad] =
// externalPtr->grad2V(r_i).  gV_i = gVe_i + sum_j gVi (action.cpp:1471, 1525, 1647, 1704) and, inside the T-matrix, dV = dVi +
// dVe_i with dVe_i = |gVe_i|, d2V = g2Vi + g2Ve_i (action.cpp:1546-1547, 1721-1722).
// ---------------------------------------------------------------------------------------------
#ifndef PIMCB_VIRIAL_UNROLL
#define PIMCB_VIRIAL_UNROLL 2
#endif
#ifndef PIMCB_VIRIAL_MINB
#define PIMCB_VIRIAL_MINB 3
#endif
// Measured on 64 C2 configurations, gsf action (tools/virial_ab.py, gpurun_out r01y): U = 1: 8.21 ms; U = 2 / 3 / 4 at two
// CTAs per SM (98 / 109 / 120 registers): 6.89 / 6.32 / 6.07 ms; U = 2 or 3 held to 80 registers (three CTAs per SM, no
// spill at U = 2): 5.35 ms.  All variants return bit-identical sums.
struct VirialParams {
    const double* dVdr; const double* d2V; int len; double dr; double extdV[2]; double extd2V[2]; int t2_parity; int M;
};

template <int ND, bool EXT>
__global__ void __launch_bounds__(256, PIMCB_VIRIAL_MINB) virial_kernel(const double* __restrict__ pos, const double* __restrict__ delta, int nslices,
                                                         int N, int Npad, BoxDev box, VirialParams vp, double* __restrict__ out,
                                                         const double* __restrict__ gext, const double* __restrict__ g2ext) {
    constexpr int NT = ND * (ND + 1) / 2;
    constexpr int U = PIMCB_VIRIAL_UNROLL;
    extern __shared__ __align__(16) double sm[];
    double* xs = sm;                                  // [ND][Npad]
    double* ds = sm + ND * Npad;                      // [ND][Npad] (delta)
    __shared__ double red[4][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int sl = blockIdx.x; sl < nslices; sl += gridDim.x) {
        const int t = sl % vp.M;
        const bool do_t2 = vp.t2_parity == -1 || (vp.t2_parity >= 0 && (t & 1) == vp.t2_parity);
        load_slice(xs, pos + static_cast<size_t>(sl) * ND * Npad, ND * Npad);
        if (delta) load_slice(ds, delta + static_cast<size_t>(sl) * ND * Npad, ND * Npad);
        __syncthreads();
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            double gV[ND], T[NT];
#pragma unroll
            for (int d = 0; d < ND; ++d) gV[d] = 0.0;
#pragma unroll
            for (int k = 0; k < NT; ++k) T[k] = 0.0;
            double gVe[ND], dVe = 0.0, g2Ve = 0.0;
            if constexpr (EXT) {
                double e2 = 0.0;
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    gVe[d] = gext ? __ldg(gext + static_cast<size_t>(sl) * ND * Npad + d * Npad + i) : 0.0;
                    e2 = __dadd_rn(e2, __dmul_rn(gVe[d], gVe[d]));
                }
                dVe = __dsqrt_rn(e2);                                   // sqrt(dot(gVe,gVe)), action.cpp:1522
                g2Ve = g2ext ? __ldg(g2ext + static_cast<size_t>(sl) * Npad + i) : 0.0;
            }
            // U partners in flight per thread, consumed in partner order (sums identical to the one-at-a-time loop): the
            // gathers are the long pole, exactly as in pair_kernel (52 % of the stall samples of the U = 1 version were a
            // warp waiting for its single outstanding table read; profiles/r01x_kernels.md)
            for (int kk0 = 1; kk0 < N; kk0 += U) {
                double sep[U][ND], r[U], dv[U], d2[U];
#pragma unroll
                for (int w = 0; w < U; ++w) {
                    int j = i + ring_partner(min(kk0 + w, N - 1), N, true);       // zig-zag ring (pair_kernel); surplus slots recompute the last partner
                    if (j >= N) j -= N;
                    r[w] = minimage_norm<ND>(xs, Npad, i, j, box, sep[w]);        // getSeparation(bead1, bead2)
                }
#pragma unroll
                for (int w = 0; w < U; ++w) {                                     // all gathers of the group are issued here
                    const int kidx = __double2int_rz(__ddiv_rn(r[w], vp.dr));
                    const bool inside = kidx > 0 && kidx < vp.len;
                    dv[w] = inside ? __ldg(vp.dVdr + kidx) : (kidx <= 0 ? vp.extdV[0] : vp.extdV[1]);
                    d2[w] = 0.0;
                    if (do_t2) d2[w] = inside ? __ldg(vp.d2V + kidx) : (kidx <= 0 ? vp.extd2V[0] : vp.extd2V[1]);
                }
#pragma unroll
                for (int w = 0; w < U; ++w) {
                    if (kk0 + w >= N) continue;
                    const double g = dv[w] / r[w];
                    double gi[ND], g2 = 0.0;
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        gi[d] = __dmul_rn(g, sep[w][d]);           // gVi as its own rounded value, then gV += gVi (action.cpp:1466, 1556)
                        gV[d] = __dadd_rn(gV[d], gi[d]);
                        g2 = fma(gi[d], gi[d], g2);
                    }
                    if (do_t2) {
                        const double dV = EXT ? sqrt(g2) + dVe : sqrt(g2);
                        const double d2V = EXT ? d2[w] + g2Ve : d2[w];
                        const double rinv = 1.0 / r[w];
                        const double a = d2V * rinv * rinv - dV * rinv * rinv * rinv;
                        const double diag = dV * rinv;
                        int k = 0;
#pragma unroll
                        for (int p = 0; p < ND; ++p)
#pragma unroll
                            for (int q = p; q < ND; ++q, ++k) T[k] = fma(sep[w][p] * sep[w][q], a, T[k]) + (p == q ? diag : 0.0);
                    }
                }
            }
            if constexpr (EXT) {
#pragma unroll
                for (int d = 0; d < ND; ++d) gV[d] += gVe[d];
            }
            double u[ND];
#pragma unroll
            for (int p = 0; p < ND; ++p) u[p] = 0.0;
            {
                int k = 0;
#pragma unroll
                for (int p = 0; p < ND; ++p)
#pragma unroll
                    for (int q = p; q < ND; ++q, ++k) {
                        u[p] = fma(T[k], gV[q], u[p]);
                        if (q != p) u[q] = fma(T[k], gV[p], u[q]);
                    }
            }
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                const double x = xs[d * Npad + i];
                acc[0] = fma(gV[d], x, acc[0]);
                acc[1] = fma(u[d], x, acc[1]);
                if (delta) {
                    const double dl = ds[d * Npad + i];
                    acc[2] = fma(gV[d], dl, acc[2]);
                    acc[3] = fma(u[d], dl, acc[3]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double v = acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[k][warp] = v;
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            double v = 0.0;
            for (int w = 0; w < (blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
            out[static_cast<size_t>(sl) * 4 + threadIdx.x] = v;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Symmetric form of the virial slice sums: every pair is visited ONCE.  The table entries of a pair are the same seen
// from either end (r_ij = r_ji bit for bit), and the per-pair terms are antisymmetric (gV) or symmetric (T), so thread
// i walks only the half ring k = 1..N/2, keeps its own side in registers and adds the partner's side into
// shared-memory accumulators acc[component][j].  Within one ring step k all partners j = i + k are distinct, so a plain
// read-add-write needs no atomics; a CTA barrier separates the steps (~N/2 barriers per slice against ~N^2/2 table
// gathers).  Half the gathers of virial_kernel -- the gathers are what both kernels wait for.  Summation order is fixed
// (own side in ring order, partner side in ring order), so results are reproducible; they differ from virial_kernel in
// the last bits only.  PPT = particles per thread (N <= 256 * PPT).
// ---------------------------------------------------------------------------------------------
#ifndef PIMCB_VSYM_U
#define PIMCB_VSYM_U 2
#endif
#ifndef PIMCB_VSYM_MINB
#define PIMCB_VSYM_MINB 3
#endif
template <int ND, int PPT>
__global__ void __launch_bounds__(256, PPT == 1 ? PIMCB_VSYM_MINB : (PPT == 2 ? 2 : 1))
virial_sym_kernel(const double* __restrict__ pos, const double* __restrict__ delta, int nslices, int N, int Npad, BoxDev box,
                  VirialParams vp, double* __restrict__ out) {
    constexpr int NT = ND * (ND + 1) / 2;
    constexpr int NC = ND + NT;                       // accumulated components per particle: gV then the upper triangle of T
    constexpr int U = PIMCB_VSYM_U;
    extern __shared__ __align__(16) double sm[];
    double* xs = sm;                                  // [ND][Npad]
    double* ds = sm + ND * Npad;                      // [ND][Npad] (delta)
    double* acc = sm + 2 * ND * Npad;                 // [NC][Npad] partner-side accumulators
    __shared__ double red[4][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int khalf = N / 2;
    const bool evenN = (N & 1) == 0;
    for (int sl = blockIdx.x; sl < nslices; sl += gridDim.x) {
        const int t = sl % vp.M;
        const bool do_t2 = vp.t2_parity == -1 || (vp.t2_parity >= 0 && (t & 1) == vp.t2_parity);
        load_slice(xs, pos + static_cast<size_t>(sl) * ND * Npad, ND * Npad);
        if (delta) load_slice(ds, delta + static_cast<size_t>(sl) * ND * Npad, ND * Npad);
        for (int k = threadIdx.x; k < NC * Npad; k += blockDim.x) acc[k] = 0.0;
        __syncthreads();
        double own[PPT][NC];
#pragma unroll
        for (int p = 0; p < PPT; ++p)
#pragma unroll
            for (int c = 0; c < NC; ++c) own[p][c] = 0.0;
        for (int kk0 = 1; kk0 <= khalf; kk0 += U) {
            double sep[U][PPT][ND], r[U][PPT], dv[U][PPT], d2[U][PPT];
            int jj[U][PPT];
#pragma unroll
            for (int w = 0; w < U; ++w)
#pragma unroll
                for (int p = 0; p < PPT; ++p) {
                    const int i = threadIdx.x + 256 * p, kk = kk0 + w;
                    // the pair (i, i + N/2) of an even ring belongs to the lower half only
                    const bool valid = i < N && kk <= khalf && !(evenN && kk == khalf && i >= khalf);
                    int j = i + kk;
                    if (j >= N) j -= N;
                    jj[w][p] = valid ? j : -1;
                    r[w][p] = 1.0;
                    if (valid) r[w][p] = minimage_norm<ND>(xs, Npad, i, j, box, sep[w][p]);     // getSeparation(bead1, bead2)
                }
#pragma unroll
            for (int w = 0; w < U; ++w)
#pragma unroll
                for (int p = 0; p < PPT; ++p) {                                   // all gathers of the group are issued here
                    dv[w][p] = 0.0;
                    d2[w][p] = 0.0;
                    if (jj[w][p] >= 0) {
                        const int kidx = __double2int_rz(__ddiv_rn(r[w][p], vp.dr));
                        const bool inside = kidx > 0 && kidx < vp.len;
                        dv[w][p] = inside ? __ldg(vp.dVdr + kidx) : (kidx <= 0 ? vp.extdV[0] : vp.extdV[1]);
                        if (do_t2) d2[w][p] = inside ? __ldg(vp.d2V + kidx) : (kidx <= 0 ? vp.extd2V[0] : vp.extd2V[1]);
                    }
                }
#pragma unroll
            for (int w = 0; w < U; ++w) {
#pragma unroll
                for (int p = 0; p < PPT; ++p) {
                    const int j = jj[w][p];
                    if (j < 0) continue;
                    const double g = dv[w][p] / r[w][p];
                    double gi[ND], g2 = 0.0;
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        gi[d] = __dmul_rn(g, sep[w][p][d]);
                        g2 = fma(gi[d], gi[d], g2);
                        own[p][d] = __dadd_rn(own[p][d], gi[d]);
                        acc[d * Npad + j] = __dsub_rn(acc[d * Npad + j], gi[d]);  // gradV(sep_ji) = -gradV(sep_ij)
                    }
                    if (do_t2) {
                        const double dV = sqrt(g2);
                        const double rinv = 1.0 / r[w][p];
                        const double a = d2[w][p] * rinv * rinv - dV * rinv * rinv * rinv;
                        const double diag = dV * rinv;
                        int k = ND;
#pragma unroll
                        for (int pp = 0; pp < ND; ++pp)
#pragma unroll
                            for (int q = pp; q < ND; ++q, ++k) {
                                const double m = fma(sep[w][p][pp] * sep[w][p][q], a, pp == q ? diag : 0.0);
                                own[p][k] += m;
                                acc[k * Npad + j] += m;
                            }
                    }
                }
                __syncthreads();                      // the next ring step targets other j: order the read-add-writes
            }
        }
        double sums[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const int i = threadIdx.x + 256 * p;
            if (i >= N) continue;
            double gV[ND], T[NT];
#pragma unroll
            for (int d = 0; d < ND; ++d) gV[d] = own[p][d] + acc[d * Npad + i];
#pragma unroll
            for (int k = 0; k < NT; ++k) T[k] = own[p][ND + k] + acc[(ND + k) * Npad + i];
            double uu[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) uu[d] = 0.0;
            {
                int k = 0;
#pragma unroll
                for (int pp = 0; pp < ND; ++pp)
#pragma unroll
                    for (int q = pp; q < ND; ++q, ++k) {
                        uu[pp] = fma(T[k], gV[q], uu[pp]);
                        if (q != pp) uu[q] = fma(T[k], gV[pp], uu[q]);
                    }
            }
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                const double x = xs[d * Npad + i];
                sums[0] = fma(gV[d], x, sums[0]);
                sums[1] = fma(uu[d], x, sums[1]);
                if (delta) {
                    const double dl = ds[d * Npad + i];
                    sums[2] = fma(gV[d], dl, sums[2]);
                    sums[3] = fma(uu[d], dl, sums[3]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double v = sums[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[k][warp] = v;
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            double v = 0.0;
            for (int w = 0; w < (blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
            out[static_cast<size_t>(sl) * 4 + threadIdx.x] = v;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Symmetric form of pair_kernel<ND, true>: on the slices that carry the gradient correction every pair is visited ONCE
// (half ring); the visiting thread keeps its own side of the force in registers and subtracts the partner's side in a
// shared-memory accumulator (distinct partners within a ring step, CTA barrier between steps) -- two table gathers per
// pair (V, dV/dr) instead of three.  Slices without the correction run the same half ring without accumulators or
// barriers.  Vint, the table indices and sepHist are bit-identical to pair_kernel (same r, same order of the V sum for
// N <= 256); gradVSquared differs in the last bits (other summation order) and is reproducible.
// ---------------------------------------------------------------------------------------------
template <int ND, int PPT>
__global__ void __launch_bounds__(256, PPT == 1 ? 4 : (PPT == 2 ? 2 : 1))
pair_sym_kernel(const double* __restrict__ pos, int nslices, int N, int Npad, BoxDev box, PairParams pp, double* __restrict__ vint,
                double* __restrict__ f2, int* __restrict__ hist) {
    constexpr int U = 2;
    extern __shared__ __align__(16) double sm[];
    double* xs = sm;                                  // [ND][Npad]
    double* acc = sm + ND * Npad;                     // [ND][Npad] partner-side force accumulators
    __shared__ double redV[8], redF[8];
    __shared__ int shist[kNPCFSEP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int khalf = N / 2;
    const bool evenN = (N & 1) == 0;
    for (int sl = blockIdx.x; sl < nslices; sl += gridDim.x) {
        const int t = sl % pp.M;
        const bool do_f2 = pp.f2_parity < 0 || (t & 1) == pp.f2_parity;
        load_slice(xs, pos + static_cast<size_t>(sl) * ND * Npad, ND * Npad);
        if (threadIdx.x < kNPCFSEP) shist[threadIdx.x] = 0;
        if (do_f2)
            for (int k = threadIdx.x; k < ND * Npad; k += blockDim.x) acc[k] = 0.0;
        __syncthreads();
        double vsum = 0.0, own[PPT][ND];
#pragma unroll
        for (int p = 0; p < PPT; ++p)
#pragma unroll
            for (int d = 0; d < ND; ++d) own[p][d] = 0.0;
        for (int kk0 = 1; kk0 <= khalf; kk0 += U) {
            double sep[U][PPT][ND], r[U][PPT], vv[U][PPT], dv[U][PPT];
            int jj[U][PPT];
#pragma unroll
            for (int w = 0; w < U; ++w)
#pragma unroll
                for (int p = 0; p < PPT; ++p) {
                    const int i = threadIdx.x + 256 * p, kk = kk0 + w;
                    const bool valid = i < N && kk <= khalf && !(evenN && kk == khalf && i >= khalf);
                    int j = i + kk;
                    if (j >= N) j -= N;
                    jj[w][p] = valid ? j : -1;
                    r[w][p] = 1.0;
                    if (valid) {
                        if (do_f2) {
                            r[w][p] = minimage_norm<ND>(xs, Npad, i, j, box, sep[w][p]);            // getSeparation(bead1,bead2), action.cpp:1211
                        } else {
                            const int lo = min(i, j), hi = max(i, j);
                            r[w][p] = minimage_norm<ND>(xs, Npad, hi, lo, box, sep[w][p]);          // getSeparation(bead2,bead1), action.cpp:934
                        }
                    }
                }
#pragma unroll
            for (int w = 0; w < U; ++w)
#pragma unroll
                for (int p = 0; p < PPT; ++p) {                                   // all gathers of the group are issued here
                    vv[w][p] = 0.0;
                    dv[w][p] = 0.0;
                    if (jj[w][p] >= 0) {
                        const int kidx = __double2int_rz(__ddiv_rn(r[w][p], pp.dr));
                        const bool inside = kidx > 0 && kidx < pp.len;
                        vv[w][p] = inside ? __ldg(pp.V + kidx) : (kidx <= 0 ? pp.extV[0] : pp.extV[1]);
                        if (do_f2) dv[w][p] = inside ? __ldg(pp.dVdr + kidx) : (kidx <= 0 ? pp.extdV[0] : pp.extdV[1]);
                    }
                }
#pragma unroll
            for (int w = 0; w < U; ++w) {
#pragma unroll
                for (int p = 0; p < PPT; ++p) {
                    const int j = jj[w][p];
                    if (j < 0) continue;
                    vsum += vv[w][p];
                    if (pp.want_hist) {
                        const int nR = __double2int_rz(__ddiv_rn(r[w][p], pp.dSep));               // action.cpp:221
                        if (nR >= 0 && nR < kNPCFSEP) atomicAdd(&shist[nR], 1);
                    }
                    if (do_f2) {
                        const double g = __ddiv_rn(dv[w][p], r[w][p]);
#pragma unroll
                        for (int d = 0; d < ND; ++d) {
                            own[p][d] = fma(g, sep[w][p][d], own[p][d]);
                            acc[d * Npad + j] = fma(-g, sep[w][p][d], acc[d * Npad + j]);          // gradV(sep_ji) = -gradV(sep_ij)
                        }
                    }
                }
                if (do_f2) __syncthreads();           // the next ring step targets other partners: order the read-add-writes
            }
        }
        double fsum = 0.0;
        if (do_f2) {
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                const int i = threadIdx.x + 256 * p;
                if (i >= N) continue;
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    double F = own[p][d] + acc[d * Npad + i];
                    if (pp.gext) F += __ldg(pp.gext + static_cast<size_t>(sl) * ND * Npad + d * Npad + i);   // action.cpp:1216
                    fsum = fma(F, F, fsum);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
            fsum += __shfl_xor_sync(0xffffffffu, fsum, o);
        }
        if (lane == 0) { redV[warp] = vsum; redF[warp] = fsum; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double v = 0.0, f = 0.0;
            for (int w = 0; w < (blockDim.x >> 5); ++w) { v += redV[w]; f += redF[w]; }
            vint[sl] = v;
            if (f2) f2[sl] = f;
        }
        if (pp.want_hist && threadIdx.x < kNPCFSEP) hist[static_cast<size_t>(sl) * kNPCFSEP + threadIdx.x] = shist[threadIdx.x];
        __syncthreads();
    }
}

}  // namespace pimcb
