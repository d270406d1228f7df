// kernels.cuh -- sm_100a device code for the pimc measurement hot path (FP64, CUDA cores).
//
// Device bead layout ("slice-major SoA"):  pos[b][t][d][Npad]  (double), Npad = N rounded up to 16,
// so that the ND coordinate rows of one time slice are one contiguous, 128-byte aligned chunk that a
// CTA pulls into shared memory with coalesced 16-byte loads.
//
// Kernels
//   rho_generic_kernel   rho_q(t) = sum_i exp(i q.r_i(t)), one sincos per (q, bead)      [FP64-pipe bound]
//   rho_lattice_kernel   same for commensurate q = 2 pi n / L: phase-power tables + sign-symmetry groups [smem-bandwidth bound]
//   isf_corr_kernel      F(q,tau) = (1/N) sum_t0 Re[rho(t0) conj rho(t0+tau)], S(q) = F(q,0)
//   ssf_direct_kernel    sum_{i<j} cos(q.minimage(r_i-r_j)) for non-commensurate q
//   pair_kernel          per-slice Vint, sum_i |F_i|^2, separation histogram (table gathers)
//   bins_accumulate_kernel, ssf_direct_finalize_kernel, aos_to_soa_kernel, fp64_peak_kernel
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pimcb {

constexpr int kNPCFSEP = 50;   // include/common.h:85 (upstream)

struct BoxDev {
    double sideInv[3];
    double pSide[3];
};

// ---------------------------------------------------------------------------------------------
// sincos for moderate arguments (|x| < ~1e5): 3-term Cody-Waite reduction by pi/2 with FMAs and the
// classic minimax kernels on [-pi/4, pi/4] (< 1 ulp each).  21 FP64 instructions, no slow path,
// no local memory.  The host refuses q-sets whose max |q.r| leaves the validity range.
// ---------------------------------------------------------------------------------------------
// Coefficients live in the constant bank so that every DFMA takes them as a c[bank][offset] operand; written as
// literals ptxas re-materialises all of them into uniform registers (30 UMOVs) on every loop iteration.
__constant__ double kSC[20] = {
    0.63661977236758134308,        // [0]  2/pi
    6755399441055744.0,            // [1]  1.5 * 2^52: rint() by add/sub
    -1.5707963267948965580e+00,    // [2]  -pi/2 split in three doubles
    -6.1232339957367660359e-17,    // [3]
    1.4973849048591698330e-33,     // [4]
    1.58969099521155010221e-10,    // [5]  sin: S6..S1
    -2.50507602534068634195e-08,   // [6]
    2.75573137070700676789e-06,    // [7]
    -1.98412698298579493134e-04,   // [8]
    8.33333333332248946124e-03,    // [9]
    -1.66666666666666324348e-01,   // [10]
    -1.13596475577881948265e-11,   // [11] cos: C6..C1
    2.08757232129817482790e-09,    // [12]
    -2.75573143513906633035e-07,   // [13]
    2.48015872894767294178e-05,    // [14]
    -1.38888888888741095749e-03,   // [15]
    4.16666666666666019037e-02,    // [16]
    -0.5, 1.0, 0.0};

__device__ __forceinline__ void sincos_fast(double x, double& s, double& c) {
    const double t = fma(x, kSC[0], kSC[1]);
    const int n = __double2loint(t);                     // quadrant = low bits of rint(x*2/pi)
    const double kd = t - kSC[1];
    double r = fma(kd, kSC[2], x);
    r = fma(kd, kSC[3], r);
    r = fma(kd, kSC[4], r);
    const double z = r * r;
    double ps = fma(z, kSC[5], kSC[6]);
    double pc = fma(z, kSC[11], kSC[12]);
    ps = fma(z, ps, kSC[7]);
    pc = fma(z, pc, kSC[13]);
    ps = fma(z, ps, kSC[8]);
    pc = fma(z, pc, kSC[14]);
    ps = fma(z, ps, kSC[9]);
    pc = fma(z, pc, kSC[15]);
    ps = fma(z, ps, kSC[10]);
    pc = fma(z, pc, kSC[16]);
    const double sr = fma(r * z, ps, r);                                   // sin(r)
    const double cr = fma(z * z, pc, fma(z, kSC[17], kSC[18]));            // cos(r)
    double sv = (n & 1) ? cr : sr;
    double cv = (n & 1) ? sr : cr;
    // sign: sin negative in quadrants 2,3; cos negative in quadrants 1,2
    const int shi = __double2hiint(sv) ^ ((n & 2) << 30);
    const int chi = __double2hiint(cv) ^ (((n + 1) & 2) << 30);
    s = __hiloint2double(shi, __double2loint(sv));
    c = __hiloint2double(chi, __double2loint(cv));
}

// Coalesced load of one slice (ND rows of Npad doubles, contiguous, 16-byte aligned) into shared memory.
__device__ __forceinline__ void load_slice(double* __restrict__ sm, const double* __restrict__ src, int count) {
    const double2* s2 = reinterpret_cast<const double2*>(src);
    double2* d2 = reinterpret_cast<double2*>(sm);
    for (int k = threadIdx.x; k < (count >> 1); k += blockDim.x) d2[k] = __ldg(s2 + k);
}

// ---------------------------------------------------------------------------------------------
// rho_q build, generic q.  One CTA per (config, slice) (grid-stride); work item = (q, particle chunk p).
// All lanes of a warp normally share the chunk, so coordinate reads are shared-memory broadcasts and
// every lane runs an independent dot + sincos + 2 accumulates per particle with no cross-lane traffic.
// rho layout: rho[slice_global][0][q] = sum cos, rho[slice_global][1][q] = sum sin.
// ---------------------------------------------------------------------------------------------
template <int ND>
__global__ void __launch_bounds__(256) rho_generic_kernel(const double* __restrict__ pos, const double* __restrict__ qsoa,
                                                           double* __restrict__ rho, int nslices, int N, int Npad, int nq,
                                                           int P, int chunk) {
    extern __shared__ __align__(16) double sm[];
    double* xs = sm;                          // [ND][Npad]
    double* part = sm + ND * Npad;            // [2][P][nq]  (only when P > 1)
    const int items = nq * P;
    for (int sl = blockIdx.x; sl < nslices; sl += gridDim.x) {
        load_slice(xs, pos + static_cast<size_t>(sl) * ND * Npad, ND * Npad);
        __syncthreads();
        for (int item = threadIdx.x; item < items; item += blockDim.x) {
            const int p = item / nq;
            const int iq = item - p * nq;
            double qv[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) qv[d] = __ldg(qsoa + d * nq + iq);
            const int i0 = p * chunk;
            const int i1 = min(N, i0 + chunk);
            double ac = 0.0, as = 0.0;
#pragma unroll 4
            for (int i = i0; i < i1; ++i) {
                double ph = qv[0] * xs[i];
#pragma unroll
                for (int d = 1; d < ND; ++d) ph = fma(qv[d], xs[d * Npad + i], ph);
                double s, c;
                sincos_fast(ph, s, c);
                ac += c;
                as += s;
            }
            if (P == 1) {
                rho[(static_cast<size_t>(sl) * 2 + 0) * nq + iq] = ac;
                rho[(static_cast<size_t>(sl) * 2 + 1) * nq + iq] = as;
            } else {
                part[p * nq + iq] = ac;
                part[(P + p) * nq + iq] = as;
            }
        }
        if (P > 1) {
            __syncthreads();
            for (int k = threadIdx.x; k < 2 * nq; k += blockDim.x) {
                const int cs = k / nq, iq = k - cs * nq;
                double acc = 0.0;
                for (int p = 0; p < P; ++p) acc += part[(cs * P + p) * nq + iq];   // fixed order: deterministic
                rho[(static_cast<size_t>(sl) * 2 + cs) * nq + iq] = acc;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// rho_q build for commensurate q = 2 pi n / L ("lattice" path).
//   exp(i q.r) = prod_d e_d^{n_d},  e_d = exp(2 pi i x_d / L_d).
// Phase 1 tabulates the powers e_d^m, m = 0..nmax_d, of every particle of the slice in shared memory (ND sincos
// + nmax complex multiplies per particle instead of nq sincos).  Phase 2 works on sign-symmetry GROUPS of
// wave-vectors: all q that share (|n_0|,..,|n_{ND-1}|) -- up to 2^ND of them -- are produced from ONE pass over
// the particles, because flipping the sign of a component only conjugates that factor:
//   3-D:  P+- = X Y^(+-),  K[sb][0..3] = sum_i (P_re Z_re, P_im Z_im, P_re Z_im, P_im Z_re)       (16 FP64 / particle)
//         rho(+a, sb b, sc c) = (K0 - sc K1) + i (sc K2 + K3),  rho(-a,..) = conj rho(+a, -sb b, -sc c)
//   2-D:  K[0..3] = sum_i (X_re Y_re, X_im Y_im, X_re Y_im, X_im Y_re);   1-D:  K[0..1] = sum_i X.
// Work item = (group, particle chunk).  The kernel is shared-memory-bandwidth bound (ND 16-byte table reads per
// (group, particle)), not FP64 bound.
// gkey: int[G][ND] = |n_d|;  gout: int[G][2^ND] = q index for sign pattern (bit d set = component d negative) or -1.
// Power table layout: tab[row][i] as double2 (re, im), row = rowoff_d + m, row stride N + 1 double2 so that
// lanes reading different rows / chunks hit different banks.
// ---------------------------------------------------------------------------------------------
template <int ND>
struct LatticeK { static constexpr int NK = ND == 3 ? 8 : (ND == 2 ? 4 : 2); static constexpr int NPAT = 1 << ND; };

template <int ND>
__global__ void __launch_bounds__(256) rho_lattice_kernel(const double* __restrict__ pos, const int* __restrict__ gkey,
                                                           const int* __restrict__ gout, double* __restrict__ rho, int nslices,
                                                           int N, int Npad, int nq, int G, int P, int chunk, int3 nmax,
                                                           double3 kphase) {
    constexpr int NK = LatticeK<ND>::NK;
    constexpr int NPAT = LatticeK<ND>::NPAT;
    extern __shared__ __align__(16) double sm[];
    const int rowoff1 = nmax.x + 1;
    const int rowoff2 = rowoff1 + (ND > 1 ? nmax.y + 1 : 0);
    const int rows = rowoff2 + (ND > 2 ? nmax.z + 1 : 0);
    const int stride = N + 1;                                  // in double2 units
    double2* tab = reinterpret_cast<double2*>(sm);             // [rows][stride]
    double* xs = sm + 2 * static_cast<size_t>(rows) * stride;  // [ND][Npad] raw coordinates
    double* part = xs + ND * Npad;                             // [NK][P*G]  (k-major: conflict-free stores and loads)
    const int items = G * P;
    for (int sl = blockIdx.x; sl < nslices; sl += gridDim.x) {
        load_slice(xs, pos + static_cast<size_t>(sl) * ND * Npad, ND * Npad);
        __syncthreads();
        // phase 1: powers of the base phases
        for (int w = threadIdx.x; w < ND * N; w += blockDim.x) {
            const int d = w / N, i = w - d * N;
            const double kp = d == 0 ? kphase.x : (d == 1 ? kphase.y : kphase.z);
            const int nm = d == 0 ? nmax.x : (d == 1 ? nmax.y : nmax.z);
            const int ro = d == 0 ? 0 : (d == 1 ? rowoff1 : rowoff2);
            double s, c;
            sincos_fast(kp * xs[d * Npad + i], s, c);
            double2* col = tab + static_cast<size_t>(ro) * stride + i;
            col[0] = make_double2(1.0, 0.0);
            double pr = c, pi = s;
            for (int m = 1; m <= nm; ++m) {
                col[static_cast<size_t>(m) * stride] = make_double2(pr, pi);
                const double nr = fma(pr, c, -pi * s);
                pi = fma(pr, s, pi * c);
                pr = nr;
            }
        }
        __syncthreads();
        // phase 2: one (group, particle chunk) per work item
        for (int item = threadIdx.x; item < items; item += blockDim.x) {
            const int p = item / G;
            const int g = item - p * G;
            const int i0 = p * chunk;
            const int i1 = min(N, i0 + chunk);
            const double2* X = tab + static_cast<size_t>(__ldg(gkey + g * ND)) * stride;
            double K[NK];
#pragma unroll
            for (int k = 0; k < NK; ++k) K[k] = 0.0;
            if constexpr (ND == 1) {
#pragma unroll 4
                for (int i = i0; i < i1; ++i) {
                    const double2 x = X[i];
                    K[0] += x.x;
                    K[1] += x.y;
                }
            } else if constexpr (ND == 2) {
                const double2* Y = tab + static_cast<size_t>(rowoff1 + __ldg(gkey + g * ND + 1)) * stride;
#pragma unroll 4
                for (int i = i0; i < i1; ++i) {
                    const double2 x = X[i], y = Y[i];
                    K[0] = fma(x.x, y.x, K[0]);
                    K[1] = fma(x.y, y.y, K[1]);
                    K[2] = fma(x.x, y.y, K[2]);
                    K[3] = fma(x.y, y.x, K[3]);
                }
            } else {
                const double2* Y = tab + static_cast<size_t>(rowoff1 + __ldg(gkey + g * ND + 1)) * stride;
                const double2* Z = tab + static_cast<size_t>(rowoff2 + __ldg(gkey + g * ND + 2)) * stride;
#pragma unroll 2
                for (int i = i0; i < i1; ++i) {
                    const double2 x = X[i], y = Y[i], z = Z[i];
                    const double m1 = x.x * y.x, m2 = x.y * y.y, m3 = x.x * y.y, m4 = x.y * y.x;
                    const double ppr = m1 - m2, ppi = m3 + m4;     // X * Y
                    const double pmr = m1 + m2, pmi = m4 - m3;     // X * conj(Y)
                    K[0] = fma(ppr, z.x, K[0]);
                    K[1] = fma(ppi, z.y, K[1]);
                    K[2] = fma(ppr, z.y, K[2]);
                    K[3] = fma(ppi, z.x, K[3]);
                    K[4] = fma(pmr, z.x, K[4]);
                    K[5] = fma(pmi, z.y, K[5]);
                    K[6] = fma(pmr, z.y, K[6]);
                    K[7] = fma(pmi, z.x, K[7]);
                }
            }
#pragma unroll
            for (int k = 0; k < NK; ++k) part[k * items + item] = K[k];
        }
        __syncthreads();
        // phase 3: fold the chunks (fixed order) and unfold the sign patterns into rho
        for (int w = threadIdx.x; w < G * NPAT; w += blockDim.x) {
            const int g = w / NPAT, pat = w - g * NPAT;
            const int iq = __ldg(gout + w);
            if (iq < 0) continue;
            const int sa = pat & 1;                                // conj of the pattern with all signs flipped
            const int sb = ND > 1 ? (((pat >> 1) & 1) ^ sa) : 0;
            const int sc = ND > 2 ? (((pat >> 2) & 1) ^ sa) : 0;
            double k0 = 0.0, k1 = 0.0, k2 = 0.0, k3 = 0.0;
            const int base = ND == 3 ? 4 * sb : 0;
            for (int p = 0; p < P; ++p) {
                const double* src = part + base * items + p * G + g;
                k0 += src[0];
                k1 += src[items];
                if (ND > 1) { k2 += src[2 * items]; k3 += src[3 * items]; }
            }
            double re, im;
            if (ND == 1) { re = k0; im = k1; }
            else {
                const int sl_ = ND == 3 ? sc : sb;                 // sign of the last multiplied factor
                re = sl_ ? k0 + k1 : k0 - k1;
                im = sl_ ? k3 - k2 : k3 + k2;
            }
            rho[(static_cast<size_t>(sl) * 2 + 0) * nq + iq] = re;
            rho[(static_cast<size_t>(sl) * 2 + 1) * nq + iq] = sa ? -im : im;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// tau-correlation:  F(q,tau) = (1/N) sum_t0 [C(t0) C(t0+tau) + S(t0) S(t0+tau)],  tau = 0..M/2, mirrored to M-tau
// (F(q,tau) = F(q,M-tau) identically), S(q) = F(q,0) for commensurate q.
// Register-tiled so that the FP64 pipe, not shared memory, is the limit: `lpq` lanes (8/16/32) share one (config,q)
// pair; a lane owns 8 consecutive tau and sweeps t0 in blocks of 8 -- per block 8 broadcast values a(t0..t0+7) and a
// 16-value window w(t0+tau0 .. +15) are pulled with 16-byte loads and feed 8x8x2 = 128 DFMAs.
// C and S of a pair are staged periodically extended (index i -> value at i mod M, length 2M+16) with 2 doubles of
// padding after every 8 so that the lanes' 64-byte-strided windows fall in different banks.
// Output per config:  cfg[b][q] = F(q,0)  (sf/N, commensurate q);  cfg[b][nq + q*M + tau] = F(q,tau)  (isf/N).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int corr_idx(int i) { return i + 2 * (i >> 3); }

__global__ void __launch_bounds__(128) isf_corr_kernel(const double* __restrict__ rho, double* __restrict__ cfg, int M, int nq,
                                                        int npairs, int lpq, double invN,
                                                        const unsigned char* __restrict__ commensurate) {
    extern __shared__ __align__(16) double sm[];
    const int len = 2 * M + 16;
    const int plen = corr_idx(len) + 2;                      // padded length of one array (even -> 16-byte aligned)
    const int ppw = 32 / lpq;                                // pairs per warp
    const int ppc = ppw * (blockDim.x >> 5);                 // pairs per CTA
    const int pair0 = blockIdx.x * ppc;
    // stage: consecutive threads take consecutive q of one slice (contiguous in rho), each value is written to
    // every periodic image i = t, t+M, t+2M < len
    for (int w = threadIdx.x; w < ppc * M; w += blockDim.x) {
        const int t = w / ppc, lp = w - t * ppc;
        const int pair = pair0 + lp;
        if (pair >= npairs) continue;
        const int b = pair / nq, iq = pair - b * nq;
        const size_t sl = static_cast<size_t>(b) * M + t;
        const double c = rho[(sl * 2 + 0) * nq + iq];
        const double sn = rho[(sl * 2 + 1) * nq + iq];
        for (int i = t; i < len; i += M) {
            sm[(2 * lp + 0) * plen + corr_idx(i)] = c;
            sm[(2 * lp + 1) * plen + corr_idx(i)] = sn;
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lp = warp * ppw + lane / lpq;
    const int pair = pair0 + lp;
    if (pair >= npairs) return;
    const int b = pair / nq, iq = pair - b * nq;
    const double* WC = sm + (2 * lp + 0) * plen;
    const double* WS = sm + (2 * lp + 1) * plen;
    const size_t cfg_stride = static_cast<size_t>(nq) + static_cast<size_t>(nq) * M;
    double* out = cfg + static_cast<size_t>(b) * cfg_stride + nq + static_cast<size_t>(iq) * M;
    const int half = M / 2;
    const int nblk = (half + 1 + 7) / 8;                     // tau blocks of 8
    for (int tb = lane % lpq; tb < nblk; tb += lpq) {
        const int tau0 = 8 * tb;
        double acc[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[v] = 0.0;
        for (int t0 = 0; t0 < M; t0 += 8) {
            double aC[8], aS[8], wC[16], wS[16];
            {
                const double2* pa = reinterpret_cast<const double2*>(WC + corr_idx(t0));
                const double2* pb = reinterpret_cast<const double2*>(WS + corr_idx(t0));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double2 x = pa[k], y = pb[k];
                    aC[2 * k] = x.x; aC[2 * k + 1] = x.y;
                    aS[2 * k] = y.x; aS[2 * k + 1] = y.y;
                }
            }
            if (t0 + 8 > M) {                                // last block of a non-multiple-of-8 M: drop t0 >= M
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (t0 + u >= M) { aC[u] = 0.0; aS[u] = 0.0; }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double2* pc = reinterpret_cast<const double2*>(WC + corr_idx(t0 + tau0 + 8 * h));
                const double2* ps = reinterpret_cast<const double2*>(WS + corr_idx(t0 + tau0 + 8 * h));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double2 x = pc[k], y = ps[k];
                    wC[8 * h + 2 * k] = x.x; wC[8 * h + 2 * k + 1] = x.y;
                    wS[8 * h + 2 * k] = y.x; wS[8 * h + 2 * k + 1] = y.y;
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    acc[v] = fma(aC[u], wC[u + v], acc[v]);
                    acc[v] = fma(aS[u], wS[u + v], acc[v]);
                }
            }
        }
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            const int tau = tau0 + v;
            if (tau > half) continue;
            const double val = acc[v] * invN;
            out[tau] = val;
            if (tau > 0 && tau < M - tau) out[M - tau] = val;
            if (tau == 0 && commensurate[iq]) cfg[static_cast<size_t>(b) * cfg_stride + iq] = val;
        }
    }
    // odd M never occurs upstream (setup.cpp:1001-1008 forces M even); for odd M, tau = (M+1)/2.. mirror as well.
}

// ---------------------------------------------------------------------------------------------
// Direct minimum-image S(q) for the non-commensurate q (reference CPU semantics,
// src/estimator.cpp:3715-3734 + include/path.h:179-184).  One CTA per (config, slice); pairs (i, i+k mod N)
// k = 1..N/2 are spread over threads; the min-image separation is computed once per pair and reused for a
// register tile of QT wave-vectors.  partial[sl][k] = sum_{i<j} cos(q_k . sep_ij).
// ---------------------------------------------------------------------------------------------
template <int ND, int QT>
__global__ void __launch_bounds__(256) ssf_direct_kernel(const double* __restrict__ pos, const double* __restrict__ qsoa,
                                                          const int* __restrict__ qidx, int nsel, int nq, double* __restrict__ partial,
                                                          int nslices, int N, int Npad, BoxDev box) {
    extern __shared__ __align__(16) double sm[];
    double* xs = sm;                       // [ND][Npad]
    __shared__ double red[QT][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int sl = blockIdx.x; sl < nslices; sl += gridDim.x) {
        load_slice(xs, pos + static_cast<size_t>(sl) * ND * Npad, ND * Npad);
        __syncthreads();
        for (int q0 = 0; q0 < nsel; q0 += QT) {
            double qv[QT][ND];
#pragma unroll
            for (int k = 0; k < QT; ++k) {
                const int iq = (q0 + k < nsel) ? __ldg(qidx + q0 + k) : -1;
#pragma unroll
                for (int d = 0; d < ND; ++d) qv[k][d] = iq >= 0 ? __ldg(qsoa + d * nq + iq) : 0.0;
            }
            double acc[QT];
#pragma unroll
            for (int k = 0; k < QT; ++k) acc[k] = 0.0;
            const int kmax = N / 2;
            for (int kk = 1; kk <= kmax; ++kk) {
                const int ilim = ((N & 1) == 0 && kk == kmax) ? N / 2 : N;
                for (int i = threadIdx.x; i < ilim; i += blockDim.x) {
                    int j = i + kk;
                    if (j >= N) j -= N;
                    double sep[ND];
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        const double s = xs[d * Npad + i] - xs[d * Npad + j];
                        sep[d] = s - box.pSide[d] * floor(fma(s, box.sideInv[d], 0.5));
                    }
#pragma unroll
                    for (int k = 0; k < QT; ++k) {
                        double ph = qv[k][0] * sep[0];
#pragma unroll
                        for (int d = 1; d < ND; ++d) ph = fma(qv[k][d], sep[d], ph);
                        double s, c;
                        sincos_fast(ph, s, c);
                        acc[k] += c;
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < QT; ++k) {
                double v = acc[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) red[k][warp] = v;
            }
            __syncthreads();
            if (threadIdx.x < QT && q0 + threadIdx.x < nsel) {
                double v = 0.0;
                for (int w = 0; w < (blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
                partial[static_cast<size_t>(sl) * nsel + q0 + threadIdx.x] = v;
            }
            __syncthreads();
        }
    }
}

// cfg[b][qidx[k]] = (M*N + 2 * sum_t partial[b*M+t][k]) / N     (sf/N, src/estimator.cpp:3726-3736)
__global__ void ssf_direct_finalize_kernel(const double* __restrict__ partial, const int* __restrict__ qidx, int nsel,
                                           double* __restrict__ cfg, int B, int M, int N, int nq) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * nsel) return;
    const int b = idx / nsel, k = idx - b * nsel;
    double acc = 0.0;
    for (int t = 0; t < M; ++t) acc += partial[(static_cast<size_t>(b) * M + t) * nsel + k];
    const size_t cfg_stride = static_cast<size_t>(nq) + static_cast<size_t>(nq) * M;
    cfg[b * cfg_stride + qidx[k]] = (static_cast<double>(M) * N + 2.0 * acc) / N;
}

// bins[j] += sum_b cfg[b][j], b ascending (deterministic).
__global__ void bins_accumulate_kernel(const double* __restrict__ cfg, double* __restrict__ bins, int B, size_t len) {
    const size_t j = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= len) return;
    double acc = bins[j];
    for (int b = 0; b < B; ++b) acc += cfg[static_cast<size_t>(b) * len + j];
    bins[j] = acc;
}

// ---------------------------------------------------------------------------------------------
// Pair potential.  One CTA per (config, slice).  Thread i walks partners j = i+k mod N.  The table index
// k = int(|sep|/dr) follows the reference's operation order with individually rounded IEEE operations
// (__d*_rn intrinsics are never contracted into FMAs) so that it is bit-identical to the CPU:
//   sep_d = r_a,d - r_b,d ; sep_d -= pSide_d*floor(sep_d*sideInv_d + 0.5)   (include/container.h:50-53)
//   r = sqrt(((0 + s0*s0) + s1*s1) + s2*s2) ; k = int(r/dr)                   (include/potential.h:249-260, 985-1003)
// WANT_F2: full j != i loop accumulating F_i = sum_j (dVdr[k]/r) sep_ij, then sum_i |F_i|^2
//          (src/action.cpp:1188-1223); V is accumulated on the k <= N/2 half so each pair counts once.
// Without WANT_F2 only the N(N-1)/2 half is visited.
// ---------------------------------------------------------------------------------------------
template <int ND>
__device__ __forceinline__ double minimage_norm(const double* __restrict__ xs, int Npad, int a, int b, const BoxDev& box, double* sep) {
    double r2 = 0.0;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        double s = __dsub_rn(xs[d * Npad + a], xs[d * Npad + b]);
        const double f = floor(__dadd_rn(__dmul_rn(s, box.sideInv[d]), 0.5));
        s = __dsub_rn(s, __dmul_rn(box.pSide[d], f));
        sep[d] = s;
        r2 = __dadd_rn(r2, __dmul_rn(s, s));
    }
    return __dsqrt_rn(r2);
}

__device__ __forceinline__ double table_direct(const double* __restrict__ tab, int len, double dr, double ext0, double ext1, double r) {
    const int k = __double2int_rz(__ddiv_rn(r, dr));
    if (k <= 0) return ext0;
    if (k >= len) return ext1;
    return __ldg(tab + k);
}

struct PairParams {
    const double* V; const double* dVdr; int len; double dr; double extV[2]; double extdV[2];
    double dSep; int want_hist; int f2_parity; int M;
};

template <int ND, bool WANT_F2>
__global__ void __launch_bounds__(256) pair_kernel(const double* __restrict__ pos, int nslices, int N, int Npad, BoxDev box,
                                                    PairParams pp, double* __restrict__ vint, double* __restrict__ f2,
                                                    int* __restrict__ hist) {
    extern __shared__ __align__(16) double sm[];
    double* xs = sm;
    __shared__ double redV[8], redF[8];
    __shared__ int shist[kNPCFSEP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int sl = blockIdx.x; sl < nslices; sl += gridDim.x) {
        const int t = sl % pp.M;
        const bool do_f2 = WANT_F2 && (pp.f2_parity < 0 || (t & 1) == pp.f2_parity);
        load_slice(xs, pos + static_cast<size_t>(sl) * ND * Npad, ND * Npad);
        if (threadIdx.x < kNPCFSEP) shist[threadIdx.x] = 0;
        __syncthreads();
        double vsum = 0.0, fsum = 0.0;
        const int khalf = N / 2;
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            double F[ND];
#pragma unroll
            for (int d = 0; d < ND; ++d) F[d] = 0.0;
            const int klast = do_f2 ? N - 1 : khalf;
            for (int kk = 1; kk <= klast; ++kk) {
                int j = i + kk;
                if (j >= N) j -= N;
                // V-half: pairs (i, i+kk), kk <= N/2, (for even N and kk == N/2 only i < N/2)
                const bool vhalf = (kk < khalf) || (kk == khalf && ((N & 1) || i < khalf));
                if (!do_f2 && !vhalf) continue;
                double sep[ND];
                double r;
                if (do_f2) {
                    r = minimage_norm<ND>(xs, Npad, i, j, box, sep);            // getSeparation(bead1,bead2), action.cpp:1211
                } else {
                    const int lo = min(i, j), hi = max(i, j);
                    r = minimage_norm<ND>(xs, Npad, hi, lo, box, sep);          // getSeparation(bead2,bead1), action.cpp:934
                }
                if (vhalf) {
                    vsum += table_direct(pp.V, pp.len, pp.dr, pp.extV[0], pp.extV[1], r);
                    if (pp.want_hist) {
                        const int nR = __double2int_rz(__ddiv_rn(r, pp.dSep));  // action.cpp:221
                        if (nR >= 0 && nR < kNPCFSEP) atomicAdd(&shist[nR], 1);
                    }
                }
                if (do_f2) {
                    const double g = __ddiv_rn(table_direct(pp.dVdr, pp.len, pp.dr, pp.extdV[0], pp.extdV[1], r), r);
#pragma unroll
                    for (int d = 0; d < ND; ++d) F[d] = fma(g, sep[d], F[d]);
                }
            }
            if (do_f2) {
#pragma unroll
                for (int d = 0; d < ND; ++d) fsum = fma(F[d], F[d], fsum);
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
            if (WANT_F2 && f2) f2[sl] = f;
        }
        if (pp.want_hist && threadIdx.x < kNPCFSEP) hist[static_cast<size_t>(sl) * kNPCFSEP + threadIdx.x] = shist[threadIdx.x];
        __syncthreads();
    }
}

// AoS double[nslices][Next][ND] (the reference's beads array, DMA'd as-is) -> pos[sl][d][Npad].
template <int ND>
__global__ void aos_to_soa_kernel(const double* __restrict__ aos, double* __restrict__ pos, int nslices, int N, int Next, int Npad) {
    extern __shared__ __align__(16) double sm[];
    for (int sl = blockIdx.x; sl < nslices; sl += gridDim.x) {
        const double* src = aos + static_cast<size_t>(sl) * Next * ND;
        for (int k = threadIdx.x; k < N * ND; k += blockDim.x) sm[k] = src[k];
        __syncthreads();
        double* dst = pos + static_cast<size_t>(sl) * ND * Npad;
        for (int k = threadIdx.x; k < ND * Npad; k += blockDim.x) {
            const int d = k / Npad, i = k - d * Npad;
            dst[k] = i < N ? sm[i * ND + d] : 0.0;
        }
        __syncthreads();
    }
}

// Register-resident DFMA chains: 8 independent accumulators per thread, `iters` x 8 x 4 FMAs.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double v = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (v == 123.456) out[0] = v;   // never true; keeps the chains live
}

}  // namespace pimcb
