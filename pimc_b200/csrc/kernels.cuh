// kernels.cuh -- sm_100a device code for the pimc measurement hot path (FP64: CUDA cores and DMMA tensor instructions).
//
// Device bead layout ("slice-major SoA"):  pos[b][t][d][Npad]  (double), Npad = N rounded up to 16,
// so that the ND coordinate rows of one time slice are one contiguous, 128-byte aligned chunk that a
// CTA pulls into shared memory with coalesced 16-byte loads.
//
// Kernels
//   rho_generic_kernel   rho_q(t) = sum_i exp(i q.r_i(t)), one sincos per (q, bead)      [FP64-pipe bound]
//   rho_lattice_kernel   same for commensurate q = 2 pi n / L: phase-power tables + sign-symmetry groups,
//                        lane = particle, warp = column of groups                       [FP64 / smem balanced]
//   rho_lattice_mma_kernel  the same particle sums as a batched small GEMM on the FP64 tensor cores (DMMA)
//   isf_corr_mma_kernel  F(q,tau) = (1/N) sum_t0 Re[rho(t0) conj rho(t0+tau)], S(q) = F(q,0), as DMMA GEMMs (M <= 510)
//   isf_corr_kernel      the same on the CUDA cores (register tiled; any M)
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

__device__ __forceinline__ void sincos_fast(double x, double& s, double& c, int& n) {
    const double t = fma(x, kSC[0], kSC[1]);
    n = __double2loint(t);                               // quadrant = low bits of rint(x*2/pi)
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
__device__ __forceinline__ void sincos_fast(double x, double& s, double& c) {
    int n;
    sincos_fast(x, s, c, n);
}

// rho_q layout ("blocks of four slices"):  rho[b][t / 4][q][{cos,sin}][t % 4]  (double).
//   * a rho kernel finishing slice t writes one 8-byte element into each of 2 nq CONSECUTIVE 32-byte sectors (the
//     warp's 32 lanes touch 8 cache lines per store; a fully pair-major layout -- M contiguous values per (b, q) --
//     touched 32 lines per store and cost the rho kernel 6 %);
//   * the tau-correlation reads, per (b, q) pair, whole 32-byte sectors (four consecutive slices of its pair), so no
//     fetched byte is wasted and four configurations of the SAME q can share a CTA (quad-summed partial rows).
__host__ __device__ inline int rho_tblocks(int M) { return (M + 3) >> 2; }
__device__ __forceinline__ size_t rho_pair_base(int b, int iq, int cs, int nq, int nTB) {   // + rho_slice_off(t)
    return (static_cast<size_t>(b) * nTB * nq * 2 + static_cast<size_t>(iq) * 2 + cs) * 4;
}
__device__ __forceinline__ size_t rho_slice_off(int t, int nq) { return static_cast<size_t>(t >> 2) * nq * 8 + (t & 3); }

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
                                                           int P, int chunk, int M) {
    extern __shared__ __align__(16) double sm[];
    double* xs = sm;                          // [ND][Npad]
    double* part = sm + ND * Npad;            // [2][P][nq]  (only when P > 1)
    const int items = nq * P;
    const int nTB = rho_tblocks(M);
    for (int sl = blockIdx.x; sl < nslices; sl += gridDim.x) {
        const int cb = sl / M, ct = sl - cb * M;      // configuration and time slice of this CTA's work item
        const size_t so = rho_slice_off(ct, nq);
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
                rho[rho_pair_base(cb, iq, 0, nq, nTB) + so] = ac;
                rho[rho_pair_base(cb, iq, 1, nq, nTB) + so] = as;
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
                rho[rho_pair_base(cb, iq, cs, nq, nTB) + so] = acc;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// rho_q build for commensurate q = 2 pi n / L ("lattice" path).
//   exp(i q.r) = prod_d e_d^{n_d},  e_d = exp(2 pi i x_d / L_d).
// Phase 1 tabulates the powers e_d^m, m = 0..nmax_d, of every particle of the slice in shared memory (ND sincos
// + nmax complex multiplies per particle instead of nq sincos).
// Phase 2 works on sign-symmetry GROUPS of wave-vectors: all q that share (|n_0|,..,|n_{ND-1}|) -- up to 2^ND of
// them -- come out of ONE pass over the particles, because flipping the sign of a component only conjugates that
// factor.  In 3-D, with X = e_x^|a|, Y = e_y^|b|, Z = e_z^|c|:
//      P+- = X Y^(+-),   K[sb][0..3] = sum_i (P_re Z_re, P_im Z_im, P_re Z_im, P_im Z_re)      (8 DFMA / particle)
//      rho(+a, sb b, sc c) = (K0 - sc K1) + i (sc K2 + K3),   rho(-a, ..) = conj rho(+a, -sb b, -sc c).
// Mapping: LANE = PARTICLE, WARP = TASK.  A task is a column (|a|,|b|) with a run of |c| entries: the warp computes
// P+- of its particles once (registers, J particles per lane) and reuses them for every entry of the run, so each
// (group, particle) costs ONE 16-byte shared-memory read (Z) instead of three -- the v2 kernel (thread = group) was
// shared-memory-bandwidth bound at 48 B per (group, particle).  The sum over particles is a halving butterfly
// (v[8] -> 1 value per lane in 9 exchanged doubles) followed by a single-lane accumulate into part[k][group].
// Tasks are distributed over the 8 warps of the CTA by a host-side LPT schedule (plan.warp_first / plan.tasks).
// Phase 3 unfolds the sign patterns:  gout[g][pattern] = q index (bit d of pattern set = component d negative) or -1.
// 2-D: column = |a|, entries = |b|, K[0..3] = sum (X_re Y_re, X_im Y_im, X_re Y_im, X_im Y_re);  1-D: K = sum X^|a|.
// Power table layout: tab[row][i] as double2 (re, im), row = rowoff_d + m, row stride N + 1 double2.
// ---------------------------------------------------------------------------------------------
template <int ND>
struct LatticeK { static constexpr int NK = ND == 3 ? 8 : (ND == 2 ? 4 : 2); static constexpr int NPAT = 1 << ND; };

struct LatticePlan {
    const int* gout;        // [G][2^ND]
    const int* ent;         // [G]     |n_last| of group g (groups are stored column by column)
    const int* tasks;       // [ntask][4] = {|a|, |b|, first group, one-past-last group}
    const int* warp_first;  // [9]     tasks of warp w: warp_first[w] .. warp_first[w+1]
    int G;
};

// Sum v[0..NK) over the 32 lanes.  On return lane L holds the total of component
// idx = bits (4,3,2) of L for NK = 8, bits (4,3) for NK = 4, bit 4 for NK = 2; lanes sharing those bits agree.
template <int NK>
__device__ __forceinline__ double butterfly_sum(double (&v)[NK], int lane, int& idx) {
    constexpr unsigned FULL = 0xffffffffu;
    if constexpr (NK == 8) {
        const bool u4 = lane & 16, u3 = lane & 8, u2 = lane & 4;
        double a[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double send = u4 ? v[k] : v[k + 4];
            const double keep = u4 ? v[k + 4] : v[k];
            a[k] = keep + __shfl_xor_sync(FULL, send, 16);
        }
        double b[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const double send = u3 ? a[k] : a[k + 2];
            const double keep = u3 ? a[k + 2] : a[k];
            b[k] = keep + __shfl_xor_sync(FULL, send, 8);
        }
        double w = (u2 ? b[1] : b[0]) + __shfl_xor_sync(FULL, u2 ? b[0] : b[1], 4);
        w += __shfl_xor_sync(FULL, w, 2);
        w += __shfl_xor_sync(FULL, w, 1);
        idx = (u4 ? 4 : 0) + (u3 ? 2 : 0) + (u2 ? 1 : 0);
        return w;
    } else if constexpr (NK == 4) {
        const bool u4 = lane & 16, u3 = lane & 8;
        double a[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const double send = u4 ? v[k] : v[k + 2];
            const double keep = u4 ? v[k + 2] : v[k];
            a[k] = keep + __shfl_xor_sync(FULL, send, 16);
        }
        double w = (u3 ? a[1] : a[0]) + __shfl_xor_sync(FULL, u3 ? a[0] : a[1], 8);
        w += __shfl_xor_sync(FULL, w, 4);
        w += __shfl_xor_sync(FULL, w, 2);
        w += __shfl_xor_sync(FULL, w, 1);
        idx = (u4 ? 2 : 0) + (u3 ? 1 : 0);
        return w;
    } else {
        const bool u4 = lane & 16;
        double w = (u4 ? v[1] : v[0]) + __shfl_xor_sync(FULL, u4 ? v[0] : v[1], 16);
        w += __shfl_xor_sync(FULL, w, 8);
        w += __shfl_xor_sync(FULL, w, 4);
        w += __shfl_xor_sync(FULL, w, 2);
        w += __shfl_xor_sync(FULL, w, 1);
        idx = u4 ? 1 : 0;
        return w;
    }
}

// Row stride (double2 units) of the phase-power table: N rounded up to whole particle blocks of 32*J, plus one
// element so that consecutive rows start in different banks.
__host__ __device__ inline int lattice_stride(int N, int J) { return (N + 32 * J - 1) / (32 * J) * (32 * J) + 1; }

#ifndef PIMCB_LATTICE_MINB
#define PIMCB_LATTICE_MINB 4
#endif
constexpr int kLatticeWarps = 4;   // warps (= concurrent tasks) per CTA of the lattice kernel

template <int ND, int J>
__global__ void __launch_bounds__(32 * kLatticeWarps, PIMCB_LATTICE_MINB) rho_lattice_kernel(const double* __restrict__ pos, LatticePlan plan,
                                                           double* __restrict__ rho, int nslices, int N, int Npad, int nq,
                                                           int3 nmax, double3 kphase, int M) {
    constexpr int NK = LatticeK<ND>::NK;
    const int nTB = rho_tblocks(M);
    constexpr int NPAT = LatticeK<ND>::NPAT;
    extern __shared__ __align__(16) double sm[];
    const int rowoff1 = nmax.x + 1;
    const int rowoff2 = rowoff1 + (ND > 1 ? nmax.y + 1 : 0);
    const int rows = rowoff2 + (ND > 2 ? nmax.z + 1 : 0);
    const int stride = lattice_stride(N, J);                   // in double2 units; rows are zero beyond N
    const int G = plan.G;
    double2* tab = reinterpret_cast<double2*>(sm);             // [rows][stride]
    double* xs = sm + 2 * static_cast<size_t>(rows) * stride;  // [ND][Npad] raw coordinates
    double* part = xs + ND * Npad;                             // [NK][G]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lastrow = ND == 3 ? rowoff2 : (ND == 2 ? rowoff1 : 0);   // row offset of the entry dimension
    const int nblk = (N + 32 * J - 1) / (32 * J);              // particle blocks of 32*J
    for (int sl = blockIdx.x; sl < nslices; sl += gridDim.x) {
        load_slice(xs, pos + static_cast<size_t>(sl) * ND * Npad, ND * Npad);
        __syncthreads();
        // phase 1: powers of the base phases; entries i >= N of every row are zero so that padded lanes add nothing
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const double kp = d == 0 ? kphase.x : (d == 1 ? kphase.y : kphase.z);
            const int nm = d == 0 ? nmax.x : (d == 1 ? nmax.y : nmax.z);
            double2* row0 = tab + static_cast<size_t>(d == 0 ? 0 : (d == 1 ? rowoff1 : rowoff2)) * stride;
            for (int i = threadIdx.x; i < nblk * 32 * J; i += blockDim.x) {
                double2* col = row0 + i;
                if (i < N) {
                    double s, c;
                    sincos_fast(kp * xs[d * Npad + i], s, c);
                    col[0] = make_double2(1.0, 0.0);
                    double pr = c, pi = s;
                    for (int m = 1; m <= nm; ++m) {
                        col[m * stride] = make_double2(pr, pi);
                        const double nr = fma(pr, c, -pi * s);
                        pi = fma(pr, s, pi * c);
                        pr = nr;
                    }
                } else {
                    for (int m = 0; m <= nm; ++m) col[m * stride] = make_double2(0.0, 0.0);
                }
            }
        }
        __syncthreads();
        // phase 2: warp = task, lane = particle
        for (int ti = __ldg(plan.warp_first + warp); ti < __ldg(plan.warp_first + warp + 1); ++ti) {
            const int4 tk = __ldg(reinterpret_cast<const int4*>(plan.tasks) + ti);
            for (int ib = 0; ib < nblk; ++ib) {
                const int ibase = ib * 32 * J + lane;
                double pa[J], pb[J], pc[J], pd[J];             // 3-D: P+ (re,im), P- (re,im); 2-D: X (re,im)
                if constexpr (ND > 1) {
                    const double2* X = tab + tk.x * stride + ibase;
                    const double2* Y = tab + (rowoff1 + tk.y) * stride + ibase;
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        const double2 x = X[32 * j];           // zero beyond N
                        if constexpr (ND == 3) {
                            const double2 y = Y[32 * j];
                            const double m1 = x.x * y.x, m2 = x.y * y.y, m3 = x.x * y.y, m4 = x.y * y.x;
                            pa[j] = m1 - m2; pb[j] = m3 + m4;  // X * Y
                            pc[j] = m1 + m2; pd[j] = m4 - m3;  // X * conj(Y)
                        } else {
                            pa[j] = x.x; pb[j] = x.y;
                        }
                    }
                }
                for (int g = tk.z; g < tk.w; ++g) {
                    const int c = __ldg(plan.ent + g);
                    const double2* Z = tab + (lastrow + c) * stride + ibase;
                    double K[NK];
#pragma unroll
                    for (int k = 0; k < NK; ++k) K[k] = 0.0;
                    if (ND == 1 || c != 0) {
#pragma unroll
                        for (int j = 0; j < J; ++j) {
                            const double2 z = Z[32 * j];
                            if constexpr (ND == 3) {
                                K[0] = fma(pa[j], z.x, K[0]); K[1] = fma(pb[j], z.y, K[1]);
                                K[2] = fma(pa[j], z.y, K[2]); K[3] = fma(pb[j], z.x, K[3]);
                                K[4] = fma(pc[j], z.x, K[4]); K[5] = fma(pd[j], z.y, K[5]);
                                K[6] = fma(pc[j], z.y, K[6]); K[7] = fma(pd[j], z.x, K[7]);
                            } else if constexpr (ND == 2) {
                                K[0] = fma(pa[j], z.x, K[0]); K[1] = fma(pb[j], z.y, K[1]);
                                K[2] = fma(pa[j], z.y, K[2]); K[3] = fma(pb[j], z.x, K[3]);
                            } else {
                                K[0] += z.x; K[1] += z.y;
                            }
                        }
                    } else {                                   // Z = 1: no table read, no multiply
#pragma unroll
                        for (int j = 0; j < J; ++j) {
                            if constexpr (ND == 3) { K[0] += pa[j]; K[3] += pb[j]; K[4] += pc[j]; K[7] += pd[j]; }
                            else { K[0] += pa[j]; K[3] += pb[j]; }
                        }
                    }
                    int idx;
                    const double w = butterfly_sum<NK>(K, lane, idx);
                    if ((lane & (32 / NK - 1)) == 0) {
                        double* dst = part + idx * G + g;
                        *dst = ib == 0 ? w : *dst + w;
                    }
                }
            }
        }
        __syncthreads();
        // phase 3: unfold the sign patterns into rho
        for (int w = threadIdx.x; w < G * NPAT; w += blockDim.x) {
            const int g = w / NPAT, pat = w - g * NPAT;
            const int iq = __ldg(plan.gout + w);
            if (iq < 0) continue;
            const int sa = pat & 1;                                // conj of the pattern with all signs flipped
            const int sb = ND > 1 ? (((pat >> 1) & 1) ^ sa) : 0;
            const int sc = ND > 2 ? (((pat >> 2) & 1) ^ sa) : 0;
            const double* src = part + (ND == 3 ? 4 * sb : 0) * G + g;
            const double k0 = src[0], k1 = src[G];
            double re, im;
            if constexpr (ND == 1) { re = k0; im = k1; }
            else {
                const double k2 = src[2 * G], k3 = src[3 * G];
                const int sl_ = ND == 3 ? sc : sb;                 // sign of the last multiplied factor
                re = sl_ ? k0 + k1 : k0 - k1;
                im = sl_ ? k3 - k2 : k3 + k2;
            }
            const int cb = sl / M, ct = sl - cb * M;
            rho[rho_pair_base(cb, iq, 0, nq, nTB) + rho_slice_off(ct, nq)] = re;
            rho[rho_pair_base(cb, iq, 1, nq, nTB) + rho_slice_off(ct, nq)] = sa ? -im : im;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// rho_q build for commensurate q on the FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64).
// The particle sum of the lattice path is a small real GEMM per time slice,
//        C[row][col] = sum_i  Lrow(i) * Rcol(i),
// with, in 3-D,  L rows = {Re,Im}(X^|a| Y^|b|), {Re,Im}(X^|a| conj Y^|b|) for every (|a|,|b|) column of the q-set and
//                R cols = {Re,Im}(Z^|c|) for every |c|;   2-D: L = {Re,Im} X^|a|, R = {Re,Im} Y^|b|;   1-D: L = 1, R = X^|a|.
// The 8 (3-D) / 4 (2-D) / 2 (1-D) real sums K of a sign-symmetry group are entries of C (see phase C), and the
// reduction over particles happens inside the tensor-core accumulation -- no shuffles, no per-particle FP64 issue
// slots: one DMMA = 256 FMAs.  A CTA (4 warps) works on one (configuration, slice) at a time; every warp owns a
// quarter of the particles and PRIVATE operand planes, so the warps only meet once per slice:
//   phase A (lane = particle, blocks of 32) evaluates ND sincos, the power recurrences and the L/R planes straight
//           into the warp's shared-memory planes (stride 36 doubles: the 8 rows of a fragment fall into distinct banks);
//   phase B 8 k-steps of MT x NT DMMAs on those planes (__syncwarp only, so one warp's tensor work overlaps another
//           warp's FP64 phase A);
//   phase C (CTA barrier) adds the four warps' partial tiles in fixed order and unfolds the sign patterns into rho.
// ---------------------------------------------------------------------------------------------
constexpr int kMmaChunk = 32;                // particles per warp block (8 k-steps x 4)
constexpr int kMmaStride = kMmaChunk + 4;    // plane stride in doubles: 32 B past a multiple of 128 B
constexpr int kMmaWarps = 4;                 // warps per CTA, each with private operand planes

struct MmaPlan {
    const unsigned* unfold;   // [2 nq]  output k = 2 q + {re, im}:  rho = (+-) C[o0] (+-) C[o1], packed by the host as
                              //         o0 | o1 << 12 | neg0 << 24 | neg1 << 25 (offsets in doubles into the staged C tiles);
                              //         the host resolves sign patterns, coinciding factors and zero planes once per q-set
    int G, nL, nR;      // groups; L rows / R cols in use INCLUDING the reserved zero plane (index nL-1 / nR-1);
                        // the kernel's MT x NT tiles cover them, everything from the zero plane on is cleared
    // small lookup tables, passed by value so that they sit in the constant bank:
    short lmap[81];     // 3-D: lmap[a*9 + b] first L row of column (a,b) or -1 (fixed stride 9); 2-D: lmap[a]; 1-D: unused
    short rmap[17];     // [nmax_last+1] first R col of |n_last| or -1
};
// Rows are stored only once when factors coincide (3-D):  b = 0: X conj(Y) = X Y (2 rows: Re, Im of X^a);
// a = 0: X conj(Y) = conj(Y^b) (2 rows: Re, Im of Y^b, Im- = -Im+);  a = b = 0: one row of ones (Im = the zero row).
// Likewise c = 0 stores the single column of ones (Im = the zero column).  Row nL-1 / column nR-1 are all zero.

__device__ __forceinline__ void dmma8x8x4(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// Persistent warps, each fully independent: a WARP owns a whole (configuration, slice) at a time -- all particle
// blocks of the slice accumulate into the warp's register tiles, so there is no cross-warp reduction and no CTA
// barrier after the prologue (the CTA-per-slice version of this kernel spent 10 % of its time at the two barriers of
// phase C and serialised the unfold behind them).  Slices are handed out by a global ticket counter (`sched[0]`;
// the last warp to retire re-arms it through `sched[1]`), which keeps the tail to one slice per warp.
// The coordinates of the NEXT particle block are fetched into registers while the current one is processed.  The
// prefetch address carries a (run-time zero) dependence on the quadrant integers of the current block's sincos:
// without it ptxas issues the prefetch LDGs right before the first use of the current coordinates, on the same
// scoreboard, and the "prefetch" waits for itself (21 % of all stall samples in profiles/r01j).
// NM > 0 (3-D only): every |n_d| <= NM, the power recurrences and the (a,b) column loop are fully unrolled with the
// powers held in registers; NM = 0: run-time loop bounds.
template <int ND, int MT, int NT, int NM, bool DIRECT = false>
__global__ void __launch_bounds__(128) rho_lattice_mma_kernel(const double* __restrict__ pos, const MmaPlan plan,
                                                               double* __restrict__ rho, int nslices, int N, int Npad, int nq,
                                                               int3 nmax, double3 kphase, unsigned* __restrict__ sched,
                                                               int zero_mask, int split, double* __restrict__ partial, int M,
                                                               unsigned ticket_wrap, int3 stride, double* __restrict__ soa_out) {
    // DIRECT = false: `pos` is the library's own pos[slice][d][Npad] layout; stride and soa_out are ignored (and this
    //   instantiation compiles exactly as it did before they existed -- as run-time values they cost the batched path 5 %).
    // DIRECT = true (single-walker graph): `pos` is the reference's beads array itself, double[slice][N_ext][ND], read over
    //   the host link from inside this kernel; stride = (N_ext ND, ND, 1) doubles per (slice, particle, dimension); the
    //   transposed copy that the other kernels expect is written to soa_out on the way.
    constexpr int ML = MT, NR = NT;                         // every tile is computed; unused rows / cols are zero planes
    constexpr int ntile = ML * NR;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) double sm[];
    constexpr int region = 8 * (ML + NR) * kMmaStride + ntile * 64;   // doubles of shared memory per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* Lp = sm + warp * region;                        // [8*ML][kMmaStride]   this warp's L planes
    double* Rp = Lp + 8 * ML * kMmaStride;                  // [8*NR][kMmaStride]   this warp's R planes
    double* Cw = Rp + 8 * NR * kMmaStride;                  // [ntile][64]          this warp's C staging
    unsigned* s_unfold = reinterpret_cast<unsigned*>(sm + kMmaWarps * region);   // [2 nq]  CTA copy of the unfold table
    const int il = lane;                                    // particle of the block this lane owns in phase A
    const int nchunk = (N + kMmaChunk - 1) / kMmaChunk;     // particle blocks of a slice

    for (int w = threadIdx.x; w < 2 * nq; w += blockDim.x) s_unfold[w] = __ldg(plan.unfold + w);
    // the reserved zero plane and the padding rows / columns (never written by phase A) must be zero
    for (int w = (plan.nL - 1) * kMmaStride + lane; w < 8 * ML * kMmaStride; w += 32) Lp[w] = 0.0;
    for (int w = (plan.nR - 1) * kMmaStride + lane; w < 8 * NR * kMmaStride; w += 32) Rp[w] = 0.0;
    __syncthreads();

    // NM > 0: plane offsets of every (a,b) column / |c| power, resolved once (register resident) instead of a constant-bank
    // lookup + multiply per particle block
    int loff[NM > 0 ? (NM + 1) * (NM + 1) : 1], roff[NM > 0 ? NM + 1 : 1];
    if constexpr (NM > 0) {
#pragma unroll
        for (int a = 0; a <= NM; ++a)
#pragma unroll
            for (int b = 0; b <= NM; ++b) {
                const int row = plan.lmap[a * 9 + b];
                loff[a * (NM + 1) + b] = row >= 0 ? row * kMmaStride + il : -1;
            }
#pragma unroll
        for (int m = 0; m <= NM; ++m) {
            const int col = plan.rmap[m];
            roff[m] = col >= 0 ? col * kMmaStride + il : -1;
        }
    }
    // requested at the start of a slice, needed at its last particle block.  atomicInc with a RUN-TIME wrap limit
    // (0xffffffff, a kernel argument): ptxas warp-aggregates every form of "add 1" -- atomicAdd, atom.add, atom.inc with a
    // literal limit -- and broadcasts the aggregated result with a shuffle right behind the atomic, which made every warp
    // sit out one atomic round trip per slice (10 % of all stall samples); this form stays a plain ATOMG.INC whose
    // result is first read ~7 particle blocks later.
    auto ticket_request = [&]() {
        unsigned v = 0;
        if (lane == 0) v = atomicInc(sched, ticket_wrap);
        return v;
    };
    auto fetch = [&](int sl, int ch, int dep, double (&x)[3]) {
        const int i = ch * kMmaChunk + il;
        x[0] = x[1] = x[2] = 0.0;
        if (sl < nslices && i < N) {
            if constexpr (DIRECT) {
                const double* ps = pos + static_cast<size_t>(sl) * stride.x + static_cast<size_t>(i + dep) * stride.y;
                x[0] = __ldg(ps);
                if constexpr (ND > 1) x[1] = __ldg(ps + stride.z);
                if constexpr (ND > 2) x[2] = __ldg(ps + 2 * stride.z);
            } else {
                const double* ps = pos + static_cast<size_t>(sl) * ND * Npad + (i + dep);
                x[0] = __ldg(ps);
                if constexpr (ND > 1) x[1] = __ldg(ps + Npad);
                if constexpr (ND > 2) x[2] = __ldg(ps + 2 * Npad);
            }
        }
    };
    // Work item = (slice, part): `split` warps (anywhere on the GPU) share the particle blocks of a slice when there
    // are too few slices to occupy every warp (small batches, one walker); part p takes blocks p, p + split, ...
    const int nitems = nslices * split;
    double xn[3];
    int item = static_cast<int>(__shfl_sync(FULL, ticket_request(), 0));
    fetch(item < nitems ? item / split : nslices, item % split, 0, xn);
    while (item < nitems) {
        const int sl = item / split, part = item - sl * split;
        const unsigned tk = ticket_request();               // next item's ticket: requested now, read at the last block
        int item_next = nitems;
        double acc[2][MT][NT][2];                           // two accumulator sets (even / odd k-steps) for DMMA ILP
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int n = 0; n < NT; ++n) acc[e][m][n][0] = acc[e][m][n][1] = 0.0;
        for (int ch = part; ch < nchunk; ch += split) {
            const double xc[3] = {xn[0], xn[1], xn[2]};
            // ---- phase A: thread = particle of the chunk -------------------------------------------------
            {
                const bool live = ch * kMmaChunk + il < N;
                if (DIRECT && ch * kMmaChunk + il < Npad) {                  // the transposed copy (zeros beyond N)
                    double* dst = soa_out + static_cast<size_t>(sl) * ND * Npad + ch * kMmaChunk + il;
                    dst[0] = xc[0];
                    if constexpr (ND > 1) dst[Npad] = xc[1];
                    if constexpr (ND > 2) dst[2 * Npad] = xc[2];
                }
                // Particles beyond N (ragged last block) must add nothing: their R entries (the powers of the last
                // dimension's phase, incl. the column of ones) are zeroed with selects, so every product L x R vanishes
                // and the L side needs no mask -- FP64 instructions are the scarce resource here, selects are not.
                const double lv = live ? 1.0 : 0.0;
                double ex_s, ex_c, ey_s = 0.0, ey_c = 1.0, ez_s = 0.0, ez_c = 1.0;
                int qx = 0, qy = 0, qz = 0;                  // quadrant integers: the prefetch below depends on them
                // (a 128-entry phasor table + 3-term Taylor sincos, 16 FP64 instructions instead of 21, was measured SLOWER
                // here: 73.9 vs 72.0 us -- the scattered 16-byte table reads cost more than the five FMAs they save)
                sincos_fast(kphase.x * xc[0], ex_s, ex_c, qx);
                if constexpr (ND > 1) sincos_fast(kphase.y * xc[1], ey_s, ey_c, qy);
                if constexpr (ND > 2) sincos_fast(kphase.z * xc[2], ez_s, ez_c, qz);
                {
                    const int dep = (qx | qy | qz) & zero_mask;   // always 0, but only known at run time
                    if (ch + split < nchunk) fetch(sl, ch + split, dep, xn);
                    else {
                        item_next = static_cast<int>(__shfl_sync(FULL, tk, 0));
                        fetch(item_next < nitems ? item_next / split : nslices, item_next % split, dep, xn);
                    }
                }
                if constexpr (ND == 3 && NM > 0) {
                    // compile-time bounds: powers of the three phases in registers, (a,b) columns unrolled
                                        double xr[NM + 1], xi[NM + 1], yr[NM + 1], yi[NM + 1], zr[NM + 1], zi[NM + 1];
                    xr[0] = 1.0; xi[0] = 0.0; yr[0] = 1.0; yi[0] = 0.0; zr[0] = lv; zi[0] = 0.0;
                    xr[1] = ex_c; xi[1] = ex_s; yr[1] = ey_c; yi[1] = ey_s;          // first powers: no multiply by (1, 0)
                    zr[1] = live ? ez_c : 0.0; zi[1] = live ? ez_s : 0.0;
#pragma unroll
                    for (int m = 2; m <= NM; ++m) {
                        if (m == 2) {                               // squares in 3 instructions: (1 - 2 s^2, 2 s c)
                            const double sx = ex_s + ex_s, sy = ey_s + ey_s, sz = zi[1] + zi[1];
                            xr[2] = fma(-sx, ex_s, 1.0); xi[2] = sx * ex_c;
                            yr[2] = fma(-sy, ey_s, 1.0); yi[2] = sy * ey_c;
                            zr[2] = fma(-sz, zi[1], zr[0]); zi[2] = sz * zr[1];      // zr[0] carries the live mask
                            continue;
                        }
                        xr[m] = fma(xr[m - 1], ex_c, -xi[m - 1] * ex_s);
                        xi[m] = fma(xr[m - 1], ex_s, xi[m - 1] * ex_c);
                        yr[m] = fma(yr[m - 1], ey_c, -yi[m - 1] * ey_s);
                        yi[m] = fma(yr[m - 1], ey_s, yi[m - 1] * ey_c);
                        zr[m] = fma(zr[m - 1], ez_c, -zi[m - 1] * ez_s);
                        zi[m] = fma(zr[m - 1], ez_s, zi[m - 1] * ez_c);
                    }
#pragma unroll
                    for (int m = 0; m <= NM; ++m) {
                        if (roff[m] >= 0) {
                            Rp[roff[m]] = zr[m];
                            if (m > 0) Rp[roff[m] + kMmaStride] = zi[m];
                        }
                    }
#pragma unroll
                    for (int a = 0; a <= NM; ++a) {
#pragma unroll
                        for (int b = 0; b <= NM; ++b) {
                            if (loff[a * (NM + 1) + b] >= 0) {
                                double* d = Lp + loff[a * (NM + 1) + b];
                                if (a > 0 && b > 0) {
                                    const double m1 = xr[a] * yr[b], m2 = xi[a] * yi[b], m3 = xr[a] * yi[b], m4 = xi[a] * yr[b];
                                    d[0] = m1 - m2;                   // Re X Y
                                    d[kMmaStride] = m3 + m4;          // Im X Y
                                    d[2 * kMmaStride] = m1 + m2;      // Re X conj(Y)
                                    d[3 * kMmaStride] = m4 - m3;      // Im X conj(Y)
                                } else if (a > 0) {
                                    d[0] = xr[a];
                                    d[kMmaStride] = xi[a];
                                } else if (b > 0) {
                                    d[0] = yr[b];
                                    d[kMmaStride] = yi[b];
                                } else {
                                    d[0] = 1.0;
                                }
                            }
                        }
                    }
                } else {
                {   // R planes: powers of the last dimension's phase
                    const double bs = ND == 3 ? ez_s : (ND == 2 ? ey_s : ex_s);
                    const double bc = ND == 3 ? ez_c : (ND == 2 ? ey_c : ex_c);
                    const int nl = ND == 3 ? nmax.z : (ND == 2 ? nmax.y : nmax.x);
                    double pr = lv, pi = 0.0;
                    for (int m = 0; m <= nl; ++m) {
                        const int col = plan.rmap[m];
                        if (col >= 0) {
                            Rp[col * kMmaStride + il] = pr;
                            if (m > 0 || ND == 1) Rp[(col + 1) * kMmaStride + il] = pi;
                        }
                        const double nr = fma(pr, bc, -pi * bs);
                        pi = fma(pr, bs, pi * bc);
                        pr = nr;
                    }
                }
                if constexpr (ND == 1) {
                    Lp[il] = 1.0;
                } else if constexpr (ND == 2) {
                    double pr = 1.0, pi = 0.0;
                    for (int a = 0; a <= nmax.x; ++a) {
                        const int row = plan.lmap[a];
                        if (row >= 0) {
                            Lp[row * kMmaStride + il] = pr;
                            if (a > 0) Lp[(row + 1) * kMmaStride + il] = pi;
                        }
                        const double nr = fma(pr, ex_c, -pi * ex_s);
                        pi = fma(pr, ex_s, pi * ex_c);
                        pr = nr;
                    }
                } else {
                    double xr = 1.0, xi = 0.0;
                    for (int a = 0; a <= nmax.x; ++a) {
                        double yr = 1.0, yi = 0.0;
                        for (int b = 0; b <= nmax.y; ++b) {
                            const int row = plan.lmap[a * 9 + b];
                            if (row >= 0) {
                                double* d = Lp + row * kMmaStride + il;
                                if (a > 0 && b > 0) {
                                    const double m1 = xr * yr, m2 = xi * yi, m3 = xr * yi, m4 = xi * yr;
                                    d[0] = m1 - m2;                   // Re X Y
                                    d[kMmaStride] = m3 + m4;          // Im X Y
                                    d[2 * kMmaStride] = m1 + m2;      // Re X conj(Y)
                                    d[3 * kMmaStride] = m4 - m3;      // Im X conj(Y)
                                } else if (a > 0) {                   // Y = 1
                                    d[0] = xr;
                                    d[kMmaStride] = xi;
                                } else if (b > 0) {                   // X = 1 (times the live mask)
                                    d[0] = yr;
                                    d[kMmaStride] = yi;
                                } else {
                                    d[0] = 1.0;
                                }
                            }
                            const double nr = fma(yr, ey_c, -yi * ey_s);
                            yi = fma(yr, ey_s, yi * ey_c);
                            yr = nr;
                        }
                        const double nr = fma(xr, ex_c, -xi * ex_s);
                        xi = fma(xr, ex_s, xi * ex_c);
                        xr = nr;
                    }
                }
                }   // run-time loop bounds
            }
            __syncwarp();
            // ---- phase B: 8 k-steps over the warp's 32 particles, all tiles ------------------------------------
            {
                const int frow = lane >> 2, fk = lane & 3;
                const double* la = Lp + frow * kMmaStride + fk;
                const double* rb = Rp + frow * kMmaStride + fk;
#pragma unroll
                for (int ks = 0; ks < kMmaChunk / 4; ++ks) {
                    double a[MT], b[NT];
#pragma unroll
                    for (int m = 0; m < MT; ++m) a[m] = la[m * 8 * kMmaStride + 4 * ks];
#pragma unroll
                    for (int n = 0; n < NT; ++n) b[n] = rb[n * 8 * kMmaStride + 4 * ks];
#pragma unroll
                    for (int m = 0; m < MT; ++m)
#pragma unroll
                        for (int n = 0; n < NT; ++n) dmma8x8x4(acc[ks & 1][m][n], a[m], b[n]);
                }
            }
            __syncwarp();
        }
        // ---- phase C (warp-local): stage the tiles, unfold the sign patterns into rho ---------------------------
        bool unfold = true;
        if (split > 1) {
            // park this part's tiles in global scratch; the warp that completes the slice adds the parts in fixed order
            // (deterministic) and unfolds.  slice_done[sl] is re-armed by that warp.
            double2* mine = reinterpret_cast<double2*>(partial + (static_cast<size_t>(sl) * split + part) * (ntile * 64)) + lane;
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int n = 0; n < NT; ++n)
                    __stcg(mine + (m * NR + n) * 32, make_double2(acc[0][m][n][0] + acc[1][m][n][0], acc[0][m][n][1] + acc[1][m][n][1]));
            __threadfence();
            __syncwarp();
            unsigned prev = 0;
            if (lane == 0) prev = atomicAdd(sched + 2 + sl, 1u);
            prev = __shfl_sync(FULL, prev, 0);
            unfold = prev == static_cast<unsigned>(split - 1);
            if (unfold) {
                __threadfence();
                if (lane == 0) sched[2 + sl] = 0u;
#pragma unroll
                for (int m = 0; m < MT; ++m)
#pragma unroll
                    for (int n = 0; n < NT; ++n) {
                        const double2* src = reinterpret_cast<const double2*>(partial + static_cast<size_t>(sl) * split * (ntile * 64)) + (m * NR + n) * 32 + lane;
                        double2 t = __ldcg(src);
                        for (int p = 1; p < split; ++p) {
                            const double2 u = __ldcg(src + static_cast<size_t>(p) * (ntile * 32));
                            t.x += u.x; t.y += u.y;
                        }
                        reinterpret_cast<double2*>(Cw + (m * NR + n) * 64)[lane] = t;
                    }
            }
        } else {
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    double2* d = reinterpret_cast<double2*>(Cw + (m * NR + n) * 64) + lane;
                    *d = make_double2(acc[0][m][n][0] + acc[1][m][n][0], acc[0][m][n][1] + acc[1][m][n][1]);
                }
        }
        __syncwarp();
        const int cb = sl / M, ct = sl - cb * M;
        const size_t out_base = rho_pair_base(cb, 0, 0, nq, rho_tblocks(M)) + rho_slice_off(ct, nq);   // + 4 (2 q + {re, im})
        for (int k = lane; unfold && k < 2 * nq; k += 32) {
            const unsigned d = s_unfold[k];
            const double a = Cw[d & 0xfffu], b = Cw[(d >> 12) & 0xfffu];
            const double sa_ = __hiloint2double(__double2hiint(a) ^ static_cast<int>((d << 7) & 0x80000000u), __double2loint(a));
            const double sb_ = __hiloint2double(__double2hiint(b) ^ static_cast<int>((d << 6) & 0x80000000u), __double2loint(b));
            rho[out_base + 4 * k] = sa_ + sb_;
        }
        __syncwarp();
        item = item_next;
    }
    // re-arm the ticket counter for the next launch: every warp takes exactly one failing ticket before it retires
    if (lane == 0) {
        __threadfence();
        const unsigned done = atomicAdd(sched + 1, 1u);
        if (done == gridDim.x * kMmaWarps - 1) { sched[0] = 0u; sched[1] = 0u; }
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

__global__ void __launch_bounds__(128, 7) isf_corr_kernel(const double* __restrict__ rho, double* __restrict__ cfg, int M, int nq,
                                                        int npairs, int lpq, int tsplit, double invN,
                                                        const unsigned char* __restrict__ commensurate) {
    // lpq lanes own the tau blocks of one pair; tsplit (1 or 2) such lane groups share the pair and split the t0
    // range, their partial sums are added with one shuffle.  Lanes per pair = lpq * tsplit.
    extern __shared__ __align__(16) double sm[];
    const int len = 2 * M + 16;
    const int plen = corr_idx(len) + 2;                      // padded length of one array (even -> 16-byte aligned)
    const int lpp = lpq * tsplit;                            // lanes per pair
    const int ppw = 32 / lpp;                                // pairs per warp
    const int ppc = ppw * (blockDim.x >> 5);                 // pairs per CTA
    const int pair0 = blockIdx.x * ppc;
    // stage: the threads of a pair walk its rho values (whole 32-byte sectors per four slices); each value goes to every periodic image
    // i = t, t+M, t+2M < len.
    {
        const int tstep = blockDim.x / ppc, lp = threadIdx.x / tstep;      // tstep consecutive threads walk one pair's rows
        const int pair = pair0 + lp;
        if (pair < npairs) {
            const int b = pair / nq, iq = pair - b * nq;
            const double* srcc = rho + rho_pair_base(b, iq, 0, nq, rho_tblocks(M));
            const double* srcs = srcc + 4;
            double* dc = sm + (2 * lp + 0) * plen;
            double* ds = sm + (2 * lp + 1) * plen;
            int t = threadIdx.x - lp * tstep;
            for (; t + 3 * tstep < M; t += 4 * tstep) {      // 8 independent loads in flight
                double c[4], sn[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    c[u] = __ldg(srcc + rho_slice_off(t + u * tstep, nq));
                    sn[u] = __ldg(srcs + rho_slice_off(t + u * tstep, nq));
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    for (int i = t + u * tstep; i < len; i += M) { dc[corr_idx(i)] = c[u]; ds[corr_idx(i)] = sn[u]; }
            }
            for (; t < M; t += tstep) {
                const double c = __ldg(srcc + rho_slice_off(t, nq));
                const double sn = __ldg(srcs + rho_slice_off(t, nq));
                for (int i = t; i < len; i += M) { dc[corr_idx(i)] = c; ds[corr_idx(i)] = sn; }
            }
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lp = warp * ppw + lane / lpp;
    const int pair = min(pair0 + lp, npairs - 1);            // surplus lanes recompute the last pair (no divergent exit before shuffles)
    const bool owner = pair0 + lp < npairs;
    const int b = pair / nq, iq = pair - b * nq;
    const double* WC = sm + (2 * (pair - pair0)) * plen;
    const double* WS = WC + plen;
    const size_t cfg_stride = static_cast<size_t>(nq) + static_cast<size_t>(nq) * M;
    double* out = cfg + static_cast<size_t>(b) * cfg_stride + nq + static_cast<size_t>(iq) * M;
    const int half = M / 2;
    const int nblk = (half + 1 + 7) / 8;                     // tau blocks of 8
    const int lin = lane % lpp;                              // lane within the pair
    const int seg = lin / lpq;                               // which part of the t0 range
    const int nt8 = (M + 7) / 8;                             // t0 blocks of 8
    const int tb0 = (nt8 * seg) / tsplit, tb1 = (nt8 * (seg + 1)) / tsplit;
    const int rounds = (nblk + lpq - 1) / lpq;
    for (int r = 0; r < rounds; ++r) {
        const int tb = r * lpq + (lin % lpq);
        const int tau0 = 8 * min(tb, nblk - 1);
        double acc[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[v] = 0.0;
        for (int t8 = tb0; t8 < tb1; ++t8) {
            const int t0 = 8 * t8;
            // cos part, then sin part: one 8-value broadcast block a(t0..t0+7) and one 16-value window per part
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                const double* W = part ? WS : WC;
                double a[8], w[16];
                const double2* pa = reinterpret_cast<const double2*>(W + corr_idx(t0));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double2 x = pa[k];
                    a[2 * k] = x.x; a[2 * k + 1] = x.y;
                }
                if (t0 + 8 > M) {                            // last block of a non-multiple-of-8 M: drop t0 >= M
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (t0 + u >= M) a[u] = 0.0;
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const double2* pw = reinterpret_cast<const double2*>(W + corr_idx(t0 + tau0 + 8 * h));
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const double2 x = pw[k];
                        w[8 * h + 2 * k] = x.x; w[8 * h + 2 * k + 1] = x.y;
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
#pragma unroll
                    for (int v = 0; v < 8; ++v) acc[v] = fma(a[u], w[u + v], acc[v]);
                }
            }
        }
        if (tsplit == 2) {
#pragma unroll
            for (int v = 0; v < 8; ++v) acc[v] += __shfl_xor_sync(0xffffffffu, acc[v], lpq);
        }
        if (owner && seg == 0 && tb < nblk) {
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
    }
    // odd M never occurs upstream (setup.cpp:1001-1008 forces M even); for odd M, tau = (M+1)/2.. mirror as well.
}

// ---------------------------------------------------------------------------------------------
// tau-correlation on the FP64 tensor cores.  With tau = 8 i + j the circular correlation of a (config, q) pair is
//        F[8 i + j] = sum_s a[(s - 8 i) mod M] * a[(s + j) mod M]         (a = C, then a = S, same accumulators)
// i.e. D = A B with A[i][s] = a[s - 8 i], B[s][j] = a[s + j], contracted over s in [0, M): one DMMA m8n8k4 per four
// s and per block of 64 tau.  One warp per pair, MTC = ceil((M/2 + 1) / 64) accumulator tiles.  The pair's C and S
// are staged periodically extended (offset OFF = 64 MTC to the left) with 4 doubles of padding after every 8, which
// puts the 8 rows of an A fragment (12 doubles apart) and the 11 consecutive elements of a B fragment in distinct banks.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int corrm_idx(int e) { return e + 4 * (e >> 3); }

// A CTA is four warps = the four configurations b0..b0+3 of one q (blockIdx = (b0 / 4) * nq + q); every warp stages
// its own pair (whole 32-byte sectors of the rho buffer) and runs independently (no CTA barrier on the per-configuration path).
// PARTIAL = false: out = cfg, one row of nq + nq*M results per configuration.
// PARTIAL = true : the four warps' results are added in fixed order in shared memory and ACCUMULATED into the CTA's own
//                  persistent row segment (out = rows[ceil(B/4)][nq][M/2 + 1]); per-configuration rows are never
//                  written and the bin accumulation costs neither a launch nor a dependency between CTAs
//                  (pimcb_measure, all q commensurate).  bins_fold_kernel folds the rows into the bin when it is read.
template <int MTC, bool PARTIAL>
__global__ void __launch_bounds__(128) isf_corr_mma_kernel(const double* __restrict__ rho, double* __restrict__ out, int M, int nq,
                                                            int B, double invN, const unsigned char* __restrict__ commensurate) {
    extern __shared__ __align__(16) double sm[];
    constexpr int OFF = 64 * MTC;
    const int Mpad = (M + 3) & ~3;
    const int ext = OFF + Mpad + 8;                          // extended length in elements
    const int plen = corrm_idx(ext) + 4;                     // padded doubles per array
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bq = blockIdx.x / nq, iq = blockIdx.x - bq * nq;
    const int b = 4 * bq + warp;
    const bool live = b < B;
    const int half = M / 2;
    double* dc = sm + (2 * warp) * plen;
    double* ds = dc + plen;
    const int fi = lane >> 2, fk = lane & 3;                 // fragment row / k index of this lane
    double acc[2][MTC][2];                                   // even / odd k-steps accumulate separately (DMMA ILP)
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int m = 0; m < MTC; ++m) acc[e][m][0] = acc[e][m][1] = 0.0;
    if (live) {
        {   // stage: element e of the periodically extended arrays is rho(t = (e - OFF) mod M)
            const double* rc = rho + rho_pair_base(b, iq, 0, nq, rho_tblocks(M));
            const double* rs = rc + 4;
            int t = ((lane - OFF) % M + M) % M;
            for (int e0 = lane; e0 < ext; e0 += 128) {
                double c[4], sn[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (e0 + 32 * u < ext) { const size_t o = rho_slice_off(t, nq); c[u] = __ldg(rc + o); sn[u] = __ldg(rs + o); }
                    t += 32;
                    while (t >= M) t -= M;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (e0 + 32 * u < ext) { const int k = corrm_idx(e0 + 32 * u); dc[k] = c[u]; ds[k] = sn[u]; }
            }
        }
        __syncwarp();
        // Operand addressing without per-element index arithmetic.  k-step ks covers s = 4 ks + fk; two k-steps advance
        // the padded index by exactly 12, and OFF is a multiple of 8, so with h = ks >> 1
        //     B[s][j = fi]        sits at  pb{0,1} + 12 h   (even / odd ks; the odd base absorbs the "+4" and any padding jump)
        //     A[i = fi + 8 m][s]  sits at  pa + 12 h (+4 for odd ks) - 96 m   (fk + 4 never crosses a group of 8)
        const int nfull = M >> 2;                            // k-steps with all four s < M
        const int nks = Mpad >> 2;
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            const double* E = part ? ds : dc;
            const double* pa = E + corrm_idx(fk - 8 * fi + OFF);
            const double* pb0 = E + corrm_idx(fk + fi + OFF);
            const double* pb1 = E + corrm_idx(4 + fk + fi + OFF);
            const int hmax = nfull >> 1;
#pragma unroll 2
            for (int h = 0; h < hmax; ++h) {
                const double bv0 = pb0[12 * h], bv1 = pb1[12 * h];
                double av0[MTC], av1[MTC];
#pragma unroll
                for (int m = 0; m < MTC; ++m) { av0[m] = pa[12 * h - 96 * m]; av1[m] = pa[12 * h + 4 - 96 * m]; }
#pragma unroll
                for (int m = 0; m < MTC; ++m) dmma8x8x4(acc[0][m], av0[m], bv0);
#pragma unroll
                for (int m = 0; m < MTC; ++m) dmma8x8x4(acc[1][m], av1[m], bv1);
            }
            for (int ks = 2 * hmax; ks < nks; ++ks) {        // at most two: a full even k-step and / or the masked tail
                const int h = ks >> 1, sidx = 4 * ks + fk;
                const double bv = (ks & 1) ? pb1[12 * h] : pb0[12 * h];
#pragma unroll
                for (int m = 0; m < MTC; ++m) {
                    const double v = pa[12 * h + 4 * (ks & 1) - 96 * m];
                    const double vm = sidx < M ? v : 0.0;    // the contraction runs over s < M only
                    if (ks & 1) dmma8x8x4(acc[1][m], vm, bv); else dmma8x8x4(acc[0][m], vm, bv);
                }
            }
        }
    }
    const size_t row_len = static_cast<size_t>(nq) + static_cast<size_t>(nq) * M;
    if constexpr (!PARTIAL) {
        if (!live) return;
        double* row = out + static_cast<size_t>(b) * row_len;
        double* dst = row + nq + static_cast<size_t>(iq) * M;
#pragma unroll
        for (int m = 0; m < MTC; ++m)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int tau = 64 * m + 8 * fi + 2 * fk + e;
                if (tau > half) continue;
                const double val = (acc[0][m][e] + acc[1][m][e]) * invN;
                dst[tau] = val;
                if (tau > 0 && tau < M - tau) dst[M - tau] = val;
                if (tau == 0 && commensurate[iq]) row[iq] = val;
            }
    } else {
        // park F(tau <= M/2) of this configuration at the start of the warp's own staging area, then add the four
        // configurations in fixed order (b0, b0+1, b0+2, b0+3) -- dead warps contribute exact zeros
        __syncwarp();
#pragma unroll
        for (int m = 0; m < MTC; ++m)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int tau = 64 * m + 8 * fi + 2 * fk + e;
                if (tau <= half) dc[tau] = (acc[0][m][e] + acc[1][m][e]) * invN;
            }
        __syncthreads();
        // out = persistent quad rows [quad][q][M/2 + 1]: this CTA is the only writer of its row segment, launch after
        // launch, so a plain read-add-write accumulates the bin without any ordering between CTAs
        double* row = out + (static_cast<size_t>(bq) * nq + iq) * (half + 1);
        for (int tau = threadIdx.x; tau <= half; tau += blockDim.x)
            row[tau] += ((sm[tau] + sm[2 * plen + tau]) + sm[4 * plen + tau]) + sm[6 * plen + tau];
    }
}

// The same correlation with PERSISTENT CTAs and the staging of the next work item overlapped with the tensor work of the
// current one.  isf_corr_mma_kernel launches one CTA per (four configurations, q) and spends 40 % of its instructions
// outside the DMMA loops: index arithmetic of the staging (modulo M, padded index, one global address per element and
// image), which every warp repeats for every pair although none of it depends on the pair; the FP64 datapath is 2/3 busy
// (profiles/r02q_kernels.md) although a DMMA loop fed from shared memory keeps it 93 % busy (tools/micro/dmma_lds.cu).
// Here a CTA walks the items blockIdx.x, blockIdx.x + gridDim.x, ... and
//   1. per item, each warp writes the pair it holds in REGISTERS (M unique values of C and of S, slice t = lane + 32 u)
//      to its staging area: both periodic images of a value from the one register copy, a dozen integer instructions
//      per value,
//   2. issues the loads of its NEXT pair into those registers (not waited for) and, on the bin path, of the row segment
//      the item will be added to,
//   3. runs the DMMA loops of the current pair, 4. parks / stores the result.
// PARTIAL: the four configurations' F(tau) meet in a small double-buffered parking area, ONE CTA barrier per item.
// Same arithmetic and the same summation order as isf_corr_mma_kernel: results are bit-identical.
// Requires M >= 64 MTC + 11 (at most two images per value); the host falls back to isf_corr_mma_kernel otherwise.
// MEASURED (C2, profiles/r02u_corr_experiments.md): 17.6 us per 64 configurations against 16.7 us for
// isf_corr_mma_kernel -- not faster, so corr mode 1 stays the default and this kernel is the A/B leg (corr mode 2).
template <int MTC, bool PARTIAL>
__global__ void __launch_bounds__(128, (MTC == 1 ? 7 : (MTC == 2 ? 6 : (MTC == 3 ? 4 : 3)))) isf_corr_mma_pipe_kernel(
    const double* __restrict__ rho, double* __restrict__ out, int M, int nq, int B, double invN,
    const unsigned char* __restrict__ commensurate, int nitems) {
    extern __shared__ __align__(16) double sm[];
    constexpr int OFF = 64 * MTC;
    constexpr int U = 4 * MTC;                               // values per lane and array: M <= 128 MTC - 2
    const int Mpad = (M + 3) & ~3;
    const int ext = OFF + Mpad + 8;                          // extended length in elements
    const int plen = corrm_idx(ext) + 4;                     // padded doubles per array
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = M / 2;
    double* dc = sm + (2 * warp) * plen;
    double* ds = dc + plen;
    double* park = sm + 8 * plen;                            // [2][4][half + 1] (PARTIAL; 0 doubles otherwise)
    const int fi = lane >> 2, fk = lane & 3;                 // fragment row / k index of this lane
    const int nfull = M >> 2;                                // k-steps with all four s < M
    const int nks = Mpad >> 2;

    // 0. the item-independent part of the staging, per lane: value u is slice t = lane + 32 u; its first periodic image is
    //    element e0 = (t + OFF) mod M = first + 32 u (- M), its second e0 + M when that is inside the extended array.
    const int first = (lane + OFF) % M;
    const int tstep = 8 * nq * 8;                            // rho_slice_off(t + 32) - rho_slice_off(t)
    const int toff0 = static_cast<int>(rho_slice_off(lane, nq));
    double pc[U], ps[U];                                     // the pair held for the next staging step
    auto prefetch = [&](int item) {
        if (item >= nitems) return;                          // no next item: the registers are never read again
        const int bq = item / nq, iq = item - bq * nq;
        const int b = min(4 * bq + warp, B - 1);             // dead warps read a valid pair and never use it
        const double* rc = rho + rho_pair_base(b, iq, 0, nq, rho_tblocks(M));
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int o = lane + 32 * u < M ? toff0 + u * tstep : 0;      // lanes past the end read slice 0 and park it
            pc[u] = __ldg(rc + o);
            ps[u] = __ldg(rc + 4 + o);
        }
    };
    prefetch(blockIdx.x);
    int parity = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, parity ^= 1) {
        const int bq = item / nq, iq = item - bq * nq;
        const int b = 4 * bq + warp;
        const bool live = b < B;
        // 1. registers -> staging area
        // (the staging indices do not depend on the item; ptxas hoists all 4 U of them out of the loop and spills them to
        // local memory -- an opaque copy of `first` per item keeps them a dozen integer instructions instead)
        int first_i = first;
        asm volatile("" : "+r"(first_i));
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int e0 = first_i + 32 * u;
            if (e0 >= M) e0 -= M;                            // first < M and 32 u <= t < M: one wrap at most
            const int e1 = e0 + M;
            int k0 = corrm_idx(e0), k1 = e1 < ext ? corrm_idx(e1) : k0;     // a missing second image repeats the first
            if (lane + 32 * u >= M) k0 = k1 = 8;             // index 8 of the padded layout is a slot nothing reads
            dc[k0] = pc[u]; ds[k0] = ps[u];
            dc[k1] = pc[u]; ds[k1] = ps[u];
        }
        __syncwarp();
        // 2. the next item's pair on its way while this one is multiplied -- and, on the bin path, this item's row segment
        //    (read-add-write below): fetched now, so that its latency is not paid behind the barrier
        prefetch(item + gridDim.x);
        constexpr int RV = (64 * MTC + 127) / 128;           // row values per thread: half + 1 <= 64 MTC
        double rowv[RV];
        if constexpr (PARTIAL) {
            const double* row = out + (static_cast<size_t>(bq) * nq + iq) * (half + 1);
#pragma unroll
            for (int r = 0; r < RV; ++r) {
                const int tau = threadIdx.x + 128 * r;
                rowv[r] = tau <= half ? __ldcg(row + tau) : 0.0;
            }
        }
        // 3. tensor work (identical to isf_corr_mma_kernel)
        double acc[2][MTC][2];
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int m = 0; m < MTC; ++m) acc[e][m][0] = acc[e][m][1] = 0.0;
        if (live) {
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                const double* E = part ? ds : dc;
                const double* pa = E + corrm_idx(fk - 8 * fi + OFF);
                const double* pb0 = E + corrm_idx(fk + fi + OFF);
                const double* pb1 = E + corrm_idx(4 + fk + fi + OFF);
                const int hmax = nfull >> 1;
#pragma unroll 2
                for (int h = 0; h < hmax; ++h) {
                    const double bv0 = pb0[12 * h], bv1 = pb1[12 * h];
                    double av0[MTC], av1[MTC];
#pragma unroll
                    for (int m = 0; m < MTC; ++m) { av0[m] = pa[12 * h - 96 * m]; av1[m] = pa[12 * h + 4 - 96 * m]; }
#pragma unroll
                    for (int m = 0; m < MTC; ++m) dmma8x8x4(acc[0][m], av0[m], bv0);
#pragma unroll
                    for (int m = 0; m < MTC; ++m) dmma8x8x4(acc[1][m], av1[m], bv1);
                }
                for (int ks = 2 * hmax; ks < nks; ++ks) {    // at most two: a full even k-step and / or the masked tail
                    const int h = ks >> 1, sidx = 4 * ks + fk;
                    const double bv = (ks & 1) ? pb1[12 * h] : pb0[12 * h];
#pragma unroll
                    for (int m = 0; m < MTC; ++m) {
                        const double v = pa[12 * h + 4 * (ks & 1) - 96 * m];
                        const double vm = sidx < M ? v : 0.0;
                        if (ks & 1) dmma8x8x4(acc[1][m], vm, bv); else dmma8x8x4(acc[0][m], vm, bv);
                    }
                }
            }
        }
        __syncwarp();                                        // every lane is done reading the staging area
        // 4. results
        if constexpr (!PARTIAL) {
            if (live) {
                const size_t row_len = static_cast<size_t>(nq) + static_cast<size_t>(nq) * M;
                double* row = out + static_cast<size_t>(b) * row_len;
                double* dst = row + nq + static_cast<size_t>(iq) * M;
#pragma unroll
                for (int m = 0; m < MTC; ++m)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int tau = 64 * m + 8 * fi + 2 * fk + e;
                        if (tau > half) continue;
                        const double val = (acc[0][m][e] + acc[1][m][e]) * invN;
                        dst[tau] = val;
                        if (tau > 0 && tau < M - tau) dst[M - tau] = val;
                        if (tau == 0 && commensurate[iq]) row[iq] = val;
                    }
            }
        } else {
            // the four configurations in fixed order (b0 .. b0+3; dead warps contribute exact zeros).  The parking area is
            // double buffered by item parity: the barrier of item n + 1 orders the adds of item n before the parking of n + 2
            double* pk = park + (parity * 4 + warp) * (half + 1);
#pragma unroll
            for (int m = 0; m < MTC; ++m)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int tau = 64 * m + 8 * fi + 2 * fk + e;
                    if (tau <= half) pk[tau] = (acc[0][m][e] + acc[1][m][e]) * invN;
                }
            __syncthreads();
            const double* p0 = park + (parity * 4) * (half + 1);
            double* row = out + (static_cast<size_t>(bq) * nq + iq) * (half + 1);
#pragma unroll
            for (int r = 0; r < RV; ++r) {
                const int tau = threadIdx.x + 128 * r;
                if (tau <= half)
                    row[tau] = rowv[r] + (((p0[tau] + p0[(half + 1) + tau]) + p0[2 * (half + 1) + tau]) + p0[3 * (half + 1) + tau]);
            }
        }
    }
}

// Folds the persistent quad rows into the bin (fixed row order), mirrors tau -> M - tau, S(q) = F(q,0) for commensurate
// q, and clears the rows.  Runs once per bin read-out, not per measurement.
__global__ void __launch_bounds__(128) bins_fold_kernel(double* __restrict__ rows, double* __restrict__ bins, int nrows, int nq, int M,
                                                         const unsigned char* __restrict__ commensurate) {
    const int half = M / 2, seg = half + 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nq * seg) return;
    const int iq = idx / seg, tau = idx - iq * seg;
    double tot = 0.0;
    for (int r = 0; r < nrows; ++r) {
        double* p = rows + (static_cast<size_t>(r) * nq) * seg + idx;
        tot += *p;
        *p = 0.0;
    }
    double* f = bins + nq + static_cast<size_t>(iq) * M;
    f[tau] += tot;
    if (tau > 0 && tau < M - tau) f[M - tau] += tot;
    if (tau == 0 && commensurate[iq]) bins[iq] += tot;
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

// bins[j] += sum_b cfg[b][j].  kBinLanes adjacent lanes share one element j and sum a contiguous slice of the
// configurations each (b ascending, loads issued together); the slices are then combined by a fixed-order shuffle
// tree, so the result is deterministic for a given batch size.
constexpr int kBinLanes = 16;
__global__ void __launch_bounds__(256) bins_accumulate_kernel(const double* __restrict__ cfg, double* __restrict__ bins, int B, size_t len) {
    const size_t gid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t j = min(gid / kBinLanes, len - 1);
    const int part = static_cast<int>(gid % kBinLanes);
    const int b0 = (B * part) / kBinLanes, b1 = (B * (part + 1)) / kBinLanes;
    double acc = 0.0;
    int b = b0;
    for (; b + 3 < b1; b += 4) {
        const double v0 = __ldg(cfg + static_cast<size_t>(b) * len + j), v1 = __ldg(cfg + static_cast<size_t>(b + 1) * len + j);
        const double v2 = __ldg(cfg + static_cast<size_t>(b + 2) * len + j), v3 = __ldg(cfg + static_cast<size_t>(b + 3) * len + j);
        acc += ((v0 + v1) + v2) + v3;
    }
    for (; b < b1; ++b) acc += __ldg(cfg + static_cast<size_t>(b) * len + j);
#pragma unroll
    for (int o = 1; o < kBinLanes; o <<= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);   // lane 0 of the group: fixed tree
    if (part == 0 && gid / kBinLanes < len) bins[j] += acc;
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

// Order in which a thread walks its ring of partners when it visits ALL of them (force loops: every pair is seen from
// both ends and both visits read the same table entry).  Zig-zag (A/B switch, off): step 2m-1 -> partner i+m, step 2m ->
// partner i-m, so the two visits of a pair are one step apart and the second could find the 32-byte sector in L1 instead
// of going back to L2 / DRAM a whole ring later.  Measured on 64 C2 configurations (gpurun_out r01zz, profiles/
// r01zz_ring_ab.txt): virial kernel 5.33 -> 5.19 ms, pair kernel 3.39 -> 3.76 ms -- the reuse does not materialise (the
// sectors of ~10^3 gathers in flight per SM do not survive in L1) and the scattered partner order costs the pair kernel
// more than it saves; the ascending ring stays.
#ifndef PIMCB_RING_ZIGZAG
#define PIMCB_RING_ZIGZAG 0
#endif
__device__ __forceinline__ int ring_partner(int step, int N, bool full_ring) {
#if PIMCB_RING_ZIGZAG
    if (full_ring) return (step & 1) ? (step + 1) >> 1 : N - (step >> 1);
#endif
    return step;
}

struct PairParams {
    const double* V; const double* dVdr; int len; double dr; double extV[2]; double extdV[2];
    double dSep; int want_hist; int f2_parity; int M;
    const double* gext;    // gradient of the external potential per bead, slice rows like pos ([sl][d][Npad]), or nullptr ("free")
};

#ifndef PIMCB_PAIR_UNROLL
#define PIMCB_PAIR_UNROLL 2
#endif
#ifndef PIMCB_PAIR_MINB
#define PIMCB_PAIR_MINB 4
#endif

// The table gathers are the long pole (46 % of the stall samples of the one-pair-at-a-time version were a warp waiting
// for its single outstanding gather): every thread works on U partners at once -- U separations, U (or 2 U) gathers
// in flight -- and then accumulates them in the original partner order, so the sums are bit-identical to the
// sequential loop.  Measured on 64 C2 configurations (gsf action): U = 1: 4.35 ms, U = 2 (64 registers, 4 CTAs/SM):
// 3.34 ms, U = 3: 3.7 ms, U = 4: 3.8 - 4.9 ms (occupancy).  At U = 2 the kernel moves 6.8 TB/s of 32-byte sectors out
// of L2; interleaving V and dV/dr into one (V, dV/dr) table to halve the sectors of the force slices was tried and is
// SLOWER (3.9 - 4.2 ms): it doubles the footprint of the V-only reads and the L2 hit rate (66 %) pays for it.  A
// persisting-L2 access-policy window on either table (79 MB available) changes nothing (3.37 ms).
template <int ND, bool WANT_F2>
__global__ void __launch_bounds__(256, PIMCB_PAIR_MINB) pair_kernel(const double* __restrict__ pos, int nslices, int N, int Npad, BoxDev box,
                                                    PairParams pp, double* __restrict__ vint, double* __restrict__ f2,
                                                    int* __restrict__ hist) {
    constexpr int U = PIMCB_PAIR_UNROLL;
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
            // V-half: pairs (i, i+kk), kk <= N/2 (for even N and kk == N/2 only i < N/2); with forces the whole ring
            const int kv = ((N & 1) || i < khalf) ? khalf : khalf - 1;       // last kk that contributes to V
            const int klast = do_f2 ? N - 1 : kv;
            for (int kk0 = 1; kk0 <= klast; kk0 += U) {
                double r[U], sep[U][ND], vv[U], dv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int kk = ring_partner(min(kk0 + u, klast), N, do_f2);   // surplus slots recompute the last partner
                    int j = i + kk;
                    if (j >= N) j -= N;
                    if (do_f2) {
                        r[u] = minimage_norm<ND>(xs, Npad, i, j, box, sep[u]);          // getSeparation(bead1,bead2), action.cpp:1211
                    } else {
                        const int lo = min(i, j), hi = max(i, j);
                        r[u] = minimage_norm<ND>(xs, Npad, hi, lo, box, sep[u]);        // getSeparation(bead2,bead1), action.cpp:934
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {                                // all gathers of the group are issued here
                    const int kidx = __double2int_rz(__ddiv_rn(r[u], pp.dr));
                    const bool inside = kidx > 0 && kidx < pp.len;
                    const bool live = kk0 + u <= klast;
                    const bool vhalf = live && ring_partner(kk0 + u, N, do_f2) <= kv;
                    vv[u] = 0.0;
                    dv[u] = 0.0;
                    if (vhalf) vv[u] = inside ? __ldg(pp.V + kidx) : (kidx <= 0 ? pp.extV[0] : pp.extV[1]);
                    if (WANT_F2 && do_f2 && live) dv[u] = inside ? __ldg(pp.dVdr + kidx) : (kidx <= 0 ? pp.extdV[0] : pp.extdV[1]);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool live = kk0 + u <= klast;
                    const bool vhalf = live && ring_partner(kk0 + u, N, do_f2) <= kv;
                    if (vhalf) {
                        vsum += vv[u];
                        if (pp.want_hist) {
                            const int nR = __double2int_rz(__ddiv_rn(r[u], pp.dSep));   // action.cpp:221
                            if (nR >= 0 && nR < kNPCFSEP) atomicAdd(&shist[nR], 1);
                        }
                    }
                    if (WANT_F2 && do_f2 && live) {
                        const double g = __ddiv_rn(dv[u], r[u]);
#pragma unroll
                        for (int d = 0; d < ND; ++d) F[d] = fma(g, sep[u][d], F[d]);
                    }
                }
            }
            if (do_f2) {
                if (pp.gext) {                                   // F += externalPtr->gradV(path(bead1)), action.cpp:1216
                    const double* ge = pp.gext + static_cast<size_t>(sl) * ND * Npad + i;
#pragma unroll
                    for (int d = 0; d < ND; ++d) F[d] += __ldg(ge + d * Npad);
                }
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
