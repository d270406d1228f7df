// kernels_pair.cuh -- pair-potential slice sums, second generation.
//
// What the first-generation kernels (pair_kernel / pair_sym_kernel, kernels.cuh / kernels_ext.cuh) taught
// (profiles/r02b_pair_probe.txt, r02a_gather_peak.txt, r02c_pipe_rates.txt):
//   * shrinking the tables 16x changes their time by 7 % -- they are bound by instruction issue, not by the gathers:
//     ~170 warp-instructions per pair, of which three IEEE divisions, one IEEE square root and three floor() go
//     through multi-instruction sequences on the slow conversion / MUFU pipe, and the force slices add a shared-memory
//     read-add-write per component plus one CTA barrier per ring step;
//   * the gathers themselves are bounded by the rate at which L2 hands out 32-byte sectors: 289 G sectors/s while the
//     footprint stays below ~53 MB, 140 G/s at 106 MB (two 53 MB tables).
// This kernel attacks the instruction side:
//   * TILES instead of a ring.  Particles are cut into groups of 32; a warp owns a home group (lane = particle, force in
//     registers) and meets the other groups one 32 x 32 tile at a time, partner m = (lane + s) & 31 at step s, so every
//     pair is visited exactly once (half ring over GROUPS, half ring inside the diagonal tile).  The partner's share of
//     a pair force travels to its lane with one warp shuffle per component and accumulates there in registers: no
//     shared-memory read-add-writes, no barrier inside a slice (one per round of four group offsets).  Each tile parks
//     the partner-side sums in its own shared-memory slot and a fixed-order fold adds them: results are reproducible.
//   * The table index without a division or a correctly rounded square root.  k = int(r/dr) and the histogram bin
//     int(r/dSep) only need r to a few ulp unless r/dr lies within 2^-20 of an integer: r comes from MUFU.RSQ64H + two
//     Goldschmidt steps (which also yield 1/r for the force), the quotient is a multiplication by the rounded
//     reciprocal, and floor() and the fractional part are read from the bits of t + 1.5 * 2^e.  The rare unsafe cases
//     (|frac - integer| < 2^-20, 2 in 10^6 pairs) and everything outside the fast path's validity are left out of the
//     fast loop and redone after it with the reference's exact operation sequence (minimage_norm / __ddiv_rn; see
//     pair_tile_redo), so k and the histogram stay BIT-IDENTICAL to the CPU (include/potential.h:249-260,
//     src/action.cpp:216-224) -- the tests hold sepHist to array_equal.
//   * The histogram bin from the table index: k = int(r/dr) pins r/dSep to an interval of width dr/dSep = 8e-6 bins, so
//     int(r/dSep) = (k * round(2^S dr/dSep)) >> S (one integer multiply-add pair) unless that product lies within the
//     interval's width of a bin edge (3 in 10^5 pairs), which again takes the exact path.
//   * Table reads without range branches: the device tables carry ext[0] in entry 0 and ext[1] in entry `len`
//     (TabulatedPotential::direct returns ext[0] for k <= 0 and ext[1] for k >= len and never reads entry 0), so the read
//     is table[min(max(k, 0), len)].
//   * Minimum image by rint(): sep = s - pSide * rint(s * sideInv) with rint() as an add/subtract of 1.5 * 2^52 (three
//     FP64 instructions per component instead of five plus an FRND); it differs from Container::putInBC's
//     floor(x + 0.5) (include/container.h:50-59) only on exact ties, where both images are equally near and r is the
//     same to rounding -- covered by the same safety margin.
#pragma once

#include "kernels.cuh"
#include "table_codec.h"

namespace pimcb {

// What both tile kernels need to turn a pair of beads into a table index (and a histogram bin).
struct TileIndexParams {
    int len; double dr, inv_dr;
    double dSep; int want_hist;
    double magic;          // 1.5 * 2^e: ulp(magic) = 2^-fb, table indices < 2^(e-1)
    int fb;                // fraction bits below the integer part in the low word of (t + magic), 21..28
    unsigned hmul_lo, hmul_hi;   // round(2^hshift dr/dSep): bin = (k * hmul) >> hshift
    int hshift;            // 32..56, or 0: no integer shortcut for the bin (every pair takes the exact path for it)
    unsigned hslop, hspan; // bin safe iff (frac32 - hslop) < hspan  (unsigned), frac32 = fraction of k dr/dSep in 2^-32
    double x[4], xh[4], x3[4];   // CodecSteps of dr
};

struct PairTileParams {
    const double* V; const double* dVdr;   // verbatim tables, len + 1 entries: [0] = ext[0], [len] = ext[1]
    TileIndexParams ix;
    int f2_parity; int M;
    const double* gext;    // gradient of the external potential per bead, slice rows like pos ([sl][d][Npad]), or nullptr
    int G;                 // particle groups of 32
    int spc;               // slices per CTA work unit (> 1 when a slice has fewer groups than the CTA has warps)
    const TableSector* VD; // (V, dV/dr) packed four entries per 32-byte sector (table_codec.h), or nullptr: verbatim tables only
    // slice selection: sel_p < 0: the launch covers every slice; sel_p = 0 / 1: only the slices with (t mod 2) == sel_p,
    // sel_cnt of them per configuration; the kernel's `nslices` then counts SELECTED slices.  A gsf call is two launches:
    // the force kernel on the slices that carry gradVSquared, the V-only kernel (fewer registers, more warps) on the rest.
    int sel_p, sel_cnt;
};
__device__ __forceinline__ int pair_slice_of(const PairTileParams& pp, int u) {
    if (pp.sel_p < 0) return u;
    const int b = u / pp.sel_cnt;
    return b * pp.M + 2 * (u - b * pp.sel_cnt) + pp.sel_p;
}

// One whole 32-byte sector per lane in one request (LDG.E.256).
__device__ __forceinline__ TableSector ldg_sector(const TableSector* p) {
    TableSector s;
    asm("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(s.w[0]), "=l"(s.w[1]), "=l"(s.w[2]), "=l"(s.w[3]) : "l"(p));
    return s;
}

constexpr int kPairWarps = 8;
constexpr int kPairRound = 4;        // group offsets per round (partner-side slots in shared memory)
constexpr double kRintMagic = 6755399441055744.0;   // 1.5 * 2^52

// r ~ sqrt(r2) and rinv ~ 1/sqrt(r2), both to ~2 ulp: MUFU.RSQ64H seed + two Goldschmidt iterations.
template <bool WANT_RINV>
__device__ __forceinline__ void rsqrt_pair(double r2, double& r, double& rinv) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r2));
    double g = r2 * y, h = 0.5 * y;
    double e = fma(-h, g, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    e = fma(-h, g, 0.5);
    r = fma(g, e, g);
    rinv = WANT_RINV ? 2.0 * fma(h, e, h) : 0.0;
}

// floor(t) and "t is safely away from an integer", from the bits of u = t + magic (magic = 1.5 * 2^e, fb = 52 - e fraction
// bits).  Returns false when t is out of the trick's range (negative, too large, NaN) or within 2^-20 of an integer.
__device__ __forceinline__ bool floor_safe(double u, double magic, int fb, int& k) {
    const unsigned lo = static_cast<unsigned>(__double2loint(u)), hi = static_cast<unsigned>(__double2hiint(u));
    const unsigned mhi = static_cast<unsigned>(__double2hiint(magic));
    k = static_cast<int>(__funnelshift_r(lo, hi & 0x7ffffu, fb));          // (mantissa below the 1.5 bit) >> fb
    const unsigned frac = lo << (32 - fb);                                  // fraction, left aligned
    // same exponent and the leading mantissa bit of 1.5 still set: u in [magic, magic + 2^(e-1)); fraction in [2^-20, 1 - 2^-20)
    return ((hi ^ mhi) < 0x80000u) && (frac - 0x1000u) < 0xffffe000u;
}

// Geometry of one pair on the fast path: minimum-image separation, r, 1/r, table index and histogram bin.
// `unsafe` = the index or the bin must be recomputed with the reference's exact operation sequence.
template <int ND, bool WANT_RINV>
__device__ __forceinline__ void pair_fast(const double (&xi)[ND], const double* __restrict__ xb, int NP, int m, const BoxDev& box,
                                          const TileIndexParams& ix, double (&sep)[ND], double& r, double& rinv, int& k, int& nR,
                                          bool& unsafe) {
    double r2 = 0.0;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        const double sd = xi[d] - xb[d * NP + m];
        const double tt = fma(sd, box.sideInv[d], kRintMagic) - kRintMagic;
        sep[d] = fma(-box.pSide[d], tt, sd);
        r2 = fma(sep[d], sep[d], r2);
    }
    rsqrt_pair<WANT_RINV>(r2, r, rinv);
    unsafe = !floor_safe(fma(r, ix.inv_dr, ix.magic), ix.magic, ix.fb, k);
    nR = 0;
    if (ix.want_hist) {
        // k dr / dSep in fixed point: 64-bit product of k with round(2^hshift dr/dSep)
        const unsigned long long prod = static_cast<unsigned long long>(static_cast<unsigned>(k)) * ix.hmul_lo +
                                        (static_cast<unsigned long long>(static_cast<unsigned>(k) * ix.hmul_hi) << 32);
        const unsigned phi = static_cast<unsigned>(prod >> 32), plo = static_cast<unsigned>(prod);
        nR = static_cast<int>(phi >> (ix.hshift - 32));
        const unsigned frac = __funnelshift_r(plo, phi, ix.hshift - 32);   // the 32 bits below the bin number
        unsafe = unsafe || !((frac - ix.hslop) < ix.hspan);
    }
}

// Table entries F[k] (and G[k] when WANT_G) with k clamped to [0, len]: one packed sector (CODEC) or the verbatim tables.
template <bool WANT_G, bool CODEC>
__device__ __forceinline__ void table_issue(const TableSector* __restrict__ packed, const double* __restrict__ F, const double* __restrict__ G,
                                            int kc, TableSector& sec, double& f, double& g) {
    if constexpr (CODEC) {
        sec = ldg_sector(packed + (kc >> 2));
    } else {
        f = __ldg(F + kc);
        g = WANT_G ? __ldg(G + kc) : 0.0;
    }
}
template <bool WANT_G, bool CODEC>
__device__ __forceinline__ void table_finish(const TileIndexParams& ix, const double* __restrict__ F, const double* __restrict__ G, int kc,
                                             const TableSector& sec, double& f, double& g) {
    if constexpr (CODEC) {
        const int j = kc & 3;
        if (sector_is_raw(sec)) {
            f = __ldg(F + kc);
            g = WANT_G ? __ldg(G + kc) : 0.0;
        } else {
            sector_decode<WANT_G>(sec, j, ix.x[j], ix.xh[j], ix.x3[j], f, g);
        }
    }
}

// One tile: the 32 particles of the home group (lane = particle i, position xi) against the 32 particles of group b
// (positions xb[d * NP + m]), steps [s_lo, s_hi), partner m = (lane + s) & 31.  FORCE: own-side force into Fi, partner-side
// force into Gv (it ends up in the lane that holds the partner: particle 32 b + lane).  CHECK = false: every pair of the
// tile exists (both groups full, not the diagonal tile) and no validity logic is compiled.
//
// Two passes.  The loop runs the fast path only; a pair whose index or bin is not provably the reference's ("unsafe":
// 1.5 in 10^5 pairs) contributes NOTHING there and sets the bit of its step in the lane's `redo` mask.  After the loop --
// in 1.5 % of the tiles -- pair_tile_redo walks the set bits in a fixed order (lane ascending, step ascending), all lanes
// recomputing the pair with the reference's exact operation sequence so that both the owning lane and the partner's lane
// can add their share: no call, no exact-path registers and no reconvergence bookkeeping inside the loop (the first cut
// called the exact path from inside the loop: +45 instructions per two pairs of call plumbing on the FAST path).
#ifndef PIMCB_PTILE_U
#define PIMCB_PTILE_U 2
#endif
template <int ND, bool FORCE>
__device__ __forceinline__ void pair_tile_redo(unsigned redo, const double* __restrict__ xsl, int NP, int ibase, int b, int lane,
                                               const BoxDev& box, const PairTileParams& pp, int* __restrict__ shist_sl, double& vsum,
                                               double (&Fi)[ND], double (&Gv)[ND]) {
    constexpr unsigned FULL = 0xffffffffu;
    const TileIndexParams& ix = pp.ix;
    unsigned lanes = __ballot_sync(FULL, redo != 0u);
    while (lanes) {
        const int l = __ffs(lanes) - 1;
        lanes &= lanes - 1;
        unsigned steps = __shfl_sync(FULL, redo, l);
        while (steps) {
            const int ss = __ffs(steps) - 1;
            steps &= steps - 1;
            const int m = (l + ss) & 31;
            // the reference's own operation sequence (putInBC -> dot -> sqrt -> r/dr -> int(); r/dSep -> int()), bit for bit
            double sx[ND];
            const double rx = minimage_norm<ND>(xsl, NP, ibase + l, 32 * b + m, box, sx);
            const int k = __double2int_rz(__ddiv_rn(rx, ix.dr));
            const int kc = max(min(k, ix.len), 0);
            const double v = __ldg(pp.V + kc);
            if (lane == l) {
                vsum += v;
                if (ix.want_hist) {
                    const int nR = __double2int_rz(__ddiv_rn(rx, ix.dSep));
                    if (static_cast<unsigned>(nR) < static_cast<unsigned>(kNPCFSEP)) atomicAdd(shist_sl + nR, 1);
                }
            }
            if constexpr (FORCE) {
                const double g = __ddiv_rn(__ldg(pp.dVdr + kc), rx);      // (dV/dr)/r, potential.h:997-1003
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    const double f = g * sx[d];
                    if (lane == l) Fi[d] += f;
                    if (lane == m) Gv[d] -= f;
                }
            }
        }
    }
}

template <int ND, bool FORCE, bool CODEC, bool CHECK>
__device__ __forceinline__ void pair_tile(const double* __restrict__ xsl, int NP, const double (&xi)[ND], int i, bool ivalid, int b,
                                          int lane, int s_lo, int s_hi, bool diag, int N, const BoxDev& box,
                                          const PairTileParams& pp, int* __restrict__ shist_sl, double& vsum, double (&Fi)[ND],
                                          double (&Gv)[ND]) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int U = PIMCB_PTILE_U;         // pairs in flight per lane (tile step counts are multiples of 16)
    const double* xb = xsl + 32 * b;
    const TileIndexParams& ix = pp.ix;
    unsigned redo = 0u;                      // bit ss: the pair of step ss is left to pair_tile_redo
    for (int s = s_lo; s < s_hi; s += U) {
        double sep[U][ND], r[U], rinv[U], vv[U], dv[U];
        int kidx[U], nR[U], kc[U];
        bool ok[U];
        TableSector sec[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int ss = s + u;
            const int m = (lane + ss) & 31;
            bool unsafe;
            pair_fast<ND, FORCE>(xi, xb, NP, m, box, ix, sep[u], r[u], rinv[u], kidx[u], nR[u], unsafe);
            if constexpr (CHECK) {
                const bool valid = ivalid && 32 * b + m < N && !(diag && ss == 16 && lane >= 16);
                unsafe = unsafe && valid;
                ok[u] = valid && !unsafe;
            } else {
                ok[u] = !unsafe;
            }
#ifndef PIMCB_COUNT_FASTPATH         // (static instruction counting of the fast path only: tools/sass_loop.py)
            if (unsafe) redo |= 1u << ss;
#endif
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {           // both table reads of the step pair are issued here
            kc[u] = max(min(kidx[u], ix.len), 0);
            table_issue<FORCE, CODEC>(pp.VD, pp.V, pp.dVdr, kc[u], sec[u], vv[u], dv[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            table_finish<FORCE, CODEC>(ix, pp.V, pp.dVdr, kc[u], sec[u], vv[u], dv[u]);
            vsum += ok[u] ? vv[u] : 0.0;
            if (ix.want_hist && ok[u] && static_cast<unsigned>(nR[u]) < static_cast<unsigned>(kNPCFSEP)) atomicAdd(shist_sl + nR[u], 1);
            if constexpr (FORCE) {
                const double g = ok[u] ? dv[u] * rinv[u] : 0.0;      // (dV/dr)/r, potential.h:997-1003
                const int src = (lane - (s + u)) & 31;  // the lane whose partner this lane is at this step
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    const double f = g * sep[u][d];
                    Fi[d] += f;
                    Gv[d] -= __shfl_sync(FULL, f, src);          // gradV(sep_ji) = -gradV(sep_ij)
                }
            }
        }
    }
#ifndef PIMCB_COUNT_FASTPATH
    if (__any_sync(FULL, redo != 0u)) pair_tile_redo<ND, FORCE>(redo, xsl, NP, i - lane, b, lane, box, pp, shist_sl, vsum, Fi, Gv);
#endif
}

#ifndef PIMCB_PTILE_MINB
#define PIMCB_PTILE_MINB 3
#endif
#ifndef PIMCB_PTILE_MINB_V
#define PIMCB_PTILE_MINB_V 4
#endif
// FK = true: the force-capable kernel (slices selected by pp.f2_parity get gradVSquared).  FK = false: V-only -- no force
// code, no force accumulators in shared memory, 64 registers: four CTAs (32 warps) per SM instead of three.
template <int ND, bool CODEC, bool FK = true>
__global__ void __launch_bounds__(32 * kPairWarps, (FK ? PIMCB_PTILE_MINB : PIMCB_PTILE_MINB_V))
pair_tile_kernel(const double* __restrict__ pos, int nslices, int N, int Npad, BoxDev box, PairTileParams pp,
                 double* __restrict__ vint, double* __restrict__ f2, int* __restrict__ hist) {
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) double sm[];
    const int G = pp.G, spc = pp.spc, NP = 32 * G;
    double* xs = sm;                                            // [spc][ND][NP]
    double* accF = xs + static_cast<size_t>(spc) * ND * NP;     // [spc][ND][NP]     total force per particle (FK only)
    double* part = accF + (FK ? static_cast<size_t>(spc) * ND * NP : 0);   // [kPairRound][spc][ND][NP] partner-side sums per offset (FK only)
    double* redV = part + (FK ? static_cast<size_t>(kPairRound) * spc * ND * NP : 0);   // [spc * G]  per-home V sums
    int* shist = reinterpret_cast<int*>(redV + spc * G + (spc * G & 1));     // [spc][kNPCFSEP]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int omax = G / 2;                                     // largest group offset (a half tile when G is even)
    const bool f2_any = FK && f2 != nullptr;
    const int nunits = (nslices + spc - 1) / spc;

    for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
        const int sl0 = unit * spc;
        const int nsl = min(spc, nslices - sl0);
        // stage the slices of this unit: rows padded to whole groups (the padding is never counted as a particle)
        for (int k = threadIdx.x; k < nsl * ND * NP; k += blockDim.x) {
            const int s = k / (ND * NP), rem = k - s * (ND * NP), d = rem / NP, i = rem - d * NP;
            xs[k] = i < Npad ? __ldg(pos + (static_cast<size_t>(pair_slice_of(pp, sl0 + s)) * ND + d) * Npad + i) : 0.0;
            if (f2_any) accF[k] = 0.0;
        }
        if (pp.ix.want_hist)
            for (int k = threadIdx.x; k < nsl * kNPCFSEP; k += blockDim.x) shist[k] = 0;
        __syncthreads();

        // rounds of kPairRound group offsets (one partner-side slot each); the first round also takes the diagonal tiles
        for (int o1 = 1; o1 == 1 || o1 <= omax; o1 += kPairRound) {
            for (int h = warp; h < nsl * G; h += kPairWarps) {
                const int sloc = h / G, a = h - sloc * G;
                const int sl = pair_slice_of(pp, sl0 + sloc), t = sl % pp.M;
                const bool do_f = f2_any && (pp.f2_parity < 0 || (t & 1) == pp.f2_parity);
                const double* xsl = xs + static_cast<size_t>(sloc) * ND * NP;
                int* shist_sl = shist + sloc * kNPCFSEP;
                const int i = 32 * a + lane;
                const bool ivalid = i < N;
                double xi[ND], Fi[ND];
#pragma unroll
                for (int d = 0; d < ND; ++d) { xi[d] = xsl[d * NP + i]; Fi[d] = 0.0; }
                double vsum = 0.0;
                for (int oi = (o1 == 1 ? -1 : 0); oi < kPairRound; ++oi) {
                    const int o = oi < 0 ? 0 : o1 + oi;         // oi = -1: the diagonal tile
                    if (o > omax) break;
                    int s_lo, s_hi;                             // steps [s_lo, s_hi) of this tile
                    if (o == 0) { s_lo = 1; s_hi = 17; }        // diagonal: distances 1..15 for every lane, 16 for lanes < 16
                    else if (2 * o == G) { s_lo = a < o ? 0 : 1; s_hi = s_lo + 16; }   // the offset shared with the opposite group
                    else { s_lo = 0; s_hi = 32; }
                    int b = a + o;
                    if (b >= G) b -= G;
                    const bool full = o != 0 && 32 * (a + 1) <= N && 32 * (b + 1) <= N;   // every pair of the tile exists
                    double Gv[ND];
#pragma unroll
                    for (int d = 0; d < ND; ++d) Gv[d] = 0.0;
                    if (FK && do_f) {
                        if (full) pair_tile<ND, true, CODEC, false>(xsl, NP, xi, i, ivalid, b, lane, s_lo, s_hi, false, N, box, pp, shist_sl, vsum, Fi, Gv);
                        else pair_tile<ND, true, CODEC, true>(xsl, NP, xi, i, ivalid, b, lane, s_lo, s_hi, o == 0, N, box, pp, shist_sl, vsum, Fi, Gv);
                        if (o == 0) {
#pragma unroll
                            for (int d = 0; d < ND; ++d) Fi[d] += Gv[d];        // the diagonal tile's partners are the home group itself
                        } else {
                            double* slot = part + ((static_cast<size_t>(oi) * spc + sloc) * ND) * NP + 32 * b + lane;
#pragma unroll
                            for (int d = 0; d < ND; ++d) slot[d * NP] = Gv[d];
                        }
                    } else {
                        if (full) pair_tile<ND, false, CODEC, false>(xsl, NP, xi, i, ivalid, b, lane, s_lo, s_hi, false, N, box, pp, shist_sl, vsum, Fi, Gv);
                        else pair_tile<ND, false, CODEC, true>(xsl, NP, xi, i, ivalid, b, lane, s_lo, s_hi, o == 0, N, box, pp, shist_sl, vsum, Fi, Gv);
                    }
                }
                if (FK && do_f) {
#pragma unroll
                    for (int d = 0; d < ND; ++d) accF[(static_cast<size_t>(sloc) * ND + d) * NP + i] += Fi[d];   // this lane is the only writer
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) vsum += __shfl_xor_sync(FULL, vsum, off);
                if (lane == 0) redV[h] = o1 == 1 ? vsum : redV[h] + vsum;
            }
            __syncthreads();
            if (f2_any) {
                // fold the partner-side slots of this round, offsets ascending (fixed order)
                for (int k = threadIdx.x; k < nsl * ND * NP; k += blockDim.x) {
                    const int sloc = k / (ND * NP);
                    const int t = pair_slice_of(pp, sl0 + sloc) % pp.M;
                    if (!(pp.f2_parity < 0 || (t & 1) == pp.f2_parity)) continue;
                    const int rem = k - sloc * (ND * NP);
                    double acc = accF[k];
                    for (int oi = 0; oi < kPairRound && o1 + oi <= omax; ++oi)
                        acc += part[(static_cast<size_t>(oi) * spc + sloc) * ND * NP + rem];
                    accF[k] = acc;
                }
                __syncthreads();
            }
        }
        // per-slice results: Vint = sum of the home sums (home order), sum_i |F_i + gradVext_i|^2
        for (int sloc = warp; sloc < nsl; sloc += kPairWarps) {
            const int sl = pair_slice_of(pp, sl0 + sloc), t = sl % pp.M;
            const bool do_f = f2_any && (pp.f2_parity < 0 || (t & 1) == pp.f2_parity);
            double fsum = 0.0;
            if (FK && do_f) {
                for (int i = lane; i < N; i += 32) {
#pragma unroll
                    for (int d = 0; d < ND; ++d) {
                        double F = accF[(static_cast<size_t>(sloc) * ND + d) * NP + i];
                        if (pp.gext) F += __ldg(pp.gext + (static_cast<size_t>(sl) * ND + d) * Npad + i);      // action.cpp:1216
                        fsum = fma(F, F, fsum);
                    }
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) fsum += __shfl_xor_sync(FULL, fsum, off);
            if (lane == 0) {
                double v = 0.0;
                for (int a = 0; a < G; ++a) v += redV[sloc * G + a];
                vint[sl] = v;
                if (f2) f2[sl] = fsum;
            }
            if (pp.ix.want_hist)
                for (int k = lane; k < kNPCFSEP; k += 32) hist[static_cast<size_t>(sl) * kNPCFSEP + k] = shist[sloc * kNPCFSEP + k];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Virial slice sums on the same tiles (see kernels_ext.cuh, virial_kernel, for the definitions and the upstream lines):
//   gV_i = sum_j (dVdr[k]/r) sep_ij,   T_i = sum_j [ sep sep^T (d2V[k]/r^2 - dV/r^3) + 1 dV/r ],  dV = |dVdr[k]|,
//   out[sl] = { sum_i gV_i.r_i, sum_i (T_i gV_i).r_i, sum_i gV_i.delta_i, sum_i (T_i gV_i).delta_i }.
// The pair terms are antisymmetric (gV) / symmetric (T): the partner's share travels to its lane by warp shuffle and
// accumulates there in registers (ND + ND (ND + 1) / 2 components), exactly as the pair force does in pair_tile_kernel.
// dV = sqrt(gVi.gVi) upstream (src/action.cpp:1544) is |dVdr[k]| up to rounding ((|dVdr|/r) |sep| with |sep| = r), which is
// what is used here; external potential "free" (the non-free case runs virial_kernel<ND, true>).
// ---------------------------------------------------------------------------------------------
struct VirialTileParams {
    const double* dVdr; const double* d2V;     // verbatim tables, len + 1 entries: [0] = ext[0], [len] = ext[1]
    TileIndexParams ix;                          // want_hist = 0
    int t2_parity; int M;
    int G; int spc; int rounds;                  // rounds = group offsets per round (partner-side slots that fit)
    const TableSector* DD;                       // (dV/dr, d2V/dr2) packed four entries per sector, or nullptr
    int sel_p, sel_cnt;                          // slice selection, as in PairTileParams: a gsf call is two launches
};
__device__ __forceinline__ int virial_slice_of(const VirialTileParams& vp, int u) {
    if (vp.sel_p < 0) return u;
    const int b = u / vp.sel_cnt;
    return b * vp.M + 2 * (u - b * vp.sel_cnt) + vp.sel_p;
}

#ifndef PIMCB_VTILE_U
#define PIMCB_VTILE_U 2
#endif
// own / vis += the gV and T-matrix terms of one pair with separation sep, r, 1/r, table values dv, d2
template <int ND, bool T2>
__device__ __forceinline__ void virial_pair_terms(const double (&sep)[ND], double rinv, double dv, double d2, double (&gi)[ND],
                                                  double (&mm)[ND * (ND + 1) / 2]) {
    const double g = dv * rinv;
#pragma unroll
    for (int d = 0; d < ND; ++d) gi[d] = g * sep[d];
    if constexpr (T2) {
        const double dV = fabs(dv);                    // |(dV/dr / r) sep| = |dV/dr| (src/action.cpp:1544)
        const double ri2 = rinv * rinv;
        const double diagv = dV * rinv;
        const double a = fma(d2, ri2, -diagv * ri2);   // d2V/r^2 - dV/r^3
        int k = 0;
#pragma unroll
        for (int p = 0; p < ND; ++p)
#pragma unroll
            for (int q = p; q < ND; ++q, ++k) mm[k] = fma(sep[p] * sep[q], a, p == q ? diagv : 0.0);
    }
}

template <int ND, bool T2>
__device__ __forceinline__ void virial_tile_redo(unsigned redo, const double* __restrict__ xsl, int NP, int ibase, int b, int lane,
                                                 const BoxDev& box, const VirialTileParams& vp,
                                                 double (&own)[ND + ND * (ND + 1) / 2], double (&vis)[ND + ND * (ND + 1) / 2]) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int NT = ND * (ND + 1) / 2;
    const TileIndexParams& ix = vp.ix;
    unsigned lanes = __ballot_sync(FULL, redo != 0u);
    while (lanes) {
        const int l = __ffs(lanes) - 1;
        lanes &= lanes - 1;
        unsigned steps = __shfl_sync(FULL, redo, l);
        while (steps) {
            const int ss = __ffs(steps) - 1;
            steps &= steps - 1;
            const int m = (l + ss) & 31;
            double sx[ND], gi[ND], mm[NT];
            const double rx = minimage_norm<ND>(xsl, NP, ibase + l, 32 * b + m, box, sx);     // the reference's exact sequence
            const int kc = max(min(__double2int_rz(__ddiv_rn(rx, ix.dr)), ix.len), 0);
            const double dv = __ldg(vp.dVdr + kc), d2 = T2 ? __ldg(vp.d2V + kc) : 0.0;
            virial_pair_terms<ND, T2>(sx, 1.0 / rx, dv, d2, gi, mm);
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                if (lane == l) own[d] += gi[d];
                if (lane == m) vis[d] -= gi[d];
            }
            if constexpr (T2) {
#pragma unroll
                for (int k = 0; k < NT; ++k) {
                    if (lane == l) own[ND + k] += mm[k];
                    if (lane == m) vis[ND + k] += mm[k];
                }
            }
        }
    }
}

template <int ND, bool T2, bool CODEC, bool CHECK>
__device__ __forceinline__ void virial_tile(const double* __restrict__ xsl, int NP, const double (&xi)[ND], int i, bool ivalid, int b,
                                            int lane, int s_lo, int s_hi, bool diag, int N, const BoxDev& box,
                                            const VirialTileParams& vp, double (&own)[ND + ND * (ND + 1) / 2],
                                            double (&vis)[ND + ND * (ND + 1) / 2]) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int NT = ND * (ND + 1) / 2;
    constexpr int U = PIMCB_VTILE_U;
    const double* xb = xsl + 32 * b;
    const TileIndexParams& ix = vp.ix;
    unsigned redo = 0u;                      // see pair_tile
    for (int s = s_lo; s < s_hi; s += U) {
        double sep[U][ND], r[U], rinv[U], dv[U], d2[U];
        int kidx[U], nR[U], kc[U];
        bool ok[U];
        TableSector sec[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int ss = s + u;
            const int m = (lane + ss) & 31;
            bool unsafe;
            pair_fast<ND, true>(xi, xb, NP, m, box, ix, sep[u], r[u], rinv[u], kidx[u], nR[u], unsafe);
            if constexpr (CHECK) {
                const bool valid = ivalid && 32 * b + m < N && !(diag && ss == 16 && lane >= 16);
                unsafe = unsafe && valid;
                ok[u] = valid && !unsafe;
            } else {
                ok[u] = !unsafe;
            }
            if (unsafe) redo |= 1u << ss;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            kc[u] = max(min(kidx[u], ix.len), 0);
            table_issue<T2, CODEC>(vp.DD, vp.dVdr, vp.d2V, kc[u], sec[u], dv[u], d2[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            table_finish<T2, CODEC>(ix, vp.dVdr, vp.d2V, kc[u], sec[u], dv[u], d2[u]);
            dv[u] = ok[u] ? dv[u] : 0.0;     // pairs that do not exist, or are left to virial_tile_redo, add exact zeros
            d2[u] = ok[u] ? d2[u] : 0.0;
            rinv[u] = ok[u] ? rinv[u] : 0.0;
            const int src = (lane - (s + u)) & 31;
            double gi[ND], mm[NT];
            virial_pair_terms<ND, T2>(sep[u], rinv[u], dv[u], d2[u], gi, mm);
#pragma unroll
            for (int d = 0; d < ND; ++d) {
                own[d] += gi[d];
                vis[d] -= __shfl_sync(FULL, gi[d], src);       // gradV(sep_ji) = -gradV(sep_ij)
            }
            if constexpr (T2) {
#pragma unroll
                for (int k = 0; k < NT; ++k) {
                    own[ND + k] += mm[k];
                    vis[ND + k] += __shfl_sync(FULL, mm[k], src);   // the T-matrix term is the same seen from either end
                }
            }
        }
    }
    if (__any_sync(FULL, redo != 0u)) virial_tile_redo<ND, T2>(redo, xsl, NP, i - lane, b, lane, box, vp, own, vis);
}

#ifndef PIMCB_VTILE_MINB
#define PIMCB_VTILE_MINB 2
#endif
#ifndef PIMCB_VTILE_MINB_G
#define PIMCB_VTILE_MINB_G 3
#endif
// T2K = true: the kernel that can carry the T-matrix terms (slices selected by vp.t2_parity).  T2K = false: gV terms only --
// ND components per particle instead of ND + ND (ND + 1) / 2 in registers, shared memory and shuffles, no d2V/dr2.
// NCK = components actually kept; the per-pair code keeps its full-size arrays, whose unused tail the compiler drops.
template <int ND, bool CODEC, bool T2K = true>
__global__ void __launch_bounds__(32 * kPairWarps, (T2K ? PIMCB_VTILE_MINB : PIMCB_VTILE_MINB_G))
virial_tile_kernel(const double* __restrict__ pos, const double* __restrict__ delta, int nslices, int N, int Npad, BoxDev box,
                   VirialTileParams vp, double* __restrict__ out) {
    constexpr int NT = ND * (ND + 1) / 2, NC = ND + NT, NCK = T2K ? NC : ND;
    extern __shared__ __align__(16) double sm[];
    const int G = vp.G, spc = vp.spc, NP = 32 * G, R = vp.rounds;
    double* xs = sm;                                            // [spc][ND][NP]
    double* acc = xs + static_cast<size_t>(spc) * ND * NP;      // [spc][NCK][NP]  gV then the upper triangle of T per particle
    double* part = acc + static_cast<size_t>(spc) * NCK * NP;   // [R][spc][NCK][NP] partner-side sums per offset
    double* red = part + static_cast<size_t>(R) * spc * NCK * NP;            // [spc][4][kPairWarps]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int omax = G / 2;
    const int nunits = (nslices + spc - 1) / spc;

    for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
        const int sl0 = unit * spc;
        const int nsl = min(spc, nslices - sl0);
        for (int k = threadIdx.x; k < nsl * ND * NP; k += blockDim.x) {
            const int s = k / (ND * NP), rem = k - s * (ND * NP), d = rem / NP, i = rem - d * NP;
            xs[k] = i < Npad ? __ldg(pos + (static_cast<size_t>(virial_slice_of(vp, sl0 + s)) * ND + d) * Npad + i) : 0.0;
        }
        for (int k = threadIdx.x; k < nsl * NCK * NP; k += blockDim.x) acc[k] = 0.0;
        __syncthreads();

        for (int o1 = 1; o1 == 1 || o1 <= omax; o1 += R) {
            for (int h = warp; h < nsl * G; h += kPairWarps) {
                const int sloc = h / G, a = h - sloc * G;
                const int t = virial_slice_of(vp, sl0 + sloc) % vp.M;
                const bool do_t2 = T2K && (vp.t2_parity == -1 || (vp.t2_parity >= 0 && (t & 1) == vp.t2_parity));
                const double* xsl = xs + static_cast<size_t>(sloc) * ND * NP;
                const int i = 32 * a + lane;
                const bool ivalid = i < N;
                double xi[ND], own[NC];
#pragma unroll
                for (int d = 0; d < ND; ++d) xi[d] = xsl[d * NP + i];
#pragma unroll
                for (int c = 0; c < NC; ++c) own[c] = 0.0;
                for (int oi = (o1 == 1 ? -1 : 0); oi < R; ++oi) {
                    const int o = oi < 0 ? 0 : o1 + oi;
                    if (o > omax) break;
                    int s_lo, s_hi;
                    if (o == 0) { s_lo = 1; s_hi = 17; }
                    else if (2 * o == G) { s_lo = a < o ? 0 : 1; s_hi = s_lo + 16; }
                    else { s_lo = 0; s_hi = 32; }
                    int b = a + o;
                    if (b >= G) b -= G;
                    double vis[NC];
#pragma unroll
                    for (int c = 0; c < NC; ++c) vis[c] = 0.0;
                    const bool full = o != 0 && 32 * (a + 1) <= N && 32 * (b + 1) <= N;   // every pair of the tile exists
                    if (T2K && do_t2) {
                        if (full) virial_tile<ND, true, CODEC, false>(xsl, NP, xi, i, ivalid, b, lane, s_lo, s_hi, false, N, box, vp, own, vis);
                        else virial_tile<ND, true, CODEC, true>(xsl, NP, xi, i, ivalid, b, lane, s_lo, s_hi, o == 0, N, box, vp, own, vis);
                    } else {
                        if (full) virial_tile<ND, false, CODEC, false>(xsl, NP, xi, i, ivalid, b, lane, s_lo, s_hi, false, N, box, vp, own, vis);
                        else virial_tile<ND, false, CODEC, true>(xsl, NP, xi, i, ivalid, b, lane, s_lo, s_hi, o == 0, N, box, vp, own, vis);
                    }
                    if (o == 0) {
#pragma unroll
                        for (int c = 0; c < NCK; ++c) own[c] += vis[c];
                    } else {
                        double* slot = part + ((static_cast<size_t>(oi) * spc + sloc) * NCK) * NP + 32 * b + lane;
#pragma unroll
                        for (int c = 0; c < NCK; ++c) slot[c * NP] = vis[c];
                    }
                }
#pragma unroll
                for (int c = 0; c < NCK; ++c) acc[(static_cast<size_t>(sloc) * NCK + c) * NP + i] += own[c];   // this lane is the only writer
            }
            __syncthreads();
            for (int k = threadIdx.x; k < nsl * NCK * NP; k += blockDim.x) {       // fold this round's slots, offsets ascending
                const int sloc = k / (NCK * NP), rem = k - sloc * (NCK * NP);
                double v = acc[k];
                for (int oi = 0; oi < R && o1 + oi <= omax; ++oi) v += part[(static_cast<size_t>(oi) * spc + sloc) * NCK * NP + rem];
                acc[k] = v;
            }
            __syncthreads();
        }
        // per slice: sum_i gV.w and (T gV).w for w = r (raw position) and w = delta
        for (int sloc = 0; sloc < nsl; ++sloc) {
            const int sl = virial_slice_of(vp, sl0 + sloc);
            const double* A = acc + static_cast<size_t>(sloc) * NCK * NP;
            const double* xsl = xs + static_cast<size_t>(sloc) * ND * NP;
            double sums[4] = {0.0, 0.0, 0.0, 0.0};
            for (int i = threadIdx.x; i < N; i += blockDim.x) {
                double gV[ND], T[NT], uu[ND];
#pragma unroll
                for (int d = 0; d < ND; ++d) { gV[d] = A[d * NP + i]; uu[d] = 0.0; }
#pragma unroll
                for (int k = 0; k < NT; ++k) T[k] = T2K ? A[(ND + k) * NP + i] : 0.0;
                int k = 0;
#pragma unroll
                for (int p = 0; p < ND; ++p)
#pragma unroll
                    for (int q = p; q < ND; ++q, ++k) {
                        uu[p] = fma(T[k], gV[q], uu[p]);
                        if (q != p) uu[q] = fma(T[k], gV[p], uu[q]);
                    }
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    const double x = xsl[d * NP + i];
                    sums[0] = fma(gV[d], x, sums[0]);
                    sums[1] = fma(uu[d], x, sums[1]);
                    if (delta) {
                        const double dl = __ldg(delta + (static_cast<size_t>(sl) * ND + d) * Npad + i);
                        sums[2] = fma(gV[d], dl, sums[2]);
                        sums[3] = fma(uu[d], dl, sums[3]);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double v = sums[k];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                if (lane == 0) red[(sloc * 4 + k) * kPairWarps + warp] = v;
            }
        }
        __syncthreads();
        for (int k = threadIdx.x; k < nsl * 4; k += blockDim.x) {
            double v = 0.0;
            for (int w = 0; w < kPairWarps; ++w) v += red[k * kPairWarps + w];
            out[static_cast<size_t>(virial_slice_of(vp, sl0 + k / 4)) * 4 + (k & 3)] = v;
        }
        __syncthreads();
    }
}

}  // namespace pimcb
