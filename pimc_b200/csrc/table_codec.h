// table_codec.h -- lossless 4-bytes-per-entry encoding of a (function, derivative) pair of lookup tables.
//
// Why.  Every pair of beads costs one read of lookupV[k] (and lookupdVdr[k] on the force slices) at an index that is
// uncorrelated with its neighbours': dr = 1e-6 rm, so one 32-byte sector spans 1.2e-5 Angstrom of separation.  What bounds
// those reads is the RATE at which L2 hands out sectors (tools/micro/gather_peak.cu, profiles/r02a_gather_peak.txt): 289 G
// sectors/s while the footprint stays below ~53 MB, 140 G/s at 106 MB (the C2 tables: 2 x 53 MB), 73 G/s from DRAM.  The
// verbatim tables therefore cost TWO slow reads per pair on the force slices (ncu, profiles/r02e_*: L2 hit rate 66 %,
// 4.9 GB of DRAM reads per launch).  This encoding packs FOUR consecutive entries of BOTH tables into ONE sector (both
// tables of C2: 53 MB in total; L2 hit rate 98 %), so that a pair costs one fast read -- and it reproduces every entry
// BIT FOR BIT, so the device still returns exactly the numbers TabulatedPotential::direct (include/potential.h:249-260)
// would.
//
// How.  For the entries k0 = 4 s .. k0 + 3 of the tables F and G, where G is (numerically) dF/dr:
//     w[0]         F[k0]                       (verbatim)
//     w[1]         G[k0]                       (verbatim)
//     w[2]         C  = dG/dr at k0: top 48 bits of the double | resG[2] << 8 | resG[1]
//     w[3] low     C2 = d2G/dr2 at k0: the HIGH word of the double (sign, exponent, 20 mantissa bits)
//     w[3] high    resF[1] | resF[2] << 8 | resF[3] << 16 | resG[3] << 24
// and entry j = 1..3 is the Taylor prediction, evaluated with the SAME correctly rounded IEEE operations on the host
// (encoder, verifier) and on the device (decoder), with x = j dr, xh = x / 2, x3 = x / 3 taken from a 3-entry table that
// the host computes once (CodecSteps),
//     Gp = fma(x, fma(xh, C2, C), G0)                          Fp = fma(x, fma(xh, fma(x3, C2, C), G0), F0)
// plus a signed 8-bit correction added to the BIT PATTERN (res = bits(actual) - bits(predicted)); j = 0 is verbatim.
// The encoder fits C and C2 to the sector's own G values and accepts the sector only if all six corrections fit; a
// sector that does not (zero crossings of F or G, the r -> 0 core, the switch of the Aziz damping function, tables whose
// second table is not the derivative of the first) is marked RAW (C2 = NaN) and the decoder reads the verbatim tables,
// which stay in HBM.  pimcb_set_pair_table verifies the whole encoding against the verbatim tables ON THE DEVICE before
// using it.  Aziz 1979, C2 box: 0.4 % RAW sectors for (V, dV/dr), 0.6 % for (dV/dr, d2V/dr2).
#ifndef PIMCB_TABLE_CODEC_H
#define PIMCB_TABLE_CODEC_H

#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define PIMCB_HD __host__ __device__ __forceinline__
#else
#define PIMCB_HD inline
#endif

namespace pimcb {

struct alignas(32) TableSector { uint64_t w[4]; };

constexpr uint32_t kRawSectorC2 = 0x7ff80000u;      // high word of a quiet NaN: the sector is not encoded

// x = j dr, x / 2, x / 3 for j = 0..3, computed ONCE on the host and handed to encoder, verifier and decoder alike.
struct CodecSteps { double x[4], xh[4], x3[4]; };
inline CodecSteps codec_steps(double dr) {
    CodecSteps s;
    for (int j = 0; j < 4; ++j) {
        volatile double x = static_cast<double>(j) * dr;     // one rounded multiplication / division each, never contracted
        volatile double xh = x * 0.5, x3 = x / 3.0;
        s.x[j] = x; s.xh[j] = xh; s.x3[j] = x3;
    }
    return s;
}

PIMCB_HD double codec_from_bits(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(static_cast<long long>(b));
#else
    double d; std::memcpy(&d, &b, 8); return d;
#endif
}
PIMCB_HD uint64_t codec_to_bits(double d) {
#if defined(__CUDA_ARCH__)
    return static_cast<uint64_t>(__double_as_longlong(d));
#else
    uint64_t b; std::memcpy(&b, &d, 8); return b;
#endif
}
PIMCB_HD double codec_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}

PIMCB_HD bool sector_is_raw(const TableSector& s) { return static_cast<uint32_t>(s.w[3]) == kRawSectorC2; }

PIMCB_HD double sector_C(const TableSector& s) { return codec_from_bits(s.w[2] & ~0xffffull); }
PIMCB_HD double sector_C2(const TableSector& s) { return codec_from_bits(static_cast<uint64_t>(static_cast<uint32_t>(s.w[3])) << 32); }
PIMCB_HD double sector_predict_G(double x, double xh, double G0, double C, double C2) {
    return codec_fma(x, codec_fma(xh, C2, C), G0);
}
PIMCB_HD double sector_predict_F(double x, double xh, double x3, double F0, double G0, double C, double C2) {
    return codec_fma(x, codec_fma(xh, codec_fma(x3, C2, C), G0), F0);
}
// corrections of entry j (0..3; j = 0 has none)
PIMCB_HD int sector_res_F(const TableSector& s, int j) {
    const uint32_t hi = static_cast<uint32_t>(s.w[3] >> 32);
    return j ? static_cast<int>(static_cast<int8_t>((hi >> (8 * (j - 1))) & 0xff)) : 0;
}
PIMCB_HD int sector_res_G(const TableSector& s, int j) {
    const uint32_t word = j == 3 ? static_cast<uint32_t>(s.w[3] >> 56) : (static_cast<uint32_t>(s.w[2]) >> (8 * ((j - 1) & 1)));
    return j ? static_cast<int>(static_cast<int8_t>(word & 0xff)) : 0;
}

// Entry j (0..3) of table F (and of table G when WANT_G) out of an ENCODED sector; x / xh / x3 = the CodecSteps entries of j.
// Bit-exact by construction (verified at upload).  For j = 0 the prediction is fma(0, ., F0) = F0 and the correction is 0.
template <bool WANT_G>
PIMCB_HD void sector_decode(const TableSector& s, int j, double x, double xh, double x3, double& F, double& G) {
    const double F0 = codec_from_bits(s.w[0]), G0 = codec_from_bits(s.w[1]), C = sector_C(s), C2 = sector_C2(s);
    F = codec_from_bits(codec_to_bits(sector_predict_F(x, xh, x3, F0, G0, C, C2)) + static_cast<uint64_t>(static_cast<int64_t>(sector_res_F(s, j))));
    if (WANT_G) G = codec_from_bits(codec_to_bits(sector_predict_G(x, xh, G0, C, C2)) + static_cast<uint64_t>(static_cast<int64_t>(sector_res_G(s, j))));
    else G = 0.0;
}

// ---- encoder (host) ---------------------------------------------------------------------------------------------------
// Encodes the sector of entries k0 .. k0+3 (entries beyond `len` are treated as absent: the sector is RAW).  Returns
// true when the sector was encoded, false when it was marked RAW.
inline bool sector_encode(const double* F, const double* G, int len, int k0, double dr, const CodecSteps& st, TableSector& out) {
    auto raw = [&]() {
        out.w[0] = k0 < len ? codec_to_bits(F[k0]) : 0;
        out.w[1] = k0 < len ? codec_to_bits(G[k0]) : 0;
        out.w[2] = 0;
        out.w[3] = kRawSectorC2;
        return false;
    };
    if (k0 + 3 >= len) return raw();
    const double f[4] = {F[k0], F[k0 + 1], F[k0 + 2], F[k0 + 3]}, g[4] = {G[k0], G[k0 + 1], G[k0 + 2], G[k0 + 3]};
    for (int j = 0; j < 4; ++j)
        if (!std::isfinite(f[j]) || !std::isfinite(g[j])) return raw();
    // fit G_j ~ G0 + x C + x^2 C2 / 2 through the sector's own values: C2 from the second difference over the widest
    // stencil (kept to the 20 mantissa bits of a double's high word), C from the end points
    const double c2 = (g[3] - g[2] - g[1] + g[0]) / (2.0 * dr * dr);
    if (!std::isfinite(c2)) return raw();
    const uint32_t c2hi = static_cast<uint32_t>((codec_to_bits(c2) + 0x80000000ull) >> 32);     // rounded to the high word
    if ((c2hi & 0x7ff00000u) == 0x7ff00000u) return raw();
    const double C2 = codec_from_bits(static_cast<uint64_t>(c2hi) << 32);
    const double c_fit = (g[3] - g[0]) / (3.0 * dr) - 1.5 * dr * C2;
    if (!std::isfinite(c_fit)) return raw();
    const uint64_t cb0 = codec_to_bits(c_fit) & ~0xffffull;
    int best_cost = 1 << 30;
    uint64_t best_w2 = 0, best_w3 = 0;
    for (int dc = -2; dc <= 2; ++dc) {                          // neighbouring 48-bit values of C: keep the best
        const uint64_t cb = cb0 + static_cast<uint64_t>(static_cast<int64_t>(dc) * 0x10000ll);
        const double C = codec_from_bits(cb);
        if (!std::isfinite(C)) continue;
        int64_t rF[4] = {0, 0, 0, 0}, rG[4] = {0, 0, 0, 0};
        int cost = 0;
        bool ok = true;
        for (int j = 1; j < 4 && ok; ++j) {
            const double Fp = sector_predict_F(st.x[j], st.xh[j], st.x3[j], f[0], g[0], C, C2);
            const double Gp = sector_predict_G(st.x[j], st.xh[j], g[0], C, C2);
            // corrections act on the bit pattern: predicted and actual value must share the sign
            if (std::signbit(Fp) != std::signbit(f[j]) || std::signbit(Gp) != std::signbit(g[j])) { ok = false; break; }
            rF[j] = static_cast<int64_t>(codec_to_bits(f[j]) - codec_to_bits(Fp));
            rG[j] = static_cast<int64_t>(codec_to_bits(g[j]) - codec_to_bits(Gp));
            if (rF[j] < -128 || rF[j] > 127 || rG[j] < -128 || rG[j] > 127) { ok = false; break; }
            const int a = static_cast<int>(rF[j] < 0 ? -rF[j] : rF[j]), b = static_cast<int>(rG[j] < 0 ? -rG[j] : rG[j]);
            cost = cost > a ? cost : a;
            cost = cost > b ? cost : b;
        }
        if (!ok || cost >= best_cost) continue;
        best_cost = cost;
        auto u8 = [](int64_t v) { return static_cast<uint64_t>(static_cast<uint8_t>(v)); };
        best_w2 = cb | (u8(rG[2]) << 8) | u8(rG[1]);
        best_w3 = static_cast<uint64_t>(c2hi) | (u8(rF[1]) << 32) | (u8(rF[2]) << 40) | (u8(rF[3]) << 48) | (u8(rG[3]) << 56);
    }
    if (best_cost == (1 << 30)) return raw();
    out.w[0] = codec_to_bits(f[0]);
    out.w[1] = codec_to_bits(g[0]);
    out.w[2] = best_w2;
    out.w[3] = best_w3;
    if (sector_is_raw(out)) return raw();
    // the decoder must reproduce every entry exactly -- checked here with the decoder itself
    for (int j = 0; j < 4; ++j) {
        double Fd, Gd;
        sector_decode<true>(out, j, st.x[j], st.xh[j], st.x3[j], Fd, Gd);
        if (codec_to_bits(Fd) != codec_to_bits(f[j]) || codec_to_bits(Gd) != codec_to_bits(g[j])) return raw();
    }
    return true;
}

}  // namespace pimcb

#endif
