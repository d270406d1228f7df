// table_codec.h -- lossless 4-bytes-per-entry encoding of a (function, derivative) pair of lookup tables.
//
// Why.  Every pair of beads costs one read of lookupV[k] (and lookupdVdr[k] on the force slices) at an index that is
// uncorrelated with its neighbours': dr = 1e-6 rm, so one 32-byte sector spans 1.2e-5 Angstrom of separation.  What bounds
// those reads is the RATE at which L2 hands out sectors (tools/micro/gather_peak.cu, profiles/r02a_gather_peak.txt): 289 G
// sectors/s while the footprint stays below ~53 MB, 140 G/s at 106 MB (the C2 tables: 2 x 53 MB), 73 G/s from DRAM.  The
// verbatim tables therefore cost TWO slow reads per pair on the force slices.  This encoding packs FOUR consecutive
// entries of BOTH tables into ONE sector (both tables of C2: 53 MB in total), so that a pair costs one fast read -- and
// it reproduces every entry BIT FOR BIT, so the device still returns exactly the numbers TabulatedPotential::direct
// (include/potential.h:249-260) would.
//
// How.  For the entries k0 = 4 s .. k0 + 3 of the tables F and G, where G is (numerically) dF/dr:
//     bytes  0.. 7   F[k0]                       (verbatim)
//     bytes  8..15   G[k0]                       (verbatim)
//     bytes 16..23   C  = dG/dr at k0, top 48 bits of the double | resF[2] << 8 | resF[1]
//     bytes 24..31   C2 = d2G/dr2 as a float | resF[3] << 32 | resG[1] << 40 | resG[2] << 48 | resG[3] << 56
// and entry j = 1..3 (x = j dr) is the Taylor prediction, evaluated with the SAME correctly rounded IEEE operations on
// the host (encoder, verifier) and on the device (decoder),
//     Gp = fma(x, fma(x, 0.5 C2, C), G0)                      Fp = fma(x, fma(x, fma(x, C2/6, 0.5 C), G0), F0)
// plus a signed 8-bit correction added to the BIT PATTERN (res = bits(actual) - bits(predicted)).  The encoder fits
// C and C2 to the sector's own G values and accepts the sector only if all six corrections fit; a sector that does not
// (zero crossings of F or G, the r -> 0 core, the switch of the Aziz damping function, tables whose second table is
// not the derivative of the first) is marked RAW (C2 = NaN) and the decoder reads the verbatim tables, which stay in
// HBM.  pimcb_set_pair_table verifies the whole encoding against the verbatim tables on the device before using it.
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

constexpr uint32_t kRawSectorC2 = 0x7fc00000u;      // quiet NaN: the sector is not encoded, read the verbatim tables

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
PIMCB_HD double codec_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;          // one correctly rounded multiplication, never contracted
    return r;
#endif
}
PIMCB_HD double codec_float_bits_to_double(uint32_t fb) {
#if defined(__CUDA_ARCH__)
    return static_cast<double>(__uint_as_float(fb));
#else
    float f; std::memcpy(&f, &fb, 4); return static_cast<double>(f);
#endif
}

PIMCB_HD bool sector_is_raw(const TableSector& s) { return static_cast<uint32_t>(s.w[3]) == kRawSectorC2; }

// Predictions of entry j (1..3) of an encoded sector; x = j * dr as computed by codec_x().
PIMCB_HD double codec_x(int j, double dr) { return codec_mul(static_cast<double>(j), dr); }

PIMCB_HD void sector_coeffs(const TableSector& s, double& F0, double& G0, double& C, double& C2) {
    F0 = codec_from_bits(s.w[0]);
    G0 = codec_from_bits(s.w[1]);
    C = codec_from_bits(s.w[2] & ~0xffffull);
    C2 = codec_float_bits_to_double(static_cast<uint32_t>(s.w[3]));
}
PIMCB_HD double sector_predict_G(double x, double G0, double C, double C2) {
    return codec_fma(x, codec_fma(x, codec_mul(0.5, C2), C), G0);
}
PIMCB_HD double sector_predict_F(double x, double F0, double G0, double C, double C2) {
    return codec_fma(x, codec_fma(x, codec_fma(x, codec_mul(C2, 1.0 / 6.0), codec_mul(0.5, C)), G0), F0);
}
PIMCB_HD int sector_res_F(const TableSector& s, int j) {      // j = 1..3
    const uint64_t v = j == 3 ? (s.w[3] >> 32) : (s.w[2] >> (8 * (j - 1)));
    return static_cast<int>(static_cast<int8_t>(v & 0xff));
}
PIMCB_HD int sector_res_G(const TableSector& s, int j) {      // j = 1..3
    return static_cast<int>(static_cast<int8_t>((s.w[3] >> (32 + 8 * j)) & 0xff));
}

// Entry k of table F (and of table G when WANT_G) out of an ENCODED sector.  Bit-exact by construction (verified at upload).
template <bool WANT_G>
PIMCB_HD void sector_decode(const TableSector& s, int j, double dr, double& F, double& G) {
    double F0, G0, C, C2;
    sector_coeffs(s, F0, G0, C, C2);
    if (j == 0) { F = F0; G = G0; return; }
    const double x = codec_x(j, dr);
    F = codec_from_bits(codec_to_bits(sector_predict_F(x, F0, G0, C, C2)) + static_cast<uint64_t>(static_cast<int64_t>(sector_res_F(s, j))));
    if (WANT_G) G = codec_from_bits(codec_to_bits(sector_predict_G(x, G0, C, C2)) + static_cast<uint64_t>(static_cast<int64_t>(sector_res_G(s, j))));
    else G = 0.0;
}

// ---- encoder (host) ---------------------------------------------------------------------------------------------------
// Encodes the sector of entries k0 .. k0+3 (entries beyond `len` are treated as absent: the sector is RAW).  Returns
// true when the sector was encoded, false when it was marked RAW.
inline bool sector_encode(const double* F, const double* G, int len, int k0, double dr, TableSector& out) {
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
    // stencil, C from the end points
    const double c2 = (g[3] - g[2] - g[1] + g[0]) / (2.0 * dr * dr);
    const float c2f = static_cast<float>(c2);
    if (!std::isfinite(c2f)) return raw();
    uint32_t c2bits;
    std::memcpy(&c2bits, &c2f, 4);
    if (c2bits == kRawSectorC2) return raw();
    const double C2 = static_cast<double>(c2f);
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
            const double x = codec_x(j, dr);
            const double Fp = sector_predict_F(x, f[0], g[0], C, C2), Gp = sector_predict_G(x, g[0], C, C2);
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
        best_w2 = cb | (static_cast<uint64_t>(static_cast<uint8_t>(rF[2])) << 8) | static_cast<uint64_t>(static_cast<uint8_t>(rF[1]));
        best_w3 = static_cast<uint64_t>(c2bits) | (static_cast<uint64_t>(static_cast<uint8_t>(rF[3])) << 32) |
                  (static_cast<uint64_t>(static_cast<uint8_t>(rG[1])) << 40) | (static_cast<uint64_t>(static_cast<uint8_t>(rG[2])) << 48) |
                  (static_cast<uint64_t>(static_cast<uint8_t>(rG[3])) << 56);
    }
    if (best_cost == (1 << 30)) return raw();
    out.w[0] = codec_to_bits(f[0]);
    out.w[1] = codec_to_bits(g[0]);
    out.w[2] = best_w2;
    out.w[3] = best_w3;
    // the decoder must reproduce every entry exactly -- checked here with the decoder itself
    for (int j = 0; j < 4; ++j) {
        double Fd, Gd;
        sector_decode<true>(out, j, dr, Fd, Gd);
        if (codec_to_bits(Fd) != codec_to_bits(f[j]) || codec_to_bits(Gd) != codec_to_bits(g[j])) return raw();
    }
    return true;
}

}  // namespace pimcb

#endif
