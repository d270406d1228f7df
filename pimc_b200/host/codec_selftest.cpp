// codec_selftest.cpp -- CPU check of pimc_b200/csrc/table_codec.h on the real lookup tables: packs (V, dV/dr) and
// (dV/dr, d2V/dr2) of the Aziz potential for a box of N particles, decodes every entry with the decoder the device uses
// and compares bit patterns.  Prints key=value lines for tests/test_host_layer.py.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../csrc/table_codec.h"
#include "aziz.h"

int main(int argc, char** argv) {
    const int N = argc > 1 ? std::atoi(argv[1]) : 64;
    const int year = argc > 2 ? std::atoi(argv[2]) : 1979;
    Prism box(0.02198, N);
    constants()->rc_ = box.side[NDIM - 1];
    AzizPotential az(year, &box);
    const TableView t = az.tableView();
    std::printf("tableLength=%d\n", t.tableLength);
    const pimcb::CodecSteps st = pimcb::codec_steps(t.dr);
    for (int pass = 0; pass < 2; ++pass) {
        const double* F = pass ? t.dVdr : t.V;
        const double* G = pass ? t.d2Vdr2 : t.dVdr;
        const int ns = (t.tableLength + 3) / 4;
        std::vector<pimcb::TableSector> sec(ns);
        long raw = 0, bad = 0;
        int maxres = 0;
        for (int s = 0; s < ns; ++s) {
            if (!pimcb::sector_encode(F, G, t.tableLength, 4 * s, t.dr, st, sec[s])) { ++raw; continue; }
            for (int j = 1; j < 4; ++j)
                maxres = std::max(maxres, std::max(std::abs(pimcb::sector_res_F(sec[s], j)), std::abs(pimcb::sector_res_G(sec[s], j))));
        }
        for (int k = 0; k < t.tableLength; ++k) {
            const pimcb::TableSector& s = sec[k >> 2];
            if (pimcb::sector_is_raw(s)) continue;
            const int j = k & 3;
            double f, g;
            pimcb::sector_decode<true>(s, j, st.x[j], st.xh[j], st.x3[j], f, g);
            if (pimcb::codec_to_bits(f) != pimcb::codec_to_bits(F[k]) || pimcb::codec_to_bits(g) != pimcb::codec_to_bits(G[k])) ++bad;
        }
        std::printf("pass%d_sectors=%d\npass%d_raw=%ld\npass%d_mismatches=%ld\npass%d_maxres=%d\n", pass, ns, pass, raw, pass, bad, pass, maxres);
    }
    // a pair of tables whose second is NOT the derivative of the first must come out (almost) entirely RAW, never wrong
    std::vector<double> A(4096), B(4096);
    for (int k = 0; k < 4096; ++k) { A[k] = 1.0 + 1e-3 * k; B[k] = 17.0 * ((k * 2654435761u) % 1000) - 3.0; }
    long raw = 0, bad = 0;
    for (int s = 0; s < 1024; ++s) {
        pimcb::TableSector sc;
        if (!pimcb::sector_encode(A.data(), B.data(), 4096, 4 * s, 1e-3, pimcb::codec_steps(1e-3), sc)) { ++raw; continue; }
        for (int j = 0; j < 4; ++j) {
            double f, g;
            const pimcb::CodecSteps s3 = pimcb::codec_steps(1e-3);
            pimcb::sector_decode<true>(sc, j, s3.x[j], s3.xh[j], s3.x3[j], f, g);
            if (f != A[4 * s + j] || g != B[4 * s + j]) ++bad;
        }
    }
    std::printf("unrelated_raw=%ld\nunrelated_mismatches=%ld\n", raw, bad);
    return 0;
}
