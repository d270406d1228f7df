// host_selftest -- GPU-free checks of the host layer: factory registration names, wave-vector generation, header and
// row formatting.  Prints a small key=value report that tests/test_host_layer.py compares with independent CPU results.
#include <iostream>

#include "aziz.h"
#include "estimator_base.h"
#include "state_file.h"

class ProbeEstimator : public EstimatorBase {
public:
    ProbeEstimator(const Path& p, ActionBase* a, const MTRand& r, double maxR) : EstimatorBase(p, a, r, maxR, 1, "probe") {}
    // getQVectors2 report: one "shell=" line per magnitude, vectors separated by ';'
    void shells(double dq, double qMax, const std::string& geometry) {
        int numq = 0;
        const auto q = getQVectors2(dq, qMax, numq, geometry);
        std::cout.precision(17);
        std::cout << "numq2=" << numq << std::endl;
        for (const auto& sh : q) {
            std::cout << "shell=";
            for (size_t k = 0; k < sh.size(); ++k) {
                for (int d = 0; d < NDIM; ++d) std::cout << sh[k][d] << (d + 1 < NDIM ? " " : "");
                if (k + 1 < sh.size()) std::cout << ";";
            }
            std::cout << std::endl;
        }
    }
    void run() {
        std::vector<dVec> q;
        getQVectors(q);
        std::cout << "nq=" << q.size() << std::endl;
        for (const dVec& v : q) std::cout << "q=" << dVecToString(v) << std::endl;
        std::cout.precision(17);
        for (const dVec& v : q) {
            std::cout << "qraw=";
            for (int d = 0; d < NDIM; ++d) std::cout << v[d] << (d + 1 < NDIM ? " " : "");
            std::cout << std::endl;
        }
    }
};

// host_selftest --state <file> N density : parse a text state file and report what the loader would hold
static int stateReport(const char* file, int N, double density) {
    PimcState st;
    std::string err;
    if (!readStateFile(file, st, err)) { std::cout << "error=" << err << std::endl; return 1; }
    Prism box(density, N);
    std::cout.precision(17);
    std::cout << "header=" << st.headerWorldLines << std::endl << "slices=" << st.numTimeSlices << std::endl
              << "worldlines=" << st.numWorldLines << std::endl << "beadsOn=" << st.numBeadsOn() << std::endl
              << "diagonal=" << st.isDiagonal() << std::endl << "leftPacked=" << st.isLeftPacked() << std::endl
              << "linksClosed=" << st.linksClosed() << std::endl;
    if (!st.isLeftPacked()) st.leftPack();
    st.putInside(box);
    if (st.nextLink.size() == st.beads.size())
        for (int s = 0; s < st.numTimeSlices; ++s)
            for (int p = 0; p < st.numWorldLines; ++p)
                std::cout << "link=" << s << " " << p << " " << st.nextLink[st.idx(s, p)][0] << " " << st.nextLink[st.idx(s, p)][1] << " "
                          << st.prevLink[st.idx(s, p)][0] << " " << st.prevLink[st.idx(s, p)][1] << " " << st.wormBeads[st.idx(s, p)] << std::endl;
    std::cout << "perSlice=";
    for (int n : st.numBeadsAtSlice) std::cout << n << " ";
    std::cout << std::endl;
    for (int s = 0; s < st.numTimeSlices; ++s)
        for (int p = 0; p < st.numBeadsAtSlice[s]; ++p) {
            std::cout << "bead=";
            for (int d = 0; d < NDIM; ++d) std::cout << st.beads[st.idx(s, p)][d] << (d + 1 < NDIM ? " " : "");
            std::cout << std::endl;
        }
    return 0;
}

int main(int argc, char** argv) {
    if (argc == 5 && std::string(argv[1]) == "--state") return stateReport(argv[2], std::atoi(argv[3]), std::atof(argv[4]));
    if (argc == 7 && std::string(argv[1]) == "--q2") {          // --q2 N density dq qMax geometry
        const int N2 = std::atoi(argv[2]);
        Prism box2(std::atof(argv[3]), N2);
        constants()->numTimeSlices_ = 4;
        constants()->initialNumParticles_ = N2;
        Path path2(&box2, 4, N2, N2);
        MTRand r2;
        ProbeEstimator(path2, nullptr, r2, 0.0).shells(std::atof(argv[4]), std::atof(argv[5]), argv[6]);
        return 0;
    }
    if (argc < 5) { std::cerr << "usage: host_selftest N density type \"wavevector\"" << std::endl; return 2; }
    const int N = std::atoi(argv[1]);
    Prism box(std::atof(argv[2]), N);
    constants()->wavevectorType_ = argv[3];
    constants()->wavevector_ = argv[4];
    constants()->numTimeSlices_ = 4;
    constants()->initialNumParticles_ = N;
    Path path(&box, 4, N, N);
    MTRand r;
    std::cout.precision(17);
    std::cout << "side=" << box.side[0] << std::endl << "maxSep=" << box.maxSep << std::endl;
    for (const std::string& n : estimatorFactory()->getNames()) std::cout << "registered=" << n << std::endl;
    ProbeEstimator(path, nullptr, r, 0.0).run();
    std::cout << "row=" << pimcb_format("%16.8E", 30864.19725) << pimcb_format("%16.8E", -6.25e-4) << std::endl;
#if NDIM == 3
    AzizPotential az(1979, &box);
    const TableView t = az.tableView();
    std::cout << "tableLength=" << t.tableLength << std::endl << "dr=" << t.dr << std::endl;
    std::cout << "V1e6=" << t.V[1000000] << std::endl << "dV1e6=" << t.dVdr[1000000] << std::endl;
    double cs = 0.0;
    for (int k = 0; k < t.tableLength; k += 997) cs += t.V[k] * 1e-3 + t.dVdr[k] * 1e-6;
    std::cout << "checksum=" << cs << std::endl;
    std::cout << "d2V1e6=" << t.d2Vdr2[1000000] << std::endl;
    double cs2 = 0.0;
    for (int k = 0; k < t.tableLength; k += 997) cs2 += t.d2Vdr2[k] * 1e-9;
    std::cout << "checksum2=" << cs2 << std::endl;
#endif
    return 0;
}
