#include "scattering_b200.h"
#include "b200_register.h"

#ifdef PIMCB_STANDALONE
#define PIMCB_FMT(spec, v) pimcb_format(spec, v)
#else
#include <boost/format.hpp>
#define PIMCB_FMT(spec, v) boost::str(boost::format(spec) % (v))
#endif

PIMCB_REGISTER_ESTIMATOR("elastic scattering", ElasticScatteringEstimatorB200)
PIMCB_REGISTER_ESTIMATOR("cylinder static structure factor", CylinderStaticStructureFactorEstimatorB200)

// ---- elastic scattering ---------------------------------------------------------------------------------------------
// src/estimator.cpp:4114-4176: q list from the command line, numq columns, header = integer column indices, norm 0.5.
ElasticScatteringEstimatorB200::ElasticScatteringEstimatorB200(const Path& _path, ActionBase* _actionPtr, const MTRand& _random,
                                                               double _maxR, int _frequency, std::string _label)
    : EstimatorBase(_path, _actionPtr, _random, _maxR, _frequency, _label) {
    getQVectors(qValues);
    numq = static_cast<int>(qValues.size());
    B200Session::get(path).setQVectors(qValues);
    initialize(numq);
    header = PIMCB_FMT("#%15d", 0);
    for (int n = 1; n < numq; n++) header.append(PIMCB_FMT("%16d", n));
    for (int n = 0; n < numq; n++) norm(n) = 0.5;
}

// estimator += es (src/estimator.cpp:4233)
void ElasticScatteringEstimatorB200::accumulate() {
    B200Session& session = B200Session::get(path);
    session.beginIfUnhooked();
    const std::vector<double>& es = session.elastic();
    for (int n = 0; n < numq; n++) estimator(n) += es[n];
}

// ---- cylinder S(q) -----------------------------------------------------------------------------------------------------
// src/estimator.cpp:5358-5403: magnitudes 0, dq, 2 dq, ... <= 4 1/A with dq = 2 pi / L_z, "line" geometry (wave-vectors
// along the axis), one column per magnitude, header = the magnitudes, norm = 1/M / (vectors in the shell).
CylinderStaticStructureFactorEstimatorB200::CylinderStaticStructureFactorEstimatorB200(const Path& _path, ActionBase* _actionPtr,
                                                                                       const MTRand& _random, double _maxR,
                                                                                       int _frequency, std::string _label)
    : EstimatorBase(_path, _actionPtr, _random, _maxR, _frequency, _label) {
    const double qMax = 4.0;
    const double dq = 2.0 * M_PI / path.boxPtr->side[NDIM - 1];
    int numq = 0;
    q = getQVectors2(dq, qMax, numq, "line");
    numq = static_cast<int>(q.size());

    std::vector<dVec> flat;
    for (const auto& shell : q) flat.insert(flat.end(), shell.begin(), shell.end());
    B200Session::get(path).setQVectors(flat);

    initialize(numq);
    header = PIMCB_FMT("#%15.6E", 0.0);
    for (int nq = 1; nq < numq; nq++) header.append(PIMCB_FMT("%16.6E", std::sqrt(dot(q[nq][0], q[nq][0]))));
    for (int nq = 0; nq < numq; nq++) {
        norm(nq) = 1.0 / constants()->numTimeSlices();
        if (!q[nq].empty()) norm(nq) /= q[nq].size();
    }
}

void CylinderStaticStructureFactorEstimatorB200::sample() {
    numSampled++;
    if (!baseSample()) return;
    B200Session& session = B200Session::get(path);
    session.beginIfUnhooked();
    int numInside = 0;
    session.ssfCylinder(maxR, numInside);
    if (numInside > 0) {
        totNumAccumulated++;
        numAccumulated++;
        accumulate();
    }
}

// estimator += sf/numParticles with numParticles = num1DParticles (src/estimator.cpp:5417, 5455)
void CylinderStaticStructureFactorEstimatorB200::accumulate() {
    int numParticles = 0;
    const std::vector<double>& raw = B200Session::get(path).ssfCylinder(maxR, numParticles);
    size_t k = 0;
    for (size_t nq = 0; nq < q.size(); nq++) {
        double sf = 0.0;
        for (size_t v = 0; v < q[nq].size(); v++) sf += raw[k++];
        estimator(nq) += sf / numParticles;
    }
}
