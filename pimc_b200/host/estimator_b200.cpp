#include "estimator_b200.h"
#include "b200_register.h"

#ifdef PIMCB_STANDALONE
#define PIMCB_FMT_INT(spec, v) pimcb_format(spec, v)
#else
#include <boost/format.hpp>
#define PIMCB_FMT_INT(spec, v) boost::str(boost::format(spec) % (v))
#endif

PIMCB_REGISTER_ESTIMATOR("static structure factor", StaticStructureFactorEstimatorB200)
PIMCB_REGISTER_ESTIMATOR("intermediate scattering function", IntermediateScatteringFunctionEstimatorB200)

// ---- S(q) ----------------------------------------------------------------------------------------------------
// Constructor contract of src/estimator.cpp:3661-3693: q list from the command line, `numq` columns, a first header
// line "# ESTINF: num_q = N; (qx,qy,qz) ..." followed by the integer column indices, norm = 1/M.
StaticStructureFactorEstimatorB200::StaticStructureFactorEstimatorB200(const Path& _path, ActionBase* _actionPtr,
                                                                       const MTRand& _random, double _maxR, int _frequency,
                                                                       std::string _label)
    : EstimatorBase(_path, _actionPtr, _random, _maxR, _frequency, _label) {
    getQVectors(qValues);
    numq = static_cast<int>(qValues.size());
    B200Session::get(path).setQVectors(qValues);
    initialize(numq);

    header = PIMCB_FMT_INT("# ESTINF: num_q = %d; ", numq);
    for (const dVec& q : qValues) header += dVecToString(q) + " ";
    header += "\n";
    header += PIMCB_FMT_INT("#%15d", 0);
    for (int n = 1; n < numq; n++) header += PIMCB_FMT_INT("%16d", n);

    for (int n = 0; n < numq; n++) norm(n) = 1.0 / constants()->numTimeSlices();
}

StaticStructureFactorEstimatorB200::~StaticStructureFactorEstimatorB200() {}

// estimator += sf/numParticles (src/estimator.cpp:3736); the sum over slices, pairs and the division by N happen on
// the device (pimcb_ssf), the 1/M of `norm` and the bin average stay in EstimatorBase::output().
void StaticStructureFactorEstimatorB200::accumulate() {
    B200Session& session = B200Session::get(path);
    session.beginIfUnhooked();
    const std::vector<double>& sf = session.ssf();
    for (int n = 0; n < numq; n++) estimator(n) += sf[n];
}

// ---- F(q,tau) ---------------------------------------------------------------------------------------------------
// src/estimator.cpp:3876-3909: numq*M columns, header = column indices 0..numq*M-1, column i*M + tau, norm = 1/M.
IntermediateScatteringFunctionEstimatorB200::IntermediateScatteringFunctionEstimatorB200(
    const Path& _path, ActionBase* _actionPtr, const MTRand& _random, double _maxR, int _frequency, std::string _label)
    : EstimatorBase(_path, _actionPtr, _random, _maxR, _frequency, _label) {
    const int numTimeSlices = constants()->numTimeSlices();
    getQVectors(qValues);
    numq = static_cast<int>(qValues.size());
    B200Session::get(path).setQVectors(qValues);
    initialize(numq * numTimeSlices);

    header = PIMCB_FMT_INT("#%15d", 0);
    for (int n = 1; n < numq * numTimeSlices; n++) header.append(PIMCB_FMT_INT("%16d", n));

    for (int n = 0; n < numq * numTimeSlices; n++) norm(n) = 1.0 / numTimeSlices;
}

IntermediateScatteringFunctionEstimatorB200::~IntermediateScatteringFunctionEstimatorB200() {}

// estimator += isf/numParticles (src/estimator.cpp:3960)
void IntermediateScatteringFunctionEstimatorB200::accumulate() {
    B200Session& session = B200Session::get(path);
    session.beginIfUnhooked();
    const std::vector<double>& f = session.isf();
    const int n = static_cast<int>(f.size());
    for (int k = 0; k < n; k++) estimator(k) += f[k];
}

// ---- bins accumulated elsewhere -------------------------------------------------------------------------------------------
// A walker batch measured with B200Session::measureBatch (and reduced over GPUs) arrives as the SUM of its accumulate()
// increments; adding it with its count leaves EstimatorBase::output (estimator * norm / numAccumulated) unchanged.
void StaticStructureFactorEstimatorB200::addBin(const double* sums, uint32 count) {
    for (int n = 0; n < numq; n++) estimator(n) += sums[n];
    numAccumulated += count;
    totNumAccumulated += count;
}

void IntermediateScatteringFunctionEstimatorB200::addBin(const double* sums, uint32 count) {
    const int n = numq * constants()->numTimeSlices();
    for (int k = 0; k < n; k++) estimator(k) += sums[k];
    numAccumulated += count;
    totNumAccumulated += count;
}
