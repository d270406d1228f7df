#include "energy_estimator.h"

REGISTER_ESTIMATOR("energy", EnergyEstimator)

// src/estimator.cpp:917-926
EnergyEstimator::EnergyEstimator(const Path& _path, ActionBase* _actionPtr, const MTRand& _random, double _maxR, int _frequency,
                                 std::string _label)
    : EstimatorBase(_path, _actionPtr, _random, _maxR, _frequency, _label) {
    endLine = false;
    initialize({"K", "V", "V_ext", "V_int", "E", "E_mu", "K/N", "V/N", "E/N"});
}

// src/estimator.cpp:940-1029
void EnergyEstimator::accumulate() {
    double totK = 0.0, totV = 0.0;
    std::array<double, 2> totVop{};
    const int numParticles = path.getTrueNumParticles();
    const int numTimeSlices = endSlice - startSlice;

    const double tailV = (1.0 * numParticles * numParticles / path.boxPtr->volume) * actionPtr->interactionPtr->tailV;
    const double kinNorm = constants()->fourLambdaTauInv() / (constants()->tau() * numTimeSlices);
    const double classicalKinetic = (0.5 * NDIM / constants()->tau()) * numParticles;

    beadLocator beadIndex;
    for (int slice = startSlice; slice < endSlice; slice++) {
        beadIndex[0] = slice;
        const int numBeads = path.numBeadsAtSlice(slice);
        for (int ptcl = 0; ptcl < numBeads; ptcl++) {
            beadIndex[1] = ptcl;
            const dVec vel = path.getVelocity(beadIndex);
            double v2 = 0.0;
            for (int i = 0; i < NDIM; ++i) v2 += vel[i] * vel[i];
            totK -= v2;
        }
    }
    totK *= kinNorm;

    double t1 = 0.0, t2 = 0.0;
    for (int slice = startSlice; slice < endDiagSlice; slice++) {
        t1 += sliceFactor[slice] * actionPtr->derivPotentialActionLambda(slice);
        t2 += sliceFactor[slice] * actionPtr->derivPotentialActionTau(slice);
        if (!(slice % actionPtr->period)) {
            const std::array<double, 2> v = actionPtr->potential(slice);
            totVop[0] += sliceFactor[slice] * v[0];
            totVop[1] += sliceFactor[slice] * v[1];
        }
    }
    t1 *= constants()->lambda() / (constants()->tau() * numTimeSlices);
    t2 /= 1.0 * numTimeSlices;
    totVop[0] /= (numTimeSlices / actionPtr->period);
    totVop[1] /= (numTimeSlices / actionPtr->period);

    totK += (classicalKinetic + t1);
    totV = t2 - t1 + tailV;
    totVop[1] += tailV;

    estimator(estIndex["K"]) += totK;
    estimator(estIndex["V"]) += totV;
    estimator(estIndex["V_ext"]) += totVop[0];
    estimator(estIndex["V_int"]) += totVop[1];
    estimator(estIndex["E"]) += totK + totV;
    estimator(estIndex["E_mu"]) += totK + totV - constants()->mu() * numParticles;
    if (numParticles > 0) {
        numPPAccumulated += 1;
        estimator(estIndex["K/N"]) += totK / (1.0 * numParticles);
        estimator(estIndex["V/N"]) += totV / (1.0 * numParticles);
        estimator(estIndex["E/N"]) += (totK + totV) / (1.0 * numParticles);
    }
    if (numAccumulated == constants()->binSize()) {
        norm(estIndex["K/N"]) = 1.0 * numAccumulated / numPPAccumulated;
        norm(estIndex["V/N"]) = 1.0 * numAccumulated / numPPAccumulated;
        norm(estIndex["E/N"]) = 1.0 * numAccumulated / numPPAccumulated;
        numPPAccumulated = 0;
    }
}
