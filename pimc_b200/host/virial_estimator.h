// virial_estimator.h -- stand-alone restatement of the reference's centroid-virial energy estimator
// (VirialEnergyEstimator, src/estimator.cpp:1045-1250) for the stand-alone tools.  Its O(N^2 M) terms are the
// ActionBase virtuals deltaDOTgradUterm1/2, rDOTgradUterm1/2, derivPotentialActionTau, virKinCorr,
// secondderivPotentialActionTau and potential(), which LocalActionB200 answers from one pimcb_pair_sums and one
// pimcb_virial_sums pass per configuration; the exchange/kinetic link sums are O(N M window) and stay on the host.
// Inside a reference checkout the reference's own class runs unchanged on top of LocalActionB200.
#ifndef PIMCB_VIRIAL_ESTIMATOR_H
#define PIMCB_VIRIAL_ESTIMATOR_H

#include "estimator_base.h"

class VirialEnergyEstimator : public EstimatorBase {
public:
    VirialEnergyEstimator(const Path& _path, ActionBase* _actionPtr, const MTRand& _random, double _maxR, int _frequency = 1,
                          std::string _label = "estimator");
    static const std::string name;
    std::string getName() const override { return name; }

private:
    void accumulate() override;
    uint32 numPPAccumulated = 0;
};

#endif
