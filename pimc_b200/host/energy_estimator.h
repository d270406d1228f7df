// energy_estimator.h -- stand-alone restatement of the reference's thermodynamic EnergyEstimator
// (src/estimator.cpp:905-1029), so that the stand-alone tools can show the default estimator of every pimc run on top
// of the device pair sums: its O(N^2 M) terms are ActionBase::potential / derivPotentialActionTau /
// derivPotentialActionLambda, which LocalActionB200 answers from one pimcb_pair_sums pass; the kinetic link sum is
// O(N M) and stays on the host.  Inside a reference checkout the reference's own class is used unchanged.
#ifndef PIMCB_ENERGY_ESTIMATOR_H
#define PIMCB_ENERGY_ESTIMATOR_H

#include "estimator_base.h"

class EnergyEstimator : public EstimatorBase {
public:
    EnergyEstimator(const Path& _path, ActionBase* _actionPtr, const MTRand& _random, double _maxR, int _frequency = 1,
                    std::string _label = "estimator");
    static const std::string name;
    std::string getName() const override { return name; }

private:
    void accumulate() override;
    uint32 numPPAccumulated = 0;
};

#endif
