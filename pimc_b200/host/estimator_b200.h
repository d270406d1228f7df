// estimator_b200.h -- B200 replacements of the scattering estimators, behind the reference's plugin API.
//
//   StaticStructureFactorEstimatorB200            replaces StaticStructureFactorEstimator (include/estimator.h:861-877,
//                                                 src/estimator.cpp:3661-3737) and ...GpuEstimator (:878-912 / :3751-3861)
//   IntermediateScatteringFunctionEstimatorB200   replaces IntermediateScatteringFunctionEstimator (:958-977 / :3876-3961)
//                                                 and ...EstimatorGpu (:979-1014 / :3976-4100)
//
// Both register under the CPU class names "static structure factor" / "intermediate scattering function"
// (src/estimator.cpp:60-61) with labels "ssfq" / "isf" and the CPU column layout (Nq resp. Nq*M columns, norm 1/M), so
// `-e "static structure factor"` and the OUTPUT/ce-ssfq-*.dat / ce-isf-*.dat files are drop-in.
#ifndef PIMCB_ESTIMATOR_B200_H
#define PIMCB_ESTIMATOR_B200_H

#ifdef PIMCB_STANDALONE
#include "estimator_base.h"
#else
#include <complex>                  // estimator.h:805 names std::complex without including it
#include "estimator.h"
#endif
#include "b200_session.h"

class StaticStructureFactorEstimatorB200 : public EstimatorBase {
public:
    StaticStructureFactorEstimatorB200(const Path&, ActionBase*, const MTRand&, double, int _frequency = 1,
                                       std::string _label = "ssfq");
    ~StaticStructureFactorEstimatorB200();
    static const std::string name;
    std::string getName() const { return name; }
    // sums of `count` accumulate() increments computed elsewhere (device-resident bins, other GPUs): estimator += sums
    void addBin(const double* sums, uint32 count);
private:
    int numq;
    std::vector<dVec> qValues;
    void accumulate();
};

class IntermediateScatteringFunctionEstimatorB200 : public EstimatorBase {
public:
    IntermediateScatteringFunctionEstimatorB200(const Path&, ActionBase*, const MTRand&, double, int _frequency = 1,
                                                std::string _label = "isf");
    ~IntermediateScatteringFunctionEstimatorB200();
    static const std::string name;
    std::string getName() const { return name; }
    void addBin(const double* sums, uint32 count);
private:
    int numq;
    std::vector<dVec> qValues;
    void accumulate();
};

#endif
