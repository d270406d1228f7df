// scattering_b200.h -- B200 replacements of the other members of the scattering family,
// behind the reference's plugin API:
//
//   ElasticScatteringEstimatorB200                 replaces ElasticScatteringEstimatorGpu (include/estimator.h:1016-1050,
//                                                  src/estimator.cpp:4114-4235), registered as "elastic scattering"
//                                                  (upstream has no CPU class; its GPU class is "elastic scattering gpu")
//   CylinderStaticStructureFactorEstimatorB200     replaces CylinderStaticStructureFactorEstimator
//                                                  (src/estimator.cpp:5358-5472) under its name
//                                                  "cylinder static structure factor", label "cyl_ssf"
#ifndef PIMCB_SCATTERING_B200_H
#define PIMCB_SCATTERING_B200_H

#ifdef PIMCB_STANDALONE
#include "estimator_base.h"
#else
#include <complex>                  // estimator.h:805 names std::complex without including it
#include "estimator.h"
#endif
#include "b200_session.h"

class ElasticScatteringEstimatorB200 : public EstimatorBase {
public:
    ElasticScatteringEstimatorB200(const Path&, ActionBase*, const MTRand&, double, int _frequency = 1, std::string _label = "es");
    static const std::string name;
    std::string getName() const { return name; }
private:
    int numq;
    std::vector<dVec> qValues;
    void accumulate();
};

class CylinderStaticStructureFactorEstimatorB200 : public EstimatorBase {
public:
    CylinderStaticStructureFactorEstimatorB200(const Path&, ActionBase*, const MTRand&, double, int _frequency = 1,
                                               std::string _label = "cyl_ssf");
    static const std::string name;
    std::string getName() const { return name; }
    void sample();                                   // only when some particle is inside the radius (:5464-5472)
private:
    std::vector<std::vector<dVec>> q;                // wave-vectors per magnitude shell (getQVectors2)
    void accumulate();
};

#endif
