// b200_register.h -- factory registration of the B200 estimator classes.
// Upstream defines REGISTER_ESTIMATOR inside src/estimator.cpp:32-34 (not in a header), next to the object it uses:
//     EstimatorFactory estimatorFactory;                                    (src/estimator.cpp:31)
//     #define REGISTER_ESTIMATOR(NAME,TYPE)  const std::string TYPE::name = NAME;
//                                            bool reg ## TYPE = estimatorFactory()->Register<TYPE>(TYPE::name);
// Factory::operator() hands out ONE function-local singleton per factory type (include/factory.h:45-53), so any
// EstimatorFactory object reaches the same registry; the adaptor translation units use a file-local one.
#ifndef PIMCB_REGISTER_H
#define PIMCB_REGISTER_H

#ifdef PIMCB_STANDALONE
#include "estimator_base.h"
#else
#include "factory.h"
#endif

namespace { EstimatorFactory pimcbEstimatorFactory; }
#define PIMCB_REGISTER_ESTIMATOR(NAME, TYPE) \
    const std::string TYPE::name = NAME;     \
    bool reg##TYPE = pimcbEstimatorFactory()->Register<TYPE>(TYPE::name);

#endif
