// action_b200.h -- B200 replacement of the whole-path pair sums behind the reference's ActionBase interface.
//
// LocalActionB200 answers the virtuals the measurement code calls --
//     potentialAction()                    (src/action.cpp:456-472)
//     potential(int slice)  = V(slice)     (include/action.h:197, src/action.cpp:902-947) incl. the sepHist side effect
//     derivPotentialActionTau(int slice)   (src/action.cpp:751-765)
//     derivPotentialActionLambda(int slice)(src/action.cpp:798-806)
//     rDOTgradUterm1/2, deltaDOTgradUterm1/2, virKinCorr, secondderivPotentialActionTau (virial estimator, pressure)
// -- from ONE device pass over all slices (pimcb_pair_sums) per configuration instead of M O(N^2) host loops, and
// removes the redundant re-evaluations the energy estimator triggers (V(slice) 1+1/period times, gradVSquared twice
// per corrected slice; src/estimator.cpp:983-988).  The external potential is evaluated on the host through the
// reference's own PotentialBase (O(N) per slice; `free` for bulk He-4).  Worm factors are 1 on the diagonal
// configurations estimators sample (src/estimator.cpp:228, include/worm.h:50).
#ifndef PIMCB_ACTION_B200_H
#define PIMCB_ACTION_B200_H

#ifdef PIMCB_STANDALONE
#include "pimc_compat.h"
#else
#include "action.h"
#endif
#include "b200_session.h"

class LocalActionB200 : public ActionBase {
public:
    LocalActionB200(const Path& path, PotentialBase* external, PotentialBase* interaction, const TableView& table,
                    const std::array<double, 2>& VFactor, const std::array<double, 2>& gradVFactor, int period);
    double potentialAction() override;
    std::array<double, 2> potential(int slice) override;
    double derivPotentialActionTau(int slice) override;
    double derivPotentialActionLambda(int slice) override;
    double gradVSquared(int slice);
    // virial / pressure terms: the O(N^2) sums come from one pimcb_virial_sums pass per configuration
    double secondderivPotentialActionTau(int slice) override;   // src/action.cpp:776-787
    double rDOTgradUterm1(int slice) override;                  // src/action.cpp:1446-1478
    double rDOTgradUterm2(int slice) override;                  // src/action.cpp:1493-1575
    double deltaDOTgradUterm1(int slice) override;              // src/action.cpp:1588-1652
    double deltaDOTgradUterm2(int slice) override;              // src/action.cpp:1667-1784
    double virKinCorr(int slice) override;                      // src/action.cpp:1786-1803
private:
    std::array<double, 2> VFactor, gradVFactor;
    bool needF2;
    int f2Parity;
    int lastSlice[4] = {1 << 30, 1 << 30, 1 << 30, 1 << 30};   // last slice served per per-slice entry point (unhooked mode)
    const B200Session::PairSums& sums();
    const B200Session::PairSums& sumsForSlice(int slice, int which);
    const double* virial(int slice);        // the four sums of `slice`
    double externalV(int slice);
    std::vector<double> gext;               // gradVext per bead ([M][N_ext][NDIM]); left empty while every gradient is zero ("free")
    const std::vector<double>* externalGradient();
};

#endif
