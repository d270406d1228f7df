// action_b200.h -- B200 replacement of the whole-slice pair sums behind the reference's LocalAction interface.
//
// LocalActionB200 IS a LocalAction (include/action.h:157-254) with the same constructor parameters, created in its place in
// Setup::action (src/setup.cpp:1258-1260): pdrive.cpp:141-159 builds ONE action object per path and hands it to the
// moves and to the estimators alike, so everything bead-level that the moves call -- potentialAction(beadLocator...),
// barePotentialAction, potentialActionCorrection, the NN-table variants -- stays the inherited upstream host code, and
// only the whole-slice virtuals the measurement code calls are overridden:
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
#include "lookuptable.h"
#include "potential.h"          // PotentialBase::tableView() (upstream.patch)
#endif
#include "b200_session.h"

class LocalActionB200 : public LocalAction {
public:
    // parameter list of LocalAction (include/action.h:157-160); the lookup tables come from interaction->tableView()
    LocalActionB200(const Path& path, LookupTable& lookup, PotentialBase* external, PotentialBase* interaction,
                    WaveFunctionBase* waveFunction, const std::array<double, 2>& VFactor,
                    const std::array<double, 2>& gradVFactor, bool local = true, std::string name = "Local",
                    double endFactor = 1.0, int period = 1);
    using LocalAction::potentialAction;              // the bead-level overloads stay upstream's
    using LocalAction::derivPotentialActionTau;      // (int, double) cut-off variants likewise
    using LocalAction::derivPotentialActionLambda;
    using LocalAction::potential;
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
    bool haveTable = false;                 // false: the interaction potential is not tabulated -> upstream host loops
    bool needF2;
    int f2Parity;
    int lastSlice[4] = {1 << 30, 1 << 30, 1 << 30, 1 << 30};   // last slice served per per-slice entry point (unhooked mode)
    const B200Session::PairSums& sums();
    const B200Session::PairSums& sumsForSlice(int slice, int which);
    const double* virial(int slice);        // the four sums of `slice`
    double externalV(int slice);
    std::vector<double> gext;               // gradVext per bead ([M][N_ext][NDIM]); left empty while every gradient is zero ("free")
    std::vector<double> g2ext;              // grad2Vext per bead ([M][N_ext]), for the T-matrix of the virial terms
    const std::vector<double>* externalGradient(bool always = false);
    const std::vector<double>* externalLaplacian();
};

#endif
