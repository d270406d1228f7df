#include "action_b200.h"

#include <cmath>
#include <cstdlib>
#include <iostream>

LocalActionB200::LocalActionB200(const Path& _path, LookupTable& _lookup, PotentialBase* _externalPtr,
                                 PotentialBase* _interactionPtr, WaveFunctionBase* _waveFunctionPtr,
                                 const std::array<double, 2>& _VFactor, const std::array<double, 2>& _gradVFactor, bool _local,
                                 std::string _name, double _endFactor, int _period)
    : LocalAction(_path, _lookup, _externalPtr, _interactionPtr, _waveFunctionPtr, _VFactor, _gradVFactor, _local, _name,
                  _endFactor, _period) {
    const TableView table = interactionPtr->tableView();
    if (!table.V || !table.dVdr || table.tableLength <= 0) {
        // reference error convention (src/estimator.cpp:450-455): the B200 action only replaces tabulated pair potentials
        std::cerr << "\nERROR: pimc_b200: the interaction potential exposes no lookup tables (PotentialBase::tableView())"
                  << std::endl;
        std::exit(EXIT_FAILURE);
    }
    haveTable = true;
    B200Session::get(path).setPairTable(table.V, table.dVdr, table.tableLength, table.dr, table.extV.data(),
                                        table.extdVdr.data());
    // which slices carry the gradient correction (src/setup.cpp:1232-1253): gsf {0, 2/9} -> odd slices only,
    // li_broughton {1/12, 1/12} -> all, primitive {0, 0} -> none
    const bool even = gradVFactor[0] > EPS, odd = gradVFactor[1] > EPS;
    needF2 = even || odd;
    f2Parity = (even && odd) ? -1 : (odd ? 1 : 0);
    if (table.d2Vdr2) B200Session::get(path).setPairTableD2(table.d2Vdr2, table.tableLength, table.extd2Vdr2.data());
}

// externalPtr->gradV(path(bead)) for every active bead, O(N M) on the host through the reference's own PotentialBase
// (src/action.cpp:1216: added to the pair force of the bead before squaring).  NULL when all gradients vanish (`free`).
const std::vector<double>* LocalActionB200::externalGradient(bool always) {
    if (!needF2 && !always) return nullptr;
    const auto ext = path.get_beads_extents();
    const size_t Next = ext[1];
    bool any = false;
    gext.assign(static_cast<size_t>(path.numTimeSlices) * Next * NDIM, 0.0);
    for (int slice = 0; slice < path.numTimeSlices; ++slice) {
        const int n = path.numBeadsAtSlice(slice);
        for (int i = 0; i < n; ++i) {
            const dVec g = externalPtr->gradV(path(slice, i));
            for (int d = 0; d < NDIM; ++d) {
                gext[(static_cast<size_t>(slice) * Next + i) * NDIM + d] = g[d];
                any = any || g[d] != 0.0;
            }
        }
    }
    return any ? &gext : nullptr;
}

// externalPtr->grad2V(path(bead)) for every active bead (src/action.cpp:1523, 1698); NULL when every Laplacian vanishes.
const std::vector<double>* LocalActionB200::externalLaplacian() {
    const auto ext = path.get_beads_extents();
    const size_t Next = ext[1];
    bool any = false;
    g2ext.assign(static_cast<size_t>(path.numTimeSlices) * Next, 0.0);
    for (int slice = 0; slice < path.numTimeSlices; ++slice) {
        const int n = path.numBeadsAtSlice(slice);
        for (int i = 0; i < n; ++i) {
            const double g2 = externalPtr->grad2V(path(slice, i));
            g2ext[static_cast<size_t>(slice) * Next + i] = g2;
            any = any || g2 != 0.0;
        }
    }
    return any ? &g2ext : nullptr;
}

const B200Session::PairSums& LocalActionB200::sums() {
    B200Session& session = B200Session::get(path);
    if (session.havePairSums(needF2)) return session.pairSums(dSep, needF2, f2Parity);      // cached for this configuration
    return session.pairSums(dSep, needF2, f2Parity, externalGradient());
}

// Per-slice entry points are called for slice = 0..M-1 in ascending order within one measurement
// (src/estimator.cpp:983-988); without the newConfiguration() hook a non-increasing slice index marks a new pass.
const B200Session::PairSums& LocalActionB200::sumsForSlice(int slice, int which) {
    if (slice <= lastSlice[which]) B200Session::get(path).beginIfUnhooked();
    lastSlice[which] = slice;
    return sums();
}

// sum_i factor_i Vext(r_i), src/action.cpp:941-943 (factor 1 on diagonal configurations)
double LocalActionB200::externalV(int slice) {
    double tot = 0.0;
    const int n = path.numBeadsAtSlice(slice);
    for (int i = 0; i < n; ++i) tot += externalPtr->V(path(slice, i));
    return tot;
}

std::array<double, 2> LocalActionB200::potential(int slice) {
    const B200Session::PairSums& s = sumsForSlice(slice, 0);
    for (int k = 0; k < NPCFSEP; ++k) sepHist(k) = s.hist[static_cast<size_t>(slice) * NPCFSEP + k];   // action.cpp:918,935
    return {externalV(slice), s.vint[slice]};
}

// sum_i | sum_j gradV(r_ij) + gradVext(r_i) |^2 on the device; the external gradients are evaluated on the host through the
// reference's PotentialBase and uploaded with the configuration (pimcb_set_external_gradient).
double LocalActionB200::gradVSquared(int slice) { return sums().f2[slice]; }

// The one whole-path entry point that is also reached from OUTSIDE the measurement loop (DEBUG_MOVE's checkMove calls it
// before and after a trial move, src/move.cpp:190-243): it never trusts a cached configuration, hooked or not.
double LocalActionB200::potentialAction() {
    B200Session::get(path).invalidate();
    const B200Session::PairSums& s = sums();
    double totU = 0.0;
    for (int slice = 0; slice < path.numTimeSlices; slice++) {
        const int eo = slice % 2;
        totU += VFactor[eo] * tau() * (externalV(slice) + s.vint[slice]);
        if (gradVFactor[eo] > EPS)
            totU += gradVFactor[eo] * tau() * tau() * tau() * constants()->lambda() * s.f2[slice];
    }
    return totU;
}

double LocalActionB200::derivPotentialActionTau(int slice) {
    const B200Session::PairSums& s = sumsForSlice(slice, 1);
    const int eo = slice % 2;
    double dU = VFactor[eo] * (externalV(slice) + s.vint[slice]);
    if (gradVFactor[eo] > EPS) dU += 3.0 * gradVFactor[eo] * tau() * tau() * constants()->lambda() * s.f2[slice];
    return dU;
}

double LocalActionB200::derivPotentialActionLambda(int slice) {
    const int eo = slice % 2;
    if (gradVFactor[eo] > EPS) return gradVFactor[eo] * tau() * tau() * tau() * sumsForSlice(slice, 2).f2[slice];
    return 0.0;
}

// ---- virial / pressure terms ----------------------------------------------------------------------------------------
// The virial estimator walks slice = 0..M-1 calling all of these per slice (src/estimator.cpp:1160-1172); they share one
// device pass.  Only the first of them (deltaDOTgradUterm1, the first call of that loop) marks a new configuration in
// unhooked mode.  External potential: its gradient and Laplacian per bead are evaluated on the host through the reference's
// own PotentialBase (O(N M)) and travel with the configuration; `free` uploads nothing.
const double* LocalActionB200::virial(int slice) {
    const int t2 = needF2 ? f2Parity : -2;
    B200Session& session = B200Session::get(path);
    if (session.haveVirialSums(constants()->virialWindow(), t2))
        return &session.virialSums(constants()->virialWindow(), t2)[static_cast<size_t>(slice) * 4];
    const std::vector<double>* ge = externalGradient(true);
    const std::vector<double>* g2 = needF2 ? externalLaplacian() : nullptr;
    return &session.virialSums(constants()->virialWindow(), t2, ge, g2)[static_cast<size_t>(slice) * 4];
}

double LocalActionB200::rDOTgradUterm1(int slice) { return VFactor[slice % 2] * tau() * virial(slice)[0]; }

double LocalActionB200::rDOTgradUterm2(int slice) {
    const int eo = slice % 2;
    if (!(gradVFactor[eo] > EPS)) return 0.0;
    return virial(slice)[1] * (2.0 * gradVFactor[eo] * std::pow(tau(), 3) * constants()->lambda());
}

double LocalActionB200::deltaDOTgradUterm1(int slice) {
    if (slice <= lastSlice[3]) B200Session::get(path).beginIfUnhooked();
    lastSlice[3] = slice;
    return VFactor[slice % 2] * tau() * virial(slice)[2];
}

double LocalActionB200::deltaDOTgradUterm2(int slice) {
    const int eo = slice % 2;
    if (!(gradVFactor[eo] > EPS)) return 0.0;
    return virial(slice)[3] * (2.0 * gradVFactor[eo] * std::pow(tau(), 3) * constants()->lambda());
}

double LocalActionB200::virKinCorr(int slice) {
    const int eo = slice % 2;
    if (!(gradVFactor[eo] > EPS)) return 0.0;
    return sums().f2[slice] * gradVFactor[eo] * std::pow(tau(), 3) * constants()->lambda();
}

double LocalActionB200::secondderivPotentialActionTau(int slice) {
    const int eo = slice % 2;
    if (!(gradVFactor[eo] > EPS)) return 0.0;
    return 6.0 * gradVFactor[eo] * tau() * constants()->lambda() * sums().f2[slice];
}
