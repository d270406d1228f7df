// ref_cpu_shim.cpp -- test infrastructure (never linked by the product): C entry points around the REFERENCE'S OWN CPU
// implementation of the measurement hot path.  The upstream function bodies -- the estimators' accumulate() loops, the
// LocalAction slice sums, getQVectors, putInBC, getSeparation, the Aziz class -- are compiled from the upstream tree
// where they lie: oracle/ref_cpu_extract.py and oracle/ref_aziz_extract.py cut them out of the upstream files into scratch
// include files at build time (nothing upstream is stored in this repository), include/array_math.h is included directly.
// This file supplies ONLY the surrounding declarations those bodies need -- stand-in classes written from scratch with
// the upstream member names (the real headers pull Boost and <mdspan>): storage, accessors, constants.  Every piece of
// arithmetic on the path (minimum image, separations, table lookups, pair loops, estimator normalisations) is upstream
// text.  Built per NDIM by `make -C oracle ref` with -O2 -ffp-contract=off into oracle/_ref/librefcpu<NDIM>d.so.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <map>
#include <numeric>
#include <sstream>
#include <fstream>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#ifndef NDIM
#define NDIM 3
#endif
// include/common.h:85-93, 96-122
#define NPCFSEP 50
#define EPS 1.0E-7
#define XXX -1
#define PIMC_ASSERT(x)
typedef unsigned long uint32;
typedef std::array<std::array<double, NDIM>, NDIM> dMat;
typedef std::array<double, NDIM> dVec;
typedef std::array<int, NDIM> iVec;
typedef std::array<int, 2> beadLocator;
enum beadState { HEADTAIL, SPECIAL, NONE };

#include "array_math.h"            // the reference's own header: dot, sum, array arithmetic

using namespace std;               // upstream sources use sqrt / cos / exp / abs unqualified

#include "common_helpers.inc"      // upstream: enumerate, apply_matrix_vector_product, all(bead, bead)

// ---- DynamicArray<T,1> stand-in (include/dynamic_array.h): resize / fill / operator() / size, += array, / scalar -------
template <class T, int Rank> class DynamicArray;
template <class T>
class DynamicArray<T, 1> {
public:
    void resize(size_t n) { v_.resize(n); }
    void fill(const T& x) { std::fill(v_.begin(), v_.end(), x); }
    T& operator()(size_t i) { return v_[i]; }
    const T& operator()(size_t i) const { return v_[i]; }
    size_t size() const { return v_.size(); }
    T* data() { return v_.data(); }
    const T* data() const { return v_.data(); }
    DynamicArray& operator+=(const DynamicArray& o) {          // elementwise, include/dynamic_array.h
        for (size_t i = 0; i < v_.size(); ++i) v_[i] += o.v_[i];
        return *this;
    }
    DynamicArray& operator/=(const T& s) {
        for (auto& x : v_) x /= s;
        return *this;
    }
    template <class U> DynamicArray operator/(const U& s) const {
        DynamicArray r(*this);
        for (auto& x : r.v_) x = x / s;
        return r;
    }
private:
    std::vector<T> v_;
};

// rank 2: what the upstream stream operators touch (extents, resize, operator()(i, j))
template <class T>
class DynamicArray<T, 2> {
public:
    void resize(size_t r, size_t c) { r_ = r; c_ = c; v_.assign(r * c, T()); }
    std::array<size_t, 2> extents() const { return {r_, c_}; }
    T& operator()(size_t i, size_t j) { return v_[i * c_ + j]; }
    const T& operator()(size_t i, size_t j) const { return v_[i * c_ + j]; }
    T* data() { return v_.data(); }
    const T* data() const { return v_.data(); }
private:
    size_t r_ = 0, c_ = 0;
    std::vector<T> v_;
};

#include "common_stream.inc"       // upstream: operator<< / operator>> of std::array and DynamicArray<T,2> (state-file text)

// ---- Container (include/container.h:24-59, src/container.cpp:84-144): data members + upstream putInBC -------------------
class Container {
public:
    std::array<unsigned int, NDIM> periodic{};
    dVec side{}, sideInv{}, pSide{};
    double volume = 0.0, maxSep = 0.0;
#include "container_members.inc"   // upstream: void putInBC(dVec& r) const
};

// ---- constants (include/constants.h): the getters the bodies call ------------------------------------------------------
class ConstantParameters {
public:
    int numTimeSlices() const { return numTimeSlices_; }
    double tau() const { return tau_; }
    double lambda() const { return lambda_; }
    double mu() const { return mu_; }
    double rc() const { return rc_; }
    double V() const { return V_; }
    int virialWindow() const { return virialWindow_; }
    double fourLambdaTauInv() const { return 0.25 / (lambda_ * tau_); }     // include/constants.h:82
    uint32 binSize() const { return binSize_; }
    std::string wavevector() const { return wavevector_; }
    std::string wavevectorType() const { return wavevectorType_; }
    int numTimeSlices_ = 0, virialWindow_ = 5;
    double tau_ = 0.0, lambda_ = 0.0, mu_ = 0.0, rc_ = 0.0, V_ = 0.0;
    uint32 binSize_ = 1u << 30;
    std::string wavevector_, wavevectorType_;
};
static ConstantParameters g_constants;
static ConstantParameters* constants() { return &g_constants; }

// ---- worm + path (include/worm.h, include/path.h): diagonal configurations, every bead state NONE ----------------------
class Worm {
public:
    beadState getState(const beadLocator&) const { return NONE; }           // diagonal configuration: no worm, no special beads
    double factor(const beadState state1, const beadLocator&) const;        // upstream body (src/worm.cpp:108-118)
    double factor(const beadState state1) const { return 1.0 - 0.5 * (state1 != NONE); }   // include/worm.h:50
    int getNumBeadsOn() const { return numBeadsOn; }
    int numBeadsOn = 0;
};
#include "worm.inc"

class Path {
public:
    Path(const Container* box, int M, int N, int Next, const double* beads_aos, const int* next_aos)
        : numTimeSlices(M), boxPtr(box), n_(N), next_(Next), b_(reinterpret_cast<const dVec*>(beads_aos)) {
        worm.numBeadsOn = M * N;
        nextLink_.assign(static_cast<size_t>(M) * Next, beadLocator{XXX, XXX});
        prevLink_.assign(static_cast<size_t>(M) * Next, beadLocator{XXX, XXX});
        for (int s = 0; s < M; ++s)
            for (int p = 0; p < N; ++p) {
                beadLocator n{(s + 1) % M, p};
                if (next_aos) n = {next_aos[(static_cast<size_t>(s) * Next + p) * 2], next_aos[(static_cast<size_t>(s) * Next + p) * 2 + 1]};
                nextLink_[static_cast<size_t>(s) * Next + p] = n;
                if (n[0] != XXX && n[1] != XXX) prevLink_[static_cast<size_t>(n[0]) * Next + n[1]] = beadLocator{s, p};
            }
    }
    const int numTimeSlices;
    const Container* boxPtr;
    Worm worm;
    int numBeadsAtSlice(int) const { return n_; }
    int getTrueNumParticles() const { return worm.getNumBeadsOn() / numTimeSlices; }                 // include/path.h:54
    const dVec& operator()(int slice, int ptcl) const { return b_[static_cast<size_t>(slice) * next_ + ptcl]; }
    const dVec& operator()(const beadLocator& b) const { return (*this)(b[0], b[1]); }
    const dVec& beads(const beadLocator& b) const { return (*this)(b); }                             // `beads(next(beadIndex))` in getVelocity
    const beadLocator& next(const beadLocator& b) const { return nextLink_[static_cast<size_t>(b[0]) * next_ + b[1]]; }
    const beadLocator& prev(const beadLocator& b) const { return prevLink_[static_cast<size_t>(b[0]) * next_ + b[1]]; }
    beadLocator next(const beadLocator& b, int numLinks) const {                                     // include/path.h:233-240
        beadLocator bI = b;
        for (int m = 0; m < numLinks; m++) bI = next(bI);
        return bI;
    }
    beadLocator prev(const beadLocator& b, int numLinks) const {                                     // include/path.h:256-263
        beadLocator bI = b;
        for (int m = 0; m < numLinks; m++) bI = prev(bI);
        return bI;
    }
    dVec getSeparation(const beadLocator& bead1, const beadLocator& bead2) const;                    // upstream bodies
    dVec getVelocity(const beadLocator& beadIndex) const;
private:
    int n_, next_;
    const dVec* b_;
    std::vector<beadLocator> nextLink_, prevLink_;
};
#include "path_inline.inc"         // upstream: Path::getSeparation, Path::getVelocity

// ---- potentials: PotentialBase stand-in + the upstream Aziz class ---------------------------------------------------------
class PotentialBase {
public:
    PotentialBase() {}
    virtual ~PotentialBase() {}
    virtual double V(const dVec&) { return 0.0; }
    virtual dVec gradV(const dVec&) { return dVec{}; }
    virtual double grad2V(const dVec&) { return 0.0; }
    double tailV = 0.0;
};
class FreePotential : public PotentialBase {};                              // include/potential.h: V = 0, gradV = 0
// a non-trivial external potential for the tests of the gradient coupling (not an upstream class): V = k r^2 / 2
class SpringTestPotential : public PotentialBase {
public:
    explicit SpringTestPotential(double k_) : k(k_) {}
    double V(const dVec& r) override { return 0.5 * k * dot(r, r); }
    dVec gradV(const dVec& r) override { return k * r; }
    double grad2V(const dVec&) override { return NDIM * k; }      // Laplacian: enters the T-matrix of the virial terms
private:
    double k;
};
#include "ref_aziz_extract.inc"    // upstream: TabulatedPotential<T>, AzizPotential

// action.cpp's file-local helper (src/action.cpp:119-156) without its timing statistics: V per position
static void evaluateExternalPotential(PotentialBase* potential, const std::vector<dVec>& positions, std::vector<double>& values) {
    values.resize(positions.size());
    for (std::size_t i = 0; i < positions.size(); ++i) values[i] = potential->V(positions[i]);
}

// ---- action classes: members as upstream (include/action.h:30-254), bodies upstream ------------------------------------------
class ActionBase {
public:
    ActionBase(const Path& p, PotentialBase* ext, PotentialBase* inter, int period_)
        : period(period_), externalPtr(ext), interactionPtr(inter), path(p) {
        sepHist.resize(NPCFSEP);
        sepHist.fill(0);
        dSep = 0.5 * sqrt(1.0 * NDIM) * path.boxPtr->side[NDIM - 1] / (1.0 * NPCFSEP);              // src/action.cpp:192
    }
    virtual ~ActionBase() {}
    virtual double potentialAction() { return 0.0; }
    virtual std::array<double, 2> potential(int) { return {0.0, 0.0}; }
    virtual double derivPotentialActionTau(int) { return 0.0; }
    virtual double derivPotentialActionLambda(int) { return 0.0; }
    virtual double secondderivPotentialActionTau(int) { return 0.0; }
    virtual double rDOTgradUterm1(int) { return 0.0; }
    virtual double rDOTgradUterm2(int) { return 0.0; }
    virtual double deltaDOTgradUterm1(int) { return 0.0; }
    virtual double deltaDOTgradUterm2(int) { return 0.0; }
    virtual double virKinCorr(int) { return 0.0; }
    const int period;
    PotentialBase* externalPtr;
    PotentialBase* interactionPtr;
    DynamicArray<int, 1> sepHist;
    double tau() const { return constants()->tau(); }
protected:
    const Path& path;
    double dSep;
    beadLocator bead2, bead3;
    dVec sep, sep2;
    inline void updateSepHist(const dVec&);
};

class LocalAction : public ActionBase {
public:
    LocalAction(const Path& p, PotentialBase* ext, PotentialBase* inter, const std::array<double, 2>& VF,
                const std::array<double, 2>& GF, int period_)
        : ActionBase(p, ext, inter, period_), VFactor(VF), gradVFactor(GF) {}
    double potentialAction();
    std::array<double, 2> potential(int slice) { return V(slice); }                                   // include/action.h:197
    double derivPotentialActionTau(int slice);
    double secondderivPotentialActionTau(int slice);
    double derivPotentialActionLambda(int slice);
    double rDOTgradUterm1(const int slice);
    double rDOTgradUterm2(const int slice);
    double deltaDOTgradUterm1(int slice) { return deltadotgradUterm1(slice); }                        // include/action.h:190-194
    double deltaDOTgradUterm2(int slice) { return deltadotgradUterm2(slice); }
    double virKinCorr(int slice) { return virialKinCorrection(slice); }
    std::array<double, 2> V(const int slice);
    double gradVSquared(const int slice);
    double deltadotgradUterm1(const int slice);
    double deltadotgradUterm2(const int slice);
    double virialKinCorrection(const int slice);
protected:
    int eo = 0;
    std::array<double, 2> VFactor, gradVFactor;
};
#include "action.inc"              // upstream bodies

// ---- estimator classes: members as upstream (include/estimator.h), bodies upstream -------------------------------------------
class EstimatorBase {
public:
    EstimatorBase(const Path& p, ActionBase* a, double _maxR) : path(p), actionPtr(a), maxR(_maxR) {
        startSlice = 0;                                   // PIMC mode (src/estimator.cpp:182-199)
        endSlice = endDiagSlice = path.numTimeSlices;
        sliceFactor.assign(path.numTimeSlices, 1.0);
    }
    virtual ~EstimatorBase() {}
    DynamicArray<double, 1> estimator, norm;
    std::map<std::string, int> estIndex;
    uint32 numAccumulated = 1;
    void initialize(int n) { estimator.resize(n); norm.resize(n); norm.fill(1.0); estimator.fill(0.0); }
    void initialize(const std::vector<std::string>& labels) {
        for (size_t i = 0; i < labels.size(); ++i) estIndex[labels[i]] = static_cast<int>(i);        // src/estimator.cpp:274-278
        initialize(static_cast<int>(labels.size()));
    }
    void getQVectors(std::vector<dVec>& qValues);
    std::vector<std::vector<dVec>> getQVectors2(double dq, double qMax, int& numq, std::string qGeometry);
protected:
    const Path& path;
    ActionBase* actionPtr;
    double maxR;
    int startSlice, endSlice, endDiagSlice;
    std::vector<double> sliceFactor;
};
struct StaticStructureFactorEstimator : EstimatorBase {
    using EstimatorBase::EstimatorBase;
    DynamicArray<double, 1> sf;
    std::vector<dVec> qValues;
    void accumulate();
};
struct IntermediateScatteringFunctionEstimator : EstimatorBase {
    using EstimatorBase::EstimatorBase;
    DynamicArray<double, 1> isf;
    DynamicArray<dVec, 1> qValues_dVec;
    int numq = 0;
    void accumulate();
};
struct EnergyEstimator : EstimatorBase {
    using EstimatorBase::EstimatorBase;
    uint32 numPPAccumulated = 0;
    void accumulate();
};
struct VirialEnergyEstimator : EstimatorBase {
    using EstimatorBase::EstimatorBase;
    uint32 numPPAccumulated = 0;
    void accumulate();
};
struct CylinderStaticStructureFactorEstimator : EstimatorBase {
    using EstimatorBase::EstimatorBase;
    DynamicArray<double, 1> sf;
    std::vector<std::vector<dVec>> q;
    void accumulate();
};
#include "estimator.inc"           // upstream bodies

// ====================================================================================================================
// C entry points
// ====================================================================================================================
namespace {

Container make_container(const double* side, const unsigned* periodic) {
    Container c;                                               // src/container.cpp:84-144 (Prism): plain data
    double acc = 0.0;
    c.volume = 1.0;
    for (int i = 0; i < NDIM; ++i) {
        c.side[i] = side[i];
        c.periodic[i] = periodic ? periodic[i] : 1u;
        c.sideInv[i] = 1.0 / side[i];
        c.pSide[i] = c.periodic[i] * side[i];
        const double h = side[i] / (c.periodic[i] + 1u);
        acc += h * h;
        c.volume *= side[i];
    }
    c.maxSep = sqrt(acc);
    return c;
}

void set_constants(int M, double tau, double lambda, double mu, double rc, double volume, int window) {
    g_constants.numTimeSlices_ = M;
    g_constants.tau_ = tau;
    g_constants.lambda_ = lambda;
    g_constants.mu_ = mu;
    g_constants.rc_ = rc;
    g_constants.V_ = volume;
    g_constants.virialWindow_ = window;
}

}  // namespace

template <class T, class S>
static int write_array_text(const S* src, int R, int C, int width, char* buf, int buflen) {
    DynamicArray<T, 2> a;
    a.resize(R, C);
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < C; ++j)
            std::memcpy(&a(i, j), src + (static_cast<size_t>(i) * C + j) * width, sizeof(T));
    std::stringstream ss;
    ss << std::setprecision(16) << a << std::endl;
    const std::string t = ss.str();
    if (static_cast<int>(t.size()) + 1 <= buflen) std::memcpy(buf, t.c_str(), t.size() + 1);
    return static_cast<int>(t.size()) + 1;
}
template <class T, class S>
static int read_array_text(const char* text, S* dst, int max_elems, int width, int* R, int* C) {
    DynamicArray<T, 2> a;
    std::istringstream is(text);
    is >> a;
    if (!is) return -1;
    const auto e = a.extents();
    *R = static_cast<int>(e[0]);
    *C = static_cast<int>(e[1]);
    if (static_cast<long>(e[0] * e[1]) * width > max_elems) return -2;
    for (size_t i = 0; i < e[0]; ++i)
        for (size_t j = 0; j < e[1]; ++j) std::memcpy(dst + (i * e[1] + j) * width, &a(i, j), sizeof(T));
    return 0;
}
extern "C" {

int refcpu_ndim(void) { return NDIM; }

// EstimatorBase::getQVectors.  Returns the number of vectors (out may be NULL to size).
int refcpu_qvectors(const char* type, const char* text, const double* side, double* out, int max_out) {
    const Container box = make_container(side, nullptr);
    g_constants.wavevector_ = text;
    g_constants.wavevectorType_ = type;
    Path path(&box, 2, 1, 1, side, nullptr);                   // getQVectors only reads path.boxPtr->side
    EstimatorBase e(path, nullptr, 0.0);
    std::vector<dVec> q;
    e.getQVectors(q);
    for (size_t k = 0; k < q.size() && static_cast<int>(k) < max_out; ++k)
        for (int d = 0; d < NDIM; ++d) out[k * NDIM + d] = q[k][d];
    return static_cast<int>(q.size());
}

// EstimatorBase::getQVectors2: flattened vectors + shell sizes; returns the number of shells.
int refcpu_qvectors2(double dq, double qMax, const char* geometry, const double* side, double* out, int max_vecs, int* sizes,
                     int max_shells) {
    const Container box = make_container(side, nullptr);
    Path path(&box, 2, 1, 1, side, nullptr);
    EstimatorBase e(path, nullptr, 0.0);
    int numq = 0;
    const auto q = e.getQVectors2(dq, qMax, numq, geometry);
    int k = 0;
    for (size_t s = 0; s < q.size(); ++s) {
        if (static_cast<int>(s) < max_shells) sizes[s] = static_cast<int>(q[s].size());
        for (const dVec& v : q[s]) {
            if (k < max_vecs) for (int d = 0; d < NDIM; ++d) out[static_cast<size_t>(k) * NDIM + d] = v[d];
            ++k;
        }
    }
    return static_cast<int>(q.size());
}

// StaticStructureFactorEstimator::accumulate: out[nq] = the increment of `estimator` (sf / N).
int refcpu_ssf(const double* side, const unsigned* periodic, const double* beads, int M, int N, int Next, const double* q, int nq,
               double* out) {
    const Container box = make_container(side, periodic);
    set_constants(M, 1.0, 1.0, 0.0, 0.0, box.volume, 5);
    Path path(&box, M, N, Next, beads, nullptr);
    StaticStructureFactorEstimator e(path, nullptr, 0.0);
    for (int k = 0; k < nq; ++k) { dVec v; for (int d = 0; d < NDIM; ++d) v[d] = q[k * NDIM + d]; e.qValues.push_back(v); }
    e.sf.resize(nq);
    e.initialize(nq);
    e.accumulate();
    for (int k = 0; k < nq; ++k) out[k] = e.estimator(k);
    return 0;
}

// IntermediateScatteringFunctionEstimator::accumulate: out[nq*M] = isf / N.
int refcpu_isf(const double* beads, const double* side, int M, int N, int Next, const double* q, int nq, double* out) {
    const Container box = make_container(side, nullptr);
    set_constants(M, 1.0, 1.0, 0.0, 0.0, box.volume, 5);
    Path path(&box, M, N, Next, beads, nullptr);
    IntermediateScatteringFunctionEstimator e(path, nullptr, 0.0);
    e.numq = nq;
    e.qValues_dVec.resize(nq);
    for (int k = 0; k < nq; ++k) for (int d = 0; d < NDIM; ++d) e.qValues_dVec(k)[d] = q[k * NDIM + d];
    e.isf.resize(static_cast<size_t>(nq) * M);
    e.initialize(nq * M);
    e.accumulate();
    for (int k = 0; k < nq * M; ++k) out[k] = e.estimator(k);
    return 0;
}

// CylinderStaticStructureFactorEstimator::accumulate over explicit shells: out[nshell] = sf / num1DParticles; returns
// num1DParticles (0: upstream would not have sampled).
int refcpu_ssf_cyl(const double* side, const unsigned* periodic, const double* beads, int M, int N, int Next, const double* q,
                   const int* shell_sizes, int nshell, double maxR, double* out) {
    const Container box = make_container(side, periodic);
    set_constants(M, 1.0, 1.0, 0.0, 0.0, box.volume, 5);
    Path path(&box, M, N, Next, beads, nullptr);
    const int n1d = num1DParticles(path, maxR);
    if (n1d == 0) return 0;
    CylinderStaticStructureFactorEstimator e(path, nullptr, maxR);
    int k = 0;
    for (int s = 0; s < nshell; ++s) {
        std::vector<dVec> shell;
        for (int v = 0; v < shell_sizes[s]; ++v, ++k) { dVec x; for (int d = 0; d < NDIM; ++d) x[d] = q[static_cast<size_t>(k) * NDIM + d]; shell.push_back(x); }
        e.q.push_back(shell);
    }
    e.sf.resize(nshell);
    e.initialize(nshell);
    e.accumulate();
    for (int s = 0; s < nshell; ++s) out[s] = e.estimator(s);
    return n1d;
}

// The text a state file holds for one array (src/pimc.cpp:955-962: `stream << std::setprecision(16) << array << std::endl`),
// written by the upstream operator<<.  kind 0: DynamicArray<dVec,2> (beads), 1: DynamicArray<beadLocator,2> (links),
// 2: DynamicArray<unsigned int,2> (worm.beads).  Returns the number of bytes needed (including the terminator).
int refcpu_write_array(int kind, const void* src, int R, int C, char* buf, int buflen) {
    if (kind == 0) return write_array_text<dVec>(static_cast<const double*>(src), R, C, NDIM, buf, buflen);
    if (kind == 1) return write_array_text<beadLocator>(static_cast<const int*>(src), R, C, 2, buf, buflen);
    return write_array_text<unsigned int>(static_cast<const unsigned int*>(src), R, C, 1, buf, buflen);
}
int refcpu_read_array(int kind, const char* text, void* dst, int max_elems, int* R, int* C) {
    if (kind == 0) return read_array_text<dVec>(text, static_cast<double*>(dst), max_elems, NDIM, R, C);
    if (kind == 1) return read_array_text<beadLocator>(text, static_cast<int*>(dst), max_elems, 2, R, C);
    return read_array_text<unsigned int>(text, static_cast<unsigned int*>(dst), max_elems, 1, R, C);
}

#if NDIM == 3
// LocalAction over the upstream Aziz class.  Per slice: vint[M] (V(slice)[1]), f2[M] (gradVSquared), sephist[M][50],
// vir[M][4] = {rDOTgradUterm1, rDOTgradUterm2, deltaDOTgradUterm1, deltaDOTgradUterm2} WITH upstream's prefactors,
// dtau[M] / dlam[M] / d2tau[M] / vkc[M] (derivPotentialActionTau / Lambda, secondderiv, virKinCorr); scalars[0] =
// potentialAction().  energy[9] / virial[19]: EnergyEstimator / VirialEnergyEstimator::accumulate increments (any
// pointer may be NULL).
int refcpu_action(int year, const double* side, const double* beads, int M, int N, int Next, const int* next, double tau,
                  double lambda, double mu, int window, const double* VF, const double* GF, int period, double* vint, double* f2,
                  int* sephist, double* vir, double* dtau, double* dlam, double* d2tau, double* vkc, double* scalars,
                  double* energy, double* virial, double spring_k) {
    const Container box = make_container(side, nullptr);
    set_constants(M, tau, lambda, mu, side[NDIM - 1], box.volume, window);       // rc defaults to side[NDIM-1] (src/setup.cpp:1128-1130)
    Path path(&box, M, N, Next, beads, next);
    FreePotential freeExternal;
    SpringTestPotential spring(spring_k);
    PotentialBase& external = spring_k != 0.0 ? static_cast<PotentialBase&>(spring) : static_cast<PotentialBase&>(freeExternal);
    AzizPotential aziz(year, &box);
    LocalAction action(path, &external, &aziz, {VF[0], VF[1]}, {GF[0], GF[1]}, period);
    for (int s = 0; s < M; ++s) {
        if (vint || sephist) {
            const std::array<double, 2> v = action.V(s);
            if (vint) vint[s] = v[1];
            if (sephist) for (int k = 0; k < NPCFSEP; ++k) sephist[static_cast<size_t>(s) * NPCFSEP + k] = action.sepHist(k);
        }
        if (f2) f2[s] = action.gradVSquared(s);
        if (vir) {
            vir[4 * s + 0] = action.rDOTgradUterm1(s);
            vir[4 * s + 1] = action.rDOTgradUterm2(s);
            vir[4 * s + 2] = action.deltaDOTgradUterm1(s);
            vir[4 * s + 3] = action.deltaDOTgradUterm2(s);
        }
        if (dtau) dtau[s] = action.derivPotentialActionTau(s);
        if (dlam) dlam[s] = action.derivPotentialActionLambda(s);
        if (d2tau) d2tau[s] = action.secondderivPotentialActionTau(s);
        if (vkc) vkc[s] = action.virKinCorr(s);
    }
    if (scalars) scalars[0] = action.potentialAction();
    if (energy) {
        EnergyEstimator e(path, &action, 0.0);
        e.initialize({"K", "V", "V_ext", "V_int", "E", "E_mu", "K/N", "V/N", "E/N"});            // src/estimator.cpp:917-926
        e.accumulate();
        for (int k = 0; k < 9; ++k) energy[k] = e.estimator(k);
    }
    if (virial) {
        VirialEnergyEstimator e(path, &action, 0.0);
        e.initialize({"K_op", "K_cv", "V_op", "V_cv", "E", "E_mu", "K_op/N", "K_cv/N", "V_op/N", "V_cv/N", "E/N", "EEcv*Beta^2",
                      "Ecv*Beta", "dEdB", "CvCov1", "CvCov2", "CvCov3", "E_th", "P"});               // src/estimator.cpp:1058-1060
        e.estimator.resize(20);        // upstream's misspelt key "cVCov2" creates map slot... with index 0; keep room anyway
        e.estimator.fill(0.0);
        e.accumulate();
        for (int k = 0; k < 19; ++k) virial[k] = e.estimator(k);
    }
    return 0;
}
#endif

}  // extern "C"
