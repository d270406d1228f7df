// pimc_compat.h -- minimal stand-ins for the reference types the B200 adaptor classes touch.
//
// Inside a reference checkout the adaptor sources (estimator_b200.*, action_b200.*) are compiled against the
// reference's own headers (common.h, path.h, container.h, constants.h, estimator.h, action.h, factory.h) and this
// file is not used.  Stand-alone (this repository: no Boost, no <mdspan>) it supplies from-scratch types with the
// same names, members and call signatures, restricted to what the measurement path needs, so that the adaptor code
// is compile-checked and runnable (tools: pimcb_measure) here.  Citations are upstream file:line.
#ifndef PIMCB_COMPAT_H
#define PIMCB_COMPAT_H

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "table_view.h"

#ifndef NDIM
#define NDIM 3                      // CMakeLists.txt:223-226 (compile-time dimension upstream)
#endif
#define NPCFSEP 50                  // include/common.h:85
#define EPS 1.0E-7                  // include/common.h:89
#define XXX -1                      // include/common.h:93 (nonsense bead index)

typedef unsigned int uint32;
typedef std::array<double, NDIM> dVec;      // include/common.h:104
typedef std::array<int, NDIM> iVec;
typedef std::array<int, 2> beadLocator;     // include/common.h:110

// Rank-1 subset of the reference's DynamicArray<T,Rank> (include/dynamic_array.h:55-386): element access with
// operator()(i), resize, fill, data, size.
template <class T, int Rank> class DynamicArray;
template <class T>
class DynamicArray<T, 1> {
public:
    void resize(size_t n) { v_.resize(n); }
    void fill(const T& x) { std::fill(v_.begin(), v_.end(), x); }
    T& operator()(size_t i) { return v_[i]; }
    const T& operator()(size_t i) const { return v_[i]; }
    T* data() { return v_.data(); }
    const T* data() const { return v_.data(); }
    size_t size() const { return v_.size(); }
private:
    std::vector<T> v_;
};

inline double dot(const dVec& a, const dVec& b) {      // include/array_math.h:210-216
    double r = 0.0;
    for (int i = 0; i < NDIM; ++i) r += a[i] * b[i];
    return r;
}

// ---- Container (include/container.h:24-59, src/container.cpp:84-144) ------------------------------------------
class Container {
public:
    dVec side{}, sideInv{}, pSide{};
    std::array<unsigned int, NDIM> periodic{};
    double maxSep = 0.0, volume = 0.0;
    void putInBC(dVec& r) const {
        for (int i = 0; i < NDIM; ++i) r[i] -= pSide[i] * std::floor(r[i] * sideInv[i] + 0.5);
    }
    // Prism::putInside, include/container.h:118-135: periodic wrap, then clamp the non-periodic dimensions
    void putInside(dVec& r) const {
        putInBC(r);
        for (int i = 0; i < NDIM; ++i)
            if (!periodic[i]) {
                if (r[i] >= 0.5 * side[i]) r[i] = 0.5 * side[i] - 2 * EPS;
                if (r[i] < -0.5 * side[i]) r[i] = -0.5 * side[i] + 2 * EPS;
            }
    }
};

class Prism : public Container {
public:
    Prism(double density, int numParticles) {
        side.fill(std::pow(1.0 * numParticles / density, 1.0 / (1.0 * NDIM)));
        periodic.fill(1u);
        finish();
    }
    Prism(const dVec& _side, const std::array<unsigned int, NDIM>& _periodic) {
        side = _side;
        periodic = _periodic;
        finish();
    }
private:
    void finish() {
        double acc = 0.0;
        volume = 1.0;
        for (int i = 0; i < NDIM; ++i) {
            sideInv[i] = 1.0 / side[i];
            pSide[i] = periodic[i] * side[i];
            const double h = side[i] / (periodic[i] + 1u);
            acc += h * h;
            volume *= side[i];
        }
        maxSep = std::sqrt(acc);
    }
};

// ---- simulation constants (include/constants.h; only the getters the path reads) ----------------------------
class ConstantParameters {
public:
    bool canonical() const { return canonical_; }
    bool restart() const { return false; }
    int initialNumParticles() const { return initialNumParticles_; }
    int numTimeSlices() const { return numTimeSlices_; }
    double tau() const { return tau_; }
    double T() const { return T_; }
    double lambda() const { return lambda_; }                  // src/constants.cpp:128: 24.24/m
    double mu() const { return mu_; }
    double rc() const { return rc_; }                          // potential cutoff; defaults to side[NDIM-1] (src/setup.cpp:1128-1130)
    double fourLambdaTauInv() const { return 0.25 / (lambda_ * tau_); }   // include/constants.h:82
    uint32 binSize() const { return binSize_; }
    int virialWindow() const { return virialWindow_; }         // --virial_window, default 5 (src/setup.cpp:450)
    double V() const { return V_; }                            // cell volume (include/constants.h:70)
    int virialWindow_ = 5;
    double V_ = 0.0;
    double mu_ = 0.0, rc_ = 0.0;
    uint32 binSize_ = 100;
    std::string wavevector() const { return wavevector_; }
    std::string wavevectorType() const { return wavevectorType_; }
    std::string id() const { return id_; }
    bool canonical_ = true;
    int initialNumParticles_ = 0, numTimeSlices_ = 0;
    double tau_ = 0.0, T_ = 0.0, lambda_ = 24.24 / 4.0030;
    std::string wavevector_, wavevectorType_ = "int", id_ = "000000000";
};
ConstantParameters* constants();

// ---- worm + path (include/worm.h, include/path.h:29-217) -----------------------------------------------------
class Worm {
public:
    bool isConfigDiagonal = true;
    int numBeadsOn = 0;
    int getNumBeadsOn() const { return numBeadsOn; }
};

class Path {
public:
    Path(const Container* box, int numTimeSlices_, int numParticles, int extent)
        : numTimeSlices(numTimeSlices_), boxPtr(box), n_(numParticles), next_(extent),
          beads_(static_cast<size_t>(numTimeSlices_) * extent), nextLink_(static_cast<size_t>(numTimeSlices_) * extent) {
        worm.numBeadsOn = numTimeSlices * numParticles;
        for (int s = 0; s < numTimeSlices; ++s)            // straight closed world lines unless setLinks() says otherwise
            for (int p = 0; p < extent; ++p)
                nextLink_[static_cast<size_t>(s) * extent + p] =
                    p < numParticles ? beadLocator{(s + 1) % numTimeSlices, p} : beadLocator{XXX, XXX};
        buildPrev();
    }
    void setLinks(const std::vector<beadLocator>& nextLink) {
        if (nextLink.size() != nextLink_.size()) return;
        nextLink_ = nextLink;
        buildPrev();
    }
    const beadLocator& next(const beadLocator& b) const { return nextLink_[static_cast<size_t>(b[0]) * next_ + b[1]]; }   // path.h:92-99
    const beadLocator& prev(const beadLocator& b) const { return prevLink_[static_cast<size_t>(b[0]) * next_ + b[1]]; }   // path.h:101-107
    beadLocator next(const beadLocator& b, int numLinks) const {                                 // path.h:233-240
        beadLocator bI = b;
        for (int m = 0; m < numLinks; m++) bI = next(bI);
        return bI;
    }
    beadLocator prev(const beadLocator& b, int numLinks) const {                                 // path.h:256-263
        beadLocator bI = b;
        for (int m = 0; m < numLinks; m++) bI = prev(bI);
        return bI;
    }
    dVec getVelocity(const beadLocator& b) const {                                              // path.h:189-203
        dVec vel;
        const beadLocator& n = next(b);
        if ((b[0] == XXX && b[1] == XXX) || (n[0] == XXX && n[1] == XXX)) { vel.fill(0.0); return vel; }
        for (int i = 0; i < NDIM; ++i) vel[i] = (*this)(n)[i] - (*this)(b)[i];
        boxPtr->putInBC(vel);
        return vel;
    }
    const int numTimeSlices;
    const Container* boxPtr;
    Worm worm;
    int numBeadsAtSlice(int) const { return n_; }
    int getTrueNumParticles() const { return worm.getNumBeadsOn() / numTimeSlices; }          // path.h:54
    const dVec& operator()(int slice, int ptcl) const { return beads_[static_cast<size_t>(slice) * next_ + ptcl]; }
    dVec& operator()(int slice, int ptcl) { return beads_[static_cast<size_t>(slice) * next_ + ptcl]; }
    const dVec& operator()(const beadLocator& b) const { return (*this)(b[0], b[1]); }
    dVec getSeparation(const beadLocator& b1, const beadLocator& b2) const {                   // path.h:179-184
        dVec sep;
        for (int i = 0; i < NDIM; ++i) sep[i] = (*this)(b1)[i] - (*this)(b2)[i];
        boxPtr->putInBC(sep);
        return sep;
    }
    const dVec* get_beads_data_pointer() const { return beads_.data(); }                        // path.h:208-210 (typed as upstream)
    double* beads_data() { return reinterpret_cast<double*>(beads_.data()); }
    std::array<size_t, 2> get_beads_extents() const { return {static_cast<size_t>(numTimeSlices), static_cast<size_t>(next_)}; }
private:
    int n_, next_;
    std::vector<dVec> beads_;       // row-major [slice][ptcl] AoS, as DynamicArray<dVec,2> (path.h:164)
    std::vector<beadLocator> nextLink_, prevLink_;
    void buildPrev() {                  // prevLink is the inverse map of nextLink on a closed configuration
        prevLink_.assign(nextLink_.size(), beadLocator{XXX, XXX});
        for (int s = 0; s < numTimeSlices; ++s)
            for (int p = 0; p < next_; ++p) {
                const beadLocator& n = nextLink_[static_cast<size_t>(s) * next_ + p];
                if (n[0] != XXX && n[1] != XXX) prevLink_[static_cast<size_t>(n[0]) * next_ + n[1]] = beadLocator{s, p};
            }
    }
};

class MTRand {};                    // the path never draws random numbers

// ---- potentials (include/potential.h:40-117) -------------------------------------------------------------------
class PotentialBase {
public:
    virtual ~PotentialBase() {}
    virtual double V(const dVec&) { return 0.0; }
    virtual void V(const dVec* pos, double* values, int count) {          // src/potential.cpp:127-132
        if (count < 0) throw std::runtime_error("negative count");
        for (int i = 0; i < count; ++i) values[i] = V(pos[i]);
    }
    virtual dVec gradV(const dVec&) { return dVec{}; }
    virtual double grad2V(const dVec&) { return 0.0; }
    virtual TableView tableView() const { return TableView{}; }      // upstream.patch adds the same virtual (no table by default)
    double tailV = 0.0;
};
class FreePotential : public PotentialBase {};
// A non-trivial external potential for exercising the gradient coupling in the stand-alone tools: V = k r^2 / 2
// (upstream's HarmonicPotential has the same form with k = omega^2 / (2 lambda), src/potential.cpp).
class SpringPotential : public PotentialBase {
public:
    explicit SpringPotential(double k_) : k(k_) {}
    double V(const dVec& r) override { return 0.5 * k * dot(r, r); }
    dVec gradV(const dVec& r) override { dVec g; for (int i = 0; i < NDIM; ++i) g[i] = k * r[i]; return g; }
    double grad2V(const dVec&) override { return NDIM * k; }
private:
    double k;
};

// ---- action (include/action.h:30-254; the members the measurement path uses) -----------------------------------
class LookupTable {};               // include/lookuptable.h: nearest-neighbour grid used by the moves, not by the measurement
class WaveFunctionBase {};          // include/wavefunction.h: PIGS trial wave functions, unused in PIMC mode

class ActionBase {
public:
    // same parameter list as upstream (include/action.h:33-35)
    ActionBase(const Path& p, LookupTable& _lookup, PotentialBase* ext, PotentialBase* inter, WaveFunctionBase* wf,
               bool _local = true, std::string _name = "Base", double _endFactor = 1.0, int _period = 1)
        : local(_local), period(_period), externalPtr(ext), interactionPtr(inter), name(_name), lookup(_lookup), path(p),
          waveFunctionPtr(wf), endFactor(_endFactor) {
        sepHist.resize(NPCFSEP);
        sepHist.fill(0);
        dSep = 0.5 * std::sqrt(1.0 * NDIM) * path.boxPtr->side[NDIM - 1] / (1.0 * NPCFSEP);      // src/action.cpp:192
    }
    virtual ~ActionBase() {}
    std::string getActionName() { return name; }
    virtual double potentialAction() { return 0.0; }
    virtual std::array<double, 2> potential(int) { return {0.0, 0.0}; }
    virtual double derivPotentialActionTau(int) { return 0.0; }
    virtual double derivPotentialActionLambda(int) { return 0.0; }
    virtual double secondderivPotentialActionTau(int) { return 0.0; }     // include/action.h:63
    virtual double rDOTgradUterm1(int) { return 0.0; }                    // include/action.h:75-85
    virtual double rDOTgradUterm2(int) { return 0.0; }
    virtual double deltaDOTgradUterm1(int) { return 0.0; }
    virtual double deltaDOTgradUterm2(int) { return 0.0; }
    virtual double virKinCorr(int) { return 0.0; }
    const bool local;
    const int period;
    PotentialBase* externalPtr;     // public upstream as well (include/action.h:108-109)
    PotentialBase* interactionPtr;
    DynamicArray<int, 1> sepHist;   // action.h:114
protected:
    std::string name;
    LookupTable& lookup;
    const Path& path;
    WaveFunctionBase* waveFunctionPtr;
    double endFactor;
    int shift = 1;                  // PIMC mode (include/action.h:139-149)
    double tau() { return shift * constants()->tau(); }
    double dSep;
};

// include/action.h:157-254: the bead-level members (moves) are upstream code and stay on the host there; the stand-alone
// measurement tools never call them
class LocalAction : public ActionBase {
public:
    LocalAction(const Path& p, LookupTable& _lookup, PotentialBase* ext, PotentialBase* inter, WaveFunctionBase* wf,
                const std::array<double, 2>& _VFactor, const std::array<double, 2>& _gradVFactor, bool _local = true,
                std::string _name = "Local", double _endFactor = 1.0, int _period = 1)
        : ActionBase(p, _lookup, ext, inter, wf, _local, _name, _endFactor, _period), VFactor(_VFactor), gradVFactor(_gradVFactor) {}
protected:
    int eo = 0;
    std::array<double, 2> VFactor, gradVFactor;
};

// ---- output files (include/communicator.h; file-name pattern src/communicator.cpp:39-44,160-167) ----------------
class File {
public:
    explicit File(const std::string& name) : name_(name), stream_(name, std::ios::out | std::ios::trunc) {}
    std::fstream& stream() { return stream_; }
    bool exists() const { return false; }
    bool prepared() const { return prepared_; }
    void prepare() { prepared_ = true; }
    const std::string& name() const { return name_; }
private:
    std::string name_;
    std::fstream stream_;
    bool prepared_ = false;
};
class Communicator {
public:
    void init(const std::string& dir, const std::string& ensemble, const std::string& dataName) {
        dir_ = dir; ensemble_ = ensemble; dataName_ = dataName;
    }
    File* file(const std::string& label);
private:
    std::string dir_ = "OUTPUT", ensemble_ = "ce", dataName_ = "run";
    std::map<std::string, std::unique_ptr<File>> files_;
};
Communicator* communicate();

// ---- factory (include/factory.h:26-95) ---------------------------------------------------------------------------
template <typename CtorSignature> class Factory;
template <class BaseType, class... ParamType>
class Factory<BaseType(ParamType...)> {
    using CreateObjectFunc = BaseType (*)(ParamType...);
public:
    std::vector<std::string> getNames() const {
        std::vector<std::string> names;
        for (auto const& kv : _create) names.push_back(kv.first);
        return names;
    }
    Factory* operator()() { static Factory f; return &f; }
    BaseType Create(std::string name, ParamType... param) {
        auto it = _create.find(name);
        return it != _create.end() ? (it->second)(param...) : nullptr;
    }
    template <class DerivedType> bool Register(std::string name) {
        _create[name] = &createObj<DerivedType>;
        return true;
    }
private:
    std::map<std::string, CreateObjectFunc> _create;
    template <class DerivedType> static BaseType createObj(ParamType... param) { return new DerivedType(param...); }
};
class EstimatorBase;
typedef Factory<EstimatorBase*(Path&, ActionBase*, MTRand&, double)> EstimatorFactory;
extern EstimatorFactory estimatorFactory;

#endif
