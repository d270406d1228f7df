// estimator_base.cpp -- see estimator_base.h.  Behaviour follows src/estimator.cpp:163-429 and :439-570 upstream.
#include "estimator_base.h"

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <iterator>
#include <sys/stat.h>

EstimatorFactory estimatorFactory;

ConstantParameters* constants() { static ConstantParameters c; return &c; }
Communicator* communicate() { static Communicator c; return &c; }

File* Communicator::file(const std::string& label) {
    auto it = files_.find(label);
    if (it == files_.end()) {
        ::mkdir(dir_.c_str(), 0755);
        // src/communicator.cpp:39-44,160-167: <dir>/<ensemble>-<label>-<dataName>.dat
        const std::string name = dir_ + "/" + ensemble_ + "-" + label + "-" + dataName_ + ".dat";
        it = files_.emplace(label, std::make_unique<File>(name)).first;
    }
    return it->second.get();
}

std::string pimcb_format(const char* spec, double v) {
    char buf[64];
    std::snprintf(buf, sizeof buf, spec, v);
    return buf;
}
std::string pimcb_format(const char* spec, int v) {
    char buf[64];
    std::snprintf(buf, sizeof buf, spec, v);
    return buf;
}

// src/estimator.cpp:163-201 (PIMC branch: averages run over all slices)
EstimatorBase::EstimatorBase(const Path& _path, ActionBase* _actionPtr, const MTRand& _random, double _maxR, int _frequency,
                             std::string _label)
    : path(_path), actionPtr(_actionPtr), random(_random), maxR(_maxR), frequency(_frequency), label(_label),
      numSampled(0), numAccumulated(0), totNumAccumulated(0), diagonal(true), endLine(true) {
    canonical = constants()->canonical();
    numBeads0 = constants()->initialNumParticles() * constants()->numTimeSlices();
    sliceFactor.assign(constants()->numTimeSlices(), 1.0);      // :182, :196-199 (PIMC: every slice, weight 1)
    startSlice = 0;
    endSlice = path.numTimeSlices;
    endDiagSlice = endSlice;
}

EstimatorBase::~EstimatorBase() {}

// src/estimator.cpp:219-236
bool EstimatorBase::baseSample() {
    numSampled++;
    if (!frequency) return false;
    if ((numSampled % frequency) != 0) return false;
    if (!(path.worm.isConfigDiagonal == diagonal)) return false;
    if (!canonical) return true;
    if (path.worm.getNumBeadsOn() == numBeads0) return true;
    return false;
}

// src/estimator.cpp:245-252
void EstimatorBase::sample() {
    if (baseSample()) {
        totNumAccumulated++;
        numAccumulated++;
        accumulate();
    }
}

// src/estimator.cpp:259-266
void EstimatorBase::initialize(int _numEst) {
    numEst = _numEst;
    estimator.resize(numEst);
    norm.resize(numEst);
    norm.fill(1.0);
    reset();
}

// src/estimator.cpp:274-289
void EstimatorBase::initialize(std::vector<std::string> estLabel) {
    for (size_t i = 0; i < estLabel.size(); ++i) estIndex[estLabel[i]] = static_cast<int>(i);
    header = "";
    for (const auto& l : estLabel) {
        char buf[64];
        std::snprintf(buf, sizeof buf, "%16s", l.c_str());
        header += buf;
    }
    initialize(static_cast<int>(estLabel.size()));
}

// src/estimator.cpp:297-321
void EstimatorBase::prepare() {
    if (frequency > 0) {
        File* f = communicate()->file(label);
        outFilePtr = &(f->stream());
        if (!constants()->restart() || !f->exists()) {
            if (!f->prepared()) {
                header.replace(header.begin(), header.begin() + 1, "#");
                f->prepare();
            }
            (*outFilePtr) << header;
            if (endLine) (*outFilePtr) << std::endl;
        }
    }
}

// src/estimator.cpp:326-329
void EstimatorBase::reset() {
    numAccumulated = 0;
    estimator.fill(0.0);
}

// src/estimator.cpp:334-338
void EstimatorBase::restart(const uint32 _numSampled, const uint32 _totNumAccumulated) {
    numSampled = _numSampled;
    totNumAccumulated = _totNumAccumulated;
    reset();
}

// src/estimator.cpp:348-362
void EstimatorBase::output() {
    for (int n = 0; n < numEst; n++) estimator(n) *= (norm(n) / (1.0 * numAccumulated));
    for (int n = 0; n < numEst; n++) (*outFilePtr) << pimcb_format("%16.8E", estimator(n));
    if (endLine) (*outFilePtr) << std::endl;
    reset();
}

// src/estimator.cpp:421-429
std::string EstimatorBase::dVecToString(const dVec& v) {
    std::string strVec = "(";
    for (int i = 0; i < NDIM; i++) {
        strVec += pimcb_format("%+15.8E", v[i]);
        if (i < NDIM - 1) strVec += ",";
    }
    return strVec + ")";
}

// Wave-vector list for the scattering estimators; same result and ORDER (it defines the output columns) as
// EstimatorBase::getQVectors upstream (src/estimator.cpp:439-570):
//   int        q_j = (2 pi / L_j) * stoi(token)                                         (:463)
//   float      q_j = stof(token)  -- float precision, then widened                       (:466)
//   max_int    all lattice vectors n with |n_j| <= nmax_j and |q| <= |q(nmax)|, enumerated from the all-negative
//              corner with the LAST dimension running fastest                            (:475-537)
//   max_float  same with nmax_j = 1 + int(|q_max| L_j / 2 pi)                             (:495-499)
//   file_int / file_float: one vector per line of the file named by `wavevector`         (:540-569)
void EstimatorBase::getQVectors(std::vector<dVec>& qValues) {
    const std::string text = constants()->wavevector();
    const std::string kind = constants()->wavevectorType();
    const dVec& L = path.boxPtr->side;
    auto lattice = [&L](const iVec& n) {
        dVec q;
        for (int d = 0; d < NDIM; ++d) q[d] = n[d] * 2.0 * M_PI / L[d];
        return q;
    };

    std::istringstream in(text);
    std::vector<std::string> tok{std::istream_iterator<std::string>{in}, std::istream_iterator<std::string>{}};
    if (tok.empty()) {
        std::cerr << "\nERROR: EstimatorBase::getQVectors: No input detected." << std::endl
                  << "Action: Ensure `wavevector` command line option is set." << std::endl;
        exit(1);
    }

    if (kind == "int" || kind == "float") {
        for (size_t base = 0; base + NDIM <= tok.size(); base += NDIM) {
            dVec q;
            for (int d = 0; d < NDIM; ++d)
                q[d] = kind == "int" ? (2.0 * M_PI / L[d]) * std::stoi(tok[base + d]) : std::stof(tok[base + d]);
            qValues.push_back(q);
        }
    } else if (kind == "max_int" || kind == "max_float") {
        iVec nmax{};
        dVec qmax{};
        for (int d = 0; d < NDIM; ++d) {
            if (kind == "max_int") {
                nmax[d] = std::abs(std::stoi(tok[d]));
                qmax[d] = nmax[d] * 2.0 * M_PI / L[d];
            } else {
                qmax[d] = std::stof(tok[d]);
            }
        }
        const double bound = std::sqrt(dot(qmax, qmax));
        if (kind == "max_float")
            for (int d = 0; d < NDIM; ++d) nmax[d] = 1 + static_cast<int>(bound * L[d] / 2.0 / M_PI);
        long total = 1;
        for (int d = 0; d < NDIM; ++d) total *= 2 * nmax[d] + 1;
        for (long idx = 0; idx < total; ++idx) {        // mixed-radix counter, last dimension least significant
            iVec n;
            long rest = idx;
            for (int d = NDIM - 1; d >= 0; --d) {
                const int radix = 2 * nmax[d] + 1;
                n[d] = static_cast<int>(rest % radix) - nmax[d];
                rest /= radix;
            }
            const dVec q = lattice(n);
            if (std::sqrt(dot(q, q)) <= bound) qValues.push_back(q);
        }
    } else if (kind == "file_int" || kind == "file_float") {
        std::ifstream file(text);
        std::string line;
        while (std::getline(file, line)) {
            std::istringstream ls(line);
            dVec q;
            if (kind == "file_int") {
                std::vector<int> v((std::istream_iterator<int>(ls)), std::istream_iterator<int>());
                if (static_cast<int>(v.size()) != NDIM) continue;
                for (int d = 0; d < NDIM; ++d) q[d] = (2.0 * M_PI / L[d]) * v[d];
            } else {
                std::vector<float> v((std::istream_iterator<float>(ls)), std::istream_iterator<float>());
                if (static_cast<int>(v.size()) != NDIM) continue;
                for (int d = 0; d < NDIM; ++d) q[d] = v[d];
            }
            qValues.push_back(q);
        }
    }
}

// src/estimator.cpp:762-837: wave-vectors grouped by magnitude.  Magnitudes 0, dq, 2 dq, ... (accumulated) up to
// qMax + EPS; the null vector alone in shell 0; every other shell starts with the vector along the last axis and, in
// three dimensions, continues over the positive octant in steps dtheta = pi/48 (geometry "sphere"; "line" uses
// dtheta = pi, i.e. no further vectors), dphi = dtheta / sin(theta).
std::vector<std::vector<dVec>> EstimatorBase::getQVectors2(double dq, double qMax, int& numq, std::string qGeometry) {
    numq = 0;
    std::vector<std::vector<dVec>> shells;
    if ((qGeometry != "line") && (qGeometry != "sphere")) {
        std::cerr << "\nERROR: A valid geometry wasn't chosen for q-space." << std::endl
                  << "Action: choose \"line\" or \"sphere\"" << std::endl;
        exit(1);
    }
    for (double cq = 0.0; cq <= qMax + EPS; cq += dq) {
        std::vector<dVec> shell;
        dVec axis{};
        if (!(std::abs(cq) < EPS)) axis[NDIM - 1] = cq;
        shell.push_back(axis);
#if NDIM == 3
        if (!(std::abs(cq) < EPS)) {
            const int numTheta = 24;
            const double dtheta = (qGeometry == "line") ? M_PI : 0.5 * M_PI / numTheta;
            for (double theta = dtheta; theta <= 0.5 * M_PI + EPS; theta += dtheta) {
                const double dphi = dtheta / sin(theta);
                for (double phi = 0.0; phi <= 0.5 * M_PI + EPS; phi += dphi) {
                    dVec qd;
                    qd[0] = cq * sin(theta) * cos(phi);
                    qd[1] = cq * sin(theta) * sin(phi);
                    qd[2] = cq * cos(theta);
                    shell.push_back(qd);
                }
            }
        }
#endif
        numq += static_cast<int>(shell.size());
        shells.push_back(shell);
    }
    std::cout << "numQ = " << numq << std::endl;
    return shells;
}
