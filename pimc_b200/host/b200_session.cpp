#include "b200_session.h"

#include <cstdlib>
#include <iostream>
#include <map>
#include <memory>

struct SessionRegistry {
    std::map<const Path*, B200Session*> map;
    ~SessionRegistry() { for (auto& kv : map) delete kv.second; }
};
static SessionRegistry& registry() { static SessionRegistry r; return r; }

// Path::get_beads_data_pointer() returns the dVec (= std::array<double, NDIM>) storage of DynamicArray<dVec,2> beads
// (include/path.h:208-210, include/dynamic_array.h:209); the C ABI takes the same bytes as flat doubles
// [M][N_ext][NDIM] -- contiguity is what upstream itself relies on when it hands this pointer to its GPU kernels
// (src/estimator.cpp:4084-4085).
static_assert(sizeof(dVec) == NDIM * sizeof(double), "dVec must be NDIM packed doubles");
static const double* beadsOf(const Path& path) { return reinterpret_cast<const double*>(path.get_beads_data_pointer()); }

// ABI failures follow the reference's error convention: message on std::cerr, then exit(EXIT_FAILURE)
// (src/estimator.cpp:450-455); nothing is ever wrapped in assert() (cf. GPU_ASSERT, include/common_gpu.h:55).
void B200Session::check(int rc, const char* what) const {
    if (rc == 0) return;
    std::cerr << "\nERROR: pimc_b200: " << what << " failed (" << rc << "): " << pimcb_last_error() << std::endl;
    std::exit(EXIT_FAILURE);
}

B200Session::B200Session(const Path& path) : path_(path) {
    int device = 0;
    if (const char* env = std::getenv("PIMCB_DEVICE")) device = std::atoi(env);
    check(pimcb_create(&ctx_, device, NDIM), "pimcb_create");
    double side[NDIM];
    unsigned periodic[NDIM];
    for (int d = 0; d < NDIM; ++d) {
        side[d] = path.boxPtr->side[d];
        periodic[d] = path.boxPtr->periodic[d];
    }
    check(pimcb_set_box(ctx_, side, periodic), "pimcb_set_box");
}

B200Session::~B200Session() {
    if (locked_ptr_) pimcb_host_unregister(const_cast<void*>(locked_ptr_));
    pimcb_destroy(ctx_);
}

B200Session& B200Session::get(const Path& path) {
    auto& m = registry().map;
    auto it = m.find(&path);
    if (it == m.end()) it = m.emplace(&path, new B200Session(path)).first;
    return *it->second;
}

void B200Session::newConfiguration(const Path& path) {
    B200Session& s = get(path);
    s.hooked_ = true;
    s.invalidate();
}

void B200Session::shutdown() {
    auto& m = registry().map;
    for (auto& kv : m) delete kv.second;
    m.clear();
}

void B200Session::setQVectors(const std::vector<dVec>& q) {
    std::vector<double> flat(q.size() * NDIM);
    for (size_t k = 0; k < q.size(); ++k)
        for (int d = 0; d < NDIM; ++d) flat[k * NDIM + d] = q[k][d];
    check(pimcb_set_qvecs(ctx_, flat.data(), static_cast<int>(q.size())), "pimcb_set_qvecs");
    nq_ = q.size();
    have_sf_ = false;
    have_es_ = have_cyl_ = false;
}

void B200Session::stageIfNeeded() {
    if (staged_) return;
    const auto ext = path_.get_beads_extents();
    const int M = path_.numTimeSlices;
    const int N = path_.getTrueNumParticles();           // diagonal configuration: N active beads on every slice
    // Page-lock Path::beads once (again whenever the array was reallocated, i.e. after particle insertions grew it):
    // staging is then a single DMA of the reference array as it lies, no host-side repacking.  A refused registration
    // only means the bounce-buffer path is used.
    const void* base = beadsOf(path_);
    const size_t bytes = sizeof(double) * static_cast<size_t>(M) * ext[1] * NDIM;
    if (base != locked_ptr_ || bytes != locked_bytes_) {
        if (locked_ptr_) pimcb_host_unregister(const_cast<void*>(locked_ptr_));
        locked_ptr_ = nullptr;
        locked_bytes_ = 0;
        if (!std::getenv("PIMCB_NO_PAGE_LOCK") && pimcb_host_register(const_cast<void*>(base), bytes) == 0) {
            locked_ptr_ = base;
            locked_bytes_ = bytes;
        }
    }
    if (fused_ssf_out_) {          // ssf()/isf() asked for the staging: one call, one synchronisation
        check(pimcb_ssf_isf_beads(ctx_, beadsOf(path_), M, N, static_cast<int>(ext[1]), fused_ssf_out_, fused_isf_out_),
              "pimcb_ssf_isf_beads");
        fused_done_ = true;
    } else {
        check(pimcb_stage_beads(ctx_, beadsOf(path_), M, N, static_cast<int>(ext[1])), "pimcb_stage_beads");
    }
    staged_ = true;
    have_sf_ = have_pair_ = false;
    have_es_ = have_cyl_ = have_vir_ = false;
}

const std::vector<double>& B200Session::ssf() {
    if (!have_sf_) {
        const int M = path_.numTimeSlices;
        ssf_.resize(nq_);
        isf_.resize(nq_ * M);
        fused_ssf_out_ = ssf_.data();
        fused_isf_out_ = isf_.data();
        fused_done_ = false;
        stageIfNeeded();                       // not yet staged: stages and evaluates in one call
        fused_ssf_out_ = fused_isf_out_ = nullptr;
        if (!fused_done_) check(pimcb_ssf_isf(ctx_, ssf_.data(), isf_.data()), "pimcb_ssf_isf");
        have_sf_ = true;
    }
    return ssf_;
}

const std::vector<double>& B200Session::isf() {
    ssf();
    return isf_;
}

void B200Session::setPairTable(const double* V, const double* dVdr, int len, double dr, const double* extV,
                               const double* extdVdr) {
    check(pimcb_set_pair_table(ctx_, V, dVdr, len, dr, extV, extdVdr), "pimcb_set_pair_table");
    have_table_ = true;
    have_pair_ = false;
}

const B200Session::PairSums& B200Session::pairSums(double dSep, bool wantF2, int f2Parity, const std::vector<double>* gext) {
    if (!(have_pair_ && (pair_has_f2_ || !wantF2))) {
        stageIfNeeded();
        if (wantF2) check(pimcb_set_external_gradient(ctx_, gext ? gext->data() : nullptr), "pimcb_set_external_gradient");
        const int M = path_.numTimeSlices;
        pair_.vint.assign(M, 0.0);
        pair_.f2.assign(M, 0.0);
        pair_.hist.assign(static_cast<size_t>(M) * NPCFSEP, 0);
        check(pimcb_pair_sums(ctx_, pair_.vint.data(), wantF2 ? pair_.f2.data() : nullptr, pair_.hist.data(), dSep, f2Parity),
              "pimcb_pair_sums");
        have_pair_ = true;
        pair_has_f2_ = wantF2;
    }
    return pair_;
}

// ---- scattering variants ---------------------------------------------------------------------------------------------
const std::vector<double>& B200Session::elastic() {
    if (!have_es_) {
        ssf();                                 // stages if needed and leaves S(q)/F(q,tau) of this configuration on the device
        es_.resize(nq_);
        check(pimcb_elastic(ctx_, es_.data()), "pimcb_elastic");
        have_es_ = true;
    }
    return es_;
}

const std::vector<double>& B200Session::ssfCylinder(double maxR, int& numInside) {
    if (!(have_cyl_ && cyl_maxR_ == maxR)) {
        stageIfNeeded();
        cyl_.resize(nq_);
        check(pimcb_ssf_cyl(ctx_, maxR, cyl_.data(), &cyl_inside_), "pimcb_ssf_cyl");
        cyl_maxR_ = maxR;
        have_cyl_ = true;
    }
    numInside = cyl_inside_;
    return cyl_;
}

// ---- virial slice sums -------------------------------------------------------------------------------------------------
void B200Session::setPairTableD2(const double* d2Vdr2, int len, const double* extd2Vdr2) {
    check(pimcb_set_pair_table_d2(ctx_, d2Vdr2, len, extd2Vdr2), "pimcb_set_pair_table_d2");
    have_d2_ = true;
    have_vir_ = false;
}

const std::vector<double>& B200Session::virialSums(int window, int t2Parity, const std::vector<double>* gext,
                                                   const std::vector<double>* g2ext) {
    if (!(have_vir_ && vir_window_ == window && vir_parity_ == t2Parity)) {
        stageIfNeeded();
        const auto ext = path_.get_beads_extents();
        const int M = path_.numTimeSlices;
        const int Next = static_cast<int>(ext[1]);
        // deviation of every bead from the centroid of its world-line window, WITHOUT mirror image for the centroid and
        // with putInBC on the result (src/action.cpp:1620-1647 == 1742-1769)
        delta_.assign(static_cast<size_t>(M) * Next * NDIM, 0.0);
        beadLocator bead1, beadNext, beadPrev, beadNextOld, beadPrevOld;
        for (bead1[0] = 0; bead1[0] < M; bead1[0]++) {
            const int numParticles = path_.numBeadsAtSlice(bead1[0]);
            for (bead1[1] = 0; bead1[1] < numParticles; bead1[1]++) {
                dVec runTotMore{}, runTotLess{}, COM{};
                const dVec pos1 = path_(bead1);
                beadNextOld = bead1;
                beadPrevOld = bead1;
                for (int gamma = 0; gamma < window; gamma++) {
                    beadNext = path_.next(bead1, gamma);
                    beadPrev = path_.prev(bead1, gamma);
                    const dVec more = path_.getSeparation(beadNext, beadNextOld);
                    const dVec less = path_.getSeparation(beadPrev, beadPrevOld);
                    for (int d = 0; d < NDIM; ++d) {
                        runTotMore[d] += more[d];
                        runTotLess[d] += less[d];
                        COM[d] += (pos1[d] + runTotMore[d]) + (pos1[d] + runTotLess[d]);
                    }
                    beadNextOld = beadNext;
                    beadPrevOld = beadPrev;
                }
                dVec delta;
                for (int d = 0; d < NDIM; ++d) {
                    COM[d] /= (2.0 * window);
                    delta[d] = pos1[d] - COM[d];
                }
                path_.boxPtr->putInBC(delta);
                double* out = &delta_[(static_cast<size_t>(bead1[0]) * Next + bead1[1]) * NDIM];
                for (int d = 0; d < NDIM; ++d) out[d] = delta[d];
            }
        }
        vir_.assign(static_cast<size_t>(M) * 4, 0.0);
        check(pimcb_set_external_gradient(ctx_, gext ? gext->data() : nullptr), "pimcb_set_external_gradient");
        check(pimcb_set_external_laplacian(ctx_, g2ext ? g2ext->data() : nullptr), "pimcb_set_external_laplacian");
        check(pimcb_virial_sums(ctx_, delta_.data(), have_d2_ ? t2Parity : -2, vir_.data()), "pimcb_virial_sums");
        vir_window_ = window;
        vir_parity_ = t2Parity;
        have_vir_ = true;
    }
    return vir_;
}

// ---- batched accumulation and the multi-GPU exchange step ------------------------------------------------------------
void B200Session::measureBatch(const double* beads, int B, int M, int N, int Next) {
    check(pimcb_stage_batch(ctx_, beads, B, M, N, Next), "pimcb_stage_batch");
    check(pimcb_measure(ctx_), "pimcb_measure");
    invalidate();                              // the staged slot no longer is "the current configuration of the path"
}

void B200Session::readBins(std::vector<double>& ssf, std::vector<double>& isf, long& count) {
    const int M = path_.numTimeSlices;
    ssf.assign(nq_, 0.0);
    isf.assign(nq_ * M, 0.0);
    check(pimcb_read_bins(ctx_, ssf.data(), isf.data(), &count), "pimcb_read_bins");
}

void B200Session::resetBins() { check(pimcb_reset_bins(ctx_), "pimcb_reset_bins"); }

void B200Session::initBins() { check(pimcb_init_bins(ctx_, path_.numTimeSlices), "pimcb_init_bins"); }

void B200Session::uniqueId(void* id128) {
    if (pimcb_comm_unique_id(id128) != 0) {
        std::cerr << "\nERROR: pimc_b200: pimcb_comm_unique_id failed: " << pimcb_last_error() << std::endl;
        std::exit(EXIT_FAILURE);
    }
}

void B200Session::commInit(int nranks, int rank, const void* id128) {
    check(pimcb_comm_init(ctx_, nranks, rank, id128), "pimcb_comm_init");
}

long B200Session::reduceBins(int root) {
    long total = 0;
    check(pimcb_reduce_bins(ctx_, root, &total), "pimcb_reduce_bins");
    return total;
}

void B200Session::reduceBinsBegin(int root) {
    check(pimcb_reduce_bins_begin(ctx_, root), "pimcb_reduce_bins_begin");
}

long B200Session::reduceBinsEnd(std::vector<double>& ssf, std::vector<double>& isf) {
    const int M = path_.numTimeSlices;
    long total = 0;
    ssf.assign(nq_, 0.0);
    isf.assign(nq_ * M, 0.0);
    check(pimcb_reduce_bins_end(ctx_, ssf.data(), isf.data(), &total), "pimcb_reduce_bins_end");
    return total;
}
