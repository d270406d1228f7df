// b200_session.h -- one pimcb_ctx per Path, shared by every B200 estimator / action object that looks at that path,
// so that the beads are staged once per measured configuration instead of once per estimator (the shipped GPU
// estimators each copy the whole padded array per call, src/estimator.cpp:3833-3842, 4074-4085).
//
// Configuration identity.  The reference has no "path generation" counter; PathIntegralMonteCarlo::step() simply
// calls est.sample() for every estimator right after the move sweep (src/pimc.cpp:732-738).  Two modes:
//   * hooked:   the driver calls B200Session::newConfiguration(path) once per MC step before the estimator loop
//               (a one-line patch in pimc.cpp); all consumers of that step share one staging
//               and one S(q)/F(q,tau) (resp. pair-sum) pass.
//   * unhooked: (default) the session is a pure cache and every entry point that could be looking at a new
//               configuration (an estimator's accumulate(), potentialAction(), a per-slice call whose slice index
//               does not increase) calls beginIfUnhooked() first: always correct, costs one extra H2D + pass per
//               estimator.
#ifndef PIMCB_SESSION_H
#define PIMCB_SESSION_H

#include <vector>

#include "../../include/pimc_b200.h"
#ifdef PIMCB_STANDALONE
#include "pimc_compat.h"
#else
#include "common.h"
#include "path.h"
#include "container.h"
#endif

class B200Session {
public:
    static B200Session& get(const Path& path);
    static void newConfiguration(const Path& path);     // hook: call once per MC step before est.sample()
    static void shutdown();                             // destroys all contexts (end of main)

    // Wave-vectors are fixed per session (all scattering estimators of a run share --wavevector).
    void setQVectors(const std::vector<dVec>& q);
    int numQ() const { return static_cast<int>(nq_); }

    // S(q)/N [nq] and F(q,tau)/N [nq*M] of the current configuration (computed once per configuration).
    const std::vector<double>& ssf();
    const std::vector<double>& isf();

    // Pair sums of the current configuration: Vint[M], gradVSquared[M], sepHist[M][NPCFSEP].
    void setPairTable(const double* V, const double* dVdr, int len, double dr, const double* extV, const double* extdVdr);
    struct PairSums { std::vector<double> vint, f2; std::vector<int> hist; };
    // gext (may be NULL = "free"): gradient of the external potential per bead in the beads' own AoS shape, added to the
    // pair force inside gradVSquared (src/action.cpp:1216)
    const PairSums& pairSums(double dSep, bool wantF2, int f2Parity, const std::vector<double>* gext = nullptr);

    // Batched, device-resident accumulation (walker batches): B configurations in the reference bead layout
    // ([B][M][N_ext][NDIM]) are staged and measured without a read-back; the sums stay in the device bin until readBins.
    // Multi-GPU: one process per GPU, commInit on every rank (the 128-byte id comes from uniqueId on one of them), one
    // reduceBins per output bin.
    void measureBatch(const double* beads, int B, int M, int N, int Next);
    void readBins(std::vector<double>& ssf, std::vector<double>& isf, long& count);
    void resetBins();
    void initBins();                                    // zeroed bin before any measurement (idle ranks join the reduce)
    static void uniqueId(void* id128);
    void commInit(int nranks, int rank, const void* id128);
    long reduceBins(int root);
    // the same exchange pipelined over output bins: reduceBinsBegin snapshots the bin and starts its reduce on the
    // library's communication stream (resetBins and the next bin's measureBatch may follow at once); reduceBinsEnd waits
    // for it and fills the global bin on the root (returns its configuration count, 0 elsewhere)
    void reduceBinsBegin(int root);
    long reduceBinsEnd(std::vector<double>& ssf, std::vector<double>& isf);

    // Scattering variants of the current configuration: the elastic-scattering increment [nq] and the
    // cylinder S(q) raw sums [nq] + the number of slice-0 beads inside maxR.
    const std::vector<double>& elastic();
    const std::vector<double>& ssfCylinder(double maxR, int& numInside);

    // Virial slice sums of the current configuration: [M][4] = {sum gV.r, sum (T gV).r, sum gV.delta,
    // sum (T gV).delta}; delta (bead minus the centroid of its world-line window, src/action.cpp:1620-1647) is
    // computed here from the path's links.  t2Parity as pimcb_virial_sums.
    void setPairTableD2(const double* d2Vdr2, int len, const double* extd2Vdr2);
    // gext / g2ext (may be NULL = "free"): gradient ([M][N_ext][NDIM]) and Laplacian ([M][N_ext]) of the external potential
    // per bead; they enter all four terms (src/action.cpp:1471, 1522-1547)
    const std::vector<double>& virialSums(int window, int t2Parity, const std::vector<double>* gext = nullptr,
                                          const std::vector<double>* g2ext = nullptr);
    bool haveVirialSums(int window, int t2Parity) const { return have_vir_ && vir_window_ == window && vir_parity_ == t2Parity; }
    bool havePairSums(bool wantF2) const { return have_pair_ && (pair_has_f2_ || !wantF2); }
    void invalidate() { staged_ = false; have_sf_ = false; have_pair_ = false; have_es_ = have_cyl_ = have_vir_ = false; }
    void beginIfUnhooked() { if (!hooked_) invalidate(); }
    bool hooked() const { return hooked_; }

private:
    explicit B200Session(const Path& path);
    ~B200Session();
    B200Session(const B200Session&) = delete;
    void stageIfNeeded();
    void check(int rc, const char* what) const;

    const Path& path_;
    pimcb_ctx* ctx_ = nullptr;
    const void* locked_ptr_ = nullptr;      // Path::beads storage currently page-locked (pimcb_host_register)
    size_t locked_bytes_ = 0;
    double* fused_ssf_out_ = nullptr;       // set while ssf() drives the staging (fused stage + evaluate call)
    double* fused_isf_out_ = nullptr;
    bool fused_done_ = false;
    size_t nq_ = 0;
    bool hooked_ = false, staged_ = false, have_sf_ = false, have_pair_ = false, have_table_ = false;
    bool pair_has_f2_ = false;
    std::vector<double> ssf_, isf_;
    PairSums pair_;
    bool have_es_ = false, have_cyl_ = false, have_vir_ = false, have_d2_ = false;
    double cyl_maxR_ = 0.0;
    int cyl_inside_ = 0, vir_window_ = 0, vir_parity_ = 0;
    std::vector<double> es_, cyl_, vir_, delta_;
    friend struct SessionRegistry;
};

#endif
