// pimc_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code).
//
// A from-scratch CPU restatement of the DelMaestroGroup/pimc measurement hot path,
// used only as the checker for the CUDA library (tests/, __graft_entry__.smoke(),
// bench.py's cpu_baseline / --impl reference legs).  The product library
// (pimc_b200/csrc) never links, loads or calls anything in this file.
//
// PARITY STATUS
//  * S(q) and F(q,tau): PINNED against the reference's own code.  The reference ships no golden vectors or
//    fixtures for this path and its CPU estimators cannot be compiled here (Boost and <mdspan> are absent), but
//    its shipped GPU implementation of the same two estimators, src/estimator_gpu.cu, compiles unmodified from
//    the upstream tree (oracle/Makefile target `ref` -> oracle/_ref/librefgpu<NDIM>d.so); tests/test_reference_gpu.py
//    holds this restatement (and the product's CUDA path) to 1e-10 against it on C1, C2, 2-D and ragged inputs.
//  * Aziz tables (parameter sets, valueV / valuedVdr / valued2Vdr2, initLookupTable, direct lookups, tail correction)
//    and elastic scattering: PINNED against the reference's own code as well -- the upstream AzizPotential /
//    TabulatedPotential classes are compiled from the upstream tree into oracle/_ref/librefaziz.so
//    (oracle/ref_aziz_extract.py + ref_aziz_shim.cpp) and this file's tables are array_equal to theirs
//    (tests/test_reference_aziz.py); the upstream gpu_es kernel pins orc_elastic (tests/test_reference_gpu.py).
//  * Every loop in this file -- S(q), F(q,tau), cylinder S(q), Vint / gradVSquared / sepHist, potentialAction and its
//    derivatives, the virial slice sums, getQVectors / getQVectors2, the energy and virial estimators -- is PINNED
//    against the reference's own CPU code: the upstream function bodies are cut out of the upstream tree at build time
//    (oracle/ref_cpu_extract.py) and compiled into oracle/_ref/librefcpu<NDIM>d.so (oracle/ref_cpu_shim.cpp);
//    tests/test_reference_cpu.py holds this file array_equal to them on q-sets, S(q), F(q,tau), Vint, gradVSquared,
//    sepHist, and within 1e-11 on the derived quantities.
//  * Still unpinned by reference code: the %16.8E row formatting (boost::format upstream); the state-file array text
//    (oracle/statefile.py) IS pinned against the upstream stream operators.  Closed-form known-answer tests (tests/test_oracle_kat.py, tests/test_variants.py) cover the rest again.
//
// Every function cites the reference file:line (relative to the upstream tree) whose
// arithmetic it follows.  Floating-point semantics: this file is compiled with
// -ffp-contract=off so that each source-level operation is individually rounded
// ("as written" IEEE semantics); the integer table index int(r/dr) of the pair
// potential is compared bit-exactly against the GPU.
//
// Build: see oracle/Makefile  (g++ -O2 -ffp-contract=off -shared -fPIC).

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr double kEPS = 1.0e-7;   // include/common.h:89
constexpr int kNPCFSEP = 50;      // include/common.h:85

// ---------------------------------------------------------------------------------
// Periodic cell.  include/container.h:50-59 (putInBC), src/container.cpp:84-109, 117-144.
// ---------------------------------------------------------------------------------
struct Box {
    int ndim;
    double side[3], sideInv[3], pSide[3];
    unsigned periodic[3];
    double maxSep;
};

Box make_box(int ndim, const double* side, const unsigned* periodic) {
    Box b{};
    b.ndim = ndim;
    double acc = 0.0;
    for (int i = 0; i < ndim; ++i) {
        b.side[i] = side[i];
        b.sideInv[i] = 1.0 / side[i];                       // container.cpp:122
        b.periodic[i] = periodic ? periodic[i] : 1u;
        b.pSide[i] = b.periodic[i] * side[i];               // container.cpp:129
        const double h = side[i] / (b.periodic[i] + 1u);    // container.cpp:136
        acc += h * h;
    }
    b.maxSep = std::sqrt(acc);
    return b;
}

template <int ND>
inline void put_in_bc(const Box& b, double* r) {
    // container.h:50-53: r[i] -= pSide[i] * floor(r[i]*sideInv[i] + 0.5)
    for (int i = 0; i < ND; ++i)
        r[i] -= b.pSide[i] * std::floor(r[i] * b.sideInv[i] + 0.5);
}

template <int ND>
inline double dot(const double* a, const double* b) {
    // include/array_math.h:210-216: result = T(); result += a[i]*b[i]
    double result = 0.0;
    for (int i = 0; i < ND; ++i) result += a[i] * b[i];
    return result;
}

// AoS bead accessor: DynamicArray<dVec,2> beads(slice,ptcl), row-major, padded to
// N_ext columns.  include/path.h:57-71,164; include/dynamic_array.h:55-386.
template <int ND>
inline const double* bead(const double* beads, int Next, int slice, int ptcl) {
    return beads + (static_cast<size_t>(slice) * Next + ptcl) * ND;
}

// include/path.h:179-184 getSeparation(bead1,bead2) = putInBC(r(bead1) - r(bead2))
template <int ND>
inline void separation(const Box& b, const double* r1, const double* r2, double* sep) {
    for (int i = 0; i < ND; ++i) sep[i] = r1[i] - r2[i];
    put_in_bc<ND>(b, sep);
}

// ---------------------------------------------------------------------------------
// Static structure factor.  src/estimator.cpp:3705-3737.
// out[q] = sf(q)/numParticles  (the increment "estimator += sf/numParticles").
// ---------------------------------------------------------------------------------
template <int ND>
void ssf_impl(const Box& box, const double* beads, int M, int N, int Next,
              const double* q, int nq, double* out) {
    for (int iq = 0; iq < nq; ++iq) {
        const double* cq = q + static_cast<size_t>(iq) * ND;
        double sf = 0.0;
        for (int t = 0; t < M; ++t) {
            for (int i = 0; i < N; ++i) {
                sf += 1.0;                                   // :3726 bead1 == bead2 part
                const double* r1 = bead<ND>(beads, Next, t, i);
                for (int j = i + 1; j < N; ++j) {
                    double sep[ND];
                    separation<ND>(box, r1, bead<ND>(beads, Next, t, j), sep);
                    sf += 2 * std::cos(dot<ND>(sep, cq));    // :3729
                }
            }
        }
        out[iq] = sf / N;                                    // :3736
    }
}

// ---------------------------------------------------------------------------------
// Intermediate scattering function, the reference's direct O(Nq M^2 N^2) loop.
// src/estimator.cpp:3923-3961.  out[q*M + tau] = isf/numParticles.
// Threading is over (q, tau) output elements only, so every output element is
// accumulated in exactly the reference's order (t0 ascending, then i, then j).
// ---------------------------------------------------------------------------------
template <int ND>
void isf_element_range(const double* beads, int M, int N, int Next, const double* q,
                       int nq, double* out, long e0, long e1) {
    (void)nq;
    for (long e = e0; e < e1; ++e) {
        const int iq = static_cast<int>(e / M);
        const int tausep = static_cast<int>(e % M);
        const double* cq = q + static_cast<size_t>(iq) * ND;
        double acc = 0.0;
        for (int t0 = 0; t0 < M; ++t0) {
            const int t1 = (t0 + tausep) % M;                 // :3941
            for (int i = 0; i < N; ++i) {
                const double lq1 = dot<ND>(cq, bead<ND>(beads, Next, t0, i));   // :3946
                const double c1 = std::cos(lq1), s1 = std::sin(lq1);
                for (int j = 0; j < N; ++j) {
                    const double lq2 = dot<ND>(cq, bead<ND>(beads, Next, t1, j)); // :3951
                    acc += (c1 * std::cos(lq2) + s1 * std::sin(lq2));             // :3953
                }
            }
        }
        out[e] = acc / N;                                     // :3960
    }
}

template <int ND>
void isf_impl(const double* beads, int M, int N, int Next, const double* q, int nq,
              double* out, int nthreads) {
    const long total = static_cast<long>(nq) * M;
    if (nthreads <= 1) {
        isf_element_range<ND>(beads, M, N, Next, q, nq, out, 0, total);
        return;
    }
    std::vector<std::thread> pool;
    for (int w = 0; w < nthreads; ++w) {
        const long e0 = total * w / nthreads, e1 = total * (w + 1) / nthreads;
        pool.emplace_back([=] { isf_element_range<ND>(beads, M, N, Next, q, nq, out, e0, e1); });
    }
    for (auto& th : pool) th.join();
}

// Factorised CPU variant: rho_q(t) = sum_i exp(i q.r_i(t)); F(q,tau) = sum_t0 Re[rho(t0) conj rho(t0+tau)].
// Algebraically identical to the loop above (cos a cos b + sin a sin b summed over i,j
// factorises exactly); validated against isf_impl in tests at small sizes and used
// as the checker at sizes where the O(M^2 N^2) loop would take hours.
template <int ND>
void isf_fact_impl(const double* beads, int M, int N, int Next, const double* q, int nq,
                   double* out) {
    std::vector<double> C(M), S(M);
    for (int iq = 0; iq < nq; ++iq) {
        const double* cq = q + static_cast<size_t>(iq) * ND;
        for (int t = 0; t < M; ++t) {
            double c = 0.0, s = 0.0;
            for (int i = 0; i < N; ++i) {
                const double lq = dot<ND>(cq, bead<ND>(beads, Next, t, i));
                c += std::cos(lq);
                s += std::sin(lq);
            }
            C[t] = c; S[t] = s;
        }
        for (int tau = 0; tau < M; ++tau) {
            double acc = 0.0;
            for (int t0 = 0; t0 < M; ++t0) {
                const int t1 = (t0 + tau) % M;
                acc += C[t0] * C[t1] + S[t0] * S[t1];
            }
            out[static_cast<size_t>(iq) * M + tau] = acc / N;
        }
    }
}

// ---------------------------------------------------------------------------------
// Aziz HFDHE2 potential.  src/potential.cpp:1741-1909; include/potential.h:958-977.
// ---------------------------------------------------------------------------------
struct Aziz {
    double rm, A, epsilon, alpha, beta, D, C6, C8, C10;

    explicit Aziz(int year) {
        if (year == 1987) {          // potential.cpp:1764-1774
            epsilon = 10.948; rm = 2.9673; D = 1.4826; alpha = 10.43329537; beta = -2.27965105;
            C6 = 1.36745214; C8 = 0.42123807; C10 = 0.17473318; A = 1.8443101E5;
        } else if (year == 1995) {   // potential.cpp:1777-1787
            epsilon = 10.956; rm = 2.9683; D = 1.438; alpha = 10.5717543; beta = -2.07758779;
            C6 = 1.35186623; C8 = 0.4149514; C10 = 0.17151143; A = 1.86924404E5;
        } else {                     // 1979 default, potential.cpp:1750-1760
            epsilon = 10.8; rm = 2.9673; D = 1.241314; alpha = 13.353384; beta = 0.0;
            C6 = 1.3732412; C8 = 0.4253785; C10 = 0.1781; A = 0.5448504E6;
        }
    }
    // include/potential.h:962-964
    double F(double x) const { return (x < D ? std::exp(-(D / x - 1.0) * (D / x - 1.0)) : 1.0); }
    // include/potential.h:967-971
    double dF(double x) const {
        const double ix = 1.0 / x;
        const double r = 2.0 * D * ix * ix * (D * ix - 1.0) * std::exp(-(D * ix - 1.0) * (D * ix - 1.0));
        return (x < D ? r : 0.0);
    }
    // include/potential.h:975-980
    double d2F(double x) const {
        const double ix = 1.0 / x;
        const double r = 2.0 * D * ix * ix * ix * (2.0 * D * D * D * ix * ix * ix - 4.0 * D * D * ix * ix - D * ix + 2.0)
                         * std::exp(-(D * ix - 1.0) * (D * ix - 1.0));
        return (x < D ? r : 0.0);
    }
    // src/potential.cpp:1822-1842
    double valueV(double r) const {
        const double x = r / rm;
        const double Urep = A * std::exp(-alpha * x + beta * x * x);
        if (x < kEPS) return 0.0;
        else if (x < 0.01) return (epsilon * Urep);
        else {
            const double ix2 = 1.0 / (x * x);
            const double ix6 = ix2 * ix2 * ix2;
            const double ix8 = ix6 * ix2;
            const double ix10 = ix8 * ix2;
            const double Uatt = -(C6 * ix6 + C8 * ix8 + C10 * ix10) * F(x);
            return (epsilon * (Urep + Uatt));
        }
    }
    // src/potential.cpp:1849-1875
    double valuedVdr(double r) const {
        const double x = r / rm;
        const double T1 = A * (-alpha + 2.0 * beta * x) * std::exp(-alpha * x + beta * x * x);
        if (x < kEPS) return 0.0;
        else if (x < 0.01) return ((epsilon / rm) * T1);
        else {
            const double ix = 1.0 / x;
            const double ix2 = ix * ix;
            const double ix6 = ix2 * ix2 * ix2;
            const double ix7 = ix6 * ix;
            const double ix8 = ix6 * ix2;
            const double ix9 = ix8 * ix;
            const double ix10 = ix8 * ix2;
            const double ix11 = ix10 * ix;
            const double T2 = (6.0 * C6 * ix7 + 8.0 * C8 * ix9 + 10.0 * C10 * ix11) * F(x);
            const double T3 = -(C6 * ix6 + C8 * ix8 + C10 * ix10) * dF(x);
            return ((epsilon / rm) * (T1 + T2 + T3));
        }
    }
    // src/potential.cpp:1882-1909
    double valued2Vdr2(double r) const {
        const double x = r / rm;
        const double abFactor2 = (alpha - 2.0 * beta * x) * (alpha - 2.0 * beta * x);
        const double T1 = A * (2 * beta + abFactor2) * std::exp(-alpha * x + beta * x * x);
        if (x < kEPS) return 0.0;
        else if (x < 0.01) return ((epsilon / rm) * T1);
        else {
            const double ix = 1.0 / x;
            const double ix2 = ix * ix;
            const double ix6 = ix2 * ix2 * ix2;
            const double ix7 = ix6 * ix;
            const double ix8 = ix6 * ix2;
            const double ix9 = ix8 * ix;
            const double ix10 = ix8 * ix2;
            const double ix11 = ix10 * ix;
            const double ix12 = ix11 * ix;
            const double T2 = -(42.0 * C6 * ix8 + 72.0 * C8 * ix10 + 110.0 * C10 * ix12) * F(x);
            const double T3 = 2.0 * (6.0 * C6 * ix7 + 8.0 * C8 * ix9 + 10.0 * C10 * ix11) * dF(x);
            const double T4 = -(C6 * ix6 + C8 * ix8 + C10 * ix10) * d2F(x);
            return ((epsilon / (rm * rm)) * (T1 + T2 + T3 + T4));
        }
    }
};

// include/potential.h:249-260 TabulatedPotential::direct
inline double table_direct(const double* table, int tableLength, double dr, const double* extVal, double r) {
    const int k = int(r / dr);
    if (k <= 0) return extVal[0];
    if (k >= tableLength) return extVal[1];
    return table[k];
}

struct PairTable {
    const double* V; const double* dVdr; int len; double dr; double extV[2]; double extdVdr[2];
};

// LocalAction::V(slice), interaction part + sepHist side effect.
// src/action.cpp:902-947 (+216-224 updateSepHist); worm factor == 1 on diagonal configurations
// (include/worm.h:50, src/worm.cpp:108-118).  Separation is getSeparation(bead2,bead1), :934.
template <int ND>
double vint_slice(const Box& box, const PairTable& tab, const double* beads, int N, int Next,
                  int slice, double dSep, int* sepHist) {
    double totVint = 0.0;
    if (sepHist) std::fill(sepHist, sepHist + kNPCFSEP, 0);    // :918
    for (int i = 0; i < N; ++i) {
        const double* r1 = bead<ND>(beads, Next, slice, i);
        for (int j = i + 1; j < N; ++j) {
            double sep[ND];
            separation<ND>(box, bead<ND>(beads, Next, slice, j), r1, sep);
            const double rnorm = std::sqrt(dot<ND>(sep, sep));
            if (sepHist) {                                       // action.cpp:221-223
                const int nR = int(rnorm / dSep);
                if (nR >= 0 && nR < kNPCFSEP) ++sepHist[nR];
            }
            totVint += 1.0 * table_direct(tab.V, tab.len, tab.dr, tab.extV, rnorm);   // potential.h:985-989
        }
    }
    return totVint;
}

// LocalAction::gradVSquared(slice), interaction part (external "free" potential has zero gradient).
// src/action.cpp:1188-1223; AzizPotential::gradV include/potential.h:997-1003.
template <int ND>
double grad_v_squared_slice(const Box& box, const PairTable& tab, const double* beads, int N, int Next, int slice,
                            const double* gext = nullptr) {
    double totF2 = 0.0;
    for (int i = 0; i < N; ++i) {
        double F[ND];
        for (int d = 0; d < ND; ++d) F[d] = 0.0;
        const double* r1 = bead<ND>(beads, Next, slice, i);
        for (int j = 0; j < N; ++j) {
            if (j == i) continue;                               // :1208
            double sep[ND];
            separation<ND>(box, r1, bead<ND>(beads, Next, slice, j), sep);   // :1211 getSeparation(bead1,bead2)
            const double rnorm = std::sqrt(dot<ND>(sep, sep));
            const double g = table_direct(tab.dVdr, tab.len, tab.dr, tab.extdVdr, rnorm) / rnorm;
            for (int d = 0; d < ND; ++d) F[d] += g * sep[d];
        }
        if (gext) {                                             // :1216  F += externalPtr->gradV(path(bead1))
            const double* ge = bead<ND>(gext, Next, slice, i);
            for (int d = 0; d < ND; ++d) F[d] += ge[d];
        }
        totF2 += dot<ND>(F, F);                                 // :1219
    }
    return totF2;
}

// ---------------------------------------------------------------------------------
// Scattering variants.
// ---------------------------------------------------------------------------------

// EstimatorBase::getQVectors2, src/estimator.cpp:762-837: magnitude shells cq = 0, dq, 2 dq, ... <= qMax + EPS
// (cq accumulated by cq += dq); shell 0 is the null vector; every other shell starts with cq along the last
// dimension; in 3-D the "sphere" geometry adds (theta, phi) points on the positive octant with
// dtheta = (pi/2)/24, dphi = dtheta/sin(theta) (for "line" dtheta = pi, so the theta loop never runs).
// out: flattened vectors [numq][ndim]; shell_sizes[k] = vectors in shell k.  Returns the number of shells, or the
// negated required count when a capacity is too small; -1000 on a bad geometry string.
int qvectors2_impl(int ndim, double dq, double qMax, const char* geom_c, double* out, int max_vecs, int* shell_sizes,
                   int max_shells, int* numq_out) {
    const std::string qGeometry(geom_c);
    if (qGeometry != "line" && qGeometry != "sphere") return -1000;
    int numq = 0, nshell = 0;
    auto push = [&](const double* v) {
        if (numq < max_vecs) for (int d = 0; d < ndim; ++d) out[static_cast<size_t>(numq) * ndim + d] = v[d];
        ++numq;
    };
    for (double cq = 0.0; cq <= qMax + kEPS; cq += dq) {
        const int before = numq;
        if (std::abs(cq) < kEPS) {
            double qd[3] = {0.0, 0.0, 0.0};
            push(qd);
        } else {
            double qd[3] = {0.0, 0.0, 0.0};
            qd[ndim - 1] = cq;
            push(qd);
            if (ndim == 3) {
                const int numTheta = 24;
                const double dtheta = (qGeometry == "line") ? M_PI : 0.5 * M_PI / numTheta;
                for (double theta = dtheta; theta <= 0.5 * M_PI + kEPS; theta += dtheta) {
                    const double dphi = dtheta / sin(theta);
                    for (double phi = 0.0; phi <= 0.5 * M_PI + kEPS; phi += dphi) {
                        double v[3];
                        v[0] = cq * sin(theta) * cos(phi);
                        v[1] = cq * sin(theta) * sin(phi);
                        v[2] = cq * cos(theta);
                        push(v);
                    }
                }
            }
        }
        if (nshell < max_shells) shell_sizes[nshell] = numq - before;
        ++nshell;
    }
    if (numq_out) *numq_out = numq;
    if (numq > max_vecs || nshell > max_shells) return -std::max(numq, nshell);
    return nshell;
}

// include(r, maxR), src/estimator.cpp:4309-4311
inline bool cyl_include(const double* r, double maxR) { return (r[0] * r[0] + r[1] * r[1] < maxR * maxR); }

// CylinderStaticStructureFactorEstimator::accumulate, src/estimator.cpp:5415-5456, for ONE wave-vector (the caller sums
// the vectors of a magnitude shell and divides by num1DParticles): sum_t sum_{i in} [1 + 2 sum_{j>i, j in} cos(sep.q)].
template <int ND>
double ssf_cyl_one(const Box& box, const double* beads, int M, int N, int Next, const double* q, double maxR) {
    double sf = 0.0;
    for (int t = 0; t < M; ++t)
        for (int i = 0; i < N; ++i) {
            const double* r1 = bead<ND>(beads, Next, t, i);
            if (!cyl_include(r1, maxR)) continue;
            sf += 1.0;
            for (int j = i + 1; j < N; ++j) {
                const double* r2 = bead<ND>(beads, Next, t, j);
                if (!cyl_include(r2, maxR)) continue;
                double sep[ND];
                separation<ND>(box, r1, r2, sep);
                sf += 2 * cos(dot<ND>(sep, q));
            }
        }
    return sf;
}

// ---------------------------------------------------------------------------------
// Virial slice sums.  Links: next[(slice*Next + ptcl)*2 + {0,1}] = (slice, ptcl) of
// the next bead (NULL = straight closed world lines); prev is the inverse map.
// ---------------------------------------------------------------------------------
struct Links {
    const int* next; std::vector<int> prev; int M, Next;
    Links(const int* n, int M_, int N, int Next_) : next(n), M(M_), Next(Next_) {
        if (!next) return;
        prev.assign(static_cast<size_t>(M) * Next * 2, -1);
        for (int s = 0; s < M; ++s)
            for (int p = 0; p < N; ++p) {
                const int ns = next[(static_cast<size_t>(s) * Next + p) * 2], np = next[(static_cast<size_t>(s) * Next + p) * 2 + 1];
                if (ns < 0 || np < 0) continue;
                prev[(static_cast<size_t>(ns) * Next + np) * 2] = s;
                prev[(static_cast<size_t>(ns) * Next + np) * 2 + 1] = p;
            }
    }
    void fwd(int& s, int& p) const {          // Path::next(bead), include/path.h:96-98
        if (!next) { s = (s + 1) % M; return; }
        const size_t k = (static_cast<size_t>(s) * Next + p) * 2;
        s = next[k]; p = next[k + 1];
    }
    void back(int& s, int& p) const {         // Path::prev(bead), include/path.h:105-107
        if (!next) { s = (s + M - 1) % M; return; }
        const size_t k = (static_cast<size_t>(s) * Next + p) * 2;
        const int a = prev[k], b = prev[k + 1];
        s = a; p = b;
    }
};

// delta = putInBC(pos1 - COM) of bead (slice, ptcl) over the virial window, src/action.cpp:1620-1647 / 1742-1769:
// path.next(bead1, gamma) / path.prev(bead1, gamma) for gamma = 0..window-1 (gamma links away from bead1, so the first
// pair is bead1 itself), running sums of minimum-image link vectors, COM /= 2*window.
template <int ND>
void virial_delta(const Box& box, const double* beads, const Links& L, int slice, int ptcl, int window, double* delta) {
    double runTotMore[ND], runTotLess[ND], COM[ND];
    for (int d = 0; d < ND; ++d) runTotMore[d] = runTotLess[d] = COM[d] = 0.0;
    const double* pos1 = bead<ND>(beads, L.Next, slice, ptcl);
    int nos = slice, nop = ptcl, pos_ = slice, pop = ptcl;          // beadNextOld, beadPrevOld
    for (int gamma = 0; gamma < window; ++gamma) {
        int ns = slice, np = ptcl, ps = slice, pp = ptcl;
        for (int m = 0; m < gamma; ++m) { L.fwd(ns, np); L.back(ps, pp); }
        double sep[ND];
        separation<ND>(box, bead<ND>(beads, L.Next, ns, np), bead<ND>(beads, L.Next, nos, nop), sep);
        for (int d = 0; d < ND; ++d) runTotMore[d] += sep[d];
        separation<ND>(box, bead<ND>(beads, L.Next, ps, pp), bead<ND>(beads, L.Next, pos_, pop), sep);
        for (int d = 0; d < ND; ++d) runTotLess[d] += sep[d];
        for (int d = 0; d < ND; ++d) COM[d] += (pos1[d] + runTotMore[d]) + (pos1[d] + runTotLess[d]);
        nos = ns; nop = np; pos_ = ps; pop = pp;
    }
    for (int d = 0; d < ND; ++d) COM[d] /= (2.0 * window);
    for (int d = 0; d < ND; ++d) delta[d] = pos1[d] - COM[d];
    put_in_bc<ND>(box, delta);
}

struct VirialTables {
    const double* dVdr; const double* d2V; int len; double dr; double extdVdr[2]; double extd2V[2];
};

// One slice: out[0] = sum_i gV_i.r_i (rDOTgradUterm1 without VFactor*tau, src/action.cpp:1446-1478),
// out[1] = sum_i (gV_i T_i).r_i (rDOTgradUterm2 without 2*gradVFactor*tau^3*lambda, :1493-1575),
// out[2], out[3] the same with delta_i in place of r_i (deltadotgradUterm1/2, :1588-1784).
// gext [M][Next][ND] = externalPtr->gradV(path(bead)), g2ext [M][Next] = externalPtr->grad2V(path(bead)) (null = "free"):
// gV = gVe + sum gVi (:1471, :1525), dV = dVi + dVe and d2V = g2Vi + g2Ve inside the T-matrix (:1546-1547).
template <int ND>
void virial_slice(const Box& box, const VirialTables& tab, const double* beads, const Links& L, int N, int slice, int window,
                  bool want_t2, double* out, const double* gext = nullptr, const double* g2ext = nullptr) {
    const int Next = L.Next;
    out[0] = out[1] = out[2] = out[3] = 0.0;
    for (int i = 0; i < N; ++i) {
        const double* r1 = bead<ND>(beads, Next, slice, i);
        double gV[ND], gVe[ND], gVsum[ND], tMat[ND][ND];       // gV: term2 order (gVe first, :1525); gVsum: the pair part alone
        for (int a = 0; a < ND; ++a) {
            gVe[a] = gext ? gext[(static_cast<size_t>(slice) * Next + i) * ND + a] : 0.0;
            gV[a] = 0.0 + gVe[a];                                                                   // gV += gVe, :1525
            gVsum[a] = 0.0;
            for (int b = 0; b < ND; ++b) tMat[a][b] = 0.0;
        }
        const double dVe = std::sqrt(dot<ND>(gVe, gVe));
        const double g2Ve = g2ext ? g2ext[static_cast<size_t>(slice) * Next + i] : 0.0;
        for (int j = 0; j < N; ++j) {
            double rDiff[ND];
            separation<ND>(box, r1, bead<ND>(beads, Next, slice, j), rDiff);
            const double rmag = std::sqrt(dot<ND>(rDiff, rDiff));
            if (j == i) continue;
            double gVi[ND];
            const double g = table_direct(tab.dVdr, tab.len, tab.dr, tab.extdVdr, rmag) / rmag;    // potential.h:997-1003
            for (int d = 0; d < ND; ++d) gVi[d] = g * rDiff[d];
            if (want_t2) {
                const double dVi = std::sqrt(dot<ND>(gVi, gVi));
                const double g2Vi = table_direct(tab.d2V, tab.len, tab.dr, tab.extd2V, rmag);      // potential.h:1010-1016
                const double dV = dVi + dVe, d2V = g2Vi + g2Ve;
                for (int a = 0; a < ND; ++a)
                    for (int b = 0; b < ND; ++b) {
                        tMat[a][b] += rDiff[a] * rDiff[b] * d2V / (rmag * rmag) - rDiff[a] * rDiff[b] * dV / pow(rmag, 3);
                        if (a == b) tMat[a][b] += dV / rmag;
                    }
            }
            for (int d = 0; d < ND; ++d) { gV[d] += gVi[d]; gVsum[d] += gVi[d]; }
        }
        double gV1[ND];                                            // term1: gV = gVe + gVi (:1471, :1647)
        for (int d = 0; d < ND; ++d) gV1[d] = gVe[d] + gVsum[d];
        double gVdotT[ND];
        for (int row = 0; row < ND; ++row) {                                                        // common.h:187-199
            double acc = 0.0;
            for (int k = 0; k < ND; ++k) acc = acc + tMat[row][k] * gV[k];
            gVdotT[row] = 0.0 + acc;
        }
        double delta[ND];
        virial_delta<ND>(box, beads, L, slice, i, window, delta);
        out[0] += dot<ND>(gV1, r1);
        out[2] += dot<ND>(gV1, delta);
        if (want_t2) {
            out[1] += dot<ND>(gVdotT, r1);
            out[3] += dot<ND>(gVdotT, delta);
        }
    }
}

// ---------------------------------------------------------------------------------
// q-vector generation.  src/estimator.cpp:439-570.
template <int ND>
int qvectors_impl(const char* type_c, const char* input_c, const double* side, double* out, int max_out) {
    const std::string inputType(type_c), input(input_c);
    std::vector<std::array<double, ND>> qValues;
    std::array<double, ND> q{};

    std::istringstream iss(input);
    std::vector<std::string> tokens{std::istream_iterator<std::string>{iss}, std::istream_iterator<std::string>{}};
    if (tokens.size() < 1) return -1;                           // :450-455 (reference exits)

    if ((inputType == "int") || (inputType == "float")) {       // :458-472
        if (tokens.size() % ND != 0) return -2;
        for (size_t i = 0; i < tokens.size(); i += ND) {
            for (int j = 0; j < ND; ++j) {
                if (inputType == "int") q[j] = (2.0 * M_PI / side[j]) * std::stoi(tokens[i + j]);
                else q[j] = std::stof(tokens[i + j]);           // float precision, :466
            }
            qValues.push_back(q);
        }
    }
    if ((inputType == "max_int") || (inputType == "max_float")) {   // :475-537
        if (static_cast<int>(tokens.size()) < ND) return -2;
        std::array<int, ND> q_max_int{}, _q_int{};
        std::array<double, ND> q_max{};
        for (int i = 0; i < ND; ++i) {
            if (inputType == "max_int") {
                q_max_int[i] = std::abs(std::stoi(tokens[i]));
                q_max[i] = q_max_int[i] * 2.0 * M_PI / side[i];
            } else {
                q_max[i] = std::stof(tokens[i]);
            }
        }
        const double q_mag_max = std::sqrt(dot<ND>(q_max.data(), q_max.data()));
        double q_mag;
        if (inputType == "max_float")
            for (int i = 0; i < ND; ++i)
                q_max_int[i] = 1 + static_cast<int>(q_mag_max * side[i] / 2.0 / M_PI);
        int n_q = 1;
        for (int i = 0; i < ND; ++i) {
            n_q *= 2 * q_max_int[i] + 1;
            _q_int[i] = -q_max_int[i];
            q[i] = _q_int[i] * 2.0 * M_PI / side[i];
        }
        q_mag = std::sqrt(dot<ND>(q.data(), q.data()));
        if (q_mag <= q_mag_max) qValues.push_back(q);
        int pos = ND - 1;
        int count = 0;
        while (count < n_q - 1) {
            if (_q_int[pos] == q_max_int[pos]) {
                _q_int[pos] = -q_max_int[pos];
                pos -= 1;
            } else {
                _q_int[pos] += 1;
                for (int i = 0; i < ND; ++i) q[i] = _q_int[i] * 2.0 * M_PI / side[i];
                q_mag = std::sqrt(dot<ND>(q.data(), q.data()));
                if (q_mag <= q_mag_max) qValues.push_back(q);
                count += 1;
                pos = ND - 1;
            }
        }
    }
    if (inputType == "file_int") {                              // :540-553
        std::ifstream file(input);
        if (!file) return -3;
        std::string line;
        while (std::getline(file, line)) {
            std::istringstream ls(line);
            std::vector<int> d((std::istream_iterator<int>(ls)), std::istream_iterator<int>());
            if (static_cast<int>(d.size()) != ND) return -2;
            for (int i = 0; i < ND; ++i) q[i] = (2.0 * M_PI / side[i]) * d[i];
            qValues.push_back(q);
        }
    }
    if (inputType == "file_float") {                            // :556-569
        std::ifstream file(input);
        if (!file) return -3;
        std::string line;
        while (std::getline(file, line)) {
            std::istringstream ls(line);
            std::vector<float> d((std::istream_iterator<float>(ls)), std::istream_iterator<float>());
            if (static_cast<int>(d.size()) != ND) return -2;
            for (int i = 0; i < ND; ++i) q[i] = d[i];
            qValues.push_back(q);
        }
    }
    const int n = static_cast<int>(qValues.size());
    if (out) {
        if (n > max_out) return -4;
        for (int k = 0; k < n; ++k)
            for (int d = 0; d < ND; ++d) out[static_cast<size_t>(k) * ND + d] = qValues[k][d];
    }
    return n;
}

}  // namespace

#define ORC_DISPATCH(ndim, CALL)                 \
    switch (ndim) {                              \
        case 1: { constexpr int ND = 1; CALL; } break; \
        case 2: { constexpr int ND = 2; CALL; } break; \
        case 3: { constexpr int ND = 3; CALL; } break; \
        default: return -100;                    \
    }

extern "C" {

// Returns maxSep (src/container.cpp:100,136).
double orc_max_sep(int ndim, const double* side, const unsigned* periodic) {
    return make_box(ndim, side, periodic).maxSep;
}

int orc_put_in_bc(int ndim, const double* side, const unsigned* periodic, double* r, int count) {
    const Box b = make_box(ndim, side, periodic);
    ORC_DISPATCH(ndim, for (int k = 0; k < count; ++k) put_in_bc<ND>(b, r + static_cast<size_t>(k) * ND));
    return 0;
}

int orc_qvectors(int ndim, const char* type, const char* input, const double* side, double* out, int max_out) {
    int n = -100;
    ORC_DISPATCH(ndim, n = qvectors_impl<ND>(type, input, side, out, max_out));
    return n;
}

int orc_ssf(int ndim, const double* side, const unsigned* periodic, const double* beads, int M, int N,
            int Next, const double* q, int nq, double* out) {
    const Box b = make_box(ndim, side, periodic);
    ORC_DISPATCH(ndim, ssf_impl<ND>(b, beads, M, N, Next, q, nq, out));
    return 0;
}

// Threaded over q for the all-host-cores CPU baseline (q is the outermost independent loop,
// src/estimator.cpp:3715); each q's accumulation order is the reference's.
int orc_ssf_mt(int ndim, const double* side, const unsigned* periodic, const double* beads, int M, int N,
               int Next, const double* q, int nq, double* out, int nthreads) {
    const Box b = make_box(ndim, side, periodic);
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> pool;
    for (int w = 0; w < nthreads; ++w) {
        const int q0 = static_cast<int>(static_cast<long>(nq) * w / nthreads);
        const int q1 = static_cast<int>(static_cast<long>(nq) * (w + 1) / nthreads);
        if (q1 <= q0) continue;
        pool.emplace_back([=, &b] {
            switch (ndim) {
                case 1: ssf_impl<1>(b, beads, M, N, Next, q + q0 * 1, q1 - q0, out + q0); break;
                case 2: ssf_impl<2>(b, beads, M, N, Next, q + q0 * 2, q1 - q0, out + q0); break;
                default: ssf_impl<3>(b, beads, M, N, Next, q + q0 * 3, q1 - q0, out + q0); break;
            }
        });
    }
    for (auto& th : pool) th.join();
    return 0;
}

int orc_isf(int ndim, const double* beads, int M, int N, int Next, const double* q, int nq, double* out,
            int nthreads) {
    ORC_DISPATCH(ndim, isf_impl<ND>(beads, M, N, Next, q, nq, out, nthreads));
    return 0;
}

// Elements [e0,e1) of the flattened isf[q*M + tau] array only (bounded CPU-baseline samples: every element
// costs the same M*N^2 terms, so a sample extrapolates linearly to the full Nq*M set).
int orc_isf_range(int ndim, const double* beads, int M, int N, int Next, const double* q, int nq, double* out,
                  long e0, long e1, int nthreads) {
    if (e0 < 0 || e1 > static_cast<long>(nq) * M || e0 > e1) return -5;
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> pool;
    for (int w = 0; w < nthreads; ++w) {
        const long a = e0 + (e1 - e0) * w / nthreads, b = e0 + (e1 - e0) * (w + 1) / nthreads;
        if (b <= a) continue;
        pool.emplace_back([=] {
            switch (ndim) {
                case 1: isf_element_range<1>(beads, M, N, Next, q, nq, out, a, b); break;
                case 2: isf_element_range<2>(beads, M, N, Next, q, nq, out, a, b); break;
                default: isf_element_range<3>(beads, M, N, Next, q, nq, out, a, b); break;
            }
        });
    }
    for (auto& th : pool) th.join();
    return (ndim < 1 || ndim > 3) ? -100 : 0;
}

int orc_isf_factorised(int ndim, const double* beads, int M, int N, int Next, const double* q, int nq,
                       double* out) {
    ORC_DISPATCH(ndim, isf_fact_impl<ND>(beads, M, N, Next, q, nq, out));
    return 0;
}

// Aziz scalar functions (which: 0=V, 1=dV/dr, 2=d2V/dr2).
int orc_aziz_values(int year, int which, const double* r, double* out, int count) {
    const Aziz az(year);
    for (int k = 0; k < count; ++k)
        out[k] = which == 0 ? az.valueV(r[k]) : which == 1 ? az.valuedVdr(r[k]) : az.valued2Vdr2(r[k]);
    return 0;
}

double orc_aziz_rm(int year) { return Aziz(year).rm; }

// TabulatedPotential::initLookupTable, include/potential.h:163-183: dr = 1e-6*rm
// (src/potential.cpp:1796), tableLength = int(maxSep/dr), r accumulated by r += dr.
// Call with V == NULL to get the table length.
int orc_aziz_table(int year, double maxSep, double* V, double* dVdr, double* d2Vdr2, int len_in, double* dr_out) {
    const Aziz az(year);
    const double dr = (1.0E-6) * az.rm;
    const int tableLength = int(maxSep / dr);
    if (dr_out) *dr_out = dr;
    if (!V) return tableLength;
    if (len_in < tableLength) return -4;
    double r = 0;
    for (int n = 0; n < tableLength; ++n) {
        V[n] = az.valueV(r);
        if (dVdr) dVdr[n] = az.valuedVdr(r);
        if (d2Vdr2) d2Vdr2[n] = az.valued2Vdr2(r);
        r += dr;
    }
    return tableLength;
}

// Tabulated potential evaluated on explicit separation vectors: the scalar V(dVec) path and the
// batched V(const dVec*, double*, int) default (src/potential.cpp:127-132) are the same loop.
int orc_table_V(int ndim, const double* table, int len, double dr, const double* ext, const double* sep,
                double* out, int count) {
    if (count < 0) return -5;   // reference throws std::runtime_error on negative count
    ORC_DISPATCH(ndim, for (int k = 0; k < count; ++k) {
        const double* s = sep + static_cast<size_t>(k) * ND;
        out[k] = table_direct(table, len, dr, ext, std::sqrt(dot<ND>(s, s)));
    });
    return 0;
}

// Per-slice pair sums for all slices: vint[M], f2[M] (may be NULL), sephist[M][50] (may be NULL).
int orc_pair_sums(int ndim, const double* side, const unsigned* periodic, const double* beads, int M, int N,
                  int Next, const double* V, const double* dVdr, int len, double dr, const double* extV,
                  const double* extdVdr, double dSep, double* vint, double* f2, int* sephist, int nthreads) {
    const Box b = make_box(ndim, side, periodic);
    PairTable tab{V, dVdr, len, dr, {extV[0], extV[1]}, {extdVdr ? extdVdr[0] : 0.0, extdVdr ? extdVdr[1] : 0.0}};
    if (nthreads < 1) nthreads = 1;
    auto work = [&](int s0, int s1) {
        for (int s = s0; s < s1; ++s) {
            int* h = sephist ? sephist + static_cast<size_t>(s) * kNPCFSEP : nullptr;
            if (vint) {
                if (ndim == 1) vint[s] = vint_slice<1>(b, tab, beads, N, Next, s, dSep, h);
                else if (ndim == 2) vint[s] = vint_slice<2>(b, tab, beads, N, Next, s, dSep, h);
                else vint[s] = vint_slice<3>(b, tab, beads, N, Next, s, dSep, h);
            }
            if (f2) {
                if (ndim == 1) f2[s] = grad_v_squared_slice<1>(b, tab, beads, N, Next, s);
                else if (ndim == 2) f2[s] = grad_v_squared_slice<2>(b, tab, beads, N, Next, s);
                else f2[s] = grad_v_squared_slice<3>(b, tab, beads, N, Next, s);
            }
        }
    };
    if (ndim < 1 || ndim > 3) return -100;
    if (nthreads == 1) { work(0, M); return 0; }
    std::vector<std::thread> pool;
    for (int w = 0; w < nthreads; ++w) {
        const int s0 = static_cast<int>(static_cast<long>(M) * w / nthreads);
        const int s1 = static_cast<int>(static_cast<long>(M) * (w + 1) / nthreads);
        if (s1 > s0) pool.emplace_back(work, s0, s1);
    }
    for (auto& th : pool) th.join();
    return 0;
}

// gradVSquared[M] with a non-trivial external potential: gext[M][Next][ndim] = externalPtr->gradV(r) per bead.
int orc_grad_v_squared_ext(int ndim, const double* side, const unsigned* periodic, const double* beads, int M, int N, int Next,
                           const double* dVdr, int len, double dr, const double* extdVdr, const double* gext, double* f2) {
    if (ndim < 1 || ndim > 3) return -100;
    const Box b = make_box(ndim, side, periodic);
    PairTable tab{nullptr, dVdr, len, dr, {0.0, 0.0}, {extdVdr ? extdVdr[0] : 0.0, extdVdr ? extdVdr[1] : 0.0}};
    for (int s = 0; s < M; ++s) {
        if (ndim == 1) f2[s] = grad_v_squared_slice<1>(b, tab, beads, N, Next, s, gext);
        else if (ndim == 2) f2[s] = grad_v_squared_slice<2>(b, tab, beads, N, Next, s, gext);
        else f2[s] = grad_v_squared_slice<3>(b, tab, beads, N, Next, s, gext);
    }
    return 0;
}

// LocalAction::potentialAction(), src/action.cpp:456-472, from per-slice sums (Vext == 0 for the
// "free" external potential): totU += VFactor[eo]*tau*(Vext+Vint); if gradVFactor[eo] > EPS
// totU += gradVFactor[eo]*tau^3*lambda*gradVSquared(slice).
double orc_potential_action(const double* vint, const double* f2, int M, const double* VFactor,
                            const double* gradVFactor, double tau, double lambda) {
    double totU = 0.0;
    for (int slice = 0; slice < M; ++slice) {
        const int eo = slice % 2;
        totU += VFactor[eo] * tau * (0.0 + vint[slice]);
        if (gradVFactor[eo] > kEPS)
            totU += gradVFactor[eo] * tau * tau * tau * lambda * f2[slice];
    }
    return totU;
}

// LocalAction::derivPotentialActionTau(slice), src/action.cpp:751-765.
double orc_deriv_potential_action_tau(double vint_s, double f2_s, int slice, const double* VFactor,
                                      const double* gradVFactor, double tau, double lambda) {
    const int eo = slice % 2;
    double dU = VFactor[eo] * (0.0 + vint_s);
    if (gradVFactor[eo] > kEPS) dU += 3.0 * gradVFactor[eo] * tau * tau * lambda * f2_s;
    return dU;
}

// LocalAction::derivPotentialActionLambda(slice), src/action.cpp:798-806.
double orc_deriv_potential_action_lambda(double f2_s, int slice, const double* gradVFactor, double tau) {
    const int eo = slice % 2;
    if (gradVFactor[eo] > kEPS) return gradVFactor[eo] * tau * tau * tau * f2_s;
    return 0.0;
}

// EstimatorBase::output(), src/estimator.cpp:348-362: estimator *= norm/numAccumulated, then
// boost::format("%16.8E") per column.  Writes the formatted row (no newline) into buf.
int orc_format_row(const double* estimator, const double* norm, int numEst, unsigned numAccumulated,
                   char* buf, int buflen) {
    int off = 0;
    for (int n = 0; n < numEst; ++n) {
        const double v = estimator[n] * (norm[n] / (1.0 * numAccumulated));
        const int w = std::snprintf(buf + off, buflen - off, "%16.8E", v);
        if (w < 0 || off + w >= buflen) return -4;
        off += w;
    }
    return off;
}

// EstimatorBase::dVecToString, src/estimator.cpp:421-429.
int orc_dvec_to_string(int ndim, const double* v, char* buf, int buflen) {
    int off = std::snprintf(buf, buflen, "(");
    for (int i = 0; i < ndim; ++i) {
        off += std::snprintf(buf + off, buflen - off, "%+15.8E", v[i]);
        if (i < ndim - 1) off += std::snprintf(buf + off, buflen - off, ",");
    }
    off += std::snprintf(buf + off, buflen - off, ")");
    return off;
}

// AzizPotential tail correction, src/potential.cpp:1798-1806 (rc = potential cutoff, default side[NDIM-1]).
double orc_aziz_tail(int year, double rc) {
    const Aziz az(year);
    const double rmorc = az.rm / rc;
    const double t2 = az.C6 * pow(rmorc, 3.0) / 3.0;
    const double t3 = az.C8 * pow(rmorc, 5.0) / 5.0;
    const double t4 = az.C10 * pow(rmorc, 7.0) / 7.0;
    return 2.0 * M_PI * az.epsilon * (-az.rm * az.rm * az.rm * (t2 + t3 + t4));
}

// EnergyEstimator::accumulate, src/estimator.cpp:940-1029 (PIMC mode: startSlice = 0, endSlice = M, sliceFactor = 1),
// on top of the per-slice pair sums (Vint, gradVSquared) of this file; external potential "free" (Vext = 0).
// Kinetic term: Path::getVelocity, include/path.h:189-203, over the links `next` ([M][Next][2], NULL = straight
// world lines).  out[9] = {K, V, V_ext, V_int, E, E_mu, K/N, V/N, E/N} of ONE configuration.
int orc_energy(int nd, const double* side, const unsigned* periodic, const double* beads, int M, int N, int Next,
               const int* next, const double* vint, const double* f2, const double* VFactor, const double* gradVFactor,
               int period, double tau, double lambda, double mu, double tailV_potential, double* out) {
    double pSide[3], sideInv[3], volume = 1.0;
    for (int d = 0; d < nd; ++d) { pSide[d] = periodic[d] * side[d]; sideInv[d] = 1.0 / side[d]; volume *= side[d]; }
    const int numParticles = N, numTimeSlices = M;
    const double tailV = (1.0 * numParticles * numParticles / volume) * tailV_potential;
    const double kinNorm = (0.25 / (lambda * tau)) / (tau * numTimeSlices);          // constants.h:82
    const double classicalKinetic = (0.5 * nd / tau) * numParticles;
    double totK = 0.0;
    for (int slice = 0; slice < M; ++slice)
        for (int ptcl = 0; ptcl < N; ++ptcl) {
            int ns = (slice + 1) % M, np = ptcl;
            if (next) { ns = next[(static_cast<size_t>(slice) * Next + ptcl) * 2]; np = next[(static_cast<size_t>(slice) * Next + ptcl) * 2 + 1]; }
            if (ns < 0 || np < 0) continue;                                           // XXX links: zero velocity
            double v2 = 0.0;
            for (int d = 0; d < nd; ++d) {
                double v = beads[(static_cast<size_t>(ns) * Next + np) * nd + d] - beads[(static_cast<size_t>(slice) * Next + ptcl) * nd + d];
                v -= pSide[d] * floor(v * sideInv[d] + 0.5);
                v2 += v * v;
            }
            totK -= v2;
        }
    totK *= kinNorm;
    double t1 = 0.0, t2 = 0.0, vop = 0.0;
    for (int slice = 0; slice < M; ++slice) {
        t1 += orc_deriv_potential_action_lambda(f2 ? f2[slice] : 0.0, slice, gradVFactor, tau);
        t2 += orc_deriv_potential_action_tau(vint[slice], f2 ? f2[slice] : 0.0, slice, VFactor, gradVFactor, tau, lambda);
        if (!(slice % period)) vop += vint[slice];
    }
    t1 *= lambda / (tau * numTimeSlices);
    t2 /= 1.0 * numTimeSlices;
    vop /= (numTimeSlices / period);
    totK += (classicalKinetic + t1);
    const double totV = t2 - t1 + tailV;
    vop += tailV;
    out[0] = totK; out[1] = totV; out[2] = 0.0; out[3] = vop; out[4] = totK + totV; out[5] = totK + totV - mu * numParticles;
    out[6] = totK / numParticles; out[7] = totV / numParticles; out[8] = (totK + totV) / numParticles;
    return 0;
}

// ---- scattering variants -------------------------------------------------------------------------------------------
int orc_qvectors2(int ndim, double dq, double qMax, const char* geometry, double* out, int max_vecs, int* shell_sizes,
                  int max_shells, int* numq) {
    return qvectors2_impl(ndim, dq, qMax, geometry, out, max_vecs, shell_sizes, max_shells, numq);
}

// Cylinder S(q) per wave-vector (raw sums, not divided) + num1DParticles (src/estimator.cpp:4318-4327, slice 0 only).
int orc_ssf_cyl(int ndim, const double* side, const unsigned* periodic, const double* beads, int M, int N, int Next,
                const double* q, int nq, double maxR, double* out, int* n_inside) {
    if (ndim < 2 || ndim > 3) return -100;
    const Box b = make_box(ndim, side, periodic);
    for (int k = 0; k < nq; ++k)
        out[k] = ndim == 2 ? ssf_cyl_one<2>(b, beads, M, N, Next, q + 2 * k, maxR) : ssf_cyl_one<3>(b, beads, M, N, Next, q + 3 * k, maxR);
    if (n_inside) {
        int tot = 0;
        for (int p = 0; p < N; ++p) if (cyl_include(beads + static_cast<size_t>(p) * ndim, maxR)) tot++;
        *n_inside = tot;
    }
    return 0;
}

// Elastic scattering: what ElasticScatteringEstimatorGpu::accumulate adds to `estimator` per measurement
// (src/estimator.cpp:4197-4235 with gpu_isf<true>, src/estimator_gpu.cu:66-164): for each q the M/2 + 1 "blocks"
// tau = 0..M/2 each add 2 * inorm * sum_{m1} sum_{n1,n2} cos(q.(r(m1+tau, n2) - r(m1, n1))), inorm = 1/(N M).
// The sum order inside a block differs from the device's tree; tau ascending here.
int orc_elastic(int ndim, const double* beads, int M, int N, int Next, const double* q, int nq, double* out, int nthreads) {
    if (ndim < 1 || ndim > 3) return -100;
    const double inorm = 1.0 / (static_cast<double>(N) * M);
    auto work = [&](int k0, int k1) {
        for (int k = k0; k < k1; ++k) {
            double es = 0.0;
            for (int tau = 0; tau <= M / 2; ++tau) {
                double blk = 0.0;
                for (int m1 = 0; m1 < M; ++m1) {
                    const int m2 = (m1 + tau) % M;
                    for (int n1 = 0; n1 < N; ++n1)
                        for (int n2 = 0; n2 < N; ++n2) {
                            double q_dot_sep = 0.0;
                            for (int d = 0; d < ndim; ++d)
                                q_dot_sep += q[k * ndim + d] * (beads[(static_cast<size_t>(m2) * Next + n2) * ndim + d] -
                                                                beads[(static_cast<size_t>(m1) * Next + n1) * ndim + d]);
                            blk += cos(q_dot_sep);
                        }
                }
                es += 2.0 * blk * inorm;
            }
            out[k] = es;
        }
    };
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> pool;
    for (int w = 0; w < nthreads; ++w) {
        const int k0 = static_cast<int>(static_cast<long>(nq) * w / nthreads), k1 = static_cast<int>(static_cast<long>(nq) * (w + 1) / nthreads);
        if (k1 > k0) pool.emplace_back(work, k0, k1);
    }
    for (auto& th : pool) th.join();
    return 0;
}

// ---- virial ----------------------------------------------------------------------------------------------------------
// delta[M][Next][ndim] (padding columns zero) for every active bead: the host-side input of pimcb_virial_sums.
int orc_virial_delta(int ndim, const double* side, const unsigned* periodic, const double* beads, int M, int N, int Next,
                     const int* next, int window, double* delta) {
    if (ndim < 1 || ndim > 3) return -100;
    const Box b = make_box(ndim, side, periodic);
    const Links L(next, M, N, Next);
    std::fill(delta, delta + static_cast<size_t>(M) * Next * ndim, 0.0);
    for (int s = 0; s < M; ++s)
        for (int p = 0; p < N; ++p) {
            double* d = delta + (static_cast<size_t>(s) * Next + p) * ndim;
            if (ndim == 1) virial_delta<1>(b, beads, L, s, p, window, d);
            else if (ndim == 2) virial_delta<2>(b, beads, L, s, p, window, d);
            else virial_delta<3>(b, beads, L, s, p, window, d);
        }
    return 0;
}

// out[M][4] per slice (see virial_slice); t2_parity: -1 all slices, 0/1 slices of that parity, -2 none.
int orc_virial_sums_ext(int ndim, const double* side, const unsigned* periodic, const double* beads, int M, int N, int Next,
                        const int* next, int window, const double* dVdr, const double* d2V, int len, double dr,
                        const double* extdVdr, const double* extd2V, int t2_parity, double* out, int nthreads,
                        const double* gext, const double* g2ext);
int orc_virial_sums(int ndim, const double* side, const unsigned* periodic, const double* beads, int M, int N, int Next,
                    const int* next, int window, const double* dVdr, const double* d2V, int len, double dr,
                    const double* extdVdr, const double* extd2V, int t2_parity, double* out, int nthreads) {
    return orc_virial_sums_ext(ndim, side, periodic, beads, M, N, Next, next, window, dVdr, d2V, len, dr, extdVdr, extd2V, t2_parity,
                               out, nthreads, nullptr, nullptr);
}
// the same with a non-trivial external potential: gext [M][Next][ndim], g2ext [M][Next] (either may be null)
int orc_virial_sums_ext(int ndim, const double* side, const unsigned* periodic, const double* beads, int M, int N, int Next,
                        const int* next, int window, const double* dVdr, const double* d2V, int len, double dr,
                        const double* extdVdr, const double* extd2V, int t2_parity, double* out, int nthreads,
                        const double* gext, const double* g2ext) {
    if (ndim < 1 || ndim > 3) return -100;
    const Box b = make_box(ndim, side, periodic);
    const Links L(next, M, N, Next);
    VirialTables tab{dVdr, d2V, len, dr, {extdVdr ? extdVdr[0] : 0.0, extdVdr ? extdVdr[1] : 0.0},
                     {extd2V ? extd2V[0] : 0.0, extd2V ? extd2V[1] : 0.0}};
    auto work = [&](int s0, int s1) {
        for (int s = s0; s < s1; ++s) {
            const bool t2 = t2_parity == -1 || (t2_parity >= 0 && (s % 2) == t2_parity);
            double* o = out + static_cast<size_t>(s) * 4;
            if (ndim == 1) virial_slice<1>(b, tab, beads, L, N, s, window, t2, o, gext, g2ext);
            else if (ndim == 2) virial_slice<2>(b, tab, beads, L, N, s, window, t2, o, gext, g2ext);
            else virial_slice<3>(b, tab, beads, L, N, s, window, t2, o, gext, g2ext);
        }
    };
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> pool;
    for (int w = 0; w < nthreads; ++w) {
        const int s0 = static_cast<int>(static_cast<long>(M) * w / nthreads), s1 = static_cast<int>(static_cast<long>(M) * (w + 1) / nthreads);
        if (s1 > s0) pool.emplace_back(work, s0, s1);
    }
    for (auto& th : pool) th.join();
    return 0;
}

// VirialEnergyEstimator::accumulate, src/estimator.cpp:1086-1250, for ONE configuration on top of the per-slice sums
// vir[M][4] (orc_virial_sums / pimcb_virial_sums), vint[M], f2[M]; external potential "free".  out[19] in the
// estimator's column order {K_op, K_cv, V_op, V_cv, E, E_mu, K_op/N, K_cv/N, V_op/N, V_cv/N, E/N, EEcv*Beta^2,
// Ecv*Beta, dEdB, CvCov1, CvCov2, CvCov3, E_th, P}.  Upstream accumulates "cVCov2" under a misspelt key
// (estIndex["cVCov2"], :1239), which lands in a NEW map slot with index 0 -- i.e. it is added to K_op.  That quirk is
// reproduced when `quirk` is non-zero (column CvCov2 stays 0, K_op gets totEcv*beta*dEdB added).
int orc_virial_energy(int nd, const double* side, const unsigned* periodic, const double* beads, int M, int N, int Next,
                      const int* next, int window, const double* vir, const double* vint, const double* f2,
                      const double* VFactor, const double* gradVFactor, double tau, double lambda, double mu,
                      double tailV_potential, int quirk, double* out) {
    if (nd < 1 || nd > 3) return -100;
    const Box b = make_box(nd, side, periodic);
    const Links L(next, M, N, Next);
    double volume = 1.0;
    for (int d = 0; d < nd; ++d) volume *= side[d];
    const int numParticles = N, numTimeSlices = M, virialWindow = window;
    const double beta = 1.0 * numTimeSlices * tau;
    const double tailV = (1.0 * numParticles * numParticles / volume) * tailV_potential;
    const double thermTerm1 = (0.5 * nd / tau) * numParticles;
    const double T1 = 0.5 * nd * numParticles / (1.0 * virialWindow * tau);
    const double exchangeNorm = 1.0 / (4.0 * virialWindow * pow(tau, 2) * lambda * numTimeSlices);
    double Pressure = nd * numParticles, P2 = 0.0, P3 = 0.0, T2 = 0.0, T3 = 0.0, T4 = 0.0, T5 = 0.0, thermE = 0.0, virKinTerm = 0.0, totVop = 0.0;
    auto sepv = [&](int s1, int p1, int s2, int p2, double* sep) {
        for (int d = 0; d < nd; ++d) {
            sep[d] = beads[(static_cast<size_t>(s1) * Next + p1) * nd + d] - beads[(static_cast<size_t>(s2) * Next + p2) * nd + d];
            sep[d] -= b.pSide[d] * std::floor(sep[d] * b.sideInv[d] + 0.5);
        }
    };
    for (int slice = 0; slice < numTimeSlices; slice++)
        for (int ptcl = 0; ptcl < N; ptcl++) {
            double vel2[3] = {0, 0, 0}, vel1[3] = {0, 0, 0}, sep[3];
            int ns = slice, np = ptcl;
            L.fwd(ns, np);
            if (ns >= 0 && np >= 0) sepv(ns, np, slice, ptcl, vel2);                                 // getVelocity, path.h:189-203
            int os = slice, op = ptcl;
            for (int gamma = 1; gamma <= virialWindow; gamma++) {
                int gs = slice, gp = ptcl;
                for (int m = 0; m < gamma; ++m) L.fwd(gs, gp);
                sepv(gs, gp, os, op, sep);
                for (int d = 0; d < nd; ++d) vel1[d] += sep[d];
                os = gs; op = gp;
            }
            double d12 = 0.0, d22 = 0.0;
            for (int d = 0; d < nd; ++d) { d12 += vel1[d] * vel2[d]; d22 += vel2[d] * vel2[d]; }
            T2 -= d12;
            thermE -= d22;
        }
    P2 = thermE;
    T2 *= exchangeNorm;
    P2 /= (2.0 * lambda * tau * numTimeSlices);
    Pressure += P2;
    for (int slice = 0; slice < numTimeSlices; slice++) {
        const int eo = slice % 2;
        const bool corr = gradVFactor[eo] > kEPS;
        const double* v = vir + static_cast<size_t>(slice) * 4;
        const double c2 = 2.0 * gradVFactor[eo] * pow(tau, 3) * lambda;
        T3 += VFactor[eo] * tau * v[2];                                                              // deltaDOTgradUterm1
        T4 += corr ? v[3] * c2 : 0.0;                                                                // deltaDOTgradUterm2
        T5 += orc_deriv_potential_action_tau(vint[slice], f2 ? f2[slice] : 0.0, slice, VFactor, gradVFactor, tau, lambda);
        if (corr) virKinTerm += (f2 ? f2[slice] : 0.0) * gradVFactor[eo] * pow(tau, 3) * lambda;     // virialKinCorrection, action.cpp:1786-1803
        if (eo == 0) totVop += 0.0 + vint[slice];
        P3 += VFactor[eo] * tau * v[0] + (corr ? v[1] * c2 : 0.0);                                   // rDOTgradUterm1 + term2
    }
    P3 *= (1.0 / (2.0 * numTimeSlices));
    Pressure -= P3;
    Pressure /= (nd * tau * volume);
    totVop /= (0.5 * numTimeSlices);
    totVop += tailV;
    T3 /= (2.0 * beta);
    T4 /= (1.0 * beta);
    T5 /= (1.0 * numTimeSlices);
    virKinTerm /= (0.5 * beta);
    thermE *= (0.25 / (lambda * tau)) / (tau * numTimeSlices);
    thermE += thermTerm1;
    thermE += T5;
    thermE += tailV;
    const double totEcv = T1 + T2 + T3 + T4 + T5 + tailV;
    const double Kcv = T1 + T2 + T3 + T4 + virKinTerm;
    double dEdB = (-1.0 * T1 - 2.0 * T2 + 2.0 * T4) / tau;
    for (int slice = 0; slice < numTimeSlices; slice++) {
        const int eo = slice % 2;
        double d2U = 0.0;
        if (gradVFactor[eo] > kEPS) d2U += 6.0 * gradVFactor[eo] * tau * lambda * (f2 ? f2[slice] : 0.0);   // action.cpp:776-787
        dEdB += d2U / (1.0 * numTimeSlices);
    }
    dEdB *= beta * beta / (1.0 * numTimeSlices);
    for (int k = 0; k < 19; ++k) out[k] = 0.0;
    out[0] = totEcv - totVop; out[1] = Kcv; out[2] = totVop; out[3] = totEcv - Kcv; out[4] = totEcv;
    out[5] = totEcv - mu * numParticles;
    if (numParticles > 0) {
        out[6] = (totEcv - totVop) / (1.0 * numParticles); out[7] = Kcv / (1.0 * numParticles);
        out[8] = totVop / (1.0 * numParticles); out[9] = (totEcv - Kcv) / (1.0 * numParticles); out[10] = totEcv / (1.0 * numParticles);
    }
    out[11] = totEcv * thermE * beta * beta;
    out[12] = totEcv * beta;
    out[13] = dEdB;
    out[14] = totEcv * thermE * beta * beta * totEcv * beta;
    if (quirk) out[0] += totEcv * beta * dEdB; else out[15] = totEcv * beta * dEdB;
    out[16] = totEcv * thermE * beta * beta * dEdB;
    out[17] = thermE;
    out[18] = Pressure;
    return 0;
}

// Number of time slices from (T, tau) or explicit P: src/setup.cpp:998-1012.  Returns M, writes tau.
int orc_time_slices(double T, double tau_in, int P_in, double* tau_out) {
    int numTimeSlices;
    double tau;
    if (P_in <= 0) {
        tau = tau_in;
        numTimeSlices = static_cast<int>(1.0 / (T * tau) + kEPS);
        if ((numTimeSlices % 2) != 0) numTimeSlices--;
    } else {
        numTimeSlices = P_in;
        if ((numTimeSlices % 2) != 0) numTimeSlices--;
        tau = 1.0 / (T * numTimeSlices);
    }
    if (tau_out) *tau_out = tau;
    return numTimeSlices;
}

}  // extern "C"
