// ref_aziz_shim.cpp -- test infrastructure (never linked by the product): C entry points around the REFERENCE'S OWN
// Aziz pair potential.  The upstream class (TabulatedPotential<T> + AzizPotential: parameter sets, valueV / valuedVdr /
// valued2Vdr2, initLookupTable, direct lookups V / gradV / grad2V, tail correction) is compiled from the upstream tree
// where it lies: oracle/ref_aziz_extract.py cuts those blocks out of include/potential.h and src/potential.cpp into a
// scratch include file at build time (nothing is stored in this repository), include/array_math.h is included
// directly.  This file only supplies the few surrounding types the blocks mention -- written from scratch with the
// upstream member names -- because the real headers pull Boost and <mdspan>:
//   EPS (include/common.h:89), dVec (include/common.h:104), DynamicArray<double,1> (include/dynamic_array.h; resize /
//   fill / operator()(i)), Container::maxSep (include/container.h), constants()->rc() (include/constants.h),
//   PotentialBase::tailV (include/potential.h:40-117).
// Built by `make -C oracle ref` with -O2 -ffp-contract=off (each source operation rounded once, like liboracle.so).
#include <array>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <vector>

#ifndef NDIM
#define NDIM 3
#endif
#define EPS 1.0E-7
typedef std::array<double, NDIM> dVec;
#include "array_math.h"        // the reference's own header: dot(), scalar * array

using namespace std;

template <class T, int Rank> class DynamicArray;
template <class T>
class DynamicArray<T, 1> {
public:
    void resize(size_t n) { v_.resize(n); }
    void fill(const T& x) { std::fill(v_.begin(), v_.end(), x); }
    T& operator()(size_t i) { return v_[i]; }
    const T& operator()(size_t i) const { return v_[i]; }
    const T* data() const { return v_.data(); }
private:
    std::vector<T> v_;
};

class Container {
public:
    double maxSep = 0.0;
};

class ConstantParameters {
public:
    double rc() const { return rc_; }
    double rc_ = 0.0;
};
static ConstantParameters g_constants;
static ConstantParameters* constants() { return &g_constants; }

class PotentialBase {
public:
    PotentialBase() {}
    virtual ~PotentialBase() {}
    virtual double V(const dVec&) { return 0.0; }
    virtual dVec gradV(const dVec&) { return dVec{}; }
    virtual double grad2V(const dVec&) { return 0.0; }
    double tailV = 0.0;
};

#include "ref_aziz_extract.inc"   // generated at build time from the upstream tree, deleted afterwards

// the tables are protected members of TabulatedPotential<AzizPotential>
struct Probe : public AzizPotential {
    Probe(int year, const Container* box) : AzizPotential(year, box) {}
    int length() const { return tableLength; }
    double step() const { return dr; }
    const double* tV() const { return lookupV.data(); }
    const double* tdV() const { return lookupdVdr.data(); }
    const double* td2V() const { return lookupd2Vdr2.data(); }
};

extern "C" {

void* refaziz_create(int year, double maxSep, double rc) {
    Container box;
    box.maxSep = maxSep;
    g_constants.rc_ = rc;
    return new Probe(year, &box);
}
void refaziz_destroy(void* h) { delete static_cast<Probe*>(h); }
int refaziz_table_length(void* h) { return static_cast<Probe*>(h)->length(); }
double refaziz_dr(void* h) { return static_cast<Probe*>(h)->step(); }
double refaziz_tail(void* h) { return static_cast<Probe*>(h)->tailV; }
void refaziz_tables(void* h, double* V, double* dV, double* d2V) {
    Probe* p = static_cast<Probe*>(h);
    const size_t n = sizeof(double) * p->length();
    if (V) std::memcpy(V, p->tV(), n);
    if (dV) std::memcpy(dV, p->tdV(), n);
    if (d2V) std::memcpy(d2V, p->td2V(), n);
}
// the analytic functions the tables are built from
void refaziz_values(void* h, const double* r, int n, double* v, double* dv, double* d2v) {
    Probe* p = static_cast<Probe*>(h);
    for (int i = 0; i < n; ++i) {
        if (v) v[i] = p->valueV(r[i]);
        if (dv) dv[i] = p->valuedVdr(r[i]);
        if (d2v) d2v[i] = p->valued2Vdr2(r[i]);
    }
}
// the direct table lookups the action calls: V(sep), gradV(sep), grad2V(sep) on explicit separation vectors
void refaziz_lookup(void* h, const double* sep, int n, double* v, double* gradv, double* grad2v) {
    Probe* p = static_cast<Probe*>(h);
    for (int i = 0; i < n; ++i) {
        dVec s;
        for (int d = 0; d < NDIM; ++d) s[d] = sep[i * NDIM + d];
        if (v) v[i] = p->V(s);
        if (gradv) {
            const dVec g = p->gradV(s);
            for (int d = 0; d < NDIM; ++d) gradv[i * NDIM + d] = g[d];
        }
        if (grad2v) grad2v[i] = p->grad2V(s);
    }
}

}  // extern "C"
