#include "aziz.h"

AzizPotential::AzizPotential(int year, const Container* box) {
    struct Params { int year; double eps, rm, D, alpha, beta, C6, C8, C10, A; };
    // parameter sets: Aziz et al. 1979 (JCP 70, 4330), 1987 (Mol. Phys. 61, 1487), 1995 (PRL 74, 1586);
    // values as tabulated upstream, src/potential.cpp:1749-1787
    static const Params sets[] = {
        {1979, 10.8, 2.9673, 1.241314, 13.353384, 0.0, 1.3732412, 0.4253785, 0.1781, 0.5448504E6},
        {1987, 10.948, 2.9673, 1.4826, 10.43329537, -2.27965105, 1.36745214, 0.42123807, 0.17473318, 1.8443101E5},
        {1995, 10.956, 2.9683, 1.438, 10.5717543, -2.07758779, 1.35186623, 0.4149514, 0.17151143, 1.86924404E5},
    };
    const Params* p = &sets[0];
    for (const Params& s : sets)
        if (s.year == year) p = &s;
    epsilon = p->eps; rm = p->rm; D = p->D; alpha = p->alpha; beta = p->beta;
    C6 = p->C6; C8 = p->C8; C10 = p->C10; A = p->A;

    // lookup tables: dr = 1e-6 rm (potential.cpp:1796), tableLength = int(maxSep/dr), abscissa accumulated by
    // repeated addition exactly as TabulatedPotential::initLookupTable does (potential.h:163-183)
    dr = (1.0E-6) * rm;
    tableLength = int(box->maxSep / dr);
    lookupV.resize(tableLength);
    lookupdVdr.resize(tableLength);
    lookupd2Vdr2.resize(tableLength);
    double r = 0;
    for (int n = 0; n < tableLength; n++) {
        lookupV[n] = valueV(r);
        lookupdVdr[n] = valuedVdr(r);
        lookupd2Vdr2[n] = valued2Vdr2(r);
        r += dr;
    }
    // tail correction (potential.cpp:1798-1806); the cutoff defaults to the box side (src/setup.cpp:1128-1130)
    const double rc = constants()->rc() > 0.0 ? constants()->rc() : box->side[NDIM - 1];
    const double rmorc = rm / rc;
    const double t2 = C6 * std::pow(rmorc, 3.0) / 3.0;
    const double t3 = C8 * std::pow(rmorc, 5.0) / 5.0;
    const double t4 = C10 * std::pow(rmorc, 7.0) / 7.0;
    tailV = 2.0 * M_PI * epsilon * (-rm * rm * rm * (t2 + t3 + t4));
}

double AzizPotential::F(double x) const {
    if (x >= D) return 1.0;
    const double u = D / x - 1.0;
    return std::exp(-u * u);
}

double AzizPotential::dF(double x) const {
    if (x >= D) return 0.0;
    const double ix = 1.0 / x;
    return 2.0 * D * ix * ix * (D * ix - 1.0) * std::exp(-(D * ix - 1.0) * (D * ix - 1.0));
}

double AzizPotential::valueV(double r) const {
    const double x = r / rm;
    const double Urep = A * std::exp(-alpha * x + beta * x * x);
    if (x < EPS) return 0.0;
    if (x < 0.01) return epsilon * Urep;
    const double ix2 = 1.0 / (x * x);
    const double ix6 = ix2 * ix2 * ix2;
    const double ix8 = ix6 * ix2;
    const double ix10 = ix8 * ix2;
    const double Uatt = -(C6 * ix6 + C8 * ix8 + C10 * ix10) * F(x);
    return epsilon * (Urep + Uatt);
}

double AzizPotential::valuedVdr(double r) const {
    const double x = r / rm;
    const double T1 = A * (-alpha + 2.0 * beta * x) * std::exp(-alpha * x + beta * x * x);
    if (x < EPS) return 0.0;
    if (x < 0.01) return (epsilon / rm) * T1;
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
    return (epsilon / rm) * (T1 + T2 + T3);
}

// second derivative of the damping function, potential.h:968-975
double AzizPotential::d2F(double x) const {
    if (x >= D) return 0.0;
    const double ix = 1.0 / x;
    const double u = D * ix - 1.0;
    return 2.0 * D * ix * ix * ix * (2.0 * D * D * D * ix * ix * ix - 4.0 * D * D * ix * ix - D * ix + 2.0) * std::exp(-u * u);
}

double AzizPotential::valued2Vdr2(double r) const {
    const double x = r / rm;
    const double ab = alpha - 2.0 * beta * x;
    const double T1 = A * (2 * beta + ab * ab) * std::exp(-alpha * x + beta * x * x);
    if (x < EPS) return 0.0;
    if (x < 0.01) return (epsilon / rm) * T1;                  // upstream's hard-core branch keeps the 1/rm prefactor
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
    return (epsilon / (rm * rm)) * (T1 + T2 + T3 + T4);
}

double AzizPotential::grad2V(const dVec& r) { return direct(lookupd2Vdr2, extd2Vdr2, std::sqrt(dot(r, r))); }

double AzizPotential::direct(const std::vector<double>& table, const std::array<double, 2>& ext, double r) const {
    const int k = int(r / dr);                 // potential.h:252
    if (k <= 0) return ext[0];
    if (k >= tableLength) return ext[1];
    return table[k];
}

double AzizPotential::V(const dVec& r) { return direct(lookupV, extV, std::sqrt(dot(r, r))); }

dVec AzizPotential::gradV(const dVec& r) {
    const double rnorm = std::sqrt(dot(r, r));
    const double g = direct(lookupdVdr, extdVdr, rnorm) / rnorm;
    dVec out;
    for (int i = 0; i < NDIM; ++i) out[i] = g * r[i];
    return out;
}

TableView AzizPotential::tableView() const {
    TableView v;
    v.V = lookupV.data();
    v.dVdr = lookupdVdr.data();
    v.d2Vdr2 = lookupd2Vdr2.data();
    v.extd2Vdr2 = extd2Vdr2;
    v.tableLength = tableLength;
    v.dr = dr;
    v.extV = extV;
    v.extdVdr = extdVdr;
    return v;
}
