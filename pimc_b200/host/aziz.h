// aziz.h -- stand-alone He-4 Aziz HFDHE2 pair potential with the reference's lookup-table semantics
// (AzizPotential, include/potential.h:933-1018 + src/potential.cpp:1741-1909; TabulatedPotential,
// include/potential.h:129-277).  Used by the stand-alone tools; inside a reference checkout the reference's own
// AzizPotential supplies the tables through tableView().
#ifndef PIMCB_AZIZ_H
#define PIMCB_AZIZ_H

#include "pimc_compat.h"

class AzizPotential : public PotentialBase {
public:
    AzizPotential(int year, const Container* box);
    double V(const dVec& r) override;          // direct lookup, potential.h:985-989
    dVec gradV(const dVec& r) override;        // potential.h:997-1003
    double valueV(double r) const;             // potential.cpp:1822-1842
    double valuedVdr(double r) const;          // potential.cpp:1849-1875
    double valued2Vdr2(double r) const;        // potential.cpp:1881-1909
    double grad2V(const dVec& r) override;     // potential.h:1010-1016 (direct lookup of d2V/dr2)
    TableView tableView() const override;      // the accessor the B200 action needs (upstream.patch adds it to the reference's class)
private:
    double rm, A, epsilon, alpha, beta, D, C6, C8, C10;
    double dr = 0.0;
    int tableLength = 0;
    std::vector<double> lookupV, lookupdVdr, lookupd2Vdr2;
    std::array<double, 2> extV{}, extdVdr{}, extd2Vdr2{};
    double F(double x) const;
    double dF(double x) const;
    double d2F(double x) const;
    double direct(const std::vector<double>& table, const std::array<double, 2>& ext, double r) const;
};

#endif
