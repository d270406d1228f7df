// table_view.h -- flat, read-only view of a TabulatedPotential's lookup tables (include/potential.h:148-157 upstream:
// lookupV / lookupdVdr / lookupd2Vdr2, dr, tableLength, extV / extdVdr / extd2Vdr2 -- all `protected` there).
// The one addition the potential classes need for the B200 pair sums is a public accessor returning this struct
// (upstream.patch: PotentialBase::tableView() virtual, default "no table"; TabulatedPotential<T>::lookupView();
// AzizPotential::tableView() override).  Depends on nothing but <array>, so potential.h can include it.
#ifndef PIMCB_TABLE_VIEW_H
#define PIMCB_TABLE_VIEW_H

#include <array>

struct TableView {
    const double* V = nullptr;
    const double* dVdr = nullptr;
    const double* d2Vdr2 = nullptr;
    int tableLength = 0;
    double dr = 0.0;
    std::array<double, 2> extV{}, extdVdr{}, extd2Vdr2{};
};

#endif
