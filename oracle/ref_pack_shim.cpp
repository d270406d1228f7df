// ref_pack_shim.cpp -- test infrastructure (never linked by the product): the REFERENCE'S OWN Path::leftPack
// (src/path.cpp:145-189), cut out of the upstream tree at build time by oracle/ref_cpu_extract.py, around a minimal mutable
// stand-in of the members it touches (beads, worm.beadOn / addBead / delBead, next / prev as references).  A separate
// translation unit from ref_cpu_shim.cpp, whose Path stand-in is a read-only view.  -> oracle/_ref/librefpack<NDIM>d.so
#include <array>
#include <cstddef>
#include <cstring>
#include <vector>

#ifndef NDIM
#define NDIM 3
#endif
#define XXX -1
typedef std::array<double, NDIM> dVec;
typedef std::array<int, 2> beadLocator;

struct BeadArray {                                   // DynamicArray<dVec,2> beads: extents() and operator()(beadLocator)
    std::vector<dVec> v;
    size_t R = 0, C = 0;
    std::array<size_t, 2> extents() const { return {R, C}; }
    dVec& operator()(const beadLocator& b) { return v[static_cast<size_t>(b[0]) * C + b[1]]; }
};

class Worm {                                         // include/worm.h: the bead-on flags
public:
    std::vector<unsigned int> on;
    size_t C = 0;
    int numBeadsOn = 0;
    int beadOn(const beadLocator& b) const { return on[static_cast<size_t>(b[0]) * C + b[1]]; }
    void addBead(const beadLocator& b) { on[static_cast<size_t>(b[0]) * C + b[1]] = 1; ++numBeadsOn; }
    void delBead(const beadLocator& b) { on[static_cast<size_t>(b[0]) * C + b[1]] = 0; --numBeadsOn; }
};

class Path {
public:
    int numTimeSlices = 0;
    BeadArray beads;
    Worm worm;
    std::vector<beadLocator> nextLink, prevLink;
    beadLocator& next(const beadLocator& b) { return nextLink[static_cast<size_t>(b[0]) * beads.C + b[1]]; }
    beadLocator& prev(const beadLocator& b) { return prevLink[static_cast<size_t>(b[0]) * beads.C + b[1]]; }
    void leftPack();
};

#include "path_leftpack.inc"       // upstream body

extern "C" int refpack_ndim(void) { return NDIM; }

// In place on beads[M][W][NDIM], next[M][W][2], prev[M][W][2], on[M][W].
extern "C" int refpack_left_pack(double* beads, int* next, int* prev, unsigned int* on, int M, int W) {
    Path p;
    const size_t n = static_cast<size_t>(M) * W;
    p.numTimeSlices = M;
    p.beads.R = M; p.beads.C = W; p.beads.v.resize(n);
    p.worm.C = W; p.worm.on.assign(on, on + n);
    p.nextLink.resize(n); p.prevLink.resize(n);
    std::memcpy(p.beads.v.data(), beads, sizeof(dVec) * n);
    std::memcpy(p.nextLink.data(), next, sizeof(beadLocator) * n);
    std::memcpy(p.prevLink.data(), prev, sizeof(beadLocator) * n);
    p.leftPack();
    std::memcpy(beads, p.beads.v.data(), sizeof(dVec) * n);
    std::memcpy(next, p.nextLink.data(), sizeof(beadLocator) * n);
    std::memcpy(prev, p.prevLink.data(), sizeof(beadLocator) * n);
    std::memcpy(on, p.worm.on.data(), sizeof(unsigned int) * n);
    return 0;
}
