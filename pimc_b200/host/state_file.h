// state_file.h -- ingestion of the reference's text state files (OUTPUT/(g)ce-state-*.dat) for offline
// re-measurement of saved path configurations.
//
// Format (writer: src/pimc.cpp:925-975; array streaming: include/common.h:238-268 out, :270-399 in):
//     <numWorldLines>\n
//     <totAccepted>\t<totAttempted>\n  ... one line per move, one per estimator ...
//     (0,M-1) x (0,W-1)\n[ (x,y,z) (x,y,z) ... \n  ... ]\n        beads      DynamicArray<dVec,2>, %.16g
//     (0,M-1) x (0,W-1)\n[ (s,p) (s,p) ... ]\n                    nextLink   DynamicArray<beadLocator,2>
//     ... prevLink ...
//     (0,M-1) x (0,W-1)\n[ 1 1 0 ... ]\n                          worm.beads DynamicArray<unsigned,2>
//     <MTRand state words>\n
// The loader (src/pimc.cpp:1105-1290) reads numWorldLines, skips lines until one starts with '(', streams the four
// arrays in, unlinks empty beads, applies Container::putInside to every bead and counts the active beads per slice.
// States are written left-packed (Path::leftPack, src/path.cpp:145-189, called at pimc.cpp:935): the active beads
// of every slice sit in columns [0, n_slice).
#ifndef PIMCB_STATE_FILE_H
#define PIMCB_STATE_FILE_H

#include <iosfwd>
#include <string>
#include <vector>

#ifdef PIMCB_STANDALONE
#include "pimc_compat.h"
#else
#include "common.h"
#include "container.h"
#endif

struct PimcState {
    int headerWorldLines = 0;                 // first line of the file
    int numTimeSlices = 0, numWorldLines = 0; // extents of the arrays
    std::vector<dVec> beads;                  // [M][W]
    std::vector<beadLocator> nextLink, prevLink;
    std::vector<unsigned> wormBeads;          // 1 = bead on
    std::vector<int> numBeadsAtSlice;         // active beads per slice (pimc.cpp:1258-1268)

    size_t idx(int slice, int ptcl) const { return static_cast<size_t>(slice) * numWorldLines + ptcl; }
    int numBeadsOn() const;
    // All slices hold the same number of active beads and every active bead is linked both ways: the
    // configuration is diagonal (closed worldlines only), the only kind estimators sample (src/estimator.cpp:228).
    bool isDiagonal() const;
    bool linksClosed() const;              // every world line closed over active beads (needed by the kinetic / virial estimators)
    // Active beads of every slice in columns [0, n): true for files written by the reference; leftPack() restores it
    // for hand-made files: stable compaction of positions and flags with the links relabelled accordingly.
    bool isLeftPacked() const;
    void leftPack();
    // Container::putInside on every bead + active count per slice (src/pimc.cpp:1258-1268).
    void putInside(const Container& box);
};

// Returns false and sets `err` on malformed input.
bool readStateText(std::istream& in, PimcState& st, std::string& err);
bool readStateFile(const std::string& fileName, PimcState& st, std::string& err);

#endif
