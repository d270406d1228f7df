#include "state_file.h"

#include <fstream>
#include <istream>
#include <sstream>

namespace {

// "(r0,r1) x (c0,c1)" then everything up to '[' (include/common.h:306-372)
bool readHeader(std::istream& is, int& rows, int& cols) {
    int r0, r1, c0, c1;
    char ch;
    if (!(is >> ch) || ch != '(') return false;
    if (!(is >> r0 >> ch) || ch != ',') return false;
    if (!(is >> r1 >> ch) || ch != ')') return false;
    if (!(is >> ch) || ch != 'x') return false;
    if (!(is >> ch) || ch != '(') return false;
    if (!(is >> c0 >> ch) || ch != ',') return false;
    if (!(is >> c1 >> ch) || ch != ')') return false;
    rows = r1 - r0 + 1;
    cols = c1 - c0 + 1;
    if (rows <= 0 || cols <= 0) return false;
    while (is >> ch)
        if (ch == '[') return true;
    return false;
}

bool readFooter(std::istream& is) {
    char ch;
    while (is >> ch)
        if (ch == ']') return true;
    return false;
}

// "(a,b,...)" (include/common.h:270-300)
template <class T, size_t K>
bool readTuple(std::istream& is, std::array<T, K>& a) {
    char ch;
    if (!(is >> ch) || ch != '(') return false;
    for (size_t i = 0; i < K; ++i) {
        if (!(is >> a[i])) return false;
        if (i + 1 < K && (!(is >> ch) || ch != ',')) return false;
    }
    return (is >> ch) && ch == ')';
}

template <class T, class Read>
bool readArray(std::istream& is, int& rows, int& cols, std::vector<T>& out, Read readOne) {
    if (!readHeader(is, rows, cols)) return false;
    out.resize(static_cast<size_t>(rows) * cols);
    for (auto& v : out)
        if (!readOne(is, v)) return false;
    return readFooter(is);
}

}  // namespace

bool readStateText(std::istream& in, PimcState& st, std::string& err) {
    st = PimcState();
    if (!(in >> st.headerWorldLines)) { err = "missing world-line count on the first line"; return false; }
    // skip the acceptance / estimator lines until a line starts with '(' (src/pimc.cpp:1136-1142)
    std::string line;
    std::getline(in, line);
    while (in.good() && in.peek() != '(') std::getline(in, line);
    if (!in.good()) { err = "no beads array found"; return false; }
    int r = 0, c = 0;
    if (!readArray(in, st.numTimeSlices, st.numWorldLines, st.beads,
                   [](std::istream& is, dVec& v) { return readTuple(is, v); })) { err = "malformed beads array"; return false; }
    if (!readArray(in, r, c, st.nextLink, [](std::istream& is, beadLocator& v) { return readTuple(is, v); }) ||
        r != st.numTimeSlices || c != st.numWorldLines) { err = "malformed nextLink array"; return false; }
    if (!readArray(in, r, c, st.prevLink, [](std::istream& is, beadLocator& v) { return readTuple(is, v); }) ||
        r != st.numTimeSlices || c != st.numWorldLines) { err = "malformed prevLink array"; return false; }
    if (!readArray(in, r, c, st.wormBeads, [](std::istream& is, unsigned& v) { return static_cast<bool>(is >> v); }) ||
        r != st.numTimeSlices || c != st.numWorldLines) { err = "malformed worm.beads array"; return false; }
    // empty beads are unlinked (src/pimc.cpp:1228-1239)
    for (size_t k = 0; k < st.wormBeads.size(); ++k)
        if (!st.wormBeads[k]) { st.nextLink[k] = {XXX, XXX}; st.prevLink[k] = {XXX, XXX}; }
    // links are followed by leftPack and by the kinetic / virial estimators: a target outside the arrays is a damaged file
    // (memory safety); a target that is inactive is a hand-made file whose links were never maintained -- accepted for the
    // position-only estimators, reported by linksClosed()
    auto inRange = [&](const beadLocator& b) {
        return (b[0] == XXX && b[1] == XXX) || (b[0] >= 0 && b[0] < st.numTimeSlices && b[1] >= 0 && b[1] < st.numWorldLines);
    };
    for (size_t k = 0; k < st.wormBeads.size(); ++k)
        if (st.wormBeads[k] && (!inRange(st.nextLink[k]) || !inRange(st.prevLink[k]))) {
            std::ostringstream msg;
            msg << "bead (" << k / st.numWorldLines << "," << k % st.numWorldLines << ") links to a bead outside the arrays";
            err = msg.str();
            return false;
        }
    st.numBeadsAtSlice.assign(st.numTimeSlices, 0);
    for (int s = 0; s < st.numTimeSlices; ++s)
        for (int p = 0; p < st.numWorldLines; ++p) st.numBeadsAtSlice[s] += st.wormBeads[st.idx(s, p)] ? 1 : 0;
    return true;
}

bool readStateFile(const std::string& fileName, PimcState& st, std::string& err) {
    std::ifstream f(fileName);
    if (!f) { err = "cannot open " + fileName; return false; }
    return readStateText(f, st, err);
}

int PimcState::numBeadsOn() const {
    int n = 0;
    for (unsigned b : wormBeads) n += b ? 1 : 0;
    return n;
}

bool PimcState::isDiagonal() const {
    if (numTimeSlices == 0) return false;
    for (int s = 0; s < numTimeSlices; ++s)
        if (numBeadsAtSlice[s] != numBeadsAtSlice[0]) return false;
    for (size_t k = 0; k < wormBeads.size(); ++k)
        if (wormBeads[k] && (nextLink[k][0] == XXX || nextLink[k][1] == XXX || prevLink[k][0] == XXX || prevLink[k][1] == XXX))
            return false;
    return true;
}

// Closed world lines: every active bead links to ACTIVE beads and prev(next(b)) == b == next(prev(b)).  What the kinetic and
// virial estimators walk; files written by the reference always satisfy it.
bool PimcState::linksClosed() const {
    if (nextLink.size() != beads.size() || prevLink.size() != beads.size()) return false;
    for (size_t k = 0; k < wormBeads.size(); ++k) {
        if (!wormBeads[k]) continue;
        const beadLocator &n = nextLink[k], &p = prevLink[k];
        if (n[0] == XXX || n[1] == XXX || p[0] == XXX || p[1] == XXX) return false;
        if (!wormBeads[idx(n[0], n[1])] || !wormBeads[idx(p[0], p[1])]) return false;
        const beadLocator me{static_cast<int>(k / numWorldLines), static_cast<int>(k % numWorldLines)};
        if (prevLink[idx(n[0], n[1])] != me || nextLink[idx(p[0], p[1])] != me) return false;
    }
    return true;
}

bool PimcState::isLeftPacked() const {
    for (int s = 0; s < numTimeSlices; ++s)
        for (int p = 0; p < numWorldLines; ++p)
            if ((wormBeads[idx(s, p)] != 0) != (p < numBeadsAtSlice[s])) return false;
    return true;
}

// Stable compaction of the active beads of every slice into the leading columns, with the link arrays relabelled to the
// new columns (what Path::leftPack, src/path.cpp:145-189, arrives at by swapping hole and bead one at a time): the
// world-line connectivity -- permutation cycles included -- survives, which the kinetic and virial estimators need.
void PimcState::leftPack() {
    const bool linked = nextLink.size() == beads.size() && prevLink.size() == beads.size();
    std::vector<int> newcol(beads.size(), XXX);
    for (int s = 0; s < numTimeSlices; ++s) {
        int w = 0;
        for (int p = 0; p < numWorldLines; ++p)
            if (wormBeads[idx(s, p)]) newcol[idx(s, p)] = w++;
    }
    auto relabel = [&](const beadLocator& b) {
        if (b[0] == XXX || b[1] == XXX || newcol[idx(b[0], b[1])] == XXX) return beadLocator{XXX, XXX};   // no link / inactive target
        return beadLocator{b[0], newcol[idx(b[0], b[1])]};
    };
    std::vector<dVec> nb(beads.size());                     // vacated columns: zero positions, no links, flag off
    for (auto& v : nb) v.fill(0.0);
    std::vector<beadLocator> nn(beads.size(), beadLocator{XXX, XXX}), np_(beads.size(), beadLocator{XXX, XXX});
    std::vector<unsigned> on(beads.size(), 0u);
    for (int s = 0; s < numTimeSlices; ++s)
        for (int p = 0; p < numWorldLines; ++p) {
            const size_t from = idx(s, p);
            if (!wormBeads[from]) continue;
            const size_t to = idx(s, newcol[from]);
            nb[to] = beads[from];
            on[to] = 1u;
            if (linked) { nn[to] = relabel(nextLink[from]); np_[to] = relabel(prevLink[from]); }
        }
    beads.swap(nb);
    wormBeads.swap(on);
    if (linked) { nextLink.swap(nn); prevLink.swap(np_); } else { nextLink.clear(); prevLink.clear(); }
}

void PimcState::putInside(const Container& box) {
    for (auto& b : beads) box.putInside(b);
}
