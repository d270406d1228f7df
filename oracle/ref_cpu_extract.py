#!/usr/bin/env python
"""Test infrastructure: cuts the reference's OWN CPU implementation of the measurement hot path out of the upstream tree,
where it lies, into scratch include files that oracle/ref_cpu_shim.cpp compiles (oracle/Makefile, target `ref`).
Nothing upstream is copied into this repository: the output directory is a scratch directory the Makefile deletes after
the compile; only the built oracle/_ref/librefcpu<NDIM>d.so (git-ignored) remains.

The upstream translation units (estimator.cpp, action.cpp, ...) cannot be compiled as they are -- common.h pulls Boost,
DynamicArray needs <mdspan> -- but the FUNCTION BODIES on the hot path only touch a handful of members.  Each body is
located by its signature and cut at the brace that closes it (comments and string literals skipped while counting):

    include/container.h   Container::putInBC                                  (member body, pasted inside the stand-in class)
    include/path.h        Path::getSeparation, Path::getVelocity
    include/common.h      enumerate, all_impl / all (bead comparison), apply_matrix_vector_product, the stream operators of
                          std::array and DynamicArray<T,2> (the text format of the state files)
    src/path.cpp          Path::leftPack (compiled by oracle/ref_pack_shim.cpp)
    src/worm.cpp          Worm::factor(state, bead)
    src/action.cpp        ActionBase::updateSepHist, LocalAction::potentialAction(), derivPotentialActionTau(int),
                          secondderivPotentialActionTau(int), derivPotentialActionLambda(int), V(int slice),
                          gradVSquared(int slice), rDOTgradUterm1/2, deltadotgradUterm1/2, virialKinCorrection
    src/estimator.cpp     EstimatorBase::getQVectors, getQVectors2, include(), num1DParticles(),
                          StaticStructureFactorEstimator::accumulate, IntermediateScatteringFunctionEstimator::accumulate,
                          EnergyEstimator::accumulate, VirialEnergyEstimator::accumulate,
                          CylinderStaticStructureFactorEstimator::accumulate

    usage: ref_cpu_extract.py <reference root> <output directory>
"""
import os
import re
import sys


def strip_code(line, in_block_comment):
    """The line with comments, string and char literals blanked out (for brace counting); returns (text, in_comment)."""
    out, i, n = [], 0, len(line)
    while i < n:
        if in_block_comment:
            j = line.find("*/", i)
            if j < 0:
                return "".join(out), True
            i, in_block_comment = j + 2, False
            continue
        c = line[i]
        if line.startswith("//", i):
            break
        if line.startswith("/*", i):
            in_block_comment, i = True, i + 2
            continue
        if c in "\"'":
            j = i + 1
            while j < n and line[j] != c:
                j += 2 if line[j] == "\\" else 1
            i = j + 1
            continue
        out.append(c)
        i += 1
    return "".join(out), in_block_comment


def body(lines, start_pat, after=0):
    """Lines of the definition whose signature matches start_pat: from that line to the brace closing its body."""
    rs = re.compile(start_pat)
    for i in range(after, len(lines)):
        if rs.search(lines[i]):
            depth, seen, in_c = 0, False, False
            for j in range(i, len(lines)):
                text, in_c = strip_code(lines[j], in_c)
                for ch in text:
                    if ch == "{":
                        depth += 1
                        seen = True
                    elif ch == "}":
                        depth -= 1
                if seen and depth == 0:
                    return j, lines[i:j + 1]
            break
    raise SystemExit(f"ref_cpu_extract: definition {start_pat!r} not found -- the upstream layout changed")


MANIFEST = {
    # output file: (upstream file, [signature patterns, in order])
    "container_members.inc": ("include/container.h", [r"^\s*void putInBC\s*\(dVec\s*&\s*r\)\s*const\s*\{"]),
    "path_inline.inc": ("include/path.h", [r"^inline dVec Path::getSeparation\s*\(", r"^inline dVec Path::getVelocity\s*\("]),
    "common_helpers.inc": ("include/common.h", [r"^constexpr auto enumerate\s*\(T && iterable\)", r"^void apply_matrix_vector_product\s*\(", r"^constexpr bool all_impl\s*\(",
                                                r"^constexpr bool all\s*\(const std::array<T, N>& a, const std::array<T, N>& b\)"]),
    "common_stream.inc": ("include/common.h", [
        r"^std::ostream& operator<<\(std::ostream& os, const std::array<T, N>& arr\)",
        r"^std::ostream& operator<<\(std::ostream& os, const DynamicArray<T, 2>& arr\)",
        r"^std::istream& operator>>\(std::istream& is, std::array<T, N>& a\)",
        r"^std::istream& operator>>\(std::istream& is, DynamicArray<T, 2>& arr\)",
    ]),
    "path_leftpack.inc": ("src/path.cpp", [r"^void Path::leftPack\(\)"]),
    "worm.inc": ("src/worm.cpp", [r"^double Worm::factor\s*\(const beadState state1, const beadLocator &bead2\)"]),
    "action.inc": ("src/action.cpp", [
        r"^inline void ActionBase::updateSepHist\s*\(",
        r"^double LocalAction::potentialAction\s*\(\s*\)\s*\{",
        r"^double LocalAction::derivPotentialActionTau\s*\(int slice\)\s*\{",
        r"^double LocalAction::secondderivPotentialActionTau\s*\(int slice\)\s*\{",
        r"^double LocalAction::derivPotentialActionLambda\s*\(int slice\)\s*\{",
        r"^std::array<double,2> LocalAction::V\s*\(const int slice\)\s*\{",
        r"^double LocalAction::gradVSquared\s*\(const int slice\)\s*\{",
        r"^double LocalAction::rDOTgradUterm1\s*\(const int slice\)\s*\{",
        r"^double LocalAction::rDOTgradUterm2\s*\(const int slice\)\s*\{",
        r"^double LocalAction::deltadotgradUterm1\s*\(const int slice\)\s*\{",
        r"^double LocalAction::deltadotgradUterm2\s*\(const int slice\)\s*\{",
        r"^double LocalAction::virialKinCorrection\s*\(const int slice\)\s*\{",
    ]),
    "estimator.inc": ("src/estimator.cpp", [
        r"^void EstimatorBase::getQVectors\s*\(std::vector<dVec> &qValues\)\s*\{",
        r"^std::vector <std::vector<dVec> > EstimatorBase::getQVectors2\s*\(",
        r"^inline bool include\s*\(const dVec &r, double maxR\)",
        r"^int num1DParticles\s*\(const Path &path, double maxR\)",
        r"^void EnergyEstimator::accumulate\s*\(\)",
        r"^void VirialEnergyEstimator::accumulate\s*\(\)",
        r"^void StaticStructureFactorEstimator::accumulate\s*\(\)",
        r"^void IntermediateScatteringFunctionEstimator::accumulate\s*\(\)",
        r"^void CylinderStaticStructureFactorEstimator::accumulate\s*\(\)",
    ]),
}
# template headers sit on the line above the signature
TEMPLATE_PREFIX = {r"^std::ostream& operator<<\(std::ostream& os, const std::array<T, N>& arr\)": 1,
                   r"^std::ostream& operator<<\(std::ostream& os, const DynamicArray<T, 2>& arr\)": 1,
                   r"^std::istream& operator>>\(std::istream& is, std::array<T, N>& a\)": 1,
                   r"^std::istream& operator>>\(std::istream& is, DynamicArray<T, 2>& arr\)": 1,
                   r"^constexpr auto enumerate\s*\(T && iterable\)": 3, r"^void apply_matrix_vector_product\s*\(": 1, r"^constexpr bool all_impl\s*\(": 1,
                   r"^constexpr bool all\s*\(const std::array<T, N>& a, const std::array<T, N>& b\)": 1}


def main(root, outdir):
    os.makedirs(outdir, exist_ok=True)
    for name, (path, pats) in MANIFEST.items():
        lines = open(os.path.join(root, path)).read().split("\n")
        parts = []
        for pat in pats:
            rs = re.compile(pat)
            start = next((i for i, l in enumerate(lines) if rs.search(l)), None)
            if start is None:
                raise SystemExit(f"ref_cpu_extract: {pat!r} not found in {path} -- the upstream layout changed")
            _, b = body(lines, pat)
            k = TEMPLATE_PREFIX.get(pat, 0)
            parts += lines[start - k:start] + b + [""]
        with open(os.path.join(outdir, name), "w") as f:
            f.write("\n".join(parts) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
