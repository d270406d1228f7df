"""Synthetic He-4 worldline configurations and wave-vector sets.

Inputs for tests and bench.py.  Layout mirrors the reference's bead storage: a row-major AoS
`double[M][N_ext][NDIM]` (`DynamicArray<dVec,2> beads`, include/path.h:164) whose first N columns
of every slice are the active beads of a diagonal configuration and whose `N_ext - N` trailing
columns are padding (src/path.cpp:274-294 grows the extent and never shrinks it).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

LAMBDA_HE4 = 24.24 / 4.0030      # constants.cpp:128 lambda = 24.24/m, m = 4.0030 amu
BASE_SEED = 139853               # src/pdrive.cpp:41 base RNG seed of the reference


@dataclass(frozen=True)
class Shape:
    """One of BASELINE.json's named configurations."""
    name: str
    ndim: int
    N: int
    M: int
    T: float
    rho: float
    nq: int

    @property
    def side(self) -> np.ndarray:
        return np.full(self.ndim, (self.N / self.rho) ** (1.0 / self.ndim))   # src/container.cpp:88

    @property
    def tau(self) -> float:
        return 1.0 / (self.T * self.M)                                        # src/setup.cpp:1009


# C1..C4 of BASELINE.json (C5 = 64 x C2).  C3's 2-D density is not given upstream; 0.0432 A^-2.
C1 = Shape("C1", 3, 16, 124, 2.0, 0.02198, 64)
C2 = Shape("C2", 3, 256, 170, 1.5, 0.02198, 64)
C3 = Shape("C3", 2, 128, 250, 1.0, 0.0432, 289)
C4 = Shape("C4", 3, 1024, 320, 1.5, 0.02198, 256)
SHAPES = {s.name: s for s in (C1, C2, C3, C4)}


def put_in_bc(r: np.ndarray, side: np.ndarray) -> np.ndarray:
    """include/container.h:50-53 on an array of vectors."""
    return r - side * np.floor(r * (1.0 / side) + 0.5)


def gen_config(N: int, M: int, ndim: int, rho: float, T: float, seed: int = BASE_SEED, pad: int = 3) -> np.ndarray:
    """Lattice + jitter start, periodic Brownian-bridge worldlines; AoS [M][N+pad][ndim], padding zeroed."""
    rng = np.random.default_rng(seed)
    L = (N / rho) ** (1.0 / ndim)
    side = np.full(ndim, L)
    n_side = int(math.ceil(N ** (1.0 / ndim) - 1e-12))
    a = L / n_side
    grid = np.stack(np.meshgrid(*([np.arange(n_side)] * ndim), indexing="ij"), axis=-1).reshape(-1, ndim)[:N]
    base = (grid + 0.5) * a - 0.5 * L + rng.uniform(-0.15 * a, 0.15 * a, size=(N, ndim))
    tau = 1.0 / (T * M)
    steps = rng.normal(0.0, math.sqrt(2.0 * LAMBDA_HE4 * tau), size=(M, N, ndim))
    steps -= steps.mean(axis=0, keepdims=True)            # closed (periodic in imaginary time) paths
    walk = np.cumsum(steps, axis=0) - steps[0]
    pos = put_in_bc(base[None, :, :] + walk, side)
    beads = np.zeros((M, N + pad, ndim))
    beads[:, :N, :] = pos
    return beads


def gen_batch(shape: Shape, B: int, first: int = 0, pad: int = 3) -> np.ndarray:
    """[B][M][N+pad][ndim]; configuration k uses seed BASE_SEED + first + k."""
    return np.stack([gen_config(shape.N, shape.M, shape.ndim, shape.rho, shape.T, BASE_SEED + first + k, pad)
                     for k in range(B)])


def lattice_indices(nq: int, ndim: int, include_zero: bool = False) -> np.ndarray:
    """First nq integer vectors n sorted by (|n|^2, lexicographic)."""
    r = int(math.ceil((nq + 1) ** (1.0 / ndim))) + 2   # cube of half-width r contains the first nq shells
    g = np.stack(np.meshgrid(*([np.arange(-r, r + 1)] * ndim), indexing="ij"), axis=-1).reshape(-1, ndim)
    if not include_zero:
        g = g[np.any(g != 0, axis=1)]
    key = [g[:, d] for d in reversed(range(ndim))] + [np.sum(g * g, axis=1)]
    order = np.lexsort(key)
    return g[order][:nq]


def commensurate_q(nq: int, side: np.ndarray, include_zero: bool = False) -> np.ndarray:
    """q = (2 pi / side_j) * n_j as `--wavevector_type int` produces them (src/estimator.cpp:463)."""
    n = lattice_indices(nq, len(side), include_zero)
    return (2.0 * math.pi / np.asarray(side)) * n


def int_wavevector_text(nq: int, ndim: int) -> str:
    """The `--wavevector "..."` string whose `int` parse gives commensurate_q(nq, .)."""
    return " ".join(str(int(v)) for v in lattice_indices(nq, ndim).reshape(-1))


def float_q(nq: int, ndim: int, seed: int = 7, qmax: float = 2.5) -> np.ndarray:
    """Non-commensurate wave-vectors, rounded through float like `--wavevector_type float` (std::stof)."""
    rng = np.random.default_rng(seed)
    q = rng.uniform(-qmax, qmax, size=(nq, ndim))
    return q.astype(np.float32).astype(np.float64)


def aziz_table_numpy(max_sep: float, year: int = 1979, second: bool = False):
    """Aziz HFDHE2 lookup tables (V, dV/dr, dr) with the reference's construction (include/potential.h:163-183,
    src/potential.cpp:1741-1909): dr = 1e-6 rm, tableLength = int(maxSep/dr), abscissa by repeated addition.
    Vectorised numpy for benchmarks; values agree with the C++ builders to the last few ulp of exp()."""
    params = {1979: (10.8, 2.9673, 1.241314, 13.353384, 0.0, 1.3732412, 0.4253785, 0.1781, 0.5448504E6),
              1987: (10.948, 2.9673, 1.4826, 10.43329537, -2.27965105, 1.36745214, 0.42123807, 0.17473318, 1.8443101E5),
              1995: (10.956, 2.9683, 1.438, 10.5717543, -2.07758779, 1.35186623, 0.4149514, 0.17151143, 1.86924404E5)}
    eps, rm, D, alpha, beta, C6, C8, C10, A = params[year]
    dr = 1.0e-6 * rm
    n = int(max_sep / dr)
    steps = np.full(n, dr)
    steps[0] = 0.0
    r = np.cumsum(steps)                      # sequential float64 accumulation, r[k] = r[k-1] + dr
    x = r / rm
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        urep = A * np.exp(-alpha * x + beta * x * x)
        ix = 1.0 / x
        ix2 = ix * ix
        ix6 = ix2 * ix2 * ix2
        ix8, ix10 = ix6 * ix2, ix6 * ix2 * ix2
        F = np.where(x < D, np.exp(-(D * ix - 1.0) ** 2), 1.0)
        dF = np.where(x < D, 2.0 * D * ix2 * (D * ix - 1.0) * np.exp(-(D * ix - 1.0) ** 2), 0.0)
        disp = C6 * ix6 + C8 * ix8 + C10 * ix10
        V = eps * (urep - disp * F)
        T1 = A * (-alpha + 2.0 * beta * x) * np.exp(-alpha * x + beta * x * x)
        T2 = (6.0 * C6 * ix6 * ix + 8.0 * C8 * ix8 * ix + 10.0 * C10 * ix10 * ix) * F
        dV = (eps / rm) * (T1 + T2 - disp * dF)
    core = x < 0.01
    V = np.where(core, eps * urep, V)
    dV = np.where(core, (eps / rm) * T1, dV)
    zero = x < 1.0e-7
    V[zero] = 0.0
    dV[zero] = 0.0
    if not second:
        return V, dV, dr
    # d2V/dr2 (src/potential.cpp:1881-1909; damping function derivatives include/potential.h:958-975)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        ab = alpha - 2.0 * beta * x
        S1 = A * (2.0 * beta + ab * ab) * np.exp(-alpha * x + beta * x * x)
        u = D * ix - 1.0
        d2F = np.where(x < D, 2.0 * D * ix ** 3 * (2.0 * D ** 3 * ix ** 3 - 4.0 * D * D * ix2 - D * ix + 2.0) * np.exp(-u * u), 0.0)
        ix12 = ix10 * ix2
        S2 = -(42.0 * C6 * ix8 + 72.0 * C8 * ix10 + 110.0 * C10 * ix12) * F
        S3 = 2.0 * (6.0 * C6 * ix6 * ix + 8.0 * C8 * ix8 * ix + 10.0 * C10 * ix10 * ix) * dF
        S4 = -disp * d2F
        d2V = (eps / (rm * rm)) * (S1 + S2 + S3 + S4)
    d2V = np.where(core, (eps / rm) * S1, d2V)
    d2V[zero] = 0.0
    return V, dV, d2V, dr
