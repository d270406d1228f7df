"""Parity criterion shared by the GPU tests (north_star: <= 1e-10 relative, FP64)."""
import numpy as np

RTOL = 1e-10          # the tolerance BASELINE.json's north_star states
ATOL_SCALE = 1e-12    # absolute floor, in units of max(1, max|cpu|) of the vector under test


def assert_parity(gpu, cpu, what=""):
    gpu = np.asarray(gpu, dtype=np.float64)
    cpu = np.asarray(cpu, dtype=np.float64)
    assert gpu.shape == cpu.shape, f"{what}: shape {gpu.shape} vs {cpu.shape}"
    scale = max(1.0, float(np.max(np.abs(cpu)))) if cpu.size else 1.0
    err = np.abs(gpu - cpu)
    vec = float(err.max() / max(np.max(np.abs(cpu)), 1e-300)) if cpu.size else 0.0
    assert vec <= RTOL, f"{what}: max|gpu-cpu|/max|cpu| = {vec:.3e} > {RTOL}"
    bad = err > RTOL * np.abs(cpu) + ATOL_SCALE * scale
    assert not bad.any(), (f"{what}: {int(bad.sum())} elements off; worst abs err {err.max():.3e} "
                           f"at {np.unravel_index(err.argmax(), err.shape)}")
    return vec
