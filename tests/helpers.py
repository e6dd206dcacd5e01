"""Shared test helpers: seeded inputs and error norms."""
import numpy as np


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def plane_wave_ic(shape, vmin=-80.0, vmax=20.0):
    """The examples' initial condition (Tests/FD/Fenton/fenton.py:119-121)."""
    U = np.full(shape, vmin, dtype=np.float32)
    U[0:2] = vmax
    return U


def random_fields(shape, seed=0, nstates=3):
    """Seeded fields exercising every branch of the ionic models: U in [-85, 25], gates in [0,1]."""
    rng = np.random.default_rng(seed)
    U = rng.uniform(-85.0, 25.0, size=shape).astype(np.float32)
    states = [rng.uniform(0.0, 1.0, size=shape).astype(np.float32) for _ in range(nstates)]
    return U, states


def smooth_fields(shape, seed=0, nstates=3):
    """Smooth seeded fields (sum of a few cosines) so that many explicit steps stay stable."""
    rng = np.random.default_rng(seed)
    grids = np.meshgrid(*[np.linspace(0.0, 1.0, n) for n in shape], indexing='ij')
    U = np.zeros(shape, dtype=np.float64)
    for _ in range(4):
        k = rng.uniform(0.5, 3.0, size=len(shape))
        ph = rng.uniform(0, 2 * np.pi)
        U += np.cos(2 * np.pi * sum(ki * g for ki, g in zip(k, grids)) + ph)
    U = (-80.0 + 100.0 * (U - U.min()) / (U.max() - U.min())).astype(np.float32)
    states = [rng.uniform(0.2, 1.0, size=shape).astype(np.float32) for _ in range(nstates)]
    return U, states
