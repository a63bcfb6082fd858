"""Synthetic point fields used by the parity tests, the golden fixtures and bench.py.

"DEM-like" = smooth trend + noise on continuous-uniform coordinates (no exact distance ties),
SURVEY.md §8d:  xy ~ U[0, E]^2,  z = 200 + 50 sin(x / 2e4) cos(y / 1.5e4) + N(0, 2^2).
"""
import numpy as np


def dem_like_field(n, extent=100_000.0, seed=0):
    """Return ``[n, 3]`` float64 rows ``[x, y, z]``."""
    rng = np.random.default_rng(seed)
    xy = rng.uniform(0.0, extent, size=(n, 2))
    z = 200.0 + 50.0 * np.sin(xy[:, 0] / 2.0e4) * np.cos(xy[:, 1] / 1.5e4) + rng.normal(0.0, 2.0, size=n)
    return np.column_stack([xy, z])


def uniform_locations(m, extent=100_000.0, seed=1):
    """Return ``[m, 2]`` float64 unknown locations."""
    rng = np.random.default_rng(seed)
    return rng.uniform(0.0, extent, size=(m, 2))
