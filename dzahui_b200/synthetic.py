"""Synthetic workloads named in BASELINE.json, generated identically on every host.

C4: one regular bar, x_i = i / (n-1).
C5: ensemble of jittered bars from a counter-based splitmix64 stream:
      x(beta, i) = (i + 0.2 (U(beta,i) - 0.5)) / (n-1) for interior i, ends exactly 0 and 1,
      U = (splitmix64(0xD2A41 * 2^32 + beta * n + i) >> 11) * 2^-53,
      ic  u_i = 1 + 14 x_i + sin(2 pi x_i (1 + beta mod 7)).
Pure numpy (workload generation, not the hot path)."""
import numpy as np

_SEED = np.uint64(0xD2A41) << np.uint64(32)


def splitmix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def regular_bar(n):
    return np.arange(n, dtype=np.float64) / np.float64(n - 1)


def ensemble_meshes(bar_ids, n):
    """meshes [len(bar_ids)][n] for the given global bar indices"""
    bar_ids = np.asarray(bar_ids, dtype=np.uint64)
    out = np.empty((bar_ids.size, n), dtype=np.float64)
    i = np.arange(n, dtype=np.uint64)
    chunk = max(1, (1 << 22) // n)
    with np.errstate(over="ignore"):
        for s in range(0, bar_ids.size, chunk):
            b = bar_ids[s:s + chunk, None]
            ctr = _SEED + b * np.uint64(n) + i[None, :]
            u = (splitmix64(ctr) >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)
            x = (i[None, :].astype(np.float64) + 0.2 * (u - 0.5)) / np.float64(n - 1)
            x[:, 0] = 0.0
            x[:, -1] = 1.0
            out[s:s + chunk] = x
    return out


def ensemble_initial_conditions(meshes, bar_ids):
    """interior initial conditions [B][n-2]"""
    bar_ids = np.asarray(bar_ids, dtype=np.uint64)
    x = meshes[:, 1:-1]
    k = (1 + (bar_ids % np.uint64(7))).astype(np.float64)[:, None]
    return 1.0 + 14.0 * x + np.sin(2.0 * np.pi * x * k)
