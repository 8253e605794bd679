"""Deterministic synthetic initial conditions for the benchmark / parity workloads.

Uniform cube mirrors examples/benchmark/c/benchmark.c:28-34 (x in [-1,1)^3, m = 1/N, G = 1); the Plummer
sphere is the standard Aarseth-Henon-Wielen recipe (a = 1, M = 1, truncated at r <= 20 a) shifted to zero
centre of mass / momentum like system_set_center_of_mass_zero / _total_momentum_zero (src/system.c:1097-1201).
numpy's PCG64 generator with a fixed seed replaces the reference's time-seeded PCG32 (src/utils.c:61-67).
"""
from __future__ import annotations

import numpy as np


def uniform_cube(n: int, seed: int = 42):
    rng = np.random.default_rng(seed)
    x = rng.random((n, 3)) * 2.0 - 1.0
    v = np.zeros((n, 3))
    m = np.full(n, 1.0 / n)
    return x, v, m, 1.0


def plummer(n: int, seed: int = 42, rmax: float = 20.0):
    rng = np.random.default_rng(seed)
    r = np.empty(0)
    while r.size < n:
        u = rng.random(2 * (n - r.size) + 16)
        u = u[u > 0.0]
        rr = 1.0 / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
        r = np.concatenate([r, rr[rr <= rmax]])
    r = r[:n]

    def iso(k):
        ct = rng.random(k) * 2.0 - 1.0
        ph = rng.random(k) * 2.0 * np.pi
        st = np.sqrt(1.0 - ct * ct)
        return np.stack([st * np.cos(ph), st * np.sin(ph), ct], axis=1)

    x = r[:, None] * iso(n)
    # speed: q = v / v_esc with density g(q) = q^2 (1 - q^2)^3.5, rejection sampled
    q = np.empty(0)
    while q.size < n:
        k = 2 * (n - q.size) + 16
        a = rng.random(k)
        b = rng.random(k) * 0.1
        q = np.concatenate([q, a[b < a * a * (1.0 - a * a) ** 3.5]])
    q = q[:n]
    vesc = np.sqrt(2.0) * (1.0 + r * r) ** (-0.25)
    v = (q * vesc)[:, None] * iso(n)
    m = np.full(n, 1.0 / n)
    x -= (m[:, None] * x).sum(0) / m.sum()
    v -= (m[:, None] * v).sum(0) / m.sum()
    return np.ascontiguousarray(x), np.ascontiguousarray(v), m, 1.0


def two_plummer(n_each: int = 30000, seed: int = 42, sep: float = 10.0, vrel: float = 0.5, impact: float = 2.0):
    """Config 4: galaxy-collision-sized pair of Plummer spheres (examples/galaxy_collision scale)."""
    x1, v1, m1, _ = plummer(n_each, seed)
    x2, v2, m2, _ = plummer(n_each, seed + 1)
    x1 = x1 + np.array([-sep / 2, -impact / 2, 0.0]); v1 = v1 + np.array([vrel / 2, 0.0, 0.0])
    x2 = x2 + np.array([sep / 2, impact / 2, 0.0]); v2 = v2 + np.array([-vrel / 2, 0.0, 0.0])
    x = np.concatenate([x1, x2]); v = np.concatenate([v1, v2]); m = np.concatenate([m1, m2]) * 0.5
    return np.ascontiguousarray(x), np.ascontiguousarray(v), m, 1.0


def clustered(n: int, seed: int = 7):
    """Deep chains and multi-particle level-21 leaves: tight pairs, exact duplicates, one far outlier."""
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n, 3))
    k = n // 8
    if k > 0:
        x[1:1 + k] = x[0] + rng.normal(size=(k, 3)) * 1e-9      # a knot far below the level-21 cell size
        x[1 + k:1 + 2 * k] = x[n - 1]                            # exact duplicates of one point
    x[n // 2] = np.array([40.0, -3.0, 7.0])                     # outlier stretching the bounding box
    m = rng.random(n) + 0.1
    v = np.zeros((n, 3))
    return np.ascontiguousarray(x), v, m, 1.0


def solar_system():
    """Config 1: the reference's built-in `solar_system` (src/system.c:932-1002; JPL DE440 GM + Horizons
    2024-01-01 state; AU, day, solar mass), dumped from the compiled reference into tests/golden/ by
    tests/golden/make_golden.py so it is available where /root/reference is not."""
    from pathlib import Path
    z = np.load(Path(__file__).resolve().parent.parent / "tests" / "golden" / "solar_system.npz")
    return z["x"].copy(), z["v"].copy(), z["m"].copy(), float(z["G"])


def asteroid_belt(k: int = 100000, seed: int = 3, ecc: float = 0.12, inc: float = 0.05, grazers: int = 0):
    """Config 3: the built-in solar system plus k massless asteroids on near-Keplerian orbits with semi-major axes in
    [2, 3.35) AU (the range of examples/kirkwood_gaps/c/kirkwood_gaps.c:49-79), speeds perturbed by up to +-ecc.
    `grazers` of them are turned into Sun-grazing / hyperbolic bodies whose Kepler solve fails with dt = 180 d, which
    exercises WHFast's removal of invalid particles (src/integrator_whfast.c:551)."""
    xs, vs, ms, G = solar_system()
    rng = np.random.default_rng(seed)
    a = rng.uniform(2.0, 3.35, k)
    ph = rng.uniform(0.0, 2.0 * np.pi, k)
    pos = np.stack([a * np.cos(ph), a * np.sin(ph), rng.normal(0.0, inc, k)], axis=1)
    vc = np.sqrt(G * ms[0] / a) * (1.0 + rng.uniform(-ecc, ecc, k))
    vel = np.stack([-vc * np.sin(ph), vc * np.cos(ph), rng.normal(0.0, 1e-4, k)], axis=1)
    if grazers:
        pos[:grazers] *= rng.uniform(0.002, 0.02, (grazers, 1))
        vel[:grazers] *= rng.uniform(0.0, 30.0, (grazers, 1))
    x = np.concatenate([xs, pos + xs[0]])
    v = np.concatenate([vs, vel + vs[0]])
    m = np.concatenate([ms, np.zeros(k)])
    return np.ascontiguousarray(x), np.ascontiguousarray(v), m, G
