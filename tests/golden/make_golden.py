"""Mint golden vectors from the UNMODIFIED reference (oracle/_ref/libgrav_sim_ref.so).

Run here (where /root/reference exists): `python tests/golden/make_golden.py`.  The reference ships no
tests or fixtures of its own, so these files are the known-answer vectors for
this path: inputs are stored next to outputs, nothing depends on RNG reproducibility.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_package  # noqa: E402
from oracle.bind import Reference, jacobi_inputs  # noqa: E402

load_package()
from gravity_simulator_b200 import ics  # noqa: E402

OUT = Path(__file__).resolve().parent
R = Reference()


def tree_fields(t, prefix):
    return {f"{prefix}_{k}": (np.asarray(v)) for k, v in t.items()}


def force_case(name, x, m, G, eps, bh=((0.5, 1), (1.0, 1)), trees=(1,), massless=False):
    d = {"x": x, "m": m, "G": np.float64(G), "eps": np.float64(eps)}
    d["a_pairwise"] = R.acceleration(x, m, G, "pairwise", eps)
    if massless:
        d["a_massless"] = R.acceleration(x, m, G, "massless", eps)
    for theta, leaf in bh:
        d[f"a_bh_t{theta}_l{leaf}"] = R.acceleration(x, m, G, "barnes_hut", eps, theta, leaf)
    for leaf in trees:
        d.update(tree_fields(R.construct_octree(x, m, leaf), f"tree_l{leaf}"))
    np.savez_compressed(OUT / f"{name}.npz", **d)
    print(name, {k: getattr(v, "shape", None) for k, v in d.items() if k.startswith("a_")})


# config 1: built-in solar system (literal known input)
x, v, m, G = R.built_in_system("solar_system")
np.savez_compressed(OUT / "solar_system.npz", x=x, v=v, m=m, G=np.float64(G))
force_case("solar_forces", x, m, G, 0.0, bh=((0.0, 1), (0.5, 1)), trees=(1,), massless=True)

# config 2 (scaled down): Plummer sphere, softened
x, v, m, G = ics.plummer(2048, seed=1)
force_case("plummer2048", x, m, G, 0.01, bh=((0.5, 1), (1.0, 1), (0.5, 8)), trees=(1, 8))

# uniform cube, unsoftened
x, v, m, G = ics.uniform_cube(1500, seed=2)
force_case("uniform1500", x, m, G, 0.0, bh=((0.5, 1), (0.0, 1)), trees=(1,))

# deep chains, duplicate points (level-21 multi-particle leaves), outlier
x, v, m, G = ics.clustered(1024, seed=3)
force_case("clustered1024", x, m, G, 0.05, bh=((0.5, 1), (0.7, 4)), trees=(1, 4))

# massless method: massive particles NOT in the array prefix -> exercises the m[rank] indexing
rng = np.random.default_rng(4)
x = rng.normal(size=(600, 3))
m = np.zeros(600)
m[[2, 5, 9, 300]] = [1.0, 0.3, 2.0, 0.7]
m[0] = 0.25   # rank 0 -> the quirk reads m[0..3]: a mix of zero and non-zero entries
force_case("massless600", x, m, 1.0, 0.0, bh=(), trees=(), massless=True)

# tiny systems
for n in (1, 2, 3):
    x, v, m, G = ics.uniform_cube(n, seed=10 + n)
    force_case(f"tiny{n}", x, m, G, 0.0, bh=((0.5, 1),), trees=(1,), massless=True)

# WHFast kernels: solar system + massless belt (kirkwood-gap-like), inputs as the WHFast caller builds them
xs, vs, ms, G = R.built_in_system("solar_system")
rng = np.random.default_rng(5)
k = 400
r = rng.uniform(2.0, 3.35, k); ph = rng.uniform(0, 2 * np.pi, k); z = rng.normal(0, 0.1, k)
belt = np.stack([r * np.cos(ph), r * np.sin(ph), z], axis=1) + xs[0]
order = np.argsort(np.concatenate([np.linalg.norm(xs[1:] - xs[0], axis=1), np.linalg.norm(belt - xs[0], axis=1)]))
x = np.concatenate([xs[:1], np.concatenate([xs[1:], belt])[order]])
m = np.concatenate([ms[:1], np.concatenate([ms[1:], np.zeros(k)])[order]])
jx, eta = jacobi_inputs(x, m)
d = {"x": x, "m": m, "G": np.float64(G), "jacobi_x": jx, "eta": eta}
for eps in (0.0, 0.01):
    d[f"a_whfast_massless_eps{eps}"] = R.whfast_acceleration(x, m, G, jx, eta, "massless", eps)
np.savez_compressed(OUT / "whfast_belt.npz", **d)
jx9, eta9 = jacobi_inputs(xs, ms)
d = {"x": xs, "m": ms, "G": np.float64(G), "jacobi_x": jx9, "eta": eta9}
for eps in (0.0, 0.01):
    d[f"a_whfast_pairwise_eps{eps}"] = R.whfast_acceleration(xs, ms, G, jx9, eta9, "pairwise", eps)
    d[f"a_whfast_massless_eps{eps}"] = R.whfast_acceleration(xs, ms, G, jx9, eta9, "massless", eps)
np.savez_compressed(OUT / "whfast_solar.npz", **d)
print("whfast ok")


# whole WHFast runs (reference whfast() through oracle/ref_whfast_probe.c, serial removal order)
import os  # noqa: E402
os.environ["OMP_NUM_THREADS"] = "1"
for name, (k, seed, grazers, method, steps, dt, eps) in {
        "whfast_run_solar": (0, 1, 0, "pairwise", 30, 5.0, 0.0),
        "whfast_run_belt": (400, 2, 0, "massless", 15, 180.0, 0.0),
        "whfast_run_removal": (600, 5, 15, "massless", 6, 180.0, 0.0)}.items():
    x, v, m, G = ics.asteroid_belt(k, seed, grazers=grazers)
    r = R.whfast_run(x, v, m, G, dt, dt * steps, method, eps, True)
    np.savez_compressed(OUT / f"{name}.npz", x=x, v=v, m=m, G=np.float64(G), dt=np.float64(dt), steps=np.int64(steps),
                        method=method, eps=np.float64(eps), out_x=r["x"], out_v=r["v"], out_m=r["m"], out_ids=r["ids"])
    print(name, "n", m.shape[0], "->", r["m"].shape[0])
