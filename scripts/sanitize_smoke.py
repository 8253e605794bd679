"""Small end-to-end run of every kernel family, meant to be executed under compute-sanitizer."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_package
gb = load_package()
import numpy as np
from gravity_simulator_b200 import ics
from oracle.bind import jacobi_inputs
for n in (9, 300, 1111, 5000):
    x, v, m, G = ics.clustered(n, n) if n != 9 else ics.solar_system()
    for method, kw in (("pairwise", {}), ("massless", {}), ("barnes_hut", dict(opening_angle=0.5, max_num_particles_per_leaf=1)),
                       ("barnes_hut", dict(opening_angle=0.7, max_num_particles_per_leaf=5))):
        a = gb.acceleration(x, m, G, method, 0.01, **kw)
        assert np.isfinite(a).all()
    gb.construct_octree(x, m, 2)
    mm = m.copy(); mm[n // 3:] = 0.0
    jx, eta = jacobi_inputs(x, mm)
    gb.whfast_acceleration(x, mm, G, jx, eta, "massless", 0.0)
    with gb.Context() as c:
        c.set_system(x, m, G, v)
        c.leapfrog_begin(1e-3, "barnes_hut", 0.01, 0.5, 1)
        c.leapfrog_steps(1e-3, 2)
        c.energy(); c.positions(); c.velocities()
        c.leapfrog_end()
# resident integrators: WHFast (graph pairs, small-n sort, removal replay), Euler / Euler-Cromer / RK4, one-shot energy
for k, grazers, steps in ((600, 0, 7), (1500, 20, 6)):
    x, v, m, G = ics.asteroid_belt(k, 3, grazers=grazers)
    with gb.Context() as c:
        c.set_system(x, m, G, v)
        c.whfast_begin(180.0, "massless", 0.0, True)
        c.whfast_steps(180.0, steps)
        c.whfast_state(snapshot=True); c.whfast_state(snapshot=False)
        c.whfast_end()
x, v, m, G = ics.two_plummer(700, seed=2)
for integ in ("euler", "euler_cromer", "rk4"):
    with gb.Context() as c:
        c.set_system(x, m, G, v)
        c.fixed_begin(integ, "barnes_hut", 0.01, 0.5, 1)
        c.fixed_steps(1e-3, 2)
        c.positions(); c.velocities()
gb.compute_energy(x, v, m, G)
# round 2: the exact per-lane walk and the one-launch-per-level build beside the defaults (cooperative walk, single-launch
# build); a coincident pair through the energy fallback; the mailbox path already ran for n = 9 above
abi, _ = gb.load()
abi.grav_b200_set_bh_exact(1)
gb.acceleration(x, m, G, "barnes_hut", 0.01, 0.5, 1)
abi.grav_b200_set_bh_exact(0)
xc = x.copy(); xc[3] = xc[700]
gb.compute_energy(xc, v, m, G)
x, v, m, G = ics.plummer(40000, 1)
abi.grav_b200_set_direct_sum_mode(0)
gb.acceleration(x, m, G, "pairwise", 0.01)      # ordered interactions: fast kernel + fix-up + special-tile kernel
abi.grav_b200_set_direct_sum_mode(-1)
gb.acceleration(x, m, G, "barnes_hut", 0.01, 0.5, 1)
# second half of round 2: the pair-once direct sum (rotation loop, cp.async slices, private arrays + finishing kernel), equal
# and unequal masses, with and without softening (masked last group), forced for small systems and chosen automatically
abi.grav_b200_set_direct_sum_mode(1)
for n in (513, 1000, 3001):
    xs, vs, ms, Gs = ics.plummer(n, n)
    for mm in (ms, ms * np.random.default_rng(n).uniform(0.5, 1.5, n)):
        for eps in (0.01, 0.0):
            assert np.isfinite(gb.acceleration(xs, mm, Gs, "pairwise", eps)).all()
abi.grav_b200_set_direct_sum_mode(-1)
xs, vs, ms, Gs = ics.plummer(12500, 4)
with gb.Context() as c:
    c.set_system(xs, ms, Gs, vs)
    c.acceleration("pairwise", 0.01); c.acceleration("pairwise", 0.01)
    assert c.direct_sum_path()[0]
    c.energy()                                   # pair-once energy kernel (mirrored blocks)
    c.accelerations()
print("sanitize_smoke done")
