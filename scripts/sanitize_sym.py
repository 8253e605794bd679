"""The kernels of the second half of round 2 (pair-once direct sum and energy), meant to be executed under compute-sanitizer
(scripts/sanitize_smoke.py runs them too, after everything else)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_package
gb = load_package()
import numpy as np
from gravity_simulator_b200 import ics
abi, _ = gb.load()
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
    c.energy()
    c.accelerations()
xs, vs, ms, Gs = ics.plummer(1500, 5)
gb.compute_energy(xs, vs, ms, Gs)
print("sanitize_sym done")
