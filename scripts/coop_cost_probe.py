"""Repeated short runs through launch_simulation_python (a new context per run): wall time per run."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_package
gb = load_package()
from gravity_simulator_b200 import ics
from oracle.bind import launch_simulation, DROPIN_SO
xb, vb, mb, Gb = ics.asteroid_belt(100000, 7)
x, v, m, G = ics.two_plummer(30000, seed=5)
launch_simulation(DROPIN_SO, x, v, m, G, tf=5e-3, integrator="leapfrog", dt=1e-3, method="barnes_hut", opening_angle=0.5)
w, b = [], []
for rep in range(6):
    t0 = time.perf_counter(); launch_simulation(DROPIN_SO, xb, vb, mb, Gb, tf=180.0 * 400, integrator="whfast", dt=180.0, method="massless"); w.append((time.perf_counter() - t0) * 1e3)
    t0 = time.perf_counter(); launch_simulation(DROPIN_SO, x, v, m, G, tf=100e-3, integrator="leapfrog", dt=1e-3, method="barnes_hut", opening_angle=0.5); b.append((time.perf_counter() - t0) * 1e3)
print("whfast 400 steps ms:", " ".join(f"{t:.0f}" for t in w), "| BH leapfrog 100 steps ms:", " ".join(f"{t:.0f}" for t in b))
