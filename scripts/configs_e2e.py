"""End-to-end wall time of the BASELINE.json configs through the reference's own entry point
(launch_simulation_python): the unmodified reference build vs the drop-in build with the B200 path.
Step counts are cut down so the CPU side stays within seconds; per-step times are what matters."""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_package  # noqa: E402
load_package()
from gravity_simulator_b200 import ics  # noqa: E402
from oracle.bind import launch_simulation, DROPIN_SO, REF_SO  # noqa: E402


def timed(lib, *a, **kw):
    t0 = time.perf_counter()
    launch_simulation(lib, *a, **kw)
    return time.perf_counter() - t0


cases = []
xs, vs, ms, G = ics.solar_system()
cases.append(("config 1: solar system N=9, IAS15 tol 1e-9, pairwise", (xs, vs, ms, G), dict(tf=20 * 365.24, integrator="ias15", tolerance=1e-9, method="pairwise"), None))
x, v, m, G2 = ics.plummer(16384, 2)
cases.append(("config 2: Plummer N=16384, leapfrog dt=1e-3, pairwise eps=0.01", (x, v, m, G2), dict(tf=20e-3, integrator="leapfrog", dt=1e-3, method="pairwise", softening_length=0.01), 20))
xb, vb, mb, Gb = ics.asteroid_belt(100000, 7)
cases.append(("config 3: 9 massive + 1e5 massless, WHFast dt=180 d, massless", (xb, vb, mb, Gb), dict(tf=180.0 * 400, integrator="whfast", dt=180.0, method="massless"), 400))
xc, vc, mc, Gc = ics.two_plummer(30000, seed=5)
cases.append(("config 4: two Plummer spheres N=60000, leapfrog dt=1e-3, Barnes-Hut theta=0.5", (xc, vc, mc, Gc), dict(tf=100e-3, integrator="leapfrog", dt=1e-3, method="barnes_hut", opening_angle=0.5), 100))
for name, sysm, kw, steps in cases:
    # warm-up with the full run: after the previous case's CPU-only reference run the GPU sits at idle clocks, and a
    # few milliseconds of kernels do not bring them back (first version of this script: 5-7x slower drop-in times)
    timed(DROPIN_SO, *sysm, **kw)
    t_gpu = min(timed(DROPIN_SO, *sysm, **kw) for _ in range(2))
    t_ref = timed(REF_SO, *sysm, **kw)
    row = {"case": name, "steps": steps, "reference_s": t_ref, "dropin_s": t_gpu, "speedup": t_ref / t_gpu,
           "host_threads": os.cpu_count()}
    print(json.dumps(row), flush=True)
