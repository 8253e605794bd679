"""Where does a config-3 run through the drop-in spend its wall time?  (context, upload, begin, 400 steps, download)"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_package
gb = load_package()
from gravity_simulator_b200 import ics
from oracle.bind import launch_simulation, DROPIN_SO
x, v, m, G = ics.asteroid_belt(100000, 7)
for rep in range(3):
    t = [time.perf_counter()]
    c = gb.Context(); t.append(time.perf_counter())
    c.set_system(x, m, G, v); c.synchronize(); t.append(time.perf_counter())
    c.whfast_begin(180.0, "massless", 0.0, False); c.synchronize(); t.append(time.perf_counter())
    c.whfast_steps(180.0, 400); c.synchronize(); t.append(time.perf_counter())
    c.whfast_state(); t.append(time.perf_counter())
    c.whfast_end(); c.close(); t.append(time.perf_counter())
    print("ctx %.1f  upload %.1f  begin %.1f  400 steps %.1f  download %.1f  end+destroy %.1f ms" % tuple((b - a) * 1e3 for a, b in zip(t, t[1:])))
kw = dict(tf=180.0 * 400, integrator="whfast", dt=180.0, method="massless")
for rep in range(3):
    t0 = time.perf_counter(); launch_simulation(DROPIN_SO, x, v, m, G, **kw); print("launch_simulation_python 400 steps: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
# the same question for a Barnes-Hut leapfrog run (cooperative sort + cooperative tree build, no graphs): stable per run?
x, v, m, G = ics.two_plummer(30000, seed=5)
kw = dict(tf=100e-3, integrator="leapfrog", dt=1e-3, method="barnes_hut", opening_angle=0.5)
for rep in range(5):
    t0 = time.perf_counter(); launch_simulation(DROPIN_SO, x, v, m, G, **kw); print("config 4, 100 steps: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
