"""Profiling / timing driver for the device-resident WHFast step (config 3).
usage: prof_whfast.py [k_asteroids] [steps] [ref_steps]   -- ref_steps > 0 also times the reference's whfast() on the host."""
import os
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_package
gb = load_package()
from gravity_simulator_b200 import ics
k = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
ref_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dt = 180.0
x, v, m, G = ics.asteroid_belt(k, 7)
with gb.Context() as c:
    c.set_system(x, m, G, v)
    c.whfast_begin(dt, "massless", 0.0, True)
    c.whfast_steps(dt, 8)
    c.synchronize()
    n0 = gb.kernel_launch_count()
    c.event_record(0)
    t0 = time.perf_counter()
    c.whfast_steps(dt, steps)
    c.event_record(1)
    ms = c.event_elapsed_ms(0, 1)
    wall = (time.perf_counter() - t0) * 1e3
    print(f"resident WHFast N={k + 9}: {ms / steps * 1e3:.1f} us/step on the device ({wall / steps * 1e3:.1f} us wall), "
          f"{(gb.kernel_launch_count() - n0) / steps:.1f} launches/step, {steps / ms * 1e3:.0f} steps/s")
if ref_steps > 0:
    from oracle.bind import Reference
    R = Reference()
    for thr in (os.cpu_count(), 1):
        os.environ["OMP_NUM_THREADS"] = str(thr)
        t0 = time.perf_counter()
        R.whfast_run(x, v, m, G, dt, dt * ref_steps, "massless", 0.0, False)
        el = time.perf_counter() - t0
        print(f"reference whfast() on the host, OMP_NUM_THREADS={thr} (set after libgomp start: may not apply): {el / ref_steps * 1e3:.2f} ms/step")
