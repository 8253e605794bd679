"""Per-call latency of the drop-in acceleration() for tiny systems (config 1: N = 9)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_package
gb = load_package()
import ctypes as C
import numpy as np
from gravity_simulator_b200 import ics
_, shim = gb.load()
for n, method in ((9, "pairwise"), (9, "barnes_hut"), (1024, "pairwise"), (16384, "pairwise")):
    x, v, m, G = ics.solar_system() if n == 9 else ics.plummer(n, 1)
    a = np.empty_like(x)
    s = gb.make_system(x, m, G); p = gb.make_param(method, 0.0, 0.5, 1)
    ap = a.ctypes.data_as(gb.c_double_p)
    for _ in range(50):
        shim.acceleration(ap, C.byref(s), C.byref(p))
    reps = 2000 if n <= 1024 else 200
    t0 = time.perf_counter()
    for _ in range(reps):
        shim.acceleration(ap, C.byref(s), C.byref(p))
    dt = (time.perf_counter() - t0) / reps
    print(f"N={n:6d} {method:10s} {dt * 1e6:8.1f} us per acceleration() call (host buffers in and out)")
