"""Timing of the pair-once direct sum for one library build (GRAV_B200_LIB): python scripts/sym_time.py N [reps]"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_package
gb = load_package()
from gravity_simulator_b200 import ics
abi, _ = gb.load()
n = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.default_rng(5)
abi.grav_b200_set_direct_sum_mode(1)
for masses in ("equal", "random"):
    x, v, m, G = ics.plummer(n, 42)
    if masses == "random":
        m = rng.random(n) / n
    with gb.Context() as c:
        c.set_system(x, m, G, v)
        best = 1e30
        for _ in range(reps):
            c.acceleration("pairwise", 0.01); c.synchronize()
            best = min(best, c.timing_ms(2))
    rate = n * (n - 1.0) / (best * 1e-3)
    print(f"N={n} masses={masses}: {best:.3f} ms  {rate / 1e9:.1f} G ordered interactions/s", flush=True)
