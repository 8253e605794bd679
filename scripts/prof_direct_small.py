"""Where does a small direct-sum force call spend its time?  event-timed call vs the kernel stage, N = 2^13 .. 2^17."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_package
gb = load_package()
from gravity_simulator_b200 import ics
import numpy as np
sizes = [int(a) for a in sys.argv[1:]] or [8192, 16384, 65536, 131072]
with gb.Context() as c:
    for n in sizes:
        x, v, m, G = ics.plummer(n, 42)
        c.set_system(x, m, G, v)
        for _ in range(5):
            c.acceleration("pairwise", 0.01)
        c.synchronize()
        tot, ker = [], []
        for _ in range(20):
            c.event_record(0); c.acceleration("pairwise", 0.01); c.event_record(1)
            tot.append(c.event_elapsed_ms(0, 1)); ker.append(c.timing_ms(2))
        ideal = n * (n - 1.0) / 1057e9 * 1e3
        print(f"N={n}: call {np.median(tot):.4f} ms  kernels {np.median(ker):.4f} ms  ideal@1057G/s {ideal:.4f} ms  -> {n*(n-1.0)/np.median(tot)/1e6:.0f} G/s")
