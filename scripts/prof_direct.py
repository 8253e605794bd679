"""Profiling driver: a few direct-sum force evaluations at moderate N (run under ncu on the GPU box)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_package
gb = load_package()
from gravity_simulator_b200 import ics
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
x, v, m, G = ics.plummer(n, 42)
with gb.Context() as c:
    c.set_system(x, m, G, v)
    best = 1e30
    for _ in range(reps):
        c.acceleration("pairwise", 0.01)
        best = min(best, c.timing_ms(2))
    rate = n * (n - 1.0) / (best * 1e-3)
    print(f"N={n} kernel {best:.3f} ms  {rate / 1e9:.1f} G inter/s  fp64-pipe-util(16 ops) {rate * 16 / (148 * 64 * 1.965e9) * 100:.1f}%")
