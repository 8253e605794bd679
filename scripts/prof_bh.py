"""Profiling driver: Barnes-Hut force evaluations (run under ncu on the GPU box)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_package
gb = load_package()
from gravity_simulator_b200 import ics
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kind = sys.argv[3] if len(sys.argv) > 3 else "plummer"
x, v, m, G = ics.plummer(n, 43) if kind == "plummer" else ics.uniform_cube(n, 43)
with gb.Context() as c:
    c.set_system(x, m, G, v)
    for _ in range(reps):
        c.acceleration("barnes_hut", 0.01, 0.5, 1)
        c.synchronize()
    print(f"N={n} {kind} total {c.timing_ms(0):.3f} ms  morton {c.timing_ms(3):.3f} sort {c.timing_ms(4):.3f} build {c.timing_ms(5):.3f} walk {c.timing_ms(2):.3f}")
