"""A/B of the Barnes-Hut walk kernels on one GPU: cooperative (default), per-lane + fast evaluation, per-lane exact.
Each variant runs in its own process (the kernel choice is read once per process).
    python scripts/walk_ab.py [n] [plummer|uniform]
"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CHILD = r'''
import sys, json
sys.path.insert(0, %r); sys.path.insert(0, %r)
from conftest import load_package
gb = load_package()
from gravity_simulator_b200 import ics
import numpy as np
n, kind = int(sys.argv[1]), sys.argv[2]
x, v, m, G = ics.plummer(n, 43) if kind == "plummer" else ics.uniform_cube(n, 43)
with gb.Context() as c:
    c.set_system(x, m, G, v)
    for _ in range(3):
        c.acceleration("barnes_hut", 0.01, 0.5, 1)
    c.synchronize()
    tot, st = [], []
    for _ in range(5):
        c.flush_l2()
        c.acceleration("barnes_hut", 0.01, 0.5, 1)
        c.synchronize()
        tot.append(c.timing_ms(0)); st.append([c.timing_ms(s) for s in (3, 4, 5, 2)])
    a = c.accelerations()
st = np.mean(np.array(st), axis=0)
print(json.dumps({"n": n, "ic": kind, "total_ms": float(np.mean(tot)), "morton": st[0], "sort": st[1], "build": st[2], "walk": st[3],
                  "checksum": float(np.abs(a).sum())}))
''' % (str(ROOT), str(ROOT / "tests"))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    kind = sys.argv[2] if len(sys.argv) > 2 else "plummer"
    variants = [("cooperative", {}), ("lane+fast", {"GRAV_B200_WALK_KERNEL": "lane"}), ("lane exact", {"GRAV_B200_BH_EXACT": "1"})]
    for lib in sorted((ROOT / "scratch").glob("libgrav_b200_*.so")):       # tuning builds (scripts/build_variant.sh)
        variants.append((lib.stem.replace("libgrav_b200_", "coop:"), {"GRAV_B200_LIB": str(lib)}))
    if len(sys.argv) > 3:
        variants = [v for v in variants if any(v[0].startswith(s) for s in sys.argv[3].split(","))]
    for name, env in variants:
        r = subprocess.run([sys.executable, "-c", CHILD, str(n), kind], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:]
        print(f"{name:12s} {line}", flush=True)


if __name__ == "__main__":
    main()
