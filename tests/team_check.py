"""In-process device team check (needs >= 2 GPUs):  python tests/team_check.py [k]

(1) grav_b200_ctx_create_team: the sharded direct sum / massless / Barnes-Hut / leapfrog / RK4 / energy driven from ONE
    thread against the single-GPU results (direct sums <= 3e-13, Barnes-Hut bit-equal: every target is walked by one warp
    in the same way whichever rank owns it).
(2) GRAV_B200_DEVICES=k behind the reference's own entry points: acceleration() and launch_simulation_python of the
    drop-in build (config 2- and config 4-shaped problems at N = 2^18) in fresh processes, one with k devices and one with
    a single device, final states compared the same way.
Prints TEAM_CHECK_PASSED on success."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_package  # noqa: E402

CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from conftest import load_package
gb = load_package()
from gravity_simulator_b200 import ics
from oracle.bind import DROPIN_SO, launch_simulation
out = sys.argv[1]
n = 1 << 18
x, v, m, G = ics.plummer(n, 5)
res = {}
res["a_pairwise"] = gb.acceleration(x, m, G, "pairwise", 0.01)
res["a_bh"] = gb.acceleration(x, m, G, "barnes_hut", 0.01, 0.5, 1)
xs, vs, ms, Gs = ics.asteroid_belt(50000, 3)
res["a_massless"] = gb.acceleration(xs, ms, Gs, "massless", 0.0)
if DROPIN_SO.exists():
    xd, vd = launch_simulation(DROPIN_SO, x, v, m, G, tf=3e-3, integrator="leapfrog", dt=1e-3, method="pairwise", softening_length=0.01)
    res["lf_pairwise_x"], res["lf_pairwise_v"] = xd, vd
    x2, v2, m2, G2 = ics.two_plummer(n // 2, seed=9)
    xd, vd = launch_simulation(DROPIN_SO, x2, v2, m2, G2, tf=3e-3, integrator="leapfrog", dt=1e-3, method="barnes_hut", softening_length=0.0, opening_angle=0.5)
    res["lf_bh_x"], res["lf_bh_v"] = xd, vd
    xd, vd = launch_simulation(DROPIN_SO, x2[:60000], v2[:60000], m2[:60000], G2, tf=2e-3, integrator="rk4", dt=1e-3, method="barnes_hut", softening_length=0.01, opening_angle=0.5)
    res["rk4_bh_x"], res["rk4_bh_v"] = xd, vd
np.savez(out, **res)
''' % (str(ROOT), str(ROOT / "tests"))


def rel(p, q):
    d = np.linalg.norm(q, axis=1)
    return float(np.max(np.linalg.norm(p - q, axis=1) / np.where(d > 0, d, 1.0)))


def main():
    gb = load_package()
    from gravity_simulator_b200 import ics
    k = int(sys.argv[1]) if len(sys.argv) > 1 else min(gb.device_count(), 4)
    if gb.device_count() < 2 or k < 2:
        print("team_check: needs >= 2 GPUs")
        return 2
    ok = True

    def check(name, cond):
        nonlocal ok
        ok = ok and bool(cond)
        print(f"[{'ok' if cond else 'FAIL'}] {name}", flush=True)

    single = gb.Context(device=0)
    team = gb.Context(team=k)
    check(f"team of {k} devices created, leader owns rank 0", team.team_size() == k)
    for n in (10007, 65536, 1 << 18):
        x, v, m, G = ics.plummer(n, n)
        single.set_system(x, m, G, v); team.set_system(x, m, G, v)
        lo, hi = team.owned_range()
        check(f"n={n}: leader's targets [{lo},{hi}) = [0, n/{k})", lo == 0 and hi == n // k)
        for method, kw in (("pairwise", dict(softening_length=0.01)), ("barnes_hut", dict(softening_length=0.01, opening_angle=0.5)),
                           ("massless", dict(softening_length=0.0))):
            single.acceleration(method, **kw); team.mark_positions_sharded(); team.acceleration(method, **kw)
            a1, aN = single.accelerations(), team.accelerations()
            if method == "barnes_hut":
                check(f"n={n} {method}: team == single GPU (bit-exact)", np.array_equal(a1, aN, equal_nan=True))
            else:
                e = rel(aN, a1)
                check(f"n={n} {method}: team vs single GPU max rel {e:.1e} <= 3e-13", e <= 3e-13)
        if n > 100000:
            continue
        dt = 1e-3
        for c in (single, team):
            c.set_system(x, m, G, v)
            c.leapfrog_begin(dt, "pairwise", 0.01)
            c.leapfrog_steps(dt, 5)
        ex, ev = rel(team.positions(), single.positions()), rel(team.velocities(), single.velocities())
        check(f"n={n} leapfrog 5 steps: team vs single max rel x {ex:.1e} v {ev:.1e} <= 1e-12", ex <= 1e-12 and ev <= 1e-12)
        e1, eN = single.energy(), team.energy()
        check(f"n={n} energy while the leapfrog runs: rel diff {abs(e1 - eN) / abs(e1):.1e}", abs(e1 - eN) <= 1e-12 * abs(e1))
        for c in (single, team):
            c.leapfrog_end()
        for c in (single, team):
            c.set_system(x, m, G, v)
            c.fixed_begin("rk4", "barnes_hut", 0.01, 0.5, 1)
            c.fixed_steps(dt, 2)
        check(f"n={n} rk4 + barnes_hut 2 steps: team == single (bit-exact)",
              np.array_equal(single.positions(), team.positions()) and np.array_equal(single.velocities(), team.velocities()))
    single.close(); team.close()

    # behind the reference's entry points, in fresh processes (the default context reads GRAV_B200_DEVICES once)
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        outs = {}
        for tag, env in (("single", {"GRAV_B200_DEVICES": "1"}), ("team", {"GRAV_B200_DEVICES": str(k)})):
            out = os.path.join(tmp, tag + ".npz")
            r = subprocess.run([sys.executable, "-c", CHILD, out], env=dict(os.environ, **env), capture_output=True, text=True, timeout=900)
            if r.returncode != 0:
                print(r.stdout[-2000:], r.stderr[-3000:])
            check(f"drop-in process with GRAV_B200_DEVICES={env['GRAV_B200_DEVICES']} finished", r.returncode == 0)
            outs[tag] = dict(np.load(out)) if r.returncode == 0 else {}
        s, t = outs["single"], outs["team"]
        if s and t:
            check(f"acceleration() pairwise N=2^18 on {k} GPUs vs 1: {rel(t['a_pairwise'], s['a_pairwise']):.1e} <= 3e-13", rel(t["a_pairwise"], s["a_pairwise"]) <= 3e-13)
            check(f"acceleration() massless N=50009 on {k} GPUs vs 1: {rel(t['a_massless'], s['a_massless']):.1e} <= 1e-13", rel(t["a_massless"], s["a_massless"]) <= 1e-13)
            check(f"acceleration() barnes_hut N=2^18 on {k} GPUs == 1 GPU (bit-exact)", np.array_equal(t["a_bh"], s["a_bh"]))
            if "lf_bh_x" in s:
                e = max(rel(t["lf_pairwise_x"], s["lf_pairwise_x"]), rel(t["lf_pairwise_v"], s["lf_pairwise_v"]))
                check(f"launch_simulation_python leapfrog + pairwise N=2^18, 3 steps, {k} GPUs vs 1: {e:.1e} <= 1e-12", e <= 1e-12)
                check(f"launch_simulation_python leapfrog + barnes_hut N=2^18, 3 steps, {k} GPUs == 1 GPU (bit-exact)",
                      np.array_equal(t["lf_bh_x"], s["lf_bh_x"]) and np.array_equal(t["lf_bh_v"], s["lf_bh_v"]))
                check(f"launch_simulation_python rk4 + barnes_hut N=60000, 2 steps, {k} GPUs == 1 GPU (bit-exact)",
                      np.array_equal(t["rk4_bh_x"], s["rk4_bh_x"]) and np.array_equal(t["rk4_bh_v"], s["rk4_bh_v"]))
    if ok:
        print("TEAM_CHECK_PASSED", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
