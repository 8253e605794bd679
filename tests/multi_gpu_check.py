"""Multi-GPU parity check, one process per GPU:  torchrun --nproc-per-node N tests/multi_gpu_check.py
Sharded direct sum / Barnes-Hut / leapfrog / Euler-Cromer / RK4 against the single-GPU result computed on every rank."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_package  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gb = load_package()
    from gravity_simulator_b200 import ics
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(gb.Context.new_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    uid = bytes(buf.cpu().numpy().tobytes())

    single = gb.Context(device=local)
    multi = gb.Context(device=local, rank=rank, world_size=world, nccl_unique_id=uid)
    ok = True

    def check(name, cond):
        nonlocal ok
        ok = ok and bool(cond)
        if rank == 0:
            print(f"[{'ok' if cond else 'FAIL'}] {name}", flush=True)

    for n in (10007, 65536):
        x, v, m, G = ics.plummer(n, n)
        single.set_system(x, m, G, v); multi.set_system(x, m, G, v)
        lo, hi = multi.owned_range()
        check(f"n={n}: owned range [{lo},{hi}) matches rank*n/world", lo == rank * n // world and hi == (rank + 1) * n // world)
        abi, _ = gb.load()
        # pairwise in both formulations (include/grav_b200.h, grav_b200_set_direct_sum_mode): ordered interactions with the
        # targets sharded, and every pair once with the pair units sharded and an all-reduce of the accelerations
        for method, kw, ds in (("pairwise", dict(softening_length=0.01), -1), ("pairwise", dict(softening_length=0.01), 0),
                               ("pairwise", dict(softening_length=0.01), 1), ("barnes_hut", dict(softening_length=0.01, opening_angle=0.5), -1),
                               ("massless", dict(softening_length=0.0), -1)):
            abi.grav_b200_set_direct_sum_mode(ds)
            single.acceleration(method, **kw); multi.mark_positions_sharded(); multi.acceleration(method, **kw)
            a1, aN = single.accelerations(), multi.accelerations()
            abi.grav_b200_set_direct_sum_mode(-1)
            if method == "barnes_hut":   # per-target serial arithmetic: independent of the partition
                check(f"n={n} {method}: sharded over {world} ranks == single GPU (bit-exact)", np.array_equal(a1, aN, equal_nan=True))
            else:                        # the split points of the work move with the shard, so only rounding differs
                err = float(np.max(np.linalg.norm(a1 - aN, axis=1) / np.linalg.norm(a1, axis=1)))
                path = {-1: "auto", 0: "forced ordered", 1: "forced pair-once"}[ds]
                if method == "pairwise":
                    po, eq = multi.direct_sum_path()
                    path += ": " + (("pair-once, equal masses" if eq else "pair-once") if po else "ordered interactions")
                check(f"n={n} {method} [{path}]: sharded over {world} ranks vs single GPU max rel {err:.1e} <= 3e-13", err <= 3e-13)
        dt = 1e-3
        for c in (single, multi):
            c.set_system(x, m, G, v)
            c.leapfrog_begin(dt, "pairwise", 0.01)
            c.leapfrog_steps(dt, 5)
        rel = lambda p, q: float(np.max(np.linalg.norm(p - q, axis=1) / np.linalg.norm(p, axis=1)))
        ex, ev = rel(single.positions(), multi.positions()), rel(single.velocities(), multi.velocities())
        check(f"n={n} leapfrog 5 steps: sharded vs single max rel x {ex:.1e} v {ev:.1e} <= 1e-12", ex <= 1e-12 and ev <= 1e-12)
        for c in (single, multi):
            c.leapfrog_end()
        # Euler-Cromer and RK4 on the sharded state with Barnes-Hut forces: bit-identical to the single-GPU run
        for integ in ("euler_cromer", "rk4"):
            for c in (single, multi):
                c.set_system(x, m, G, v)
                c.fixed_begin(integ, "barnes_hut", 0.01, 0.5, 1)
                c.fixed_steps(dt, 3)
            check(f"n={n} {integ} + barnes_hut 3 steps: sharded == single (bit-exact)",
                  np.array_equal(single.positions(), multi.positions()) and np.array_equal(single.velocities(), multi.velocities()))
        e1, eN = single.energy(), multi.energy()
        check(f"n={n} energy: sharded vs single rel diff {abs(e1 - eN) / abs(e1):.1e}", abs(e1 - eN) <= 1e-13 * abs(e1))
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    single.close(); multi.close()
    dist.barrier(); dist.destroy_process_group()
    if t.item() < 1.0:
        sys.exit(1)
    if rank == 0:
        print("MULTI_GPU_CHECK_PASSED", flush=True)


if __name__ == "__main__":
    main()
