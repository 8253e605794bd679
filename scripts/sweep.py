"""Scaling sweep (BASELINE.json configs[4]): direct sum N = 2^16..2^20 and Barnes-Hut N = 2^20..2^24 (theta = 0.5,
leaf = 1), uniform and Plummer, on however many ranks torchrun started (1 when run directly).  Device-timed with CUDA
events, max over ranks.  `--cpu` additionally times the compiled reference's Barnes-Hut on the host (all cores) at 2^20.

    python scripts/sweep.py [--cpu]                                   # 1 GPU
    torchrun --nproc-per-node 8 scripts/sweep.py                      # 8 GPUs
"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_package  # noqa: E402

gb = load_package()
from gravity_simulator_b200 import ics  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist = None
uid = None
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(gb.Context.new_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    uid = bytes(buf.cpu().numpy().tobytes())


def rmax(v):
    if dist is None:
        return v
    import torch
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


ctx = gb.Context(device=local, rank=rank, world_size=world, nccl_unique_id=uid)
rows = []


def timed(n, kind, method, reps, **kw):
    x, v, m, G = ics.plummer(n, 42) if kind == "plummer" else ics.uniform_cube(n, 42)
    ctx.set_system(x, m, G, v)
    for _ in range(2):
        ctx.mark_positions_sharded(); ctx.acceleration(method, **kw)
    ctx.synchronize()
    if dist is not None:
        dist.barrier()
    best, stages = 1e30, None
    for _ in range(reps):
        ctx.flush_l2(); ctx.mark_positions_sharded()
        ctx.event_record(0); ctx.acceleration(method, **kw); ctx.event_record(1)
        ms = rmax(ctx.event_elapsed_ms(0, 1))
        if ms < best:
            best, stages = ms, [ctx.timing_ms(s) for s in (1, 3, 4, 5, 2)]
    return best, stages


# optional subsets (an 8-GPU box is charged 8x): SWEEP_DS_EXP="16,20" SWEEP_BH_EXP="20,24"
DS_EXP = [int(t) for t in os.environ["SWEEP_DS_EXP"].split(",") if t] if "SWEEP_DS_EXP" in os.environ else list(range(16, 21))
BH_EXP = [int(t) for t in os.environ["SWEEP_BH_EXP"].split(",") if t] if "SWEEP_BH_EXP" in os.environ else list(range(20, 25))
for kind in ("plummer", "uniform"):
    for e in DS_EXP:
        n = 1 << e
        ms, st = timed(n, kind, "pairwise", 3 if e >= 19 else 6, softening_length=0.01)
        rows.append({"path": "direct_sum", "ic": kind, "n": n, "gpus": world, "ms": ms, "G_interactions_per_s": n * (n - 1.0) / ms / 1e6,
                     "allgather_ms": st[0]})
    for e in BH_EXP:
        n = 1 << e
        ms, st = timed(n, kind, "barnes_hut", 2, softening_length=0.01, opening_angle=0.5, max_num_particles_per_leaf=1)
        rows.append({"path": "barnes_hut", "ic": kind, "n": n, "gpus": world, "ms": ms,
                     "stage_ms": dict(zip(("gather", "bbox_morton", "sort", "build", "walk"), st))})
ctx.close()

if "--cpu" in sys.argv and rank == 0:
    from oracle.bind import Reference
    if Reference.available():
        R = Reference()
        for kind in ("uniform", "plummer"):
            n = 1 << 20
            x, v, m, G = ics.plummer(n, 42) if kind == "plummer" else ics.uniform_cube(n, 42)
            t0 = time.perf_counter()
            R.acceleration(x, m, G, "barnes_hut", 0.01, 0.5, 1)
            rows.append({"path": "barnes_hut_cpu_reference", "ic": kind, "n": n, "threads": os.cpu_count(), "ms": (time.perf_counter() - t0) * 1e3})

if rank == 0:
    for r in rows:
        print(json.dumps(r), flush=True)
if dist is not None:
    dist.barrier(); dist.destroy_process_group()
