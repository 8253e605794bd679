#!/usr/bin/env python
"""Benchmark of the B200 acceleration path (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path, host cores

Headline workload (BASELINE.json configs[4], the configuration the metric is quoted on): FP64 direct-sum
force evaluation of a synthetic Plummer sphere, N = 2^20, softening 0.01; one "step" = one full force
evaluation of all N particles.  N>1 GPUs: targets are sharded over ranks (strong scaling: total work fixed)
and each step starts with the NCCL all-gather of the owned position shards.

value  = ordered pair interactions per second, N(N-1)/t, particle state already resident in HBM
e2e    = same metric through the reference-facing call with HOST buffers (H2D of x, m and D2H of a
         inside the timed region): at N=1 literally the drop-in `acceleration()` symbol
Also reported on the same line: roofline (FP64 pipe), cpu_baseline (compiled reference on the host, bounded
sample), clocks, gpu_launches, a sampled parity check against the reference arithmetic ("parity"), and -- inside
`config`, which the driver's parser keeps whole -- the second half of BASELINE.json's metric: Barnes-Hut full
force evaluation, s per step, at N = 2^24 (the north-star size) and 2^20, Plummer and uniform, theta = 0.5, with its
own cpu_baseline (the reference's OpenMP walk on the host cores), end-to-end time through acceleration_barnes_hut()
with host buffers, per-stage times and HBM-roofline fractions.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "fp64_direct_sum_pair_interactions_per_s"
UNIT = "G interactions/s"
FLOP_PER_INTERACTION = 20      # GPU-Gems-3 convention: 18 + rsqrt counted as 2
# FP64-pipe instructions the kernels actually issue per ORDERED interaction: 16 in the ordered-interaction kernel
# (direct_sum.cu); the pair-once kernel (direct_sum_sym.cu) spends 20 per unordered pair = 10 (18 = 9 with equal masses)
FP64_OPS = {"ordered": 16.0, "pair_once": 10.0, "pair_once_equal_mass": 9.0}
NOMINAL_FP64_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12   # 37.2, used only if the live measurement fails
NCU_DRAM_BYTES_PER_LAUNCH = {"ordered": 52571392 + 51328512,   # dram read + write, profiles/r1_direct_sum_n1m_final.txt
                             "pair_once": None}                 # filled from profiles/r2_sym_n1m.txt when that capture exists


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--n", type=int, default=1 << 20, help="particles of the direct-sum workload")
    p.add_argument("--eps", type=float, default=0.01)
    p.add_argument("--ic", default="plummer", choices=["plummer", "uniform"])
    p.add_argument("--bh-n", type=int, default=1 << 24, help="largest Barnes-Hut size (north star: 2^24; 0 = skip); 2^20 is always the second point")
    p.add_argument("--bh-cpu-n", type=int, default=1 << 20, help="particles of the Barnes-Hut CPU baseline (reference, OpenMP; 0 = skip)")
    p.add_argument("--no-parity", action="store_true")
    p.add_argument("--whfast-n", type=int, default=100000,
                   help="massless asteroids of the WHFast side measurement (config 3; 0 = skip; single GPU only)")
    p.add_argument("--cpu-n", type=int, default=1 << 16, help="particles of the bounded CPU-baseline sample (0 = skip)")
    p.add_argument("--no-e2e", action="store_true")
    return p.parse_args()


def make_ic(kind, n, seed=42):
    from conftest import load_package
    load_package()
    from gravity_simulator_b200 import ics
    return ics.plummer(n, seed) if kind == "plummer" else ics.uniform_cube(n, seed)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.device)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for k, nm in enumerate(names):
                    if f[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm), power_note="see reasons")
        out["reasons"] = sorted(reasons)
        out.pop("power_note", None)
        return out


# ---------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (oracle/_ref) on the host cores
# ---------------------------------------------------------------------------------------------------------

def cpu_reference_time(n, eps, kind, reps=1):
    """Seconds per pairwise force evaluation of the unmodified reference at size n (1 thread: the reference's
    direct-sum code has no OpenMP).  Falls back to the oracle port if _ref is absent."""
    from oracle.bind import Oracle, Reference
    x, v, m, G = make_ic(kind, n)
    if Reference.available():
        impl, kname = Reference(), "reference"
    else:
        impl, kname = Oracle(), "port"
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        impl.acceleration(x, m, G, "pairwise", eps)
        best = min(best, time.perf_counter() - t0)
    return best, kname


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.cpu_n if args.cpu_n > 0 else 1 << 15
    # bound the whole run to a few minutes: one evaluation at 2^16 takes ~13 s on one core
    while n > 4096 and (args.steps + args.warmup) * 13.0 * (n / 65536.0) ** 2 > 150.0:
        n //= 2
    from oracle.bind import Oracle, Reference
    x, v, m, G = make_ic(args.ic, n)
    impl, kname = (Reference(), "reference") if Reference.available() else (Oracle(), "port")
    for _ in range(args.warmup):
        impl.acceleration(x, m, G, "pairwise", args.eps)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        impl.acceleration(x, m, G, "pairwise", args.eps)
    dt = (time.perf_counter() - t0) / args.steps
    val = n * (n - 1) / dt / 1e9
    sample = (f"pairwise force evaluation at N={n} ({args.ic}, eps={args.eps}), the largest size whose {args.steps}+{args.warmup} "
              f"evaluations fit the few-minute budget on one core; interactions/s is size-independent for the O(N^2) loop")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the workload named here is the one actually timed; the b200 arm's N is reported separately
        "config": {"workload": f"direct-sum pairwise FP64 force evaluation, {args.ic} N={n}, eps={args.eps}",
                   "b200_arm_n": args.n,
                   "extrapolated_ms_per_step_at_b200_arm_n": dt * 1e3 * (args.n / n) ** 2,
                   "threads": "1 (src/acceleration.c has no OpenMP in the pairwise loop)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": kname, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------


# ---------------------------------------------------------------------------------------------------------
# checkers and the Barnes-Hut half of the metric
# ---------------------------------------------------------------------------------------------------------

def direct_sum_parity(x, m, G, eps, a, k=256):
    """max relative error of k sampled rows of `a` against the reference's operation order for those targets
    (oracle.pairwise_targets, pinned bit-equal to the compiled reference in tests/test_oracle.py)."""
    from oracle.bind import Oracle
    n = m.shape[0]
    r = np.linalg.norm(x, axis=1)
    tg = np.unique(np.concatenate([np.argsort(r)[:32], np.random.default_rng(5).choice(n, min(k, n), replace=False)])).astype(np.int32)
    ref = Oracle().pairwise_targets(x, m, G, eps, tg)
    err = float(np.max(np.linalg.norm(a[tg] - ref, axis=1) / np.linalg.norm(ref, axis=1)))
    return {"what": f"direct sum N={n}: {tg.shape[0]} sampled targets x all sources vs the reference's summation order (CPU oracle)",
            "max_rel": err, "tol": 1e-12, "ok": bool(err <= 1e-12)}


BH_STAGE_BYTES = {"bbox_morton": 60.0, "sort": 204.0, "build": 79.0, "walk": 127.0}   # SURVEY.md section 8d, per particle


def _ncu_bh_traffic():
    """per-launch DRAM bytes of the BH stages from the committed ncu captures (profiles/r2_bh_dram_bytes.json), if present"""
    try:
        return json.load(open(ROOT / "profiles" / "r2_bh_dram_bytes.json"))
    except Exception:
        return None


def run_barnes_hut(args, gb, ctx, rank, world, barrier, max_over_ranks, idle_until_rank0):
    from oracle.bind import Oracle, Reference
    theta, leaf, eps = 0.5, 1, args.eps
    sizes = sorted({args.bh_n, min(args.bh_n, 1 << 20)}, reverse=True)
    runs = []
    parity = None
    traffic = _ncu_bh_traffic()
    for n in sizes:
        for ic in ("plummer", "uniform"):
            x, v, m, G = make_ic(ic, n, seed=43)
            ctx.set_system(x, m, G, v)
            for _ in range(2):
                ctx.mark_positions_sharded()
                ctx.acceleration("barnes_hut", eps, theta, leaf)
            ctx.synchronize()
            barrier()
            tt, stages = [], []
            for _ in range(3):
                ctx.flush_l2()
                ctx.mark_positions_sharded()
                ctx.event_record(2)
                ctx.acceleration("barnes_hut", eps, theta, leaf)
                ctx.event_record(3)
                tt.append(ctx.event_elapsed_ms(2, 3))
                stages.append([ctx.timing_ms(s) for s in (1, 3, 4, 5, 2)])
            ms = max_over_ranks(float(np.mean(tt)))
            st = np.mean(np.array(stages), axis=0)
            run = {"n": n, "ic": ic, "value": ms * 1e-3, "unit": "s",
                   "stage_ms": {"gather": st[0], "bbox_morton": st[1], "sort": st[2], "build": st[3], "walk": st[4]},
                   "hbm_frac_algorithmic": (470.0 * n / (ms * 1e-3) / 1e9) / _hbm_peak(),
                   "stage_hbm_frac": {k: (b * n / (max(t_ms, 1e-6) * 1e-3) / 1e9) / _hbm_peak()
                                      for (k, b), t_ms in zip(BH_STAGE_BYTES.items(), st[1:])}}
            if traffic and traffic.get("n") == n and traffic.get("ic") == ic and world == 1:
                run["ncu_dram_bytes_per_stage"] = traffic["stages"]
            # sampled parity + items/s at the size the CPU port builds in under a second
            a_all = ctx.accelerations()
            if rank == 0 and n <= (1 << 20) and not args.no_parity:
                with Oracle().tree(x, m, leaf) as T:
                    starts = np.random.default_rng(0).choice(n // 32, min(64, n // 32), replace=False) * 32
                    pos = (starts[:, None] + np.arange(32)[None, :]).ravel()
                    ref, stats = T.walk_targets(G, eps, theta, pos, stats=True)
                    ids = T.to_dict()["sorted_indices"][pos]
                err = float(np.max(np.linalg.norm(a_all[ids] - ref, axis=1) / np.linalg.norm(ref, axis=1)))
                items = float(stats[:, 0].mean() + stats[:, 3].mean())
                run["items_per_target_sampled"] = items
                run["items_per_s"] = items * n / (st[4] * 1e-3)     # node visits + leaf particles of all targets per second of walk
                run["parity"] = {"what": f"{pos.shape[0]} sampled targets vs the CPU port of the reference walk (bit-pinned to the compiled reference)",
                                 "max_rel": err, "tol": 1e-12, "ok": bool(err <= 1e-12)}
                parity = run["parity"] if parity is None or not run["parity"]["ok"] else parity
            # end to end through the reference-facing call with HOST buffers
            if not args.no_e2e:
                def e2e_run():     # the single-threaded drop-in call; with world > 1 it drives all GPUs (GRAV_B200_DEVICES, set above)
                    a_host = np.empty((n, 3))
                    pinned = []
                    for arr in (x, m, a_host):
                        try:
                            gb.host_register(arr); pinned.append(arr)
                        except Exception:
                            pass
                    f = lambda: gb.acceleration(x, m, G, "barnes_hut", eps, theta, leaf, out=a_host)
                    f()
                    t0 = time.perf_counter()
                    f()
                    dt = time.perf_counter() - t0
                    for arr in pinned:
                        gb.host_unregister(arr)
                    return dt
                ctx.synchronize()
                barrier()
                e2e_s = idle_until_rank0(e2e_run) or 0.0
                barrier()
                e2e_s = max_over_ranks(e2e_s)
                run["e2e"] = {"value": e2e_s, "unit": "s", "h2d_bytes_per_step": int(x.nbytes + m.nbytes) * world, "d2h_bytes_per_step": int(3 * 8 * n),
                              "api": "acceleration() [method = barnes_hut] in libgrav_sim_b200.so" + (f", GRAV_B200_DEVICES={world}" if world > 1 else ""),
                              "host_memory": "pinned (cudaHostRegister on the caller's arrays)"}
            runs.append(run)
    # CPU baseline: the compiled reference, OpenMP walk on all host cores (its build phase is serial)
    cpu = None
    if rank == 0 and args.bh_cpu_n > 0:
        impl, kname = (Reference(), "reference") if Reference.available() else (Oracle(), "port")
        try:   # torchrun exports OMP_NUM_THREADS=1; the baseline is the reference's OpenMP walk on ALL host cores
            import ctypes
            ctypes.CDLL("libgomp.so.1").omp_set_num_threads(os.cpu_count())
        except OSError:
            pass
        cpu = {"unit": "s", "cores": os.cpu_count() if kname == "reference" else 1, "kind": kname, "n": args.bh_cpu_n, "runs": []}
        for ic in ("plummer", "uniform"):
            x, v, m, G = make_ic(ic, args.bh_cpu_n, seed=43)
            t0 = time.perf_counter()
            impl.acceleration(x, m, G, "barnes_hut", eps, theta, leaf)
            cpu["runs"].append({"ic": ic, "value": time.perf_counter() - t0})
        cpu["value"] = cpu["runs"][0]["value"]
        cpu["sample"] = (f"one acceleration_barnes_hut() call of the {'compiled reference (OpenMP walk, serial tree build)' if kname == 'reference' else 'C port (serial)'} "
                         f"at N={args.bh_cpu_n}, theta={theta}, {cpu['cores']} host threads; value = plummer")
    head = runs[0]
    return {"metric": "barnes_hut_force_eval_s_per_step", "value": head["value"], "unit": "s", "higher_is_better": False,
            "workload": f"Barnes-Hut full force evaluation (bbox, Morton keys, sort, octree + moments, walk), plummer N={head['n']}, theta={theta}, "
                        f"leaf={leaf}, eps={eps}; reference walk semantics, cooperative walk kernel (<= 1e-12 vs the reference)",
            "n_gpus": world, "scaling": "strong", "partition": "replicated build, walk sharded in interleaved 2048-position chunks, all-gather of the results" if world > 1 else "single GPU",
            "runs": runs, "cpu_baseline": cpu, "parity": parity,
            "stage_bytes_per_particle": BH_STAGE_BYTES,
            "roofline_note": "north_star names the HBM roofline for sort/build/walk: stage_hbm_frac = SURVEY 8d algorithmic bytes / stage time / measured HBM peak; "
                             "the walk is L2-latency / issue bound (L2 hit rate 99.8 %, profiles/r2_walk_coop_*.txt), items_per_s is its figure of merit"}


def run_b200_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("for --gpus > 1 launch under torch.distributed.run (one process per GPU)")

    from conftest import load_package
    gb = load_package()

    dist = None
    uid = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(gb.Context.new_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        uid = bytes(buf.cpu().numpy().tobytes())

    def barrier():
        if dist is not None:
            dist.barrier()

    _tags = [0]

    def idle_until_rank0(work):
        """rank 0 runs work(); the other ranks wait ON THE CPU (TCP store), leaving their GPUs idle.  A dist.barrier() would
        park an NCCL kernel on every waiting GPU, and a GPU shared by two processes is time-sliced: the device team that
        rank 0's single-threaded acceleration() call drives would get half of each of the other GPUs."""
        if dist is None:
            return work()
        _tags[0] += 1
        key = f"rank0_done_{_tags[0]}"
        store = dist.distributed_c10d._get_default_store()
        out = None
        if rank == 0:
            try:
                out = work()
            finally:
                store.set(key, "1")
        else:
            store.wait([key])
        return out

    def max_over_ranks(val):
        if dist is None:
            return val
        import torch
        t = torch.tensor([val], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.n
    x, v, m, G = make_ic(args.ic, n)
    ctx = gb.Context(device=local_rank, rank=rank, world_size=world, nccl_unique_id=uid)
    ctx.set_system(x, m, G, v)
    ctx.synchronize()

    def step():
        ctx.mark_positions_sharded()          # as after a drift: the gather is part of every force call
        ctx.acceleration("pairwise", args.eps)

    for _ in range(args.warmup):
        step()
    ctx.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = gb.kernel_launch_count()
    barrier()
    ctx.synchronize()
    dev_ms, kern_ms, gather_ms = 0.0, [], []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.flush_l2()                        # inputs (32 MiB) are smaller than L2: flush between timed iterations
        ctx.event_record(0)
        step()
        ctx.event_record(1)
        dev_ms += ctx.event_elapsed_ms(0, 1)  # CUDA events on the launching stream
        kern_ms.append(ctx.timing_ms(2))
        gather_ms.append(ctx.timing_ms(1))
    ctx.synchronize()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = gb.kernel_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = max_over_ranks(dev_ms)
    wall_ms = max_over_ranks(wall_ms)
    ms_per_step = dev_ms / args.steps
    inter = float(n) * float(n - 1)
    value = inter / (ms_per_step * 1e-3) / 1e9

    # roofline of the dominant kernel (direct_sum_kernel): FP64 pipe
    lo, hi = ctx.owned_range()
    k_ms = float(np.mean(kern_ms))
    k_inter = float(hi - lo) * float(n - 1)
    try:
        peak_tf, peak_mhz = gb.measure_fp64_peak(local_rank)
        peak_src = "measured live: register-resident DFMA loop (grav_b200_measure_fp64_peak)"
    except Exception as e:  # pragma: no cover
        peak_tf, peak_mhz, peak_src = NOMINAL_FP64_TFLOPS, 1965.0, f"nominal fallback ({e})"
    ach_tf = FLOP_PER_INTERACTION * k_inter / (k_ms * 1e-3) / 1e12
    pair_once, equal_mass = ctx.direct_sum_path()
    path = ("pair_once_equal_mass" if equal_mass else "pair_once") if pair_once else "ordered"
    ops = FP64_OPS[path]
    traffic = sym_ncu_traffic() if pair_once else NCU_DRAM_BYTES_PER_LAUNCH["ordered"]
    roofline = {
        "bound": "fp64", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
        "traffic": traffic if (world == 1 and n == (1 << 20)) else None,
        "traffic_note": ("dram__bytes_read+write of one launch at N=2^20, ncu --set full (profiles/r2_sym_n1m.txt): almost all of it the "
                         "RED.ADD.F64 updates of the per-CTA private accumulation arrays; " if pair_once else
                         "dram__bytes_read+write of one launch at N=2^20, ncu --set full (profiles/r1_direct_sum_n1m_final.txt; includes the 28 MB of partial sums of split target blocks); ")
                        + "algorithmic bytes = 32 B x N sources + 24 B x N results = 58.7 MB",
        "kernel": ("direct_sum_sym_kernel (every unordered pair once, Newton-3 like the reference's i<j loop; + the finishing kernel, <0.3% of the stage)"
                   if pair_once else "direct_sum_kernel<4,false,false> (+ the fix-up kernel of split target blocks, <0.1% of the stage)"),
        "kernel_ms": k_ms, "path": path,
        "convention": f"{FLOP_PER_INTERACTION} flop per ORDERED interaction, N(N-1) of them per force evaluation (SURVEY 8d), whatever the kernel really executes; "
                      f"peak = {peak_src} (= {peak_mhz:.0f} MHz x 148 SM x 64 DFMA lanes x 2)",
        "fp64_instr_per_ordered_interaction": ops,
        "fp64_pipe_util": ops * k_inter / (k_ms * 1e-3) / (peak_tf * 1e12 / 2.0),
        "fp64_pipe_util_note": f"{ops:g} FP64-pipe instructions really issued per ordered interaction over the measured DFMA issue rate: the honest "
                               "utilisation figure; frac above uses the 20-flop convention and therefore credits the pair-once kernel for the half of the r^-3 work it does not repeat",
    }

    # end to end through the reference-facing call with host buffers
    e2e = None
    if not args.no_e2e:
        a_host = np.empty((n, 3))
        for arr in (x, m, a_host):
            try:
                gb.host_register(arr)
            except Exception:
                pass
        # The call a user of the reference makes: the drop-in acceleration() symbol, host arrays in, host array out.  With
        # N > 1 GPUs that single-threaded call drives all N devices itself (GRAV_B200_DEVICES=N: an in-process device team,
        # include/grav_b200.h); rank 0 makes the call, the other ranks' processes wait at the barrier with their GPUs idle.
        os.environ["GRAV_B200_DEVICES"] = str(world)      # read when the library creates its default context (first one-shot call)

        def e2e_step():
            return gb.acceleration(x, m, G, "pairwise", args.eps, out=a_host)
        def e2e_run():
            e2e_step()
            t0 = time.perf_counter()
            ke = max(1, min(args.steps, 3))
            for _ in range(ke):
                e2e_step()
            return (time.perf_counter() - t0) / ke
        ctx.synchronize()
        barrier()
        e2e_s = idle_until_rank0(e2e_run) or 0.0
        barrier()
        e2e_s = max_over_ranks(e2e_s)
        e2e = {"value": inter / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(x.nbytes + m.nbytes) * world,
               "d2h_bytes_per_step": int(a_host.nbytes), "ms_per_step": e2e_s * 1e3,
               "api": "acceleration() in libgrav_sim_b200.so" + (f", GRAV_B200_DEVICES={world} (one calling thread drives {world} GPUs)" if world > 1 else ""),
               "host_memory": "pinned (cudaHostRegister on the caller's arrays)"}

    # sampled parity of what was just timed, against the reference's arithmetic (checker only, outside every timed region)
    parity = None
    if not args.no_parity:
        a_all = ctx.accelerations()
        if rank == 0:
            parity = direct_sum_parity(x, m, G, args.eps, a_all)

    # the same workload with unequal masses (the named Plummer sphere has equal masses, which lets the pair-once kernel
    # factor the mass out: 9 instead of 10 FP64 instructions per ordered interaction)
    gm = None
    if not args.no_parity:
        mg = m * np.random.default_rng(7).uniform(0.5, 1.5, n)
        ctx.set_system(x, mg, G, v)
        step(); ctx.synchronize()
        g_ms = 0.0
        kg = max(1, min(args.steps, 3))
        barrier()
        for _ in range(kg):
            ctx.flush_l2()
            ctx.event_record(0)
            step()
            ctx.event_record(1)
            g_ms += ctx.event_elapsed_ms(0, 1)
        g_ms = max_over_ranks(g_ms) / kg
        gpath = ctx.direct_sum_path()
        gm = {"value": inter / (g_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": g_ms, "steps": kg,
              "workload": "same positions, masses drawn uniformly from [0.5, 1.5] / N",
              "path": ("pair_once_equal_mass" if gpath[1] else "pair_once") if gpath[0] else "ordered"}
        ctx.set_system(x, m, G, v)

    # Barnes-Hut (second half of the metric): s per full force evaluation, theta = 0.5, leaf = 1
    bh = None
    if args.bh_n > 0:
        try:
            bh = run_barnes_hut(args, gb, ctx, rank, world, barrier, max_over_ranks, idle_until_rank0)
        except gb.GravB200Error as e:
            bh = {"unavailable": str(e)[:200]}

    # massless direct sum at config 3's size: HBM streaming, in practice launch-latency bound
    ml = None
    if args.whfast_n > 0 and world == 1:
        try:
            from gravity_simulator_b200 import ics as _ics
            xw, vw, mw, Gw = _ics.asteroid_belt(args.whfast_n, 7)
            ctx.set_system(xw, mw, Gw, vw)
            for _ in range(3):
                ctx.acceleration("massless", 0.0)
            ctx.synchronize()
            reps = 20
            ctx.event_record(6)
            for _ in range(reps):
                ctx.acceleration("massless", 0.0)
            ctx.event_record(7)
            ml_ms = ctx.event_elapsed_ms(6, 7) / reps
            nb = int(mw.shape[0])
            ml = {"metric": "massless_force_eval_ms", "value": ml_ms, "unit": "ms", "n": nb, "massive": int((mw != 0).sum()),
                  "algorithmic_bytes": 56 * nb, "achieved_GBps": 56.0 * nb / (ml_ms * 1e-3) / 1e9,
                  "hbm_frac": (56.0 * nb / (ml_ms * 1e-3) / 1e9) / _hbm_peak(),
                  "note": "9 x 1e5 interactions: a few launches of ~us each, bound by launch latency, not HBM"}
        except gb.GravB200Error as e:
            ml = {"unavailable": str(e)[:200]}

    # energy diagnostic: the O(N^2) potential sum behind compute_energy(), resident state
    en = None
    if world == 1:
        try:
            ne = 1 << 17
            xe, ve, me_, Ge = make_ic(args.ic, ne, seed=44)
            ctx.set_system(xe, me_, Ge, ve)
            ctx.energy()
            ctx.event_record(6)
            ctx.energy()
            ctx.event_record(7)
            e_ms = ctx.event_elapsed_ms(6, 7)
            en = {"metric": "energy_eval_ms", "value": e_ms, "unit": "ms", "n": ne, "G_pairs_per_s": ne * (ne - 1.0) / (e_ms * 1e-3) / 1e9,
                  "note": "every unordered pair once, 13 FP64-pipe instructions per pair (rsqrt seed + one correction): the direct sum's tile loop over mirrored target blocks with a scalar reduction; G_pairs_per_s counts ordered-pair equivalents N(N-1)"}
        except gb.GravB200Error as e:
            en = {"unavailable": str(e)[:200]}

    # WHFast side measurement (config 3): device-resident steps vs the reference's whfast() on the host cores
    wh = None
    if args.whfast_n > 0 and world == 1:
        try:
            from gravity_simulator_b200 import ics as _ics
            xw, vw, mw, Gw = _ics.asteroid_belt(args.whfast_n, 7)
            dtw, ksteps = 180.0, 200
            ctx.set_system(xw, mw, Gw, vw)
            ctx.whfast_begin(dtw, "massless", 0.0, True)
            ctx.whfast_steps(dtw, 9)
            ctx.synchronize()
            l0 = gb.kernel_launch_count()
            ctx.event_record(4)
            ctx.whfast_steps(dtw, ksteps)
            ctx.event_record(5)
            wh_ms = ctx.event_elapsed_ms(4, 5) / ksteps
            wh = {"metric": "whfast_steps_per_s", "value": 1e3 / wh_ms, "unit": "steps/s", "ms_per_step": wh_ms,
                  "n": int(mw.shape[0]), "massive": int((mw != 0).sum()), "dt_days": dtw, "acceleration": "massless",
                  "remove_invalid_particles": True, "launches_per_step": (gb.kernel_launch_count() - l0) / ksteps,
                  "state": "device-resident (sort, eta, Kepler drift, Jacobi transforms, acceleration, kick); bit-identical to the reference"}
            ctx.whfast_end()
            from oracle.bind import Reference
            if Reference.available():
                R = Reference()
                t0 = time.perf_counter()
                R.whfast_run(xw, vw, mw, Gw, dtw, dtw * 10, "massless", 0.0, False)
                wh["cpu_reference_ms_per_step"] = (time.perf_counter() - t0) / 10 * 1e3
                wh["cpu_reference_note"] = (f"the reference's whfast() (OpenMP build, {os.cpu_count()} host cores; its sort and both "
                                            "coordinate transforms are serial), 10 steps")
        except gb.GravB200Error as e:
            wh = {"unavailable": str(e)[:200]}

    cpu = None
    if rank == 0 and args.cpu_n > 0:
        t_cpu, kname = cpu_reference_time(args.cpu_n, args.eps, args.ic)
        cpu = {"value": args.cpu_n * (args.cpu_n - 1) / t_cpu / 1e9, "unit": UNIT, "cores": 1, "kind": kname,
               "sample": f"one pairwise force evaluation at N={args.cpu_n} ({args.ic}, eps={args.eps}), {t_cpu:.1f} s on 1 core "
                         f"(the reference's direct sum is single-threaded); rate is size-independent"}

    ctx.close()
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"direct-sum pairwise FP64 force evaluation, {args.ic} N={n}, eps={args.eps}",
                       "partition": (f"pair units sharded evenly over {world} rank(s): NCCL all-gather of positions, all-reduce of the accelerations per step" if pair_once else
                                     f"targets sharded over {world} rank(s), NCCL all-gather of positions per step") if world > 1 else "single GPU",
                       "l2": "256 MiB L2 flush between timed iterations",
                       # second half of BASELINE.json's metric ("... & BH force-eval s/step"): kept inside config so the
                       # driver's parser, which keeps config whole, carries it into BENCH / SCALE
                       "barnes_hut": bh},
            "wall_ms_per_step": wall_ms / args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "parity": parity, "unequal_masses": gm, "massless": ml, "whfast": wh, "energy": en, "gpu_launches": int(launches),
            "clocks": clocks, "allgather_ms": float(np.mean(gather_ms)) if world > 1 else 0.0,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def sym_ncu_traffic():
    """dram read + write bytes of one direct_sum_sym_kernel launch at N = 2^20 (profiles/r2_sym_n1m.txt), or None."""
    try:
        rd = wr = None
        for ln in open(ROOT / "profiles" / "r2_sym_n1m.txt"):
            f = ln.split()
            if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                val = float(f[1]) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[f[2]]
                if f[0].endswith("read.sum") and rd is None:
                    rd = val
                if f[0].endswith("write.sum") and wr is None:
                    wr = val
        return int(rd + wr) if rd is not None and wr is not None else None
    except Exception:
        return None


def _hbm_peak():
    try:
        return float(json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"])
    except Exception:
        return 6650.0   # B200_PROFILING.md fallback


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
