"""GPU tests of the reference-facing boundary beyond acceleration(): benchmark_acceleration (H-1,
src/acceleration.c:369-507) driven exactly like examples/benchmark/c/benchmark.c:37-53 drives it, side by side with the
compiled reference on the same system."""
import ctypes as C
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

_libc = C.CDLL(None)


def _run_benchmark(lib, status_t, system_t, param_t, x, m, G, params, times, capfd):
    """Calls lib.benchmark_acceleration(system, params[], n, times[]) and returns (return_code, captured stdout)."""
    n = m.shape[0]
    s = system_t()
    s.num_particles = n
    s.particle_ids = None
    s.x = x.ctypes.data_as(C.POINTER(C.c_double)); s.v = None; s.m = m.ctypes.data_as(C.POINTER(C.c_double)); s.G = G
    arr = (param_t * len(params))()
    for k, (method, theta, eps, leaf) in enumerate(params):
        arr[k].method = method; arr[k].opening_angle = theta; arr[k].softening_length = eps
        arr[k].max_num_particles_per_leaf = leaf
    nt = (C.c_int * len(times))(*times)
    lib.benchmark_acceleration.restype = status_t
    lib.benchmark_acceleration.argtypes = [C.POINTER(system_t), C.POINTER(param_t), C.c_int, C.POINTER(C.c_int)]
    capfd.readouterr()
    st = lib.benchmark_acceleration(C.byref(s), arr, len(params), nt)
    _libc.fflush(None)
    out = capfd.readouterr().out
    return st.return_code, out


def _parse(out):
    """[(test index, 'skipped' | (method, number of times, MAE string))]"""
    res = []
    for blk in re.split(r"(?=Test \d+:)", out):
        m = re.match(r"Test (\d+):\s+(.*)", blk, re.S)
        if not m:
            continue
        body = m.group(2)
        if body.startswith("Skipped"):
            res.append((int(m.group(1)), "skipped"))
        else:
            res.append((int(m.group(1)), (re.search(r"Method: (\S+)", body).group(1), int(re.search(r"Number of times: (\d+)", body).group(1)),
                                          re.search(r"MAE: (\S+)", body).group(1), re.search(r"Avg time: (\S+) \(\+- (\S+)\) s", body) is not None)))
    return res


def test_benchmark_acceleration_behaves_like_the_reference(gb, reference, ics, capfd):
    """Same system, same parameter sets, through the reference's benchmark_acceleration and through the drop-in's:
    * the first run of set 0 is the comparison vector (its MAE prints 0, every later set prints sum|da| over the three
      components / N against it, :452-465) -- the printed MAEs agree with the reference's to the printed 3 digits;
    * sets with num_times <= 0 are reported as skipped and not run (:399-403);
    * method names, repetition counts and the mean (+- std) line are printed per set;
    * an unknown method ends the run with GRAV_VALUE_ERROR (code 2, src/error.h:14-21)."""
    from oracle.bind import RefAccelerationParam, RefErrorStatus, RefSystem
    _, shim = gb.load()
    x, v, m, G = ics.plummer(3000, 6)
    m[100:] = 0.0   # massive prefix, as every shipped example has it: massless == pairwise on the massive block
    params = [(1, 1.0, 0.01, 1), (2, 1.0, 0.01, 1), (3, 0.5, 0.01, 1), (3, 0.0, 0.01, 1), (3, 1.0, 0.01, 4)]
    times = [2, 1, 3, 0, 1]
    rc_r, out_r = _run_benchmark(reference.L, RefErrorStatus, RefSystem, RefAccelerationParam, x, m, G, params, times, capfd)
    rc_g, out_g = _run_benchmark(shim, gb.ErrorStatus, gb.System, gb.AccelerationParam, x, m, G, params, times, capfd)
    assert rc_r == 0 and rc_g == 0
    assert out_g.startswith("Benchmarking acceleration...")
    pr, pg = _parse(out_r), _parse(out_g)
    assert len(pg) == len(params)
    strip = lambda r: [(i, t if t == "skipped" else (t[0], t[1], t[3])) for i, t in r]
    assert strip(pg) == strip(pr), (out_g, out_r)
    assert pg[3] == (3, "skipped")
    # printed MAEs (3 significant digits) agree with the reference's up to the rounding noise of the two force vectors
    scale = np.abs(reference.acceleration(x, m, G, "pairwise", 0.01)).sum() / m.shape[0]
    for (i, tg), (_, tr) in zip(pg, pr):
        if tg != "skipped":
            assert abs(float(tg[2]) - float(tr[2])) <= 5e-3 * float(tr[2]) + 1e-12 * scale, (i, tg, tr)
    assert pg[0][1][:2] == ("Pairwise", 2) and pg[0][1][2] == "0"            # the comparison vector against itself
    assert pg[1][1][0] == "Massless" and pg[2][1][:2] == ("Barnes-Hut", 3) and float(pg[2][1][2]) > 0.0
    # the MAE formula itself, recomputed from the drop-in's own accelerations
    a0 = gb.acceleration(x, m, G, "pairwise", 0.01)
    a2 = gb.acceleration(x, m, G, "barnes_hut", 0.01, 0.5, 1)
    assert abs(float(pg[2][1][2]) - np.abs(a0 - a2).sum() / m.shape[0]) <= 5e-3 * float(pg[2][1][2])
    # unknown method
    rc_g, out_g = _run_benchmark(shim, gb.ErrorStatus, gb.System, gb.AccelerationParam, x, m, G, [(1, 1.0, 0.0, 1), (7, 1.0, 0.0, 1)], [1, 1], capfd)
    assert rc_g == 2


def test_benchmark_acceleration_skips_everything_without_touching_the_gpu(gb, ics, capfd):
    _, shim = gb.load()
    x, v, m, G = ics.uniform_cube(64, 1)
    n0 = gb.kernel_launch_count()
    rc, out = _run_benchmark(shim, gb.ErrorStatus, gb.System, gb.AccelerationParam, x, m, G, [(1, 1.0, 0.0, 1), (3, 1.0, 0.0, 1)], [0, -2], capfd)
    assert rc == 0 and out.count("Skipped since num_times") == 2 and gb.kernel_launch_count() == n0
