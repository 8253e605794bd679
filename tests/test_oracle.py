"""CPU tests: the oracle restatement (oracle/grav_oracle.c) against the golden vectors minted from the
unmodified reference, and -- where oracle/_ref is present -- against the reference library itself."""
import numpy as np
import pytest

from conftest import max_rel_err

TREE_KEYS = ["keys", "sorted_indices", "num_particles", "num_children", "first_particle", "first_child", "mass",
             "com_x", "com_y", "com_z"]
FORCE_CASES = ["solar_forces", "plummer2048", "uniform1500", "clustered1024", "massless600", "tiny1", "tiny2", "tiny3"]


def parse_bh(key):
    # a_bh_t0.5_l1
    _, _, t, l = key.split("_")
    return float(t[1:]), int(l[1:])


@pytest.mark.parametrize("case", FORCE_CASES)
def test_oracle_matches_golden_forces(oracle, golden, case):
    g = golden(case)
    x, m, G, eps = g["x"], g["m"], float(g["G"]), float(g["eps"])
    for key in g.files:
        if key == "a_pairwise":
            a = oracle.acceleration(x, m, G, "pairwise", eps)
        elif key == "a_massless":
            a = oracle.acceleration(x, m, G, "massless", eps)
        elif key.startswith("a_bh_"):
            theta, leaf = parse_bh(key)
            a = oracle.acceleration(x, m, G, "barnes_hut", eps, theta, leaf)
        else:
            continue
        # same operations in the same order as the reference: bit-exact
        assert np.array_equal(a, g[key], equal_nan=True), (case, key, max_rel_err(a, g[key]))


@pytest.mark.parametrize("case", ["solar_forces", "plummer2048", "uniform1500", "clustered1024", "tiny1", "tiny2", "tiny3"])
def test_oracle_matches_golden_trees(oracle, golden, case):
    g = golden(case)
    leaves = sorted({int(k.split("_")[1][1:]) for k in g.files if k.startswith("tree_l")})
    assert leaves
    for leaf in leaves:
        t = oracle.construct_octree(g["x"], g["m"], leaf)
        assert t["num_nodes"] == int(g[f"tree_l{leaf}_num_nodes"])
        assert t["box_width"] == float(g[f"tree_l{leaf}_box_width"])
        for k in TREE_KEYS:
            assert np.array_equal(t[k], g[f"tree_l{leaf}_{k}"], equal_nan=True), (case, leaf, k)


def test_oracle_matches_golden_whfast(oracle, golden):
    for case in ("whfast_belt", "whfast_solar"):
        g = golden(case)
        for key in g.files:
            if not key.startswith("a_whfast_"):
                continue
            method = key.split("_")[2]
            eps = float(key.split("eps")[1])
            a = oracle.whfast_acceleration(g["x"], g["m"], float(g["G"]), g["jacobi_x"], g["eta"], method, eps)
            assert np.array_equal(a, g[key], equal_nan=True), (case, key, max_rel_err(a[1:], g[key][1:]))


def test_oracle_vs_reference_live(oracle, reference, ics):
    """Fresh inputs not in the golden set (bigger, other seeds); needs the compiled reference."""
    for name, (x, v, m, G) in {"uniform": ics.uniform_cube(5000, 11), "plummer": ics.plummer(6000, 12),
                               "clustered": ics.clustered(3000, 13)}.items():
        for leaf in (1, 3):
            to, tr = oracle.construct_octree(x, m, leaf), reference.construct_octree(x, m, leaf)
            assert to["num_nodes"] == tr["num_nodes"] and to["box_width"] == tr["box_width"]
            for k in TREE_KEYS:
                assert np.array_equal(to[k], tr[k], equal_nan=True), (name, leaf, k)
        for theta in (0.3, 0.5, 1.0):
            ao = oracle.acceleration(x, m, G, "barnes_hut", 0.01, theta, 1)
            ar = reference.acceleration(x, m, G, "barnes_hut", 0.01, theta, 1)
            assert np.array_equal(ao, ar, equal_nan=True), (name, theta)
        assert np.array_equal(oracle.acceleration(x, m, G, "pairwise", 0.0), reference.acceleration(x, m, G, "pairwise", 0.0),
                              equal_nan=True)   # exact duplicates with eps = 0 give NaN in both
        assert oracle.energy(x, v, m, G) == reference.energy(x, v, m, G)


def test_known_answer_invariants(oracle, golden):
    """Invariants the reference itself obeys: theta=0 BH == pairwise to rounding,
    Newton-3 momentum conservation, massless == pairwise when no mass is zero."""
    g = golden("uniform1500")
    x, m, G = g["x"], g["m"], float(g["G"])
    ap = oracle.acceleration(x, m, G, "pairwise", 0.0)
    assert max_rel_err(g["a_bh_t0.0_l1"], ap) < 1e-12
    assert np.abs((m[:, None] * ap).sum(0)).max() < 1e-13 * np.abs(m[:, None] * ap).sum()
    assert np.array_equal(oracle.acceleration(x, m, G, "massless", 0.0), ap)


def test_fixed_mode_is_more_accurate(oracle, golden):
    """The opt-in corrected walk is a real improvement over the reference semantics."""
    g = golden("plummer2048")
    x, m, G, eps = g["x"], g["m"], float(g["G"]), float(g["eps"])
    exact = g["a_pairwise"]
    ref_mode = g["a_bh_t0.5_l1"]
    fixed = oracle.acceleration(x, m, G, "barnes_hut", eps, 0.5, 1, fixed=True)
    err = lambda a: np.mean(np.linalg.norm(a - exact, axis=1) / np.linalg.norm(exact, axis=1))
    assert err(fixed) < 0.02 < err(ref_mode)


# ---- WHFast step pieces (oracle/whfast_oracle.c), SURVEY.md section 8f row N2 -----------------------------------
@pytest.mark.parametrize("case", ["whfast_run_solar", "whfast_run_belt", "whfast_run_removal"])
def test_oracle_whfast_run_matches_golden(oracle, golden, case):
    """Whole whfast() runs minted from the reference (sort, eta, Kepler drift with removal, both transforms, kick)."""
    g = golden(case)
    dt, steps = float(g["dt"]), int(g["steps"])
    o = oracle.whfast_integrate(g["x"], g["v"], g["m"], float(g["G"]), dt, dt * steps, str(g["method"]), float(g["eps"]), True)
    if case == "whfast_run_removal":
        assert g["out_m"].shape[0] < g["m"].shape[0]
    for k in ("x", "v", "m", "ids"):
        assert np.array_equal(o[k], g[f"out_{k}"], equal_nan=True), (case, k)


def test_oracle_whfast_vs_reference_live(oracle, reference, ics, monkeypatch):
    """Fresh systems through the reference's own whfast() and its static stage functions (ref_whfast_probe.c)."""
    monkeypatch.setenv("OMP_NUM_THREADS", "1")
    for k, seed, grazers, method, steps, dt, eps in [(1500, 31, 0, "massless", 8, 180.0, 0.0), (30, 32, 0, "pairwise", 6, 3.0, 0.01),
                                                     (2500, 33, 60, "massless", 5, 180.0, 0.0)]:
        x, v, m, G = ics.asteroid_belt(k, seed, grazers=grazers)
        r = reference.whfast_run(x, v, m, G, dt, dt * steps, method, eps, True)
        o = oracle.whfast_integrate(x, v, m, G, dt, dt * steps, method, eps, True)
        assert r["m"].shape == o["m"].shape and (grazers == 0 or r["m"].shape[0] < m.shape[0])
        for q in ("x", "v", "m", "ids"):
            assert np.array_equal(r[q], o[q], equal_nan=True), (k, method, q)
    # stage by stage on a distance-sorted belt
    x, v, m, G = ics.asteroid_belt(2000, 34)
    order = np.argsort(np.linalg.norm(x - x[0], axis=1), kind="stable")
    x, v, m = x[order], v[order], m[order]
    sr, so = reference.whfast_stages(x, v, m, G, 180.0), oracle.whfast_stages(x, v, m, G, 180.0)
    for q in sr:
        assert np.array_equal(sr[q], so[q]), q
    for z in (0.0, 0.05, -0.09, 0.3, -7.5, 123.456, -1e4, 3e7):
        assert np.array_equal(reference.stumpff(z), oracle.stumpff(z)), z


def test_sampled_target_functions_are_the_full_loops(oracle, reference, golden, ics):
    """The sampled-target entries (used for reference comparisons at N = 2^20 / 2^24, where the full CPU loops take an
    hour) give bit for bit the rows of the full restatement AND of the compiled reference."""
    rng = np.random.default_rng(5)
    for x, v, m, G in (ics.plummer(3000, 11), ics.clustered(2000, 12)):
        for eps in (0.0, 0.01):
            tg = rng.choice(m.shape[0], 200, replace=False)
            full = reference.acceleration(x, m, G, "pairwise", eps)
            got = oracle.pairwise_targets(x, m, G, eps, tg)
            assert np.array_equal(got, full[tg], equal_nan=True)
            ld = oracle.pairwise_targets(x, m, G, eps, tg, long_double=True)
            ok = np.isfinite(full[tg]).all(axis=1)
            assert max_rel_err(ld[ok], full[tg][ok]) < 1e-12
        for theta, leaf, fixed in ((0.5, 1, False), (1.0, 3, False), (0.3, 1, True)):
            with oracle.tree(x, m, leaf) as t:
                pos = rng.choice(m.shape[0], 300, replace=False)
                a, st = t.walk_targets(G, 0.01, theta, pos, fixed=fixed, stats=True)
                perm = t.to_dict()["sorted_indices"]
            full = oracle.acceleration(x, m, G, "barnes_hut", 0.01, theta, leaf, fixed=fixed)
            assert np.array_equal(a, full[perm[pos]], equal_nan=True)
            if not fixed:
                assert np.array_equal(a, reference.acceleration(x, m, G, "barnes_hut", 0.01, theta, leaf)[perm[pos]], equal_nan=True)
            assert (st[:, 0] == st[:, 1] + st[:, 2] + (st[:, 0] - st[:, 1] - st[:, 2])).all() and (st[:, 0] > 0).all()
