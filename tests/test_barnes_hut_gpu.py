"""GPU parity for the Barnes-Hut path, through the drop-in `construct_octree()` / `acceleration()` symbols.

Bars (BASELINE.json north_star): Morton keys, sort permutation and octree node layout bit-exact; moments
bit-exact (ordered non-contracted sums); accelerations compared with the reference's own output in both walk
settings (include/grav_b200.h, grav_b200_set_bh_exact):
  exact = 1  per-lane walk in the reference's depth-first order with IEEE operations: BIT EQUALITY asserted;
  exact = 0  (default) warp-cooperative walk, same per-particle decisions, different summation order: <= 1e-12 max
             relative error per particle and the same NaN rows (the tolerance north_star states for this path)."""
import numpy as np
import pytest

from conftest import assert_forces_close, bh_exact, max_rel_err

TOL = 1e-12


def check_forces(a, ref, exact, ctx=""):
    if exact:
        assert np.array_equal(a, ref, equal_nan=True), (ctx, max_rel_err(np.nan_to_num(a), np.nan_to_num(ref)))
    else:
        assert_forces_close(a, ref, TOL, ctx)

pytestmark = pytest.mark.gpu

TREE_KEYS = ["keys", "sorted_indices", "num_particles", "num_children", "first_particle", "first_child", "mass",
             "com_x", "com_y", "com_z"]
CASES = ["solar_forces", "plummer2048", "uniform1500", "clustered1024", "tiny1", "tiny2", "tiny3"]


def assert_tree_equal(t, ref, ctx=""):
    assert t["num_nodes"] == ref["num_nodes"], ctx
    assert t["box_width"] == ref["box_width"], ctx
    for k in TREE_KEYS:
        assert np.array_equal(t[k], ref[k], equal_nan=True), (ctx, k)


def golden_tree(g, leaf):
    d = {k: g[f"tree_l{leaf}_{k}"] for k in TREE_KEYS}
    d["num_nodes"] = int(g[f"tree_l{leaf}_num_nodes"])
    d["box_width"] = float(g[f"tree_l{leaf}_box_width"])
    return d


@pytest.mark.parametrize("case", CASES)
def test_tree_matches_golden(gb, golden, case):
    g = golden(case)
    for leaf in sorted({int(k.split("_")[1][1:]) for k in g.files if k.startswith("tree_l")}):
        assert_tree_equal(gb.construct_octree(g["x"], g["m"], leaf), golden_tree(g, leaf), f"{case} leaf={leaf}")


@pytest.mark.parametrize("exact", [0, 1])
@pytest.mark.parametrize("case", CASES)
def test_bh_acceleration_matches_golden(gb, golden, case, exact):
    g = golden(case)
    keys = [k for k in g.files if k.startswith("a_bh_")]
    assert keys
    with bh_exact(gb, exact):
        for key in keys:
            _, _, t, l = key.split("_")
            a = gb.acceleration(g["x"], g["m"], float(g["G"]), "barnes_hut", float(g["eps"]), float(t[1:]), int(l[1:]))
            check_forces(a, g[key], exact, (case, key))


def test_morton_keys_stage(gb, oracle, ics):
    for x in (ics.uniform_cube(10000, 3)[0], ics.plummer(7777, 4)[0], ics.clustered(3000, 5)[0], np.zeros((5, 3)),
              np.array([[1.0, 2.0, 3.0]])):
        k, c, w = gb.morton_keys(x)
        ko, co, wo = oracle.morton_keys(x)
        assert np.array_equal(k, ko) and np.array_equal(c, co, equal_nan=True) and (w == wo)
    # the particle on the max edge of the widest axis wraps to cell 0 on that axis (reference quirk)
    x = ics.uniform_cube(1000, 6)[0]
    k, c, w = gb.morton_keys(x)
    widest = np.argmax(x.max(0) - x.min(0))
    i = np.argmax(x[:, widest])
    assert (k[i] >> widest) & 0x1249249249249249 == 0


@pytest.mark.parametrize("n,leaf", [(1 << 15, 1), (50000, 2), (1 << 17, 1), (100003, 8),
                                    (131073, 1), (300007, 1), (1 << 19, 3)])      # > 131072: one-kernel-per-pass sort
def test_tree_vs_oracle_large(gb, oracle, ics, n, leaf):
    for name, (x, v, m, G) in {"plummer": ics.plummer(n, n), "uniform": ics.uniform_cube(n, n + 1)}.items():
        m = m * np.random.default_rng(n).uniform(0.5, 1.5, n)
        assert_tree_equal(gb.construct_octree(x, m, leaf), oracle.construct_octree(x, m, leaf), f"{name} n={n} leaf={leaf}")


def test_tree_vs_oracle_clustered_and_fixed_box(gb, oracle, ics):
    x, v, m, G = ics.clustered(20000, 17)
    for leaf in (1, 4):
        assert_tree_equal(gb.construct_octree(x, m, leaf), oracle.construct_octree(x, m, leaf), f"clustered leaf={leaf}")
    # caller-supplied box (construct_octree's box_center / box_width arguments, src/linear_octree.c:856-867)
    x, v, m, G = ics.uniform_cube(5000, 2)
    c, w = np.array([0.1, -0.2, 0.05]), 2.5
    assert_tree_equal(gb.construct_octree(x, m, 1, c, w), oracle.construct_octree(x, m, 1, c, w), "fixed box")


@pytest.mark.parametrize("exact", [0, 1])
@pytest.mark.parametrize("n,theta,eps,leaf", [(20000, 0.5, 0.0, 1), (30000, 1.0, 0.01, 1), (16384, 0.3, 0.01, 4), (60000, 0.5, 0.0, 1),
                                              (9, 0.5, 0.0, 1), (33, 0.7, 0.0, 8), (5000, 0.0, 0.01, 1)])
def test_bh_acceleration_vs_oracle(gb, oracle, ics, n, theta, eps, leaf, exact):
    """Includes config 4 (two-Plummer galaxy collision, N=60000, theta=0.5, eps=0)."""
    x, v, m, G = ics.two_plummer(n // 2, seed=n) if n == 60000 else ics.plummer(n, seed=n)
    with bh_exact(gb, exact):
        a = gb.acceleration(x, m, G, "barnes_hut", eps, theta, leaf)
    ref = oracle.acceleration(x, m, G, "barnes_hut", eps, theta, leaf)
    check_forces(a, ref, exact, (n, theta, eps, leaf))


@pytest.mark.parametrize("exact", [0, 1])
def test_bh_clustered_duplicates_and_big_leaves(gb, oracle, ics, exact):
    """Deep chains, a knot below the level-21 cell size, exact duplicates (multi-particle level-21 leaves; NaN rows with
    eps = 0 exactly where the reference has them) and leaves of up to 8 particles."""
    x, v, m, G = ics.clustered(6000, 21)
    for eps, leaf in ((0.0, 1), (1e-3, 1), (1e-3, 8)):
        with bh_exact(gb, exact):
            a = gb.acceleration(x, m, G, "barnes_hut", eps, 0.5, leaf)
        check_forces(a, oracle.acceleration(x, m, G, "barnes_hut", eps, 0.5, leaf), exact, ("clustered", eps, leaf))


def test_bh_theta_zero_equals_direct_sum(gb, ics):
    x, v, m, G = ics.uniform_cube(4000, 9)
    a_bh = gb.acceleration(x, m, G, "barnes_hut", 0.0, 0.0, 1)
    a_ds = gb.acceleration(x, m, G, "pairwise", 0.0)
    assert max_rel_err(a_bh, a_ds) < 1e-12


def test_bh_fixed_mode(gb, oracle, ics):
    """Opt-in corrected walk: matches the oracle's fixed walk bit for bit and is far closer to the direct sum."""
    abi, _ = gb.load()
    x, v, m, G = ics.plummer(20000, 33)
    exact = gb.acceleration(x, m, G, "pairwise", 0.01)
    ref_mode = gb.acceleration(x, m, G, "barnes_hut", 0.01, 0.5, 1)
    gb.check_rc(abi.grav_b200_set_bh_mode(gb.BH_FIXED))
    try:
        fixed = gb.acceleration(x, m, G, "barnes_hut", 0.01, 0.5, 1)
        with bh_exact(gb, 1):
            fixed_exact = gb.acceleration(x, m, G, "barnes_hut", 0.01, 0.5, 1)
    finally:
        gb.check_rc(abi.grav_b200_set_bh_mode(gb.BH_REFERENCE))
    fixed_ref = oracle.acceleration(x, m, G, "barnes_hut", 0.01, 0.5, 1, fixed=True)
    assert np.array_equal(fixed_exact, fixed_ref)
    assert_forces_close(fixed, fixed_ref, TOL, "fixed mode, cooperative walk")
    err = lambda a: float(np.mean(np.linalg.norm(a - exact, axis=1) / np.linalg.norm(exact, axis=1)))
    assert err(fixed) < 0.02 < err(ref_mode)


def test_bh_full_size_properties(gb, ics):
    """N = 2^20 (the BH bench size): determinism and tree invariants that do not need the CPU oracle."""
    n = 1 << 20
    x, v, m, G = ics.plummer(n, 77)
    t = gb.construct_octree(x, m, 1)
    assert np.all(np.diff(t["keys"]) >= 0)                                   # sortedness
    assert np.array_equal(np.sort(t["sorted_indices"]), np.arange(n))        # a permutation
    same = t["keys"][1:] == t["keys"][:-1]
    assert np.all(t["sorted_indices"][1:][same] > t["sorted_indices"][:-1][same])   # stable
    nch, fc, npart = t["num_children"], t["first_child"], t["num_particles"]
    internal = np.nonzero(nch > 0)[0]
    assert t["num_nodes"] == 1 + nch.sum()
    assert npart[0] == n and abs(t["mass"][0] - m.sum()) < 1e-12
    # children partition their parent
    csum = np.add.reduceat(npart[1:], (fc[internal] - 1)[np.argsort(fc[internal])])
    assert np.array_equal(np.sort(csum), np.sort(npart[internal]))
    a1 = gb.acceleration(x, m, G, "barnes_hut", 0.01, 0.5, 1)
    a2 = gb.acceleration(x, m, G, "barnes_hut", 0.01, 0.5, 1)
    assert np.array_equal(a1, a2) and np.isfinite(a1).all()


@pytest.mark.parametrize("kind", ["plummer", "uniform"])
def test_bh_bench_size_vs_compiled_reference(gb, reference, ics, kind):
    """N = 2^20, theta = 0.5 -- the size bench.py times -- against the UNMODIFIED reference compiled to oracle/_ref
    (OpenMP walk, a few seconds): keys, permutation, all node arrays and moments bit-equal; accelerations bit-equal in
    exact mode and <= 1e-12 in the default cooperative mode."""
    n = 1 << 20
    x, v, m, G = ics.plummer(n, 43) if kind == "plummer" else ics.uniform_cube(n, 43)
    assert_tree_equal(gb.construct_octree(x, m, 1), reference.construct_octree(x, m, 1), f"{kind} n=2^20")
    ref = reference.acceleration(x, m, G, "barnes_hut", 0.01, 0.5, 1)
    for exact in (0, 1):
        with bh_exact(gb, exact):
            check_forces(gb.acceleration(x, m, G, "barnes_hut", 0.01, 0.5, 1), ref, exact, (kind, "2^20", exact))


def test_bh_max_size_2_24(gb, oracle, ics):
    """N = 2^24 = GRAV_B200_MAX_PARTICLES (the north-star Barnes-Hut size; stresses the 26-bit count field, the 32-bit
    unit indices and the sort's status words): the whole tree against the C port, and the accelerations of 4096 sampled
    targets (128 runs of 32 consecutive sorted positions) against the port's walk, in both walk settings."""
    n = 1 << 24
    x, v, m, G = ics.plummer(n, 45)
    with oracle.tree(x, m, 1) as T:
        ref_tree = T.to_dict()
        starts = np.random.default_rng(0).choice(n // 32, 128, replace=False) * 32
        pos = (starts[:, None] + np.arange(32)[None, :]).ravel()
        ref = T.walk_targets(G, 0.01, 0.5, pos)
    t = gb.construct_octree(x, m, 1)
    assert_tree_equal(t, ref_tree, "plummer n=2^24")
    ids = ref_tree["sorted_indices"][pos]
    del t, ref_tree
    for exact in (0, 1):
        with bh_exact(gb, exact):
            a = gb.acceleration(x, m, G, "barnes_hut", 0.01, 0.5, 1)
        assert np.isfinite(a).all()
        check_forces(a[ids], ref, exact, ("2^24 sampled", exact))
