"""Device-resident WHFast against the oracle's restatement of whfast()
(oracle/whfast_oracle.c, pinned bit-exact against the reference in tests/test_oracle.py): all arithmetic is IEEE in
the reference's order, so every comparison is np.array_equal -- particle order and ids included."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _same(got, ref):
    assert got["ids"].shape == ref["ids"].shape, (got["ids"].shape, ref["ids"].shape)
    for k in ("ids", "m", "x", "v"):
        assert np.array_equal(got[k], ref[k], equal_nan=True), (k, np.nanmax(np.abs(got[k] - ref[k])))


def _run_gpu(gb, x, v, m, G, dt, steps, method, eps, remove, chunks=(None,), ids=None):
    with gb.Context() as c:
        c.set_system(x, m, G, v)
        c.whfast_begin(dt, method, eps, remove, ids=ids)
        left = steps
        for ch in chunks:
            k = left if ch is None else min(ch, left)
            c.whfast_steps(dt, k)
            left -= k
        assert left == 0
        out = c.whfast_state(snapshot=False)
        c.whfast_end()
    return out


@pytest.mark.parametrize("k,seed,steps,method,eps,dt", [
    (0, 1, 40, "pairwise", 0.0, 5.0),          # config 1's system, every body massive
    (0, 1, 25, "massless", 0.0, 5.0),
    (60, 2, 8, "pairwise", 0.01, 2.0),         # massless particles through the all-pairs kernel, softened
    (500, 3, 20, "massless", 0.0, 180.0),
    (5000, 4, 12, "massless", 0.0, 180.0),
    (20000, 5, 70, "massless", 0.0, 30.0),     # several optimistic batches
])
def test_resident_whfast_matches_oracle(gb, oracle, ics, k, seed, steps, method, eps, dt):
    x, v, m, G = ics.asteroid_belt(k, seed)
    ref = oracle.whfast_integrate(x, v, m, G, dt, dt * steps, method, eps, True)
    got = _run_gpu(gb, x, v, m, G, dt, steps, method, eps, True, chunks=(3, 1, None))
    _same(got, ref)


def test_resident_whfast_above_the_small_sort_limit(gb, oracle, ics):
    """More than 131072 particles: the distance sort takes the large-n (one kernel per pass) path."""
    x, v, m, G = ics.asteroid_belt(150000, 8)
    ref = oracle.whfast_integrate(x, v, m, G, 180.0, 180.0 * 3, "massless", 0.0, True)
    _same(_run_gpu(gb, x, v, m, G, 180.0, 3, "massless", 0.0, True), ref)


def test_resident_whfast_config3_size(gb, oracle, ics):
    """Config 3 at its full size: Sun + 8 planets + 1e5 massless asteroids, dt = 180 d."""
    x, v, m, G = ics.asteroid_belt(100000, 7)
    ref = oracle.whfast_integrate(x, v, m, G, 180.0, 180.0 * 5, "massless", 0.0, True)
    got = _run_gpu(gb, x, v, m, G, 180.0, 5, "massless", 0.0, True)
    _same(got, ref)


@pytest.mark.parametrize("k,grazers,seed,steps", [(2000, 40, 5, 6), (7000, 100, 6, 4), (3000, 25, 9, 40)])
def test_resident_whfast_removes_invalid_particles(gb, oracle, ics, k, grazers, seed, steps):
    """Kepler solves that fail flag particles for removal (:551); the optimistic batch is replayed and the flagged step
    finished with a stable compaction -- same survivors, same order, same state as the serial reference."""
    x, v, m, G = ics.asteroid_belt(k, seed, grazers=grazers)
    ref = oracle.whfast_integrate(x, v, m, G, 180.0, 180.0 * steps, "massless", 0.0, True)
    assert ref["m"].shape[0] < m.shape[0]                      # the case does remove particles
    got = _run_gpu(gb, x, v, m, G, 180.0, steps, "massless", 0.0, True)
    _same(got, ref)


def test_resident_whfast_removal_message(gb, oracle, ics, capfd):
    """With the caller's Settings.verbose at GRAV_VERBOSITY_VERBOSE a removal prints the reference's line
    (src/integrator_whfast.c:609-623): the count and the ids of the removed particles in index order.  The ids named over
    the whole run are exactly the particles the oracle's run lost; nothing is printed at a lower level."""
    import re
    x, v, m, G = ics.asteroid_belt(2000, 5, grazers=40)
    ref = oracle.whfast_integrate(x, v, m, G, 180.0, 180.0 * 6, "massless", 0.0, True)
    lost = sorted(set(range(m.shape[0])) - set(int(i) for i in ref["ids"]))
    assert lost
    for level in (3, 2):
        with gb.Context() as c:
            c.set_system(x, m, G, v)
            c.whfast_begin(180.0, "massless", 0.0, True)
            c.whfast_set_verbose(level)
            capfd.readouterr()
            c.whfast_steps(180.0, 6)
            got = c.whfast_state(snapshot=False)
            c.whfast_end()
        err = capfd.readouterr().err
        _same(got, ref)
        lines = re.findall(r"whfast_drift: Removing (\d+) invalid particles\. Particle IDs: \[([0-9, ]+)\]", err)
        if level < 3:
            assert not lines
            continue
        named = [int(t) for _, ids in lines for t in ids.split(",")]
        assert sorted(named) == lost and all(int(cnt) == len(ids.split(",")) for cnt, ids in lines)


def test_resident_whfast_without_removal(gb, oracle, ics):
    """whfast_remove_invalid_particles = false: no checkpoints, no replays, one long queue of steps."""
    x, v, m, G = ics.asteroid_belt(3000, 21)
    ref = oracle.whfast_integrate(x, v, m, G, 60.0, 60.0 * 50, "massless", 0.0, False)
    got = _run_gpu(gb, x, v, m, G, 60.0, 50, "massless", 0.0, False)
    _same(got, ref)


def test_resident_whfast_snapshot_convention(gb, oracle, ics):
    """snapshot=True: v kicked back by -dt/2 through a second Jacobi->Cartesian conversion (:346-351); the reference
    leaves system->x/v in that state, and continues from the untouched Jacobi state."""
    x, v, m, G = ics.asteroid_belt(800, 11)
    dt = 90.0
    with gb.Context() as c:
        c.set_system(x, m, G, v)
        c.whfast_begin(dt, "massless", 0.0, True)
        c.whfast_steps(dt, 6)
        half = c.whfast_state(snapshot=False)
        snap = c.whfast_state(snapshot=True)
        c.whfast_steps(dt, 4)
        end = c.whfast_state(snapshot=False)
    r6 = oracle.whfast_integrate(x, v, m, G, dt, dt * 6, "massless", 0.0, True)
    _same(half, r6)
    _same(snap, oracle.whfast_integrate(x, v, m, G, dt, dt * 6, "massless", 0.0, True, snapshot=True))
    assert np.array_equal(snap["x"], half["x"]) and not np.array_equal(snap["v"], half["v"])
    _same(end, oracle.whfast_integrate(x, v, m, G, dt, dt * 10, "massless", 0.0, True))


def test_resident_whfast_ids_and_errors(gb, oracle, ics):
    x, v, m, G = ics.asteroid_belt(300, 13)
    n = m.shape[0]
    perm = np.random.default_rng(1).permutation(n)           # shuffled input order: the Sun (id 0) is not at index 0,
    assert perm[0] != 0                                      # so the primary is found by search (src/system.c:1241-1251)
    xs, vs, ms, ids = x[perm].copy(), v[perm].copy(), m[perm].copy(), perm.astype(np.int32)
    ref = oracle.whfast_integrate(xs, vs, ms, G, 100.0, 500.0, "massless", 0.0, True, ids=ids)
    assert ref["ids"][0] == 0 and ref["m"].shape[0] == n
    got = _run_gpu(gb, xs, vs, ms, G, 100.0, 5, "massless", 0.0, True, ids=ids)
    _same(got, ref)
    with gb.Context() as c:
        c.set_system(x, m, G, v)
        with pytest.raises(gb.GravB200Error, match="Only pairwise and massless"):
            c.whfast_begin(1.0, "barnes_hut")
        with pytest.raises(gb.GravB200Error, match="Primary particle ID not found"):
            c.whfast_begin(1.0, "massless", ids=np.arange(1, n + 1, dtype=np.int32))
        with pytest.raises(gb.GravB200Error, match="whfast_begin"):
            c.whfast_steps(1.0, 1)


@pytest.mark.parametrize("skel_max_k,pair_max_k", [(4, 64), (1024, 4), (0, 0)])
def test_resident_whfast_large_k_paths(gb, oracle, ics, monkeypatch, skel_max_k, pair_max_k):
    """The paths for many massive bodies (massive list by flag/scan instead of the one-CTA ranking; massive targets and
    straddling-pair sums per thread instead of per warp / per gap), forced with the nine bodies of the solar system."""
    monkeypatch.setenv("GRAV_B200_WHFAST_SKEL_MAX_K", str(skel_max_k))
    monkeypatch.setenv("GRAV_B200_WHFAST_PAIR_MAX_K", str(pair_max_k))
    x, v, m, G = ics.asteroid_belt(700, 17, grazers=6)
    ref = oracle.whfast_integrate(x, v, m, G, 180.0, 180.0 * 9, "massless", 0.0, True)
    assert ref["m"].shape[0] < m.shape[0]
    _same(_run_gpu(gb, x, v, m, G, 180.0, 9, "massless", 0.0, True), ref)


@pytest.mark.parametrize("n", [1, 2, 3])
def test_resident_whfast_tiny_systems(gb, oracle, ics, n):
    xs, vs, ms, G = ics.solar_system()
    x, v, m = xs[:n].copy(), vs[:n].copy(), ms[:n].copy()
    for method in ("pairwise", "massless"):
        ref = oracle.whfast_integrate(x, v, m, G, 2.0, 2.0 * 5, method, 0.0, True)
        _same(_run_gpu(gb, x, v, m, G, 2.0, 5, method, 0.0, True), ref)


def test_resident_whfast_context_reuse(gb, oracle, ics):
    """One context, several integrations back to back (different method, size and step): cached step graphs and the
    massive-list cache must not leak from one run into the next."""
    runs = [(ics.asteroid_belt(40, 31), "massless", 0.0, 20.0, 9), (ics.asteroid_belt(40, 31), "pairwise", 0.01, 20.0, 9),
            (ics.asteroid_belt(900, 32), "massless", 0.0, 180.0, 7), (ics.asteroid_belt(40, 33), "massless", 0.0, 20.0, 9)]
    with gb.Context() as c:
        for (x, v, m, G), method, eps, dt, steps in runs:
            c.set_system(x, m, G, v)
            c.whfast_begin(dt, method, eps, True)
            c.whfast_steps(dt, steps)
            got = c.whfast_state()
            _same(got, oracle.whfast_integrate(x, v, m, G, dt, dt * steps, method, eps, True))
            # a massless direct-sum call on the same context afterwards uses a fresh massive list
            c.set_system(got["x"], got["m"], G, got["v"])
            c.acceleration("massless", 0.0)
            a = c.accelerations()
            assert np.array_equal(a, gb.acceleration(got["x"], got["m"], G, "massless", 0.0))


@pytest.mark.parametrize("seed", range(16))
def test_resident_whfast_fuzz(gb, oracle, seed):
    """Random small systems: 1-12 massive bodies of very different masses around a dominant primary, 0-400 massless
    ones, bound and unbound orbits, close approaches, exact duplicates of positions (distance ties), random dt and
    method.  State, order and survivors must match the oracle bit for bit.  (Runs whose oracle state contains NaN are
    skipped: the reference's qsort comparator is not an ordering for NaN distances.)"""
    rng = np.random.default_rng(1000 + seed)
    nm = int(rng.integers(1, 13)); nl = int(rng.integers(0, 401))
    method = "massless" if nl and rng.random() < 0.75 else ("pairwise" if nm + nl <= 60 else "massless")
    G = 0.00029591220828411951
    m = np.concatenate([[1.0], 10.0 ** rng.uniform(-9, -3, nm - 1), np.zeros(nl)])
    n = m.shape[0]
    a = np.concatenate([[0.0], 10.0 ** rng.uniform(-0.7, 1.5, n - 1)])
    ph = rng.uniform(0, 2 * np.pi, n); inc = rng.normal(0, 0.2, n)
    x = np.stack([a * np.cos(ph), a * np.sin(ph), a * np.sin(inc)], axis=1)
    vc = np.sqrt(G / np.maximum(a, 1e-3)) * rng.uniform(0.3, 1.6, n)      # 1.41 = escape: some are unbound
    v = np.stack([-vc * np.sin(ph), vc * np.cos(ph), vc * rng.normal(0, 0.1, n)], axis=1)
    v[0] = 0.0
    if n > 6:                                   # exact duplicates of a position: equal distances
        x[n - 1] = x[n - 2]; x[n - 3] = x[n - 2]
    order = rng.permutation(n - 1) + 1          # the primary stays first, everything else shuffled
    x[1:], v[1:], m[1:] = x[order], v[order], m[order]
    dt = float(10.0 ** rng.uniform(-1, 2.3)); steps = int(rng.integers(2, 8)); eps = float(rng.choice([0.0, 1e-3]))
    # max_steps: ceil((dt * steps) / dt) can round up to steps + 1 for a random dt (the reference then takes one more,
    # overshoot-shortened step); the comparison is about `steps` full steps
    ref = oracle.whfast_integrate(x, v, m, G, dt, dt * steps, method, eps, True, max_steps=steps)
    if not (np.isfinite(ref["x"]).all() and np.isfinite(ref["v"]).all()):
        pytest.skip("oracle state is not finite for this seed")
    _same(_run_gpu(gb, x, v, m, G, dt, steps, method, eps, True), ref)
