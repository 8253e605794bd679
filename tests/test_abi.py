"""CPU tests: the C-ABI libraries load and export every symbol the headers declare; host-side validation
logic of the shim behaves like the reference; compute entries fail loudly without a GPU."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols(header: Path, pattern: str):
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    return sorted(set(re.findall(pattern, text)))


def test_abi_header_symbols_exported(gb):
    abi, shim = gb.load()
    names = declared_symbols(ROOT / "include" / "grav_b200.h", r"\b(grav_b200_[a-z0-9_]+)\s*\(")
    assert set(names) == set(gb.ABI_SYMBOLS), set(names) ^ set(gb.ABI_SYMBOLS)
    for n in names:
        assert hasattr(abi, n), n


def test_shim_exports_reference_symbols(gb):
    _, shim = gb.load()
    for n in gb.SHIM_SYMBOLS:
        assert hasattr(shim, n), n


def test_struct_layouts_match_reference_abi(gb):
    # x86-64 SysV layouts of the reference structs
    assert C.sizeof(gb.AccelerationParam) == 32
    assert gb.AccelerationParam.opening_angle.offset == 8 and gb.AccelerationParam.max_num_particles_per_leaf.offset == 24
    assert C.sizeof(gb.ErrorStatus) == 24
    assert C.sizeof(gb.System) == 48 and gb.System.G.offset == 40
    assert C.sizeof(gb.LinearOctree) == 96


def test_param_defaults_and_validation(gb):
    _, shim = gb.load()
    p = shim.get_new_acceleration_param()
    assert (p.method, p.opening_angle, p.softening_length, p.max_num_particles_per_leaf) == (1, 1.0, 0.0, -1)
    p = gb.make_param("barnes_hut")
    assert p.max_num_particles_per_leaf == 1          # -1 -> 1 for Barnes-Hut only
    assert gb.make_param("pairwise").max_num_particles_per_leaf == -1
    for kw, msg in [(dict(method=7), "Unknown acceleration method"),
                    (dict(method="pairwise", softening_length=-1.0), "Softening length is negative"),
                    (dict(method="barnes_hut", opening_angle=-0.5), "Opening angle is negative"),
                    (dict(method="barnes_hut", max_num_particles_per_leaf=0), "must be positive")]:
        with pytest.raises(gb.GravB200Error, match=msg):
            gb.make_param(**kw)
    # a negative opening angle is only an error for Barnes-Hut (src/acceleration.c:97-108)
    gb.make_param("pairwise", opening_angle=-1.0)


def test_unknown_method_in_dispatch(gb):
    _, shim = gb.load()
    x = np.zeros((2, 3)); m = np.ones(2); a = np.zeros((2, 3))
    s = gb.make_system(x, m, 1.0)
    p = gb.make_param("pairwise")
    p.method = 42
    st = shim.acceleration(a.ctypes.data_as(gb.c_double_p), C.byref(s), C.byref(p))
    assert st.return_code == 2 and "Unknown acceleration method. Got: 42" in gb.take_traceback(st)


def test_check_if_included(gb):
    _, shim = gb.load()
    a = 0b101_011_000 << (3 * 18)
    b = 0b101_011_111 << (3 * 18)
    assert shim.linear_octree_check_if_included(a, b, 2) and not shim.linear_octree_check_if_included(a, b, 3)
    assert shim.linear_octree_check_if_included(a, b, 0)


def test_no_cpu_fallback(gb):
    """Without a GPU every compute entry must fail loudly (GRAV_FAILURE + device message), never compute on the host."""
    if gb.device_count() > 0:
        pytest.skip("a GPU is present")
    x = np.random.default_rng(0).normal(size=(16, 3)); m = np.ones(16)
    for method in ("pairwise", "massless", "barnes_hut"):
        with pytest.raises(gb.GravB200Error, match="no CUDA device|CUDA error"):
            gb.acceleration(x, m, 1.0, method)
    with pytest.raises(gb.GravB200Error):
        gb.Context()
    with pytest.raises(gb.GravB200Error):
        gb.construct_octree(x, m)


@pytest.mark.parametrize("n", [512, 513, 777, 1000, 2049, 4099, 8192, 20001, 65536])
@pytest.mark.parametrize("ctas", [1, 7, 148, 296, 1184])
def test_pair_once_decomposition_covers_every_pair_once(gb, n, ctas):
    """Host code of direct_sum_sym.cu through its test hooks (no GPU): over all CTAs of the decomposition (one GPU: 148;
    8 GPUs: 1184) every (row, group beyond the row) unit is handed out exactly once, i.e. every unordered pair of particles
    in different rows is evaluated exactly once (pairs inside a row are the finishing kernel's diagonal blocks); and the
    finishing kernel's contributor test (sym_cta_touches) names exactly the CTAs that added something to a block, so the
    private arrays are read completely and left zeroed."""
    abi, _ = gb.load()
    row = abi.grav_b200_debug_pair_once_row()
    gpr = row // 32
    NR, NG = -(-n // row), -(-n // 32)
    cap = 4096
    rows = (C.c_int * cap)(); gb_ = (C.c_int * cap)(); ge = (C.c_int * cap)()
    seen = np.zeros((NR, NG), dtype=np.int32)
    touched = np.zeros((ctas, NR), dtype=bool)
    for c in range(ctas):
        k = abi.grav_b200_debug_pair_once_segments(n, ctas, c, cap, rows, gb_, ge)
        assert 0 <= k <= cap
        for s in range(k):
            A, g0, g1 = rows[s], gb_[s], ge[s]
            assert 0 <= A < NR - 1 and gpr * (A + 1) <= g0 < g1 <= NG
            seen[A, g0:g1] += 1
            touched[c, A] = True
            touched[c, g0 // gpr:(g1 - 1) // gpr + 1] = True
    expect = np.zeros_like(seen)
    for A in range(NR - 1):
        expect[A, gpr * (A + 1):] = 1
    assert np.array_equal(seen, expect)
    claimed = np.array([[abi.grav_b200_debug_pair_once_touches(n, ctas, c, b) for b in range(NR)] for c in range(ctas if ctas <= 296 else 64)])
    assert np.array_equal(claimed.astype(bool), touched[:claimed.shape[0]])


def test_zeroed_malloc_for_the_reference_whfast():
    """oracle/bind.py zeroed_malloc: while it is active glibc hands out zero-filled blocks (what the reference's whfast()
    implicitly relies on for a[0])."""
    from oracle.bind import zeroed_malloc
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.free.argtypes = [C.c_void_p]
    blocks = []
    for _ in range(50):                      # dirty a few free chunks first
        p = libc.malloc(12216)
        C.memset(p, 0x5A, 12216)
        blocks.append(p)
    for p in blocks:
        libc.free(p)
    with zeroed_malloc():
        p = libc.malloc(12216)
        buf = (C.c_ubyte * 12216).from_address(p)
        assert not any(buf)
        libc.free(p)
