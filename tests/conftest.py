import importlib.util
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def load_package():
    """Import gravity-simulator_b200/ (hyphenated directory) as module `gravity_simulator_b200`."""
    name = "gravity_simulator_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg = ROOT / "gravity-simulator_b200"
    spec = importlib.util.spec_from_file_location(name, pkg / "__init__.py", submodule_search_locations=[str(pkg)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "cooperative_walk: run with the default Barnes-Hut walk in a module that otherwise uses the exact one")


@pytest.fixture(scope="session")
def gb():
    return load_package()


@pytest.fixture(scope="session")
def ics(gb):
    import importlib
    return importlib.import_module("gravity_simulator_b200.ics")


@pytest.fixture(scope="session")
def oracle():
    from oracle.bind import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle.bind import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libgrav_sim_ref.so not built (needs /root/reference at build time)")
    return Reference()


GOLDEN = ROOT / "tests" / "golden"


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(GOLDEN / f"{name}.npz")
    return load


def max_rel_err(a, ref):
    """max_i |a_i - ref_i|_2 / |ref_i|_2 (SURVEY.md section 8d parity gate)."""
    import numpy as np
    num = np.linalg.norm(a - ref, axis=1)
    den = np.linalg.norm(ref, axis=1)
    den = np.where(den > 0, den, 1.0)
    return float(np.max(num / den))


def assert_forces_close(a, ref, tol=1e-12, ctx=""):
    """Parity gate for accelerations whose summation order differs from the reference: the same
    NaN / inf rows as the reference, and max_i |a_i - ref_i|_2 / |ref_i|_2 <= tol on the finite ones."""
    import numpy as np
    fin = np.isfinite(ref).all(axis=1)
    assert np.array_equal(np.isfinite(a).all(axis=1), fin), (ctx, "finite pattern differs")
    if fin.any():
        err = max_rel_err(a[fin], ref[fin])
        assert err <= tol, (ctx, err)


class bh_exact:
    """with bh_exact(gb, 1): ...   -- Barnes-Hut walk arithmetic for the one-shot entries and contexts created inside
    (include/grav_b200.h, grav_b200_set_bh_exact): 1 = bit-identical per-lane walk, 0 = warp-cooperative walk (default)."""

    def __init__(self, gb, on):
        self.abi, _ = gb.load()
        self.on = int(on)

    def __enter__(self):
        self.prev = int(self.abi.grav_b200_get_bh_exact())
        self.abi.grav_b200_set_bh_exact(self.on)
        return self

    def __exit__(self, *exc):
        self.abi.grav_b200_set_bh_exact(self.prev)


class ds_mode:
    """with ds_mode(gb, 1): ...   -- formulation of the pairwise direct sum (include/grav_b200.h,
    grav_b200_set_direct_sum_mode): 1 = every unordered pair once wherever the system has two rows, 0 = ordered
    interactions always, -1 = automatic (default)."""

    def __init__(self, gb, mode):
        self.abi, _ = gb.load()
        self.mode = int(mode)

    def __enter__(self):
        self.prev = int(self.abi.grav_b200_get_direct_sum_mode())
        self.abi.grav_b200_set_direct_sum_mode(self.mode)
        return self

    def __exit__(self, *exc):
        self.abi.grav_b200_set_direct_sum_mode(self.prev)
