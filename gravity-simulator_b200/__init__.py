"""ctypes host bindings for the B200 acceleration path (product side only).

Loads the two in-tree shared libraries

* ``libgrav_b200.so``      CUDA kernels + the C ABI of ``include/grav_b200.h``
* ``libgrav_sim_b200.so``  the reference-facing C shim exporting ``acceleration()`` & co.
                           (``include/grav_sim_abi.h``; mirrors reference src/acceleration.h:33-103,
                           src/linear_octree.h:64-101)

and mirrors the reference's ctypes conventions (grav_sim/utils.py:12-58, grav_sim/simulator.py:66-102):
numpy float64 buffers passed by pointer, ``ErrorStatus`` returned by value.  Nothing here falls back
to a CPU implementation: if the libraries are missing, or no GPU is present when a compute entry is
called, a ``RuntimeError`` is raised.

The directory name contains a hyphen, so import it with :func:`importlib` (see ``tests/conftest.py``)
under the module name ``gravity_simulator_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent

METHOD_PAIRWISE, METHOD_MASSLESS, METHOD_BARNES_HUT = 1, 2, 3
METHODS = {"pairwise": 1, "massless": 2, "barnes_hut": 3}  # grav_sim/parameters.py:27-32
BH_REFERENCE, BH_FIXED = 0, 1

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_int64_p = C.POINTER(C.c_int64)


class ErrorStatus(C.Structure):  # src/error.h:31-36
    _fields_ = [("return_code", C.c_int), ("traceback", C.c_void_p), ("traceback_code_", C.c_int)]


class System(C.Structure):  # src/system.h:12-20
    _fields_ = [
        ("num_particles", C.c_int),
        ("particle_ids", c_int_p),
        ("x", c_double_p),
        ("v", c_double_p),
        ("m", c_double_p),
        ("G", C.c_double),
    ]


class AccelerationParam(C.Structure):  # src/acceleration.h:20-26
    _fields_ = [
        ("method", C.c_int),
        ("opening_angle", C.c_double),
        ("softening_length", C.c_double),
        ("max_num_particles_per_leaf", C.c_int),
    ]


class LinearOctree(C.Structure):  # src/linear_octree.h:20-57
    _fields_ = [
        ("box_width", C.c_double),
        ("num_internal_nodes", C.c_int),
        ("particle_morton_indices_deepest_level", c_int64_p),
        ("sorted_indices", c_int_p),
        ("tree_num_particles", c_int_p),
        ("tree_num_internal_children", c_int_p),
        ("tree_first_particle_sorted_idx", c_int_p),
        ("tree_first_internal_children_idx", c_int_p),
        ("tree_mass", c_double_p),
        ("tree_center_of_mass_x", c_double_p),
        ("tree_center_of_mass_y", c_double_p),
        ("tree_center_of_mass_z", c_double_p),
    ]


class GravB200Error(RuntimeError):
    pass


def _dp(a: np.ndarray):
    return a.ctypes.data_as(c_double_p)


def as_f64(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


# every symbol include/grav_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "grav_b200_last_error", "grav_b200_device_count", "grav_b200_acceleration_pairwise",
    "grav_b200_acceleration_massless", "grav_b200_acceleration_barnes_hut",
    "grav_b200_whfast_acceleration_pairwise", "grav_b200_whfast_acceleration_massless",
    "grav_b200_construct_octree", "grav_b200_morton_keys", "grav_b200_set_bh_mode", "grav_b200_get_bh_mode",
    "grav_b200_set_bh_exact", "grav_b200_get_bh_exact", "grav_b200_set_direct_sum_mode", "grav_b200_get_direct_sum_mode", "grav_b200_debug_pair_once_segments",
    "grav_b200_debug_pair_once_touches", "grav_b200_debug_pair_once_row",
    "grav_b200_ctx_create_team", "grav_b200_ctx_create_auto",
    "grav_b200_ctx_team_size",
    "grav_b200_ctx_create", "grav_b200_ctx_destroy", "grav_b200_nccl_unique_id", "grav_b200_ctx_set_system",
    "grav_b200_ctx_set_positions", "grav_b200_ctx_num_particles", "grav_b200_ctx_owned_range",
    "grav_b200_ctx_acceleration", "grav_b200_ctx_get_positions", "grav_b200_ctx_get_velocities",
    "grav_b200_ctx_get_accelerations", "grav_b200_ctx_leapfrog_begin", "grav_b200_ctx_leapfrog_steps", "grav_b200_ctx_leapfrog_end",
    "grav_b200_ctx_energy", "grav_b200_ctx_synchronize", "grav_b200_ctx_last_timing_ms",
    "grav_b200_kernel_launch_count", "grav_b200_measure_fp64_peak", "grav_b200_ctx_event_record",
    "grav_b200_ctx_event_elapsed_ms", "grav_b200_ctx_flush_l2", "grav_b200_ctx_mark_positions_sharded", "grav_b200_ctx_direct_sum_path",
    "grav_b200_host_register", "grav_b200_host_unregister",
    "grav_b200_compute_energy", "grav_b200_ctx_fixed_begin", "grav_b200_ctx_fixed_steps", "grav_b200_ctx_whfast_begin", "grav_b200_ctx_whfast_set_verbose", "grav_b200_ctx_whfast_steps", "grav_b200_ctx_whfast_get_state", "grav_b200_ctx_whfast_end",
]
SHIM_SYMBOLS = [
    "get_new_acceleration_param", "finalize_acceleration_param", "acceleration", "acceleration_barnes_hut",
    "benchmark_acceleration", "get_new_linear_octree", "construct_octree", "free_linear_octree",
    "linear_octree_check_if_included", "grav_b200_shim_whfast_acceleration",
]

_libs = None


def lib_paths():
    # GRAV_B200_LIB: alternative build of the CUDA library (kernel tuning experiments only)
    return Path(os.environ.get("GRAV_B200_LIB", HERE / "libgrav_b200.so")), HERE / "libgrav_sim_b200.so"


def load():
    """Return (abi, shim) ctypes handles; raises if the in-tree libraries are not built."""
    global _libs
    if _libs is not None:
        return _libs
    p_abi, p_shim = lib_paths()
    for p in (p_abi, p_shim):
        if not p.exists():
            raise GravB200Error(f"{p} is missing: run `make -C {HERE}` (or __graft_entry__.build()); there is no CPU fallback")
    abi = C.CDLL(str(p_abi), mode=C.RTLD_GLOBAL)
    shim = C.CDLL(str(p_shim))
    abi.grav_b200_last_error.restype = C.c_char_p
    abi.grav_b200_kernel_launch_count.restype = C.c_int64
    abi.grav_b200_ctx_destroy.restype = None
    abi.grav_b200_ctx_owned_range.restype = None
    abi.grav_b200_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p]
    abi.grav_b200_ctx_destroy.argtypes = [C.c_void_p]
    abi.grav_b200_ctx_create_team.argtypes = [C.POINTER(C.c_void_p), C.c_int, c_int_p]
    abi.grav_b200_ctx_create_auto.argtypes = [C.POINTER(C.c_void_p)]
    abi.grav_b200_ctx_team_size.argtypes = [C.c_void_p]
    abi.grav_b200_ctx_set_system.argtypes = [C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p, C.c_double]
    abi.grav_b200_ctx_set_positions.argtypes = [C.c_void_p, c_double_p]
    abi.grav_b200_ctx_num_particles.argtypes = [C.c_void_p]
    abi.grav_b200_ctx_owned_range.argtypes = [C.c_void_p, c_int_p, c_int_p]
    abi.grav_b200_ctx_acceleration.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int]
    for f in ("positions", "velocities", "accelerations"):
        getattr(abi, f"grav_b200_ctx_get_{f}").argtypes = [C.c_void_p, c_double_p]
    abi.grav_b200_ctx_leapfrog_begin.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double]
    abi.grav_b200_ctx_leapfrog_end.argtypes = [C.c_void_p]
    abi.grav_b200_ctx_leapfrog_steps.argtypes = [C.c_void_p, C.c_double, C.c_int64]
    abi.grav_b200_ctx_energy.argtypes = [C.c_void_p, c_double_p]
    abi.grav_b200_ctx_fixed_begin.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
    abi.grav_b200_ctx_fixed_steps.argtypes = [C.c_void_p, C.c_double, C.c_int64]
    abi.grav_b200_ctx_whfast_begin.argtypes = [C.c_void_p, c_int_p, C.c_int, C.c_double, C.c_double, C.c_int]
    abi.grav_b200_ctx_whfast_steps.argtypes = [C.c_void_p, C.c_double, C.c_int64]
    abi.grav_b200_ctx_whfast_get_state.argtypes = [C.c_void_p, C.c_int, c_int_p, c_int_p, c_double_p, c_double_p, c_double_p]
    abi.grav_b200_ctx_whfast_end.argtypes = [C.c_void_p]
    abi.grav_b200_ctx_whfast_set_verbose.argtypes = [C.c_void_p, C.c_int]
    abi.grav_b200_ctx_synchronize.argtypes = [C.c_void_p]
    abi.grav_b200_ctx_last_timing_ms.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
    abi.grav_b200_measure_fp64_peak.argtypes = [C.c_int, c_double_p, c_double_p]
    abi.grav_b200_ctx_event_record.argtypes = [C.c_void_p, C.c_int]
    abi.grav_b200_ctx_event_elapsed_ms.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
    abi.grav_b200_ctx_flush_l2.argtypes = [C.c_void_p]
    abi.grav_b200_ctx_mark_positions_sharded.argtypes = [C.c_void_p]
    abi.grav_b200_debug_pair_once_segments.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p]
    abi.grav_b200_debug_pair_once_touches.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    abi.grav_b200_ctx_direct_sum_path.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    abi.grav_b200_host_register.argtypes = [C.c_void_p, C.c_uint64]
    abi.grav_b200_host_unregister.argtypes = [C.c_void_p]
    abi.grav_b200_nccl_unique_id.argtypes = [C.c_void_p]
    abi.grav_b200_acceleration_pairwise.argtypes = [c_double_p, C.c_int, c_double_p, c_double_p, C.c_double, C.c_double]
    abi.grav_b200_acceleration_massless.argtypes = abi.grav_b200_acceleration_pairwise.argtypes
    abi.grav_b200_acceleration_barnes_hut.argtypes = [c_double_p, C.c_int, c_double_p, c_double_p, C.c_double, C.c_double,
                                                      C.c_double, C.c_int]
    wh = [c_double_p, C.c_int, c_double_p, c_double_p, C.c_double, c_double_p, c_double_p, C.c_double]
    abi.grav_b200_whfast_acceleration_pairwise.argtypes = wh
    abi.grav_b200_whfast_acceleration_massless.argtypes = wh
    abi.grav_b200_compute_energy.argtypes = [c_double_p, C.c_int, c_double_p, c_double_p, c_double_p, C.c_double]
    abi.grav_b200_morton_keys.argtypes = [C.c_int, c_double_p, c_int64_p, c_double_p, c_double_p]

    shim.get_new_acceleration_param.restype = AccelerationParam
    shim.get_new_linear_octree.restype = LinearOctree
    shim.free_linear_octree.restype = None
    shim.linear_octree_check_if_included.restype = C.c_bool
    shim.linear_octree_check_if_included.argtypes = [C.c_int64, C.c_int64, C.c_int]
    for name in ("finalize_acceleration_param", "acceleration", "acceleration_barnes_hut", "benchmark_acceleration",
                 "construct_octree", "grav_b200_shim_whfast_acceleration"):
        getattr(shim, name).restype = ErrorStatus
    shim.acceleration.argtypes = [c_double_p, C.POINTER(System), C.POINTER(AccelerationParam)]
    shim.acceleration_barnes_hut.argtypes = shim.acceleration.argtypes
    shim.finalize_acceleration_param.argtypes = [C.POINTER(AccelerationParam)]
    shim.construct_octree.argtypes = [C.POINTER(LinearOctree), C.POINTER(System), C.POINTER(AccelerationParam), c_double_p,
                                      C.c_double]
    shim.free_linear_octree.argtypes = [C.POINTER(LinearOctree)]
    shim.grav_b200_shim_whfast_acceleration.argtypes = [c_double_p, C.POINTER(System), c_double_p, c_double_p,
                                                        C.POINTER(AccelerationParam)]
    _libs = (abi, shim)
    return _libs


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def take_traceback(status: ErrorStatus) -> str:
    """Copy and free the malloc()ed traceback of an ErrorStatus (src/error.c:457-492 ownership rule)."""
    if not status.traceback:
        return ""
    msg = C.string_at(status.traceback).decode(errors="replace")
    _libc.free(status.traceback)
    status.traceback = None
    return msg


def check_status(status: ErrorStatus):
    if status.return_code != 0:
        raise GravB200Error(f"grav_sim error {status.return_code}: {take_traceback(status)}")


def check_rc(rc: int):
    if rc != 0:
        abi, _ = load()
        raise GravB200Error(f"grav_b200 error {rc}: {abi.grav_b200_last_error().decode(errors='replace')}")


def make_system(x: np.ndarray, m: np.ndarray, G: float, v: np.ndarray | None = None) -> System:
    """A reference `System` aliasing numpy buffers (as launch_simulation_python does, src/python_interface.c:127-168).
    The arrays must outlive the struct."""
    s = System()
    s.num_particles = int(m.shape[0])
    s.particle_ids = None
    s.x = _dp(x)
    s.v = _dp(v) if v is not None else None
    s.m = _dp(m)
    s.G = float(G)
    return s


def make_param(method, softening_length=0.0, opening_angle=1.0, max_num_particles_per_leaf=-1, finalize=True) -> AccelerationParam:
    _, shim = load()
    p = shim.get_new_acceleration_param()
    p.method = METHODS[method] if isinstance(method, str) else int(method)
    p.opening_angle = float(opening_angle)
    p.softening_length = float(softening_length)
    p.max_num_particles_per_leaf = int(max_num_particles_per_leaf)
    if finalize:
        check_status(shim.finalize_acceleration_param(C.byref(p)))
    return p


def acceleration(x, m, G, method="pairwise", softening_length=0.0, opening_angle=1.0, max_num_particles_per_leaf=-1, out=None):
    """Host-buffer force evaluation through the drop-in `acceleration()` symbol. Returns a[N,3] (written into `out` if given)."""
    _, shim = load()
    x = as_f64(x).reshape(-1, 3)
    m = as_f64(m).reshape(-1)
    a = np.empty_like(x) if out is None else out
    sys_ = make_system(x, m, G)
    prm = make_param(method, softening_length, opening_angle, max_num_particles_per_leaf)
    check_status(shim.acceleration(_dp(a), C.byref(sys_), C.byref(prm)))
    return a


def whfast_acceleration(x, m, G, jacobi_x, eta, method="pairwise", softening_length=0.0, a0=None):
    _, shim = load()
    x = as_f64(x).reshape(-1, 3)
    m = as_f64(m).reshape(-1)
    jx = as_f64(jacobi_x).reshape(-1, 3)
    eta = as_f64(eta).reshape(-1)
    a = np.zeros_like(x) if a0 is None else as_f64(a0).reshape(-1, 3).copy()
    sys_ = make_system(x, m, G)
    prm = make_param(method, softening_length)
    check_status(shim.grav_b200_shim_whfast_acceleration(_dp(a), C.byref(sys_), _dp(jx), _dp(eta), C.byref(prm)))
    return a


def construct_octree(x, m, max_num_particles_per_leaf=1, box_center=None, box_width=-1.0):
    """Build the linear octree through the drop-in `construct_octree()`; returns a dict of numpy arrays."""
    _, shim = load()
    x = as_f64(x).reshape(-1, 3)
    m = as_f64(m).reshape(-1)
    n = m.shape[0]
    sys_ = make_system(x, m, 1.0)
    prm = make_param("barnes_hut", 0.0, 1.0, max_num_particles_per_leaf)
    tree = shim.get_new_linear_octree()
    bc = _dp(as_f64(box_center)) if box_center is not None else None
    check_status(shim.construct_octree(C.byref(tree), C.byref(sys_), C.byref(prm), bc, float(box_width)))
    try:
        out = tree_to_dict(tree, n)
    finally:
        shim.free_linear_octree(C.byref(tree))
    return out


def tree_to_dict(tree: LinearOctree, n: int) -> dict:
    M = tree.num_internal_nodes
    f = np.ctypeslib.as_array
    return {
        "box_width": float(tree.box_width),
        "num_nodes": int(M),
        "keys": f(tree.particle_morton_indices_deepest_level, (n,)).copy(),
        "sorted_indices": f(tree.sorted_indices, (n,)).copy(),
        "num_particles": f(tree.tree_num_particles, (M,)).copy(),
        "num_children": f(tree.tree_num_internal_children, (M,)).copy(),
        "first_particle": f(tree.tree_first_particle_sorted_idx, (M,)).copy(),
        "first_child": f(tree.tree_first_internal_children_idx, (M,)).copy(),
        "mass": f(tree.tree_mass, (M,)).copy(),
        "com_x": f(tree.tree_center_of_mass_x, (M,)).copy(),
        "com_y": f(tree.tree_center_of_mass_y, (M,)).copy(),
        "com_z": f(tree.tree_center_of_mass_z, (M,)).copy(),
    }


def morton_keys(x):
    abi, _ = load()
    x = as_f64(x).reshape(-1, 3)
    keys = np.empty(x.shape[0], dtype=np.int64)
    center = np.empty(3)
    width = np.empty(1)
    check_rc(abi.grav_b200_morton_keys(x.shape[0], _dp(x), keys.ctypes.data_as(c_int64_p), _dp(center), _dp(width)))
    return keys, center, float(width[0])


def compute_energy(x, v, m, G) -> float:
    """compute_energy (src/utils.c:27-59) on the GPU through the one-shot C ABI."""
    abi, _ = load()
    x = as_f64(x).reshape(-1, 3); v = as_f64(v).reshape(-1, 3); m = as_f64(m).reshape(-1)
    e = C.c_double()
    check_rc(abi.grav_b200_compute_energy(C.byref(e), m.shape[0], _dp(x), _dp(v), _dp(m), float(G)))
    return e.value


def device_count() -> int:
    abi, _ = load()
    return int(abi.grav_b200_device_count())


class Context:
    """Device-resident particle state (include/grav_b200.h, family 2)."""

    def __init__(self, device=0, rank=0, world_size=1, nccl_unique_id: bytes | None = None, team=None):
        """team: list of device ids (or a count) -> an in-process device team led by this object (grav_b200_ctx_create_team)."""
        self.abi, _ = load()
        self.h = C.c_void_p()
        if team is not None:
            devs = list(range(team)) if isinstance(team, int) else list(team)
            arr = (C.c_int * len(devs))(*devs)
            check_rc(self.abi.grav_b200_ctx_create_team(C.byref(self.h), len(devs), arr))
        else:
            uid = C.create_string_buffer(nccl_unique_id, 128) if nccl_unique_id is not None else None
            check_rc(self.abi.grav_b200_ctx_create(C.byref(self.h), device, rank, world_size, uid))
        self.n = 0

    def team_size(self) -> int:
        return int(self.abi.grav_b200_ctx_team_size(self.h))

    @staticmethod
    def new_unique_id() -> bytes:
        abi, _ = load()
        buf = C.create_string_buffer(128)
        check_rc(abi.grav_b200_nccl_unique_id(buf))
        return buf.raw

    def close(self):
        if self.h:
            self.abi.grav_b200_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_system(self, x, m, G, v=None):
        x = as_f64(x).reshape(-1, 3)
        m = as_f64(m).reshape(-1)
        vv = as_f64(v).reshape(-1, 3) if v is not None else None
        self.n = m.shape[0]
        check_rc(self.abi.grav_b200_ctx_set_system(self.h, self.n, _dp(x), _dp(vv) if vv is not None else None, _dp(m), float(G)))

    def set_positions(self, x):
        x = as_f64(x).reshape(-1, 3)
        check_rc(self.abi.grav_b200_ctx_set_positions(self.h, _dp(x)))

    def owned_range(self):
        lo, hi = C.c_int(), C.c_int()
        self.abi.grav_b200_ctx_owned_range(self.h, C.byref(lo), C.byref(hi))
        return lo.value, hi.value

    def acceleration(self, method="pairwise", softening_length=0.0, opening_angle=1.0, max_num_particles_per_leaf=-1):
        meth = METHODS[method] if isinstance(method, str) else int(method)
        check_rc(self.abi.grav_b200_ctx_acceleration(self.h, meth, float(softening_length), float(opening_angle),
                                                     int(max_num_particles_per_leaf)))

    def _get(self, what, out=None):
        a = np.empty((self.n, 3)) if out is None else out
        check_rc(getattr(self.abi, f"grav_b200_ctx_get_{what}")(self.h, _dp(a)))
        return a

    def positions(self, out=None):
        return self._get("positions", out)

    def velocities(self, out=None):
        return self._get("velocities", out)

    def accelerations(self, out=None):
        return self._get("accelerations", out)

    def leapfrog_begin(self, dt, method="pairwise", softening_length=0.0, opening_angle=1.0, max_num_particles_per_leaf=-1):
        meth = METHODS[method] if isinstance(method, str) else int(method)
        check_rc(self.abi.grav_b200_ctx_leapfrog_begin(self.h, meth, float(softening_length), float(opening_angle),
                                                       int(max_num_particles_per_leaf), float(dt)))

    def leapfrog_end(self):
        check_rc(self.abi.grav_b200_ctx_leapfrog_end(self.h))

    def leapfrog_steps(self, dt, num_steps):
        check_rc(self.abi.grav_b200_ctx_leapfrog_steps(self.h, float(dt), int(num_steps)))

    # Euler / Euler-Cromer / RK4 on the resident state (src/integrator.c:281-892)
    def fixed_begin(self, integrator, method="pairwise", softening_length=0.0, opening_angle=1.0, max_num_particles_per_leaf=-1):
        code = {"euler": 1, "euler_cromer": 2, "rk4": 3}[integrator] if isinstance(integrator, str) else int(integrator)
        meth = METHODS[method] if isinstance(method, str) else int(method)
        check_rc(self.abi.grav_b200_ctx_fixed_begin(self.h, code, meth, float(softening_length), float(opening_angle),
                                                    int(max_num_particles_per_leaf)))

    def fixed_steps(self, dt, num_steps):
        check_rc(self.abi.grav_b200_ctx_fixed_steps(self.h, float(dt), int(num_steps)))

    # device-resident WHFast (src/integrator_whfast.c:200-407)
    def whfast_begin(self, dt, method="massless", softening_length=0.0, remove_invalid_particles=True, ids=None):
        meth = METHODS[method] if isinstance(method, str) else int(method)
        pid = None
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.int32)
            pid = ids.ctypes.data_as(c_int_p)
        check_rc(self.abi.grav_b200_ctx_whfast_begin(self.h, pid, meth, float(softening_length), float(dt),
                                                     int(bool(remove_invalid_particles))))

    def whfast_steps(self, dt, num_steps):
        check_rc(self.abi.grav_b200_ctx_whfast_steps(self.h, float(dt), int(num_steps)))
        self.n = int(self.abi.grav_b200_ctx_num_particles(self.h))

    def whfast_state(self, snapshot=False):
        """dict(x, v, m, ids) in the current particle order; snapshot=True applies the reference's output convention."""
        n = int(self.abi.grav_b200_ctx_num_particles(self.h))
        x = np.empty((n, 3)); v = np.empty((n, 3)); m = np.empty(n); ids = np.empty(n, dtype=np.int32)
        nn = C.c_int()
        check_rc(self.abi.grav_b200_ctx_whfast_get_state(self.h, int(bool(snapshot)), C.byref(nn), ids.ctypes.data_as(c_int_p),
                                                         _dp(x), _dp(v), _dp(m)))
        assert nn.value == n
        return {"x": x, "v": v, "m": m, "ids": ids}

    def whfast_set_verbose(self, level):
        check_rc(self.abi.grav_b200_ctx_whfast_set_verbose(self.h, int(level)))

    def whfast_end(self):
        check_rc(self.abi.grav_b200_ctx_whfast_end(self.h))

    def energy(self) -> float:
        e = C.c_double()
        check_rc(self.abi.grav_b200_ctx_energy(self.h, C.byref(e)))
        return e.value

    def synchronize(self):
        check_rc(self.abi.grav_b200_ctx_synchronize(self.h))

    def event_record(self, slot):
        check_rc(self.abi.grav_b200_ctx_event_record(self.h, slot))

    def event_elapsed_ms(self, a, b) -> float:
        ms = C.c_float()
        check_rc(self.abi.grav_b200_ctx_event_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    def flush_l2(self):
        check_rc(self.abi.grav_b200_ctx_flush_l2(self.h))

    def mark_positions_sharded(self):
        check_rc(self.abi.grav_b200_ctx_mark_positions_sharded(self.h))

    def direct_sum_path(self):
        """(pair_once, equal_mass) of the last pairwise force evaluation."""
        a, b = C.c_int(0), C.c_int(0)
        check_rc(self.abi.grav_b200_ctx_direct_sum_path(self.h, C.byref(a), C.byref(b)))
        return bool(a.value), bool(b.value)

    def timing_ms(self, stage=0) -> float:
        ms = C.c_float()
        check_rc(self.abi.grav_b200_ctx_last_timing_ms(self.h, stage, C.byref(ms)))
        return ms.value


def measure_fp64_peak(device=0):
    abi, _ = load()
    tf, mhz = C.c_double(), C.c_double()
    check_rc(abi.grav_b200_measure_fp64_peak(device, C.byref(tf), C.byref(mhz)))
    return tf.value, mhz.value


def host_register(arr: np.ndarray):
    abi, _ = load()
    check_rc(abi.grav_b200_host_register(arr.ctypes.data, arr.nbytes))


def host_unregister(arr: np.ndarray):
    abi, _ = load()
    check_rc(abi.grav_b200_host_unregister(arr.ctypes.data))


def kernel_launch_count() -> int:
    abi, _ = load()
    return int(abi.grav_b200_kernel_launch_count())
