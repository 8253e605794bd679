"""ctypes bindings for the CPU oracle.  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never by the product package.  Two libraries:

* ``oracle/libgrav_oracle.so``        our C restatement (grav_oracle.c)                     -> :class:`Oracle`
* ``oracle/_ref/libgrav_sim_ref.so``  the UNMODIFIED reference compiled by oracle/Makefile  -> :class:`Reference`
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_SO = HERE / "libgrav_oracle.so"
REF_SO = HERE / "_ref" / "libgrav_sim_ref.so"
PROBE_SO = HERE / "_ref" / "libwhfast_probe.so"

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
i64p = C.POINTER(C.c_int64)


def _d(a):
    return a.ctypes.data_as(dp)


def _f64(a, shape):
    return np.ascontiguousarray(a, dtype=np.float64).reshape(shape)


def build(force=False):
    """Compile the restatement (always possible: gcc only) and, where /root/reference exists, the reference."""
    if force or not ORACLE_SO.exists() or ORACLE_SO.stat().st_mtime < (HERE / "grav_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "libgrav_oracle.so"], check=True, capture_output=True)
    if not REF_SO.exists() and Path("/root/reference/src").is_dir():
        subprocess.run(["make", "-C", str(HERE), "ref"], check=True, capture_output=True)


class OracleTree(C.Structure):
    _fields_ = [("box_width", C.c_double), ("n", C.c_int), ("num_nodes", C.c_int), ("keys", i64p), ("perm", ip),
                ("num_particles", ip), ("num_children", ip), ("first_particle", ip), ("first_child", ip),
                ("mass", dp), ("com_x", dp), ("com_y", dp), ("com_z", dp)]


def _tree_dict(keys, perm, npart, nch, first, fc, mass, cx, cy, cz, n, M, box_width):
    f = np.ctypeslib.as_array
    return {"box_width": float(box_width), "num_nodes": int(M), "keys": f(keys, (n,)).copy(),
            "sorted_indices": f(perm, (n,)).copy(), "num_particles": f(npart, (M,)).copy(),
            "num_children": f(nch, (M,)).copy(), "first_particle": f(first, (M,)).copy(),
            "first_child": f(fc, (M,)).copy(), "mass": f(mass, (M,)).copy(), "com_x": f(cx, (M,)).copy(),
            "com_y": f(cy, (M,)).copy(), "com_z": f(cz, (M,)).copy()}


class Oracle:
    def __init__(self):
        build()
        L = self.L = C.CDLL(str(ORACLE_SO))
        L.oracle_pairwise.argtypes = [dp, C.c_int, dp, dp, C.c_double, C.c_double]
        L.oracle_pairwise.restype = None
        L.oracle_massless.argtypes = L.oracle_pairwise.argtypes
        L.oracle_whfast_pairwise.argtypes = [dp, C.c_int, dp, dp, C.c_double, dp, dp, C.c_double]
        L.oracle_whfast_pairwise.restype = None
        L.oracle_whfast_massless.argtypes = L.oracle_whfast_pairwise.argtypes
        L.oracle_bounding_box.argtypes = [dp, dp, C.c_int, dp]
        L.oracle_bounding_box.restype = None
        L.oracle_morton_keys.argtypes = [i64p, C.c_int, dp, dp, C.c_double]
        L.oracle_morton_keys.restype = None
        L.oracle_sort_keys.argtypes = [i64p, ip, C.c_int]
        L.oracle_build_tree.argtypes = [C.POINTER(OracleTree), C.c_int, dp, dp, C.c_int, dp, C.c_double]
        L.oracle_free_tree.argtypes = [C.POINTER(OracleTree)]
        L.oracle_free_tree.restype = None
        L.oracle_barnes_hut.argtypes = [dp, C.c_int, dp, dp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
        L.oracle_energy.argtypes = [C.c_int, dp, dp, dp, C.c_double]
        L.oracle_energy.restype = C.c_double
        for f in (L.oracle_pairwise_targets, L.oracle_pairwise_targets_ld):
            f.argtypes = [dp, C.c_int, ip, C.c_int, dp, dp, C.c_double, C.c_double]
            f.restype = None
        L.oracle_bh_walk_targets.argtypes = [dp, C.POINTER(C.c_longlong), C.POINTER(OracleTree), dp, dp, C.c_double,
                                             C.c_double, C.c_double, C.c_int, C.c_int, ip]
        L.oracle_bh_walk_targets.restype = None
        L.oracle_sort_by_distance.argtypes = [C.c_int, ip, dp, dp, dp, C.c_int]
        L.oracle_whfast_eta.argtypes = [dp, C.c_int, dp]
        L.oracle_whfast_eta.restype = None
        L.oracle_cartesian_to_jacobi.argtypes = [dp, dp, C.c_int, dp, dp, dp, dp]
        L.oracle_cartesian_to_jacobi.restype = None
        L.oracle_jacobi_to_cartesian.argtypes = [dp, dp, C.c_int, dp, dp, dp, dp]
        L.oracle_jacobi_to_cartesian.restype = None
        L.oracle_stumpff.argtypes = [dp, C.c_double]
        L.oracle_stumpff.restype = None
        L.oracle_whfast_drift.argtypes = [C.c_int, dp, dp, ip, dp, dp, dp, dp, C.c_double, C.c_double, C.c_int]
        L.oracle_whfast_integrate.argtypes = [C.c_int, ip, dp, dp, dp, C.c_double, C.c_double, C.c_double, C.c_int,
                                              C.c_double, C.c_int, C.c_int64, C.c_int, dp]

    def acceleration(self, x, m, G, method="pairwise", softening_length=0.0, opening_angle=1.0,
                     max_num_particles_per_leaf=1, fixed=False):
        x = _f64(x, (-1, 3)); m = _f64(m, (-1,)); n = m.shape[0]
        a = np.zeros_like(x)
        if method == "pairwise":
            self.L.oracle_pairwise(_d(a), n, _d(x), _d(m), G, softening_length)
        elif method == "massless":
            assert self.L.oracle_massless(_d(a), n, _d(x), _d(m), G, softening_length) == 0
        elif method == "barnes_hut":
            leaf = 1 if max_num_particles_per_leaf == -1 else max_num_particles_per_leaf
            assert self.L.oracle_barnes_hut(_d(a), n, _d(x), _d(m), G, softening_length, opening_angle, leaf, int(fixed)) == 0
        else:
            raise ValueError(method)
        return a

    def whfast_acceleration(self, x, m, G, jacobi_x, eta, method="pairwise", softening_length=0.0, a0=None):
        x = _f64(x, (-1, 3)); m = _f64(m, (-1,)); jx = _f64(jacobi_x, (-1, 3)); eta = _f64(eta, (-1,))
        a = np.zeros_like(x) if a0 is None else _f64(a0, (-1, 3)).copy()
        fn = self.L.oracle_whfast_pairwise if method == "pairwise" else self.L.oracle_whfast_massless
        fn(_d(a), m.shape[0], _d(x), _d(m), G, _d(jx), _d(eta), softening_length)
        return a

    def whfast_integrate(self, x, v, m, G, dt, tf, method="massless", softening_length=0.0, remove_invalid=True,
                         max_steps=-1, ids=None, snapshot=False):
        """whfast() with output disabled (src/integrator_whfast.c:200-407): returns dict(x, v, m, ids, a) as the
        reference leaves them (v at the half step, particles in the last distance-sorted order)."""
        x = _f64(x, (-1, 3)).copy(); v = _f64(v, (-1, 3)).copy(); m = _f64(m, (-1,)).copy(); n = m.shape[0]
        ids = np.arange(n, dtype=np.int32) if ids is None else np.ascontiguousarray(ids, dtype=np.int32).copy()
        a = np.zeros_like(x)
        n2 = self.L.oracle_whfast_integrate(n, ids.ctypes.data_as(ip), _d(x), _d(v), _d(m), G, dt, tf, METHODS[method],
                                            softening_length, int(remove_invalid), max_steps, int(snapshot), _d(a))
        if n2 < 0:
            raise RuntimeError(f"oracle_whfast_integrate -> {n2}")
        return {"x": x[:n2], "v": v[:n2], "m": m[:n2], "ids": ids[:n2], "a": a[:n2]}

    def whfast_stages(self, x, v, m, G, dt):
        """Same sequence as Reference.whfast_stages with the restated functions."""
        x = _f64(x, (-1, 3)).copy(); v = _f64(v, (-1, 3)).copy(); m = _f64(m, (-1,)).copy(); n = m.shape[0]
        eta, jx, jv = self.jacobi_state(x, v, m)
        out = {"eta": eta.copy(), "jx0": jx.copy(), "jv0": jv.copy()}
        ids = np.arange(n, dtype=np.int32)
        assert self.L.oracle_whfast_drift(n, _d(jx), _d(jv), ids.ctypes.data_as(ip), _d(x), _d(v), _d(m), _d(eta), G, dt, 0) == n
        out["jx1"], out["jv1"] = jx.copy(), jv.copy()
        x2 = np.empty_like(x); v2 = np.empty_like(v)
        self.L.oracle_jacobi_to_cartesian(_d(x2), _d(v2), n, _d(jx), _d(jv), _d(m), _d(eta))
        out["x1"], out["v1"] = x2, v2
        return out

    def stumpff(self, z):
        c = np.empty(4)
        self.L.oracle_stumpff(_d(c), float(z))
        return c

    def jacobi_state(self, x, v, m):
        """eta, jacobi_x, jacobi_v (src/integrator_whfast.c:1266-1279, :682-724)."""
        x = _f64(x, (-1, 3)); v = _f64(v, (-1, 3)); m = _f64(m, (-1,)); n = m.shape[0]
        eta = np.empty(n); jx = np.zeros_like(x); jv = np.zeros_like(v)
        self.L.oracle_whfast_eta(_d(eta), n, _d(m))
        self.L.oracle_cartesian_to_jacobi(_d(jx), _d(jv), n, _d(x), _d(v), _d(m), _d(eta))
        return eta, jx, jv

    def morton_keys(self, x):
        x = _f64(x, (-1, 3)); n = x.shape[0]
        c = np.empty(3); w = np.empty(1)
        self.L.oracle_bounding_box(_d(c), _d(w), n, _d(x))
        keys = np.empty(n, dtype=np.int64)
        self.L.oracle_morton_keys(keys.ctypes.data_as(i64p), n, _d(x), _d(c), float(w[0]))
        return keys, c, float(w[0])

    def construct_octree(self, x, m, max_num_particles_per_leaf=1, box_center=None, box_width=-1.0):
        x = _f64(x, (-1, 3)); m = _f64(m, (-1,)); n = m.shape[0]
        t = OracleTree()
        bc = _d(_f64(box_center, (3,))) if box_center is not None else None
        assert self.L.oracle_build_tree(C.byref(t), n, _d(x), _d(m), max_num_particles_per_leaf, bc, box_width) == 0
        try:
            return _tree_dict(t.keys, t.perm, t.num_particles, t.num_children, t.first_particle, t.first_child,
                              t.mass, t.com_x, t.com_y, t.com_z, n, t.num_nodes, t.box_width)
        finally:
            self.L.oracle_free_tree(C.byref(t))

    def energy(self, x, v, m, G):
        x = _f64(x, (-1, 3)); v = _f64(v, (-1, 3)); m = _f64(m, (-1,))
        return float(self.L.oracle_energy(m.shape[0], _d(x), _d(v), _d(m), G))

    def pairwise_targets(self, x, m, G, softening_length, targets, long_double=False):
        """Rows `targets` of the pairwise direct sum: the reference's operation order per target (bit-identical to
        acceleration(..., "pairwise")[targets]), or the same sums in long double (the rounding-noise yardstick)."""
        x = _f64(x, (-1, 3)); m = _f64(m, (-1,))
        t = np.ascontiguousarray(targets, dtype=np.int32)
        a = np.empty((t.shape[0], 3))
        fn = self.L.oracle_pairwise_targets_ld if long_double else self.L.oracle_pairwise_targets
        fn(_d(a), t.shape[0], t.ctypes.data_as(ip), m.shape[0], _d(x), _d(m), G, softening_length)
        return a

    def tree(self, x, m, max_num_particles_per_leaf=1):
        """A built tree kept alive for several sampled walks (free it with .close() or use as a context manager)."""
        return OracleTreeHandle(self, x, m, max_num_particles_per_leaf)


class OracleTreeHandle:
    def __init__(self, oracle, x, m, leaf):
        self.O = oracle
        self.x = _f64(x, (-1, 3)); self.m = _f64(m, (-1,)); self.n = self.m.shape[0]
        self.t = OracleTree()
        assert oracle.L.oracle_build_tree(C.byref(self.t), self.n, _d(self.x), _d(self.m), leaf, None, -1.0) == 0

    def to_dict(self):
        t = self.t
        return _tree_dict(t.keys, t.perm, t.num_particles, t.num_children, t.first_particle, t.first_child,
                          t.mass, t.com_x, t.com_y, t.com_z, self.n, t.num_nodes, t.box_width)

    def walk_targets(self, G, softening_length, opening_angle, positions, fixed=False, stats=False):
        """Accelerations of the particles at the given SORTED positions (row k belongs to particle
        sorted_indices[positions[k]]); stats=True also returns (visits, accepts, opened, leaf particles) per target."""
        pos = np.ascontiguousarray(positions, dtype=np.int32)
        a = np.empty((pos.shape[0], 3))
        st = np.zeros((pos.shape[0], 4), dtype=np.int64) if stats else None
        self.O.L.oracle_bh_walk_targets(_d(a), st.ctypes.data_as(C.POINTER(C.c_longlong)) if stats else None, C.byref(self.t),
                                        _d(self.x), _d(self.m), G, softening_length, opening_angle, int(fixed), pos.shape[0],
                                        pos.ctypes.data_as(ip))
        return (a, st) if stats else a

    def close(self):
        if self.t.keys:
            self.O.L.oracle_free_tree(C.byref(self.t))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


# ---- the unmodified reference ---------------------------------------------------------------------------

class RefErrorStatus(C.Structure):
    _fields_ = [("return_code", C.c_int), ("traceback", C.c_char_p), ("traceback_code_", C.c_int)]


class RefSystem(C.Structure):
    _fields_ = [("num_particles", C.c_int), ("particle_ids", ip), ("x", dp), ("v", dp), ("m", dp), ("G", C.c_double)]


class RefAccelerationParam(C.Structure):
    _fields_ = [("method", C.c_int), ("opening_angle", C.c_double), ("softening_length", C.c_double),
                ("max_num_particles_per_leaf", C.c_int)]


class RefLinearOctree(C.Structure):
    _fields_ = [("box_width", C.c_double), ("num_internal_nodes", C.c_int), ("keys", i64p), ("sorted_indices", ip),
                ("tree_num_particles", ip), ("tree_num_internal_children", ip), ("tree_first_particle_sorted_idx", ip),
                ("tree_first_internal_children_idx", ip), ("tree_mass", dp), ("com_x", dp), ("com_y", dp), ("com_z", dp)]


METHODS = {"pairwise": 1, "massless": 2, "barnes_hut": 3}


class Reference:
    """The compiled upstream library, driven through its own public C API."""

    @staticmethod
    def available() -> bool:
        build()
        return REF_SO.exists()

    def __init__(self, path: Path | None = None):
        build()
        L = self.L = C.CDLL(str(path or REF_SO))
        L.acceleration.restype = RefErrorStatus
        L.acceleration.argtypes = [dp, C.POINTER(RefSystem), C.POINTER(RefAccelerationParam)]
        L.construct_octree.restype = RefErrorStatus
        L.construct_octree.argtypes = [C.POINTER(RefLinearOctree), C.POINTER(RefSystem), C.POINTER(RefAccelerationParam), dp, C.c_double]
        L.get_new_linear_octree.restype = RefLinearOctree
        L.free_linear_octree.argtypes = [C.POINTER(RefLinearOctree)]
        L.free_linear_octree.restype = None
        L.compute_energy.restype = C.c_double
        L.compute_energy.argtypes = [C.POINTER(RefSystem)]
        L.load_built_in_system_python.argtypes = [C.c_char_p, ip, C.POINTER(ip), C.POINTER(dp), C.POINTER(dp), C.POINTER(dp), dp]

    @staticmethod
    def _sys(x, m, G, v=None):
        s = RefSystem()
        s.num_particles = m.shape[0]
        s.particle_ids = None
        s.x = _d(x); s.m = _d(m); s.v = _d(v) if v is not None else None
        s.G = G
        return s

    @staticmethod
    def _param(method, eps, theta, leaf):
        p = RefAccelerationParam()
        p.method = METHODS[method]; p.opening_angle = theta; p.softening_length = eps
        p.max_num_particles_per_leaf = 1 if leaf == -1 else leaf
        return p

    def acceleration(self, x, m, G, method="pairwise", softening_length=0.0, opening_angle=1.0, max_num_particles_per_leaf=1):
        x = _f64(x, (-1, 3)); m = _f64(m, (-1,))
        a = np.empty_like(x)
        s = self._sys(x, m, G); p = self._param(method, softening_length, opening_angle, max_num_particles_per_leaf)
        st = self.L.acceleration(_d(a), C.byref(s), C.byref(p))
        if st.return_code != 0:
            raise RuntimeError(f"reference error {st.return_code}: {st.traceback}")
        return a

    def construct_octree(self, x, m, max_num_particles_per_leaf=1, box_center=None, box_width=-1.0):
        x = _f64(x, (-1, 3)); m = _f64(m, (-1,)); n = m.shape[0]
        s = self._sys(x, m, 1.0); p = self._param("barnes_hut", 0.0, 1.0, max_num_particles_per_leaf)
        t = self.L.get_new_linear_octree()
        bc = _d(_f64(box_center, (3,))) if box_center is not None else None
        st = self.L.construct_octree(C.byref(t), C.byref(s), C.byref(p), bc, box_width)
        if st.return_code != 0:
            raise RuntimeError(f"reference error {st.return_code}: {st.traceback}")
        try:
            d = _tree_dict(t.keys, t.sorted_indices, t.tree_num_particles, t.tree_num_internal_children,
                           t.tree_first_particle_sorted_idx, t.tree_first_internal_children_idx, t.tree_mass,
                           t.com_x, t.com_y, t.com_z, n, t.num_internal_nodes, t.box_width)
        finally:
            self.L.free_linear_octree(C.byref(t))
        d["first_child"] = np.where(d["num_children"] > 0, d["first_child"], -1)   # uninitialised for leaves upstream
        return d

    def energy(self, x, v, m, G):
        x = _f64(x, (-1, 3)); v = _f64(v, (-1, 3)); m = _f64(m, (-1,))
        s = self._sys(x, m, G, v)
        return float(self.L.compute_energy(C.byref(s)))

    def whfast_acceleration(self, x, m, G, jacobi_x, eta, method="pairwise", softening_length=0.0, a0=None):
        """The reference's static whfast_acceleration(), reached through oracle/ref_whfast_probe.c."""
        if not hasattr(self, "P"):
            self.P = C.CDLL(str(PROBE_SO))
            self.P.probe_whfast_acceleration.argtypes = [dp, C.c_int, dp, dp, C.c_double, dp, dp, C.c_int, C.c_double]
        x = _f64(x, (-1, 3)); m = _f64(m, (-1,)); jx = _f64(jacobi_x, (-1, 3)); eta = _f64(eta, (-1,))
        a = np.zeros_like(x) if a0 is None else _f64(a0, (-1, 3)).copy()
        rc = self.P.probe_whfast_acceleration(_d(a), m.shape[0], _d(x), _d(m), G, _d(jx), _d(eta), METHODS[method], softening_length)
        if rc != 0:
            raise RuntimeError(f"reference whfast_acceleration -> {rc}")
        return a

    def _probe(self):
        if not hasattr(self, "P"):
            self.P = C.CDLL(str(PROBE_SO))
            self.P.probe_whfast_acceleration.argtypes = [dp, C.c_int, dp, dp, C.c_double, dp, dp, C.c_int, C.c_double]
        P = self.P
        if not hasattr(P, "_whfast_ready"):
            P.probe_whfast_run.argtypes = [C.c_int, ip, dp, dp, dp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double,
                                           C.c_int, C.c_int]
            P.probe_whfast_c2j.argtypes = [dp, dp, C.c_int, dp, dp, dp, dp]
            P.probe_whfast_c2j.restype = None
            P.probe_whfast_j2c.argtypes = [dp, dp, C.c_int, dp, dp, dp, dp]
            P.probe_whfast_j2c.restype = None
            P.probe_stumpff.argtypes = [dp, C.c_double]
            P.probe_stumpff.restype = None
            P.probe_whfast_drift.argtypes = [C.c_int, ip, dp, dp, dp, C.c_double, dp, dp, dp, C.c_double, C.c_int]
            P._whfast_ready = True
        return P

    def whfast_run(self, x, v, m, G, dt, tf, method="massless", softening_length=0.0, remove_invalid=True, ids=None):
        """The reference's whfast() (src/integrator_whfast.c:200-407), output disabled, through ref_whfast_probe.c.
        Run with OMP_NUM_THREADS=1 when particles can be removed: the OpenMP build fills remove_idx_list racily (:553-562)."""
        P = self._probe()
        x = _f64(x, (-1, 3)).copy(); v = _f64(v, (-1, 3)).copy(); m = _f64(m, (-1,)).copy(); n = m.shape[0]
        ids = np.arange(n, dtype=np.int32) if ids is None else np.ascontiguousarray(ids, dtype=np.int32).copy()
        with zeroed_malloc():     # the reference reads a[0] of a malloc'ed array it never writes
            n2 = P.probe_whfast_run(n, ids.ctypes.data_as(ip), _d(x), _d(v), _d(m), G, dt, tf, METHODS[method], softening_length,
                                    int(remove_invalid), 0)
        if n2 < 0:
            raise RuntimeError(f"reference whfast() -> {n2}")
        return {"x": x[:n2], "v": v[:n2], "m": m[:n2], "ids": ids[:n2]}

    def whfast_stages(self, x, v, m, G, dt):
        """eta, cartesian_to_jacobi, whfast_drift (no removal) and jacobi_to_cartesian of the reference, one after the
        other on the given (already distance-sorted) system; returns every intermediate."""
        P = self._probe()
        x = _f64(x, (-1, 3)).copy(); v = _f64(v, (-1, 3)).copy(); m = _f64(m, (-1,)).copy(); n = m.shape[0]
        eta = np.cumsum(m)     # sequential double additions, same as :1266-1279
        jx = np.zeros_like(x); jv = np.zeros_like(v)
        P.probe_whfast_c2j(_d(jx), _d(jv), n, _d(x), _d(v), _d(m), _d(eta))
        out = {"eta": eta.copy(), "jx0": jx.copy(), "jv0": jv.copy()}
        ids = np.arange(n, dtype=np.int32)
        assert P.probe_whfast_drift(n, ids.ctypes.data_as(ip), _d(x), _d(v), _d(m), G, _d(jx), _d(jv), _d(eta), dt, 0) == n
        out["jx1"], out["jv1"] = jx.copy(), jv.copy()
        x2 = np.empty_like(x); v2 = np.empty_like(v)
        P.probe_whfast_j2c(_d(x2), _d(v2), n, _d(m), _d(jx), _d(jv), _d(eta))
        out["x1"], out["v1"] = x2, v2
        return out

    def stumpff(self, z):
        c = np.empty(4)
        self._probe().probe_stumpff(_d(c), float(z))
        return c

    def built_in_system(self, name: str):
        n = C.c_int(); ids = ip(); x = dp(); v = dp(); m = dp(); G = C.c_double()
        rc = self.L.load_built_in_system_python(name.encode(), C.byref(n), C.byref(ids), C.byref(x), C.byref(v), C.byref(m), C.byref(G))
        if rc != 0:
            raise RuntimeError(f"load_built_in_system_python({name}) -> {rc}")
        f = np.ctypeslib.as_array
        return f(x, (n.value, 3)).copy(), f(v, (n.value, 3)).copy(), f(m, (n.value,)).copy(), G.value


def jacobi_inputs(x, m):
    """eta (prefix masses, src/integrator_whfast.c:1266-1279) and Jacobi positions (cartesian_to_jacobi, :682-727
    restated with numpy) -- the extra inputs of the WHFast acceleration kernels."""
    x = _f64(x, (-1, 3)); m = _f64(m, (-1,))
    eta = np.cumsum(m)
    n = m.shape[0]
    jx = np.zeros_like(x)
    xcm = m[0] * x[0]
    for i in range(1, n):
        jx[i] = x[i] - xcm / eta[i - 1]
        xcm = xcm * (1.0 + m[i] / eta[i - 1]) + m[i] * jx[i]
    jx[0] = xcm / eta[n - 1]
    return jx, eta


# ---- the reference's Python-facing entry point, usable with the reference build AND the drop-in build ------------

DROPIN_SO = HERE / "_ref" / "libgrav_sim_dropin.so"
INTEGRATORS = {"euler": 1, "euler_cromer": 2, "rk4": 3, "leapfrog": 4, "rkf45": 5, "dopri": 6, "dverk": 7, "rkf78": 8,
               "ias15": 9, "whfast": 10}   # src/integrator.h:17-26


class zeroed_malloc:
    """While the REFERENCE's whfast() runs: make glibc hand out zero-filled blocks (mallopt(M_PERTURB, 255): malloc'ed memory
    is filled with 255 ^ 0xff = 0).  The reference kicks jacobi_v[0] with a[0], an element of a malloc'ed array it never
    writes (src/integrator_whfast.c:268-273, :333-340); on a fresh heap that is 0.0 -- the value the oracle and the GPU
    define -- but inside a long-lived test process it is whatever an earlier allocation left there (seen: every velocity of
    the reference run offset by (-5628, 1870, -15863)).  Test infrastructure only."""
    M_PERTURB = -6

    def __enter__(self):
        self.libc = C.CDLL(None)
        self.libc.mallopt(C.c_int(self.M_PERTURB), C.c_int(255))
        return self

    def __exit__(self, *exc):
        self.libc.mallopt(C.c_int(self.M_PERTURB), C.c_int(0))


def launch_simulation(lib_path, x, v, m, G, tf, integrator="leapfrog", dt=1e-3, tolerance=1e-9, method="pairwise",
                      softening_length=0.0, opening_angle=1.0, max_num_particles_per_leaf=1, remove_invalid=False,
                      full=False, output_dir=None, output_interval=None):
    """Run launch_simulation_python (src/python_interface.c:94-193) of the given libgrav_sim build exactly as
    grav_sim/simulator.py:66-102 calls it; returns the final (x, v).  Output is disabled unless output_dir is given
    (then: CSV snapshots, initial one included, every output_interval, doubles)."""
    L = C.CDLL(str(lib_path))
    x = _f64(x, (-1, 3)).copy(); v = _f64(v, (-1, 3)).copy(); m = _f64(m, (-1,)).copy()
    n = C.c_int32(m.shape[0])
    ids = np.arange(m.shape[0], dtype=np.int32)
    new_ids = C.POINTER(C.c_int32)(); new_x = dp(); new_v = dp(); new_m = dp()
    is_exit = C.c_bool(False)
    L.launch_simulation_python.restype = C.c_int
    # (not for the drop-in build's resident loop, which never runs the reference's whfast(): keeps its timings undisturbed)
    if integrator == "whfast" and (str(lib_path) != str(DROPIN_SO) or os.environ.get("GRAV_B200_RESIDENT") == "0"):
        with zeroed_malloc():
            return _launch_simulation_call(L, n, ids, x, v, m, new_ids, new_x, new_v, new_m, G, integrator, dt, tolerance, remove_invalid,
                                           method, opening_angle, softening_length, max_num_particles_per_leaf, output_dir,
                                           output_interval, is_exit, tf, full)
    return _launch_simulation_call(L, n, ids, x, v, m, new_ids, new_x, new_v, new_m, G, integrator, dt, tolerance, remove_invalid,
                                   method, opening_angle, softening_length, max_num_particles_per_leaf, output_dir,
                                   output_interval, is_exit, tf, full)


def _launch_simulation_call(L, n, ids, x, v, m, new_ids, new_x, new_v, new_m, G, integrator, dt, tolerance, remove_invalid, method,
                            opening_angle, softening_length, max_num_particles_per_leaf, output_dir, output_interval, is_exit, tf, full):
    rc = L.launch_simulation_python(
        C.byref(n), ids.ctypes.data_as(C.POINTER(C.c_int32)), _d(x), _d(v), _d(m),
        C.byref(new_ids), C.byref(new_x), C.byref(new_v), C.byref(new_m),
        C.c_double(G), C.c_int32(INTEGRATORS[integrator]), C.c_double(dt), C.c_double(tolerance), C.c_double(-1.0),
        C.c_bool(remove_invalid), C.c_int32(METHODS[method]), C.c_double(opening_angle), C.c_double(softening_length),
        C.c_int32(max_num_particles_per_leaf), C.c_int32(1 if output_dir is None else 2),
        b"" if output_dir is None else str(output_dir).encode(), C.c_bool(output_dir is not None),
        C.c_double(tf if output_interval is None else output_interval), C.c_int32(2),
        C.c_int32(2), C.c_int32(2), C.c_int32(0), C.c_bool(False), C.byref(is_exit), C.c_double(tf))
    if rc != 0:
        raise RuntimeError(f"launch_simulation_python -> {rc}")
    if full:      # whfast reorders (and may remove) particles: n, ids and m change too
        k = n.value
        return {"x": x[:k], "v": v[:k], "m": m[:k], "ids": ids[:k]}
    return x, v


def leapfrog_reference_loop(accel, x, v, m, G, dt, nsteps, energy=None, every=0):
    """The reference leapfrog (src/integrator.c:963-1094) restated with numpy around a force callback: same update
    formulas and operation order (numpy multiplies and adds separately, no FMA).  Returns x, synchronised v, and the
    energies sampled every `every` steps (with the snapshot velocity convention, :1048-1056)."""
    x = _f64(x, (-1, 3)).copy(); v = _f64(v, (-1, 3)).copy()
    xc = np.zeros_like(x); vc = np.zeros_like(v)
    a = accel(x)
    vc += 0.5 * a * dt
    tv = v.copy(); v = tv + vc; vc += tv - v
    es = []
    if energy is not None and every:
        es.append(energy(x, v - 0.5 * a * dt))
    for s in range(1, nsteps + 1):
        xc += v * dt
        tx = x.copy(); x = tx + xc; xc += tx - x
        a = accel(x)
        vc += a * dt
        tv = v.copy(); v = tv + vc; vc += tv - v
        if energy is not None and every and s % every == 0:
            es.append(energy(x, v - 0.5 * a * dt))
    return x, v - 0.5 * a * dt, np.array(es)
