"""ctypes binding of the CPU oracle (oracle/liborc.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, by bench.py's cpu_baseline / --impl reference legs
and by __graft_entry__.smoke() as the checker.  Nothing under pressurefieldcontact.jl_b200/ may
import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_d = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64 = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "pfc_oracle.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liborc.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_sat.restype = C.c_int
        L.orc_sat.argtypes = [_d] * 8
        L.orc_clip_in_tet_coordinates.restype = C.c_int
        L.orc_clip_in_tet_coordinates.argtypes = [C.c_int, _d, _d]
        L.orc_clip_plane_tet.restype = C.c_int
        L.orc_clip_plane_tet.argtypes = [_d, _d, _d]
        L.orc_poly_centroid.restype = C.c_double
        L.orc_poly_centroid.argtypes = [C.c_int, _d, _d, _d]
        L.orc_kat.restype = C.c_int
        L.orc_kat.argtypes = [C.c_int, _d, _d]
        L.orc_zero_small_coordinates.restype = None
        L.orc_zero_small_coordinates.argtypes = [C.c_int, _d]
        L.orc_calc_clamped_piecewise.restype = C.c_double
        L.orc_calc_clamped_piecewise.argtypes = [C.c_double] * 5
        L.orc_traction_regularized.restype = None
        L.orc_traction_regularized.argtypes = [C.c_double, C.c_double, C.c_double, _d, C.c_double, _d]
        L.orc_traction_bristle.restype = None
        L.orc_traction_bristle.argtypes = [C.c_double] * 5 + [_d, C.c_double, _d]
        L.orc_decompose_K.restype = None
        L.orc_decompose_K.argtypes = [_d, C.c_double, _d, _d]
        L.orc_decompose_K_dual6.restype = None
        L.orc_decompose_K_dual6.argtypes = [_d, C.c_double, _d, _d]
        L.orc_inv44.restype = None
        L.orc_inv44.argtypes = [_d, _d]
        L.orc_tet_volume.restype = C.c_double
        L.orc_tet_volume.argtypes = [_d]
        L.orc_fit_tri_obb.restype = None
        L.orc_fit_tri_obb.argtypes = [_d, _d, _d, _d]
        L.orc_fit_tet_obb.restype = C.c_int
        L.orc_fit_tet_obb.argtypes = [_d, _d, _d, _d, _d]
        L.orc_merge_obb.restype = None
        L.orc_merge_obb.argtypes = [_d] * 9
        L.orc_patch_stiffness.restype = None
        L.orc_patch_stiffness.argtypes = [C.c_int64, _d, C.c_double, _d, _d, _d]
        L.orc_scene_create.restype = C.c_void_p
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_add_mesh.restype = C.c_int
        L.orc_add_mesh.argtypes = [C.c_void_p, C.c_int, C.c_int64, _d, C.c_int64, _i32, C.c_void_p, C.c_double, C.c_int64, _d, _d, _d, _i32, _i32,
                                   _i32]
        L.orc_add_instruction.restype = C.c_int
        L.orc_add_instruction.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, _d, C.c_int]
        L.orc_n_bristle.restype = C.c_int
        L.orc_n_bristle.argtypes = [C.c_void_p]
        L.orc_n_visited.restype = C.c_int64
        L.orc_n_visited.argtypes = [C.c_void_p]
        L.orc_eval_f64.restype = C.c_int
        L.orc_eval_f64.argtypes = [C.c_void_p, C.c_int64, _d, _d, C.c_void_p, _d, C.c_void_p, _i64, _i32, C.c_int, C.c_int]
        L.orc_eval_dual6.restype = C.c_int
        L.orc_eval_dual6.argtypes = [C.c_void_p, C.c_int64, _d, _d, _d, C.c_void_p, _d, C.c_void_p, _i64, _i32, C.c_int]
        L.orc_get_pairs.restype = C.c_int64
        L.orc_get_pairs.argtypes = [C.c_void_p, C.c_int64, C.c_int, _i32, C.c_int64]
        L.orc_get_traction.restype = C.c_int64
        L.orc_get_traction.argtypes = [C.c_void_p, C.c_int64, C.c_int, _d, C.c_int64]
        L.orc_max_threads.restype = C.c_int
        L.orc_eval_slice_regularized.restype = C.c_int
        L.orc_eval_slice_regularized.argtypes = [C.c_void_p, _d, _d, C.c_int, C.c_int, C.c_int, _d, _i64]
        L.orc_count_work.restype = C.c_int
        L.orc_count_work.argtypes = [C.c_void_p, C.c_int64, _d, _d, C.c_void_p, _i64]
        _LIB = L
    return _LIB


def _a(x, dt=np.float64):
    return np.ascontiguousarray(x, dtype=dt)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---- unit-level helpers --------------------------------------------------------------------------
def sat(a, b, R_a_b, t_a_b) -> bool:
    """a, b = (c, e, R3x3); R_a_b 3x3, t_a_b 3 -- BB_BB_intersect (src/obb/bb_intersection.jl:2-12)."""
    col = lambda R: _a(np.asarray(R, dtype=np.float64).T.reshape(9))
    return bool(lib().orc_sat(_a(a[0]), _a(a[1]), col(a[2]), _a(b[0]), _a(b[1]), col(b[2]), col(R_a_b), _a(t_a_b)))


def clip_in_tet_coordinates(zeta):
    zeta = _a(zeta).reshape(-1, 4)
    out = np.zeros((8, 4))
    n = lib().orc_clip_in_tet_coordinates(len(zeta), zeta, out)
    if n < 0:
        raise RuntimeError("Non-finite vertex likely" if n == -1 else "something is wrong")
    return out[:n].copy()


def clip_plane_tet(plane, tet_vertices):
    A = np.ones((4, 4))
    A[:3, :] = np.asarray(tet_vertices, dtype=np.float64).T  # columns [v; 1]
    out = np.zeros((4, 3))
    n = lib().orc_clip_plane_tet(_a(plane), _a(A.T.reshape(16)), out)
    return out[:n].copy()


def poly_centroid(v, nhat):
    v = _a(v).reshape(-1, 3)
    c = np.zeros(3)
    area = lib().orc_poly_centroid(len(v), v, _a(nhat), c)
    return area, c


def kat(which: int, values, n_out: int):
    """Small kernels of the path (oracle_capi.cpp::orc_kat): which 0 weightPoly, 1 vec_sub_vec_proj, 2 a_dot_one_pad_b, 3 triangle kernels,
    4 getTriQuadRule, 5 basic_dh algebra."""
    out = np.zeros(n_out)
    if lib().orc_kat(int(which), _a(values).ravel(), out) != 0:
        raise ValueError("orc_kat: unknown kernel")
    return out


def zero_small_coordinates(zeta):
    z = _a(zeta).reshape(-1, 4).copy()
    lib().orc_zero_small_coordinates(len(z), z)
    return z


def calc_clamped_piecewise(x, x1, x2, y1, y2):
    return lib().orc_calc_clamped_piecewise(x, x1, x2, y1, y2)


def traction_regularized(v_c, mu_s, mu_d, vel_t, p_dA):
    out = np.zeros(3)
    lib().orc_traction_regularized(v_c, mu_s, mu_d, _a(vel_t), p_dA, out)
    return out


def traction_bristle(tau, k_bar, mu_s, mu_d, magic, Ts, p_dA):
    out = np.zeros(3)
    lib().orc_traction_bristle(tau, k_bar, mu_s, mu_d, magic, _a(Ts), p_dA, out)
    return out


def decompose_K(K, magic):
    Sinv, Khalf = np.zeros(6), np.zeros((6, 6))
    lib().orc_decompose_K(_a(K).reshape(36), magic, Sinv, Khalf.reshape(36))
    return Sinv, Khalf


def decompose_K_dual6(K7, magic):
    Sinv, Khalf = np.zeros((6, 7)), np.zeros((6, 6, 7))
    lib().orc_decompose_K_dual6(_a(K7).reshape(-1), magic, Sinv.reshape(-1), Khalf.reshape(-1))
    return Sinv, Khalf


def inv44(A):
    out = np.zeros(16)
    lib().orc_inv44(_a(np.asarray(A).T.reshape(16)), out)
    return out.reshape(4, 4).T.copy()


def fit_tri_obb(p):
    c, e, R = np.zeros(3), np.zeros(3), np.zeros(9)
    lib().orc_fit_tri_obb(_a(p).reshape(9), c, e, R)
    return c, e, R.reshape(3, 3).T.copy()


def fit_tet_obb(p, eps):
    c, e, R = np.zeros(3), np.zeros(3), np.zeros(9)
    if lib().orc_fit_tet_obb(_a(p).reshape(12), _a(eps), c, e, R) != 0:
        raise ValueError("inverted tet")
    return c, e, R.reshape(3, 3).T.copy()


def patch_stiffness(traction8, k_bar):
    t = _a(traction8).reshape(-1, 8)
    cop, w, K = np.zeros(3), np.zeros(6), np.zeros((6, 6))
    lib().orc_patch_stiffness(len(t), t, k_bar, cop, w, K.reshape(36))
    return cop, w, K


# ---- scene-level backend (same protocol as pfc_b200.capi.Context) -----------------------------------
class OracleContext:
    """CPU-oracle backend with the same add_mesh / add_instruction / eval_* protocol as the CUDA
    library binding, so tests can run one scenario description through both."""

    name = "oracle"

    def __init__(self, n_threads: int = 1):
        self._h = C.c_void_p(lib().orc_scene_create())
        self.n_ins = 0
        self.n_threads = n_threads
        self._keep = []

    def close(self):
        if self._h:
            lib().orc_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_mesh(self, kind, xyz, idx, eps, Ebar, tree) -> int:
        xyz = _a(xyz)
        idx = _a(idx, np.int32)
        eps_a = None if eps is None else _a(eps)
        r = lib().orc_add_mesh(self._h, kind, len(xyz), xyz, len(idx), idx, _ptr(eps_a), float(Ebar or 0.0), tree.n_node, _a(tree.c), _a(tree.e),
                               _a(tree.R), _a(tree.left, np.int32), _a(tree.right, np.int32), _a(tree.leaf_id, np.int32))
        if r < 0:
            raise RuntimeError("orc_add_mesh failed")
        return r

    def add_instruction(self, mesh_1, mesh_2, chi, model, params, n_quad_rule) -> int:
        r = lib().orc_add_instruction(self._h, mesh_1, mesh_2, chi, model, _a(params), n_quad_rule)
        if r < 0:
            raise RuntimeError("orc_add_instruction failed")
        self.n_ins = r + 1
        return r

    def finalize(self, max_env: int = 1):
        pass

    @property
    def n_bristle(self) -> int:
        return lib().orc_n_bristle(self._h)

    def eval_f64(self, X, twist, s=None, keep=False):
        X = _a(X).reshape(-1, self.n_ins, 16)
        n_env = X.shape[0]
        twist = _a(twist).reshape(n_env, self.n_ins, 6)
        nb = self.n_bristle
        s_a = None if nb == 0 else _a(s).reshape(n_env, nb, 6)
        wrench = np.zeros((n_env, self.n_ins, 6))
        sdot = np.zeros((n_env, nb, 6)) if nb else None
        n_pairs = np.zeros((n_env, self.n_ins), np.int64)
        flags = np.zeros((n_env, self.n_ins), np.int32)
        lib().orc_eval_f64(self._h, n_env, X, twist, _ptr(s_a), wrench, _ptr(sdot), n_pairs, flags, self.n_threads, int(keep))
        return dict(wrench=wrench, sdot=sdot, n_pairs=n_pairs, flags=flags)

    def eval_dual6(self, X_bp, X7, twist7, s7=None):
        X_bp = _a(X_bp).reshape(-1, self.n_ins, 16)
        n_env = X_bp.shape[0]
        X7 = _a(X7).reshape(n_env, self.n_ins, 16, 7)
        twist7 = _a(twist7).reshape(n_env, self.n_ins, 6, 7)
        nb = self.n_bristle
        s_a = None if nb == 0 else _a(s7).reshape(n_env, nb, 6, 7)
        wrench = np.zeros((n_env, self.n_ins, 6, 7))
        sdot = np.zeros((n_env, nb, 6, 7)) if nb else None
        n_pairs = np.zeros((n_env, self.n_ins), np.int64)
        flags = np.zeros((n_env, self.n_ins), np.int32)
        lib().orc_eval_dual6(self._h, n_env, X_bp, X7, twist7, _ptr(s_a), wrench, _ptr(sdot), n_pairs, flags, self.n_threads)
        return dict(wrench=wrench, sdot=sdot, n_pairs=n_pairs, flags=flags)

    def get_pairs(self, env: int, ins: int) -> np.ndarray:
        n = lib().orc_get_pairs(self._h, env, ins, np.zeros((1, 2), np.int32), 0)
        out = np.zeros((max(n, 1), 2), np.int32)
        lib().orc_get_pairs(self._h, env, ins, out, n)
        return out[:n]

    def get_traction(self, env: int, ins: int) -> np.ndarray:
        n = lib().orc_get_traction(self._h, env, ins, np.zeros((1, 8)), 0)
        out = np.zeros((max(n, 1), 8))
        lib().orc_get_traction(self._h, env, ins, out, n)
        return out[:n]

    def count_work(self, X, twist, s=None):
        """Algorithmic work of the reference algorithm for this batch (instrumented scalar)."""
        X = _a(X).reshape(-1, self.n_ins, 16)
        n_env = X.shape[0]
        twist = _a(twist).reshape(n_env, self.n_ins, 6)
        s_a = None if self.n_bristle == 0 else _a(s).reshape(n_env, self.n_bristle, 6)
        out = np.zeros(5, np.int64)
        lib().orc_count_work(self._h, n_env, X, twist, _ptr(s_a), out)
        return dict(flops_broad=int(out[0]), flops_narrow=int(out[1]), node_pairs=int(out[2]), candidate_pairs=int(out[3]),
                    traction_points=int(out[4]))

    def eval_slice_regularized(self, X, twist, ins, rank, world):
        """Partial wrench of instruction `ins` (environment 0) over rank's slice of the pair list."""
        w, n = np.zeros(6), np.zeros(1, np.int64)
        rc = lib().orc_eval_slice_regularized(self._h, _a(X).reshape(-1), _a(twist).reshape(-1), ins, rank, world, w, n)
        if rc != 0:
            raise RuntimeError("slice evaluation needs a regularized instruction")
        return w, int(n[0])

    def n_visited(self) -> int:
        return lib().orc_n_visited(self._h)
