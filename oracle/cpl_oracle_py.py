"""ctypes view of oracle/libcpl_oracle.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs (see cpl_oracle.h). The product package never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcpl_oracle.so")

ENV_NONE, ENV_GROUND, ENV_SUPERQUADRIC = 0, 1, 2
BLOCK_COM, BLOCK_F, BLOCK_P, BLOCK_N = 0, 1, 2, 3


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, seconds)."""
    src = os.path.join(_HERE, "cpl_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libcpl_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        L.cpl_oracle_new.restype = C.c_void_p
        L.cpl_oracle_new.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_double]
        L.cpl_oracle_free.argtypes = [C.c_void_p]
        L.cpl_oracle_dims.argtypes = [C.c_void_p, ip, ip, ip]
        L.cpl_oracle_sorted_order.argtypes = [C.c_void_p, ip]
        L.cpl_oracle_structure.argtypes = [C.c_void_p, ip, ip]
        L.cpl_oracle_var_bounds.argtypes = [C.c_void_p, dp, dp]
        L.cpl_oracle_con_bounds.argtypes = [C.c_void_p, dp, dp]
        L.cpl_oracle_set_mass.argtypes = [C.c_void_p, C.c_double]
        L.cpl_oracle_set_wrench.argtypes = [C.c_void_p, dp]
        L.cpl_oracle_set_mu.argtypes = [C.c_void_p, C.c_double]
        L.cpl_oracle_set_ground_z.argtypes = [C.c_void_p, C.c_double]
        L.cpl_oracle_set_superquadric.argtypes = [C.c_void_p, dp, dp, dp]
        L.cpl_oracle_set_force_threshold.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.cpl_oracle_set_com_ref.argtypes = [C.c_void_p, dp]
        L.cpl_oracle_set_com_weight.argtypes = [C.c_void_p, C.c_double]
        L.cpl_oracle_set_pos_ref.argtypes = [C.c_void_p, C.c_int, dp]
        L.cpl_oracle_set_force_ref.argtypes = [C.c_void_p, C.c_int, dp]
        L.cpl_oracle_set_pos_weight.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.cpl_oracle_set_force_weight.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.cpl_oracle_set_var_bounds.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, dp]
        L.cpl_oracle_set_reduction_order.argtypes = [C.c_void_p, C.c_int]
        L.cpl_oracle_set_call_all_pairs.argtypes = [C.c_void_p, C.c_int]
        L.cpl_oracle_eval.argtypes = [C.c_void_p, dp, dp, dp, dp, dp]
        L.cpl_oracle_eval_batch.restype = C.c_int
        L.cpl_oracle_eval_batch.argtypes = [C.c_void_p, C.c_longlong, dp, dp, dp, dp, dp, C.c_int]
        L.cpl_oracle_env_value.argtypes = [C.c_void_p, dp, dp]
        L.cpl_oracle_env_gradient.argtypes = [C.c_void_p, dp, dp]
        L.cpl_oracle_env_normal.argtypes = [C.c_void_p, dp, dp]
        L.cpl_oracle_env_normal_jacobian.argtypes = [C.c_void_p, dp, dp]
        _lib = L
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _vec(v, n):
    a = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(n))
    return a


class Oracle:
    """One CplProblem (src/CplProblem.cpp:6-82) evaluated by the C restatement."""

    def __init__(self, names, env_kind=ENV_GROUND, mass=100.0):
        self.names = [str(s) for s in names]
        self.nc = len(self.names)
        arr = (C.c_char_p * self.nc)(*[s.encode() for s in self.names])
        self._h = lib().cpl_oracle_new(self.nc, arr, int(env_kind), float(mass))
        if not self._h:
            raise ValueError("cpl_oracle_new failed")
        n, m, nnz = C.c_int(), C.c_int(), C.c_int()
        lib().cpl_oracle_dims(self._h, C.byref(n), C.byref(m), C.byref(nnz))
        self.n, self.m, self.nnz = n.value, m.value, nnz.value
        self.env_kind = int(env_kind)
        self.sq = (np.array([0.0, 0.0, 10.0]), np.array([10.0] * 3), np.array([10.0] * 3))  # Superquadric.cpp:7-9

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.cpl_oracle_free(h)
            self._h = None

    def index(self, name):
        return self.names.index(name)

    # -- layout ------------------------------------------------------------------------------
    def sorted_order(self):
        p = np.zeros(self.nc, dtype=np.int32)
        lib().cpl_oracle_sorted_order(self._h, p.ctypes.data_as(C.POINTER(C.c_int)))
        return p

    def structure(self):
        r = np.zeros(self.nnz, dtype=np.int32)
        c = np.zeros(self.nnz, dtype=np.int32)
        ip = C.POINTER(C.c_int)
        lib().cpl_oracle_structure(self._h, r.ctypes.data_as(ip), c.ctypes.data_as(ip))
        return r, c

    def var_bounds(self):
        lb, ub = np.zeros(self.n), np.zeros(self.n)
        lib().cpl_oracle_var_bounds(self._h, _dp(lb), _dp(ub))
        return lb, ub

    def con_bounds(self):
        lb, ub = np.zeros(self.m), np.zeros(self.m)
        lib().cpl_oracle_con_bounds(self._h, _dp(lb), _dp(ub))
        return lb, ub

    # -- parameters --------------------------------------------------------------------------
    def set_mass(self, m):
        lib().cpl_oracle_set_mass(self._h, float(m))

    def set_wrench(self, w):
        lib().cpl_oracle_set_wrench(self._h, _dp(_vec(w, 6)))

    def set_mu(self, mu):
        lib().cpl_oracle_set_mu(self._h, float(mu))

    def set_ground_z(self, z):
        lib().cpl_oracle_set_ground_z(self._h, float(z))

    def set_superquadric(self, Cc, R, P):
        self.sq = (_vec(Cc, 3).copy(), _vec(R, 3).copy(), _vec(P, 3).copy())
        lib().cpl_oracle_set_superquadric(self._h, _dp(_vec(Cc, 3)), _dp(_vec(R, 3)), _dp(_vec(P, 3)))

    def set_force_threshold(self, name, thr):
        lib().cpl_oracle_set_force_threshold(self._h, self.index(name), float(thr))

    def set_com_ref(self, r):
        lib().cpl_oracle_set_com_ref(self._h, _dp(_vec(r, 3)))

    def set_com_weight(self, w):
        lib().cpl_oracle_set_com_weight(self._h, float(w))

    def set_pos_ref(self, name, r):
        lib().cpl_oracle_set_pos_ref(self._h, self.index(name), _dp(_vec(r, 3)))

    def set_force_ref(self, name, r):
        lib().cpl_oracle_set_force_ref(self._h, self.index(name), _dp(_vec(r, 3)))

    def set_contact_pos_weight(self, name, w):
        lib().cpl_oracle_set_pos_weight(self._h, self.index(name), float(w))

    def set_contact_force_weight(self, name, w):
        lib().cpl_oracle_set_force_weight(self._h, self.index(name), float(w))

    def set_pos_weight(self, w):
        for k in range(self.nc):
            lib().cpl_oracle_set_pos_weight(self._h, k, float(w))

    def set_force_weight(self, w):
        for k in range(self.nc):
            lib().cpl_oracle_set_force_weight(self._h, k, float(w))

    def set_var_bounds(self, block, name, lb, ub):
        k = 0 if block == BLOCK_COM else self.index(name)
        lib().cpl_oracle_set_var_bounds(self._h, int(block), k, _dp(_vec(lb, 3)), _dp(_vec(ub, 3)))

    def set_reduction_order(self, order):
        lib().cpl_oracle_set_reduction_order(self._h, int(order))

    def set_call_all_pairs(self, on):
        lib().cpl_oracle_set_call_all_pairs(self._h, int(bool(on)))

    # -- evaluation --------------------------------------------------------------------------
    def eval(self, x, want=("g", "jac", "cost", "grad")):
        x = _vec(x, self.n)
        g = np.zeros(self.m) if "g" in want else None
        jac = np.zeros(self.nnz) if "jac" in want else None
        cost = np.zeros(1) if "cost" in want else None
        grad = np.zeros(self.n) if "grad" in want else None
        lib().cpl_oracle_eval(self._h, _dp(x), _dp(g), _dp(jac), _dp(cost), _dp(grad))
        return {"g": g, "jac": jac, "cost": None if cost is None else float(cost[0]), "grad": grad}

    def eval_batch(self, X, want=("g", "jac", "cost", "grad"), nthreads=1, out=None):
        """X: (N, n) instance-major. Returns instance-major arrays."""
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        N = X.shape[0]
        assert X.shape == (N, self.n)
        out = out or {}
        g = out.get("g", np.zeros((N, self.m))) if "g" in want else None
        jac = out.get("jac", np.zeros((N, self.nnz))) if "jac" in want else None
        cost = out.get("cost", np.zeros(N)) if "cost" in want else None
        grad = out.get("grad", np.zeros((N, self.n))) if "grad" in want else None
        used = lib().cpl_oracle_eval_batch(self._h, N, _dp(X), _dp(g), _dp(jac), _dp(cost), _dp(grad), int(nthreads))
        return {"g": g, "jac": jac, "cost": cost, "grad": grad, "threads": used}

    # -- environment-level -------------------------------------------------------------------
    def env_value(self, p):
        v = np.zeros(1)
        lib().cpl_oracle_env_value(self._h, _dp(_vec(p, 3)), _dp(v))
        return float(v[0])

    def env_gradient(self, p):
        v = np.zeros(3)
        lib().cpl_oracle_env_gradient(self._h, _dp(_vec(p, 3)), _dp(v))
        return v

    def env_normal(self, p):
        v = np.zeros(3)
        lib().cpl_oracle_env_normal(self._h, _dp(_vec(p, 3)), _dp(v))
        return v

    def env_normal_jacobian(self, p):
        v = np.zeros(9)
        lib().cpl_oracle_env_normal_jacobian(self._h, _dp(_vec(p, 3)), _dp(v))
        return v.reshape(3, 3)
