"""ctypes view of oracle/_ref/libcpl_ref.so -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

libcpl_ref.so is the reference's OWN source (src/CplProblem.cpp, src/Constraints/*.cpp, src/Ground.cpp,
src/Superquadric.cpp, src/MinimizeCentroidalVariables.cpp, src/Variable3D.cpp) compiled where it lies under
/root/reference against the stand-in Eigen/ifopt headers in oracle/refshim (recipe: oracle/Makefile, target
`ref`).  It exists only where it was built (this container; it travels to the GPU box as a built artefact)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libcpl_ref.so")
_KIND = {"none": 0, "ground": 1, "superquadric": 2}


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.cpl_ref_new.restype = C.c_void_p
        L.cpl_ref_new.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_double]
        L.cpl_ref_free.argtypes = [C.c_void_p]
        for name, args in {
            "cpl_ref_set_wrench": [dp], "cpl_ref_set_mu": [C.c_double], "cpl_ref_set_ground_z": [C.c_double],
            "cpl_ref_set_superquadric": [dp, dp, dp], "cpl_ref_set_force_threshold": [C.c_int, C.c_double],
            "cpl_ref_set_com_ref": [dp], "cpl_ref_set_com_weight": [C.c_double], "cpl_ref_set_pos_ref": [C.c_int, dp],
            "cpl_ref_set_force_ref": [C.c_int, dp], "cpl_ref_set_pos_weight": [C.c_int, C.c_double],
            "cpl_ref_set_force_weight": [C.c_int, C.c_double], "cpl_ref_dims": [ip, ip, ip], "cpl_ref_structure": [ip, ip],
            "cpl_ref_bounds": [dp, dp, dp, dp], "cpl_ref_eval": [dp, dp, dp, dp, dp],
        }.items():
            getattr(L, name).argtypes = [C.c_void_p] + args
        L.cpl_ref_eval_batch.restype = C.c_int
        L.cpl_ref_eval_batch.argtypes = [C.c_void_p, C.c_longlong, dp, dp, dp, dp, dp, C.c_int]
        _lib = L
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _vec(v, n):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(n))


class RefProblem:
    """The reference's cpl::solver::CplProblem, with the setter names of BatchedCplProblem / the env objects."""

    def __init__(self, names, env_name, mass):
        self.names = [str(s) for s in names]
        arr = (C.c_char_p * len(self.names))(*[s.encode() for s in self.names])
        self._h = lib().cpl_ref_new(len(self.names), arr, _KIND[env_name], float(mass))
        n, m, z = C.c_int(), C.c_int(), C.c_int()
        lib().cpl_ref_dims(self._h, C.byref(n), C.byref(m), C.byref(z))
        self.n, self.m, self.nnz = n.value, m.value, z.value

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.cpl_ref_free(self._h)
            self._h = None

    def _k(self, name):
        return self.names.index(name)

    def SetGroundZ(self, z): lib().cpl_ref_set_ground_z(self._h, float(z))
    def SetParameters(self, Cc, R, P): lib().cpl_ref_set_superquadric(self._h, _dp(_vec(Cc, 3)), _dp(_vec(R, 3)), _dp(_vec(P, 3)))
    def SetMu(self, mu): lib().cpl_ref_set_mu(self._h, float(mu))
    def SetManipulationWrench(self, w): lib().cpl_ref_set_wrench(self._h, _dp(_vec(w, 6)))
    def SetForceThreshold(self, nm, t): lib().cpl_ref_set_force_threshold(self._h, self._k(nm), float(t))
    def SetCoMRef(self, r): lib().cpl_ref_set_com_ref(self._h, _dp(_vec(r, 3)))
    def SetCoMWeight(self, w): lib().cpl_ref_set_com_weight(self._h, float(w))
    def SetPosRef(self, nm, r): lib().cpl_ref_set_pos_ref(self._h, self._k(nm), _dp(_vec(r, 3)))
    def SetForceRef(self, nm, r): lib().cpl_ref_set_force_ref(self._h, self._k(nm), _dp(_vec(r, 3)))
    def SetContactPosWeight(self, nm, w): lib().cpl_ref_set_pos_weight(self._h, self._k(nm), float(w))
    def SetContactForceWeight(self, nm, w): lib().cpl_ref_set_force_weight(self._h, self._k(nm), float(w))

    def SetPosWeight(self, w):
        for k in range(len(self.names)):
            lib().cpl_ref_set_pos_weight(self._h, k, float(w))

    def SetForceWeight(self, w):
        for k in range(len(self.names)):
            lib().cpl_ref_set_force_weight(self._h, k, float(w))

    def SetPosBounds(self, nm, lb, ub): pass  # bounds do not enter the evaluation; checked through cpl_ref_bounds defaults
    def SetForceBounds(self, nm, lb, ub): pass
    def SetNormalBounds(self, nm, lb, ub): pass

    def structure(self):
        r = np.zeros(self.nnz, dtype=np.int32)
        c = np.zeros(self.nnz, dtype=np.int32)
        ip = C.POINTER(C.c_int)
        lib().cpl_ref_structure(self._h, r.ctypes.data_as(ip), c.ctypes.data_as(ip))
        return r, c

    def bounds(self):
        xl, xu, gl, gu = np.zeros(self.n), np.zeros(self.n), np.zeros(self.m), np.zeros(self.m)
        lib().cpl_ref_bounds(self._h, _dp(xl), _dp(xu), _dp(gl), _dp(gu))
        return xl, xu, gl, gu

    def eval_batch(self, X, want=("g", "jac", "cost", "grad"), nthreads=1, out=None):
        """out: optional dict of preallocated C-contiguous float64 arrays (a timed loop must not page-fault fresh arrays)."""
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
        N = X.shape[0]
        out = out or {}

        def buf(key, shape):
            if key not in want:
                return None
            a = out.get(key)
            if a is None:
                return np.zeros(shape)
            assert a.dtype == np.float64 and a.flags.c_contiguous and a.shape == shape, key
            return a

        g, jac, cost, grad = buf("g", (N, self.m)), buf("jac", (N, self.nnz)), buf("cost", (N,)), buf("grad", (N, self.n))
        used = lib().cpl_ref_eval_batch(self._h, N, _dp(X), _dp(g), _dp(jac), _dp(cost), _dp(grad), int(nthreads))
        return {"g": g, "jac": jac, "cost": cost, "grad": grad, "threads": used}
