"""ctypes binding of `oracle/libpmc_oracle.so` (the plain-C restatement in `oracle/pmc_oracle.c`).

TEST INFRASTRUCTURE ONLY -- see `oracle/__init__.py`.  Every method maps 1:1 onto a function declared in
`oracle/pmc_oracle.h`, which cites the reference file:line it follows.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpmc_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only; no reference sources are used)."""
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "pmc_oracle.c"))):
        subprocess.check_call(["make", "-C", _HERE, "libpmc_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Yarn5(C.Structure):
    _fields_ = [("a", C.c_int32 * 5), ("r", C.c_int32 * 5)]


class _Stats(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("estimate", "ml_estimator_variance", "bias2", "actual_mse", "eps2",
                                          "alpha", "alpha_abs", "beta", "gamma")]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.po_create.restype = C.c_void_p
        L.po_create.argtypes = [C.c_int]
        L.po_destroy.argtypes = [C.c_void_p]
        L.po_set_tolerances.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int]
        L.po_uniformoo.restype = C.c_double
        L.po_uniformoo.argtypes = [C.c_int32]
        L.po_inv_Phi.restype = C.c_double
        L.po_inv_Phi.argtypes = [C.c_double]
        L.po_yarn5_next.restype = C.c_int32
        L.po_yarn5_jump.argtypes = [C.POINTER(_Yarn5), C.c_uint64]
        L.po_yarn5_split.argtypes = [C.POINTER(_Yarn5), C.c_uint, C.c_uint]
        L.po_yarn5_fill_int.argtypes = [C.POINTER(_Yarn5), C.c_int64, C.POINTER(C.c_int32)]
        L.po_normal_fill.argtypes = [C.POINTER(_Yarn5), C.c_double, C.c_double, C.c_int64, _dp]
        L.po_normal_map.argtypes = [C.c_int64, C.POINTER(C.c_int32), C.c_double, C.c_double, _dp]
        L.po_det_eval.argtypes = [C.c_int, C.c_int64, _dp, _dp]
        L.po_set_sampler_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _ip, _ip, _dp,
                                           _dp, C.c_int, _ip, _ip, _dp, C.c_double, C.c_double, C.c_int]
        L.po_set_darcy_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _ip, _ip, _dp,
                                         _ip, _dp, _dp, _dp, C.c_int, _ip, _ip, _dp]
        L.po_set_field_transfer.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip, _ip, _dp, _dp]
        L.po_sampler_eval.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, C.c_int, _ip]
        L.po_darcy_solve.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp, _ip]
        L.po_mlmc_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_double, C.c_double,
                                    _dp, _dp, C.c_int, C.POINTER(C.c_int64)]
        L.po_set_observations.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, C.c_double]
        L.po_bayes_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_double, C.c_double, _dp, _dp, C.c_int]
        L.po_exp_w_regression.restype = C.c_double
        L.po_exp_w_regression.argtypes = [_dp, _dp, C.c_int, C.c_int]
        L.po_mlmc_compute.argtypes = [C.c_int, _dp, _ip, _dp, _dp, C.c_double, C.c_double] + [_dp] * 10 + [
            _ip, C.POINTER(_Stats)]
        L.po_mc_compute.argtypes = [_dp, C.c_int, _dp, C.c_double, C.c_double, _dp, _dp, _dp, _dp, _ip,
                                    C.POINTER(_Stats)]
        _lib = L
    return _lib


class Yarn5:
    """trng::yarn5 as restated in the oracle (default parameter set and seed)."""

    def __init__(self):
        self.g = _Yarn5()
        lib().po_yarn5_init(C.byref(self.g))

    def jump(self, n: int):
        lib().po_yarn5_jump(C.byref(self.g), C.c_uint64(n))
        return self

    def split(self, s: int, n: int):
        lib().po_yarn5_split(C.byref(self.g), s, n)
        return self

    def ints(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.int32)
        lib().po_yarn5_fill_int(C.byref(self.g), n, out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out

    def normals(self, n: int, mu: float = 0.0, sigma: float = 1.0) -> np.ndarray:
        out = np.empty(n, dtype=np.float64)
        lib().po_normal_fill(C.byref(self.g), mu, sigma, n, _d(out))
        return out

    @property
    def state(self):
        return list(self.g.r), list(self.g.a)


def normal_map(engine, mu: float = 0.0, sigma: float = 1.0) -> np.ndarray:
    """Normal deviates of given engine outputs: inv_Phi(uniformoo(x)) * sigma + mu (`po_normal_map`)."""
    engine = np.ascontiguousarray(engine, dtype=np.int32)
    out = np.empty(engine.shape[0], dtype=np.float64)
    lib().po_normal_map(engine.shape[0], engine.ctypes.data_as(C.POINTER(C.c_int32)), mu, sigma, _d(out))
    return out


DET_FUNCS = {"exp": 0, "log": 1, "erf": 2, "erfc": 3, "inv_Phi": 4, "inv_Phi_libm": 5}


def det_eval(which: str, x) -> np.ndarray:
    """The deterministic elementary functions of parelagmc_b200/csrc/detmath.h as compiled into the oracle, and inv_Phi
    over them ("inv_Phi") or over glibc's libm ("inv_Phi_libm")."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    lib().po_det_eval(DET_FUNCS[which], x.shape[0], _d(x), _d(y))
    return y


def _csr_args(m):
    if m is None:
        return None, None, None, []
    rp = np.ascontiguousarray(m.indptr, dtype=np.int32)
    ci = np.ascontiguousarray(m.indices, dtype=np.int32)
    v = np.ascontiguousarray(m.data, dtype=np.float64)
    return _i(rp), _i(ci), _d(v), [rp, ci, v]


class OracleProblem:
    """po_problem: sampler + Darcy levels and the per-sample path (see pmc_oracle.h)."""

    def __init__(self, sampler_levels, darcy_levels, alpha: float, g: float, lognormal: bool = True):
        L = lib()
        self.nlevels = len(sampler_levels)
        self.h = L.po_create(self.nlevels)
        self.Ne = [s.Ne for s in sampler_levels]
        self.Nf = [s.Nf for s in sampler_levels]
        for l, s in enumerate(sampler_levels):
            mr, mc, mv, k1 = _csr_args(s.M)
            br, bc, bv, k2 = _csr_args(s.B)
            pr, pc, pv, k3 = _csr_args(s.P)
            wd = np.ascontiguousarray(s.Wdiag, dtype=np.float64)
            rc = L.po_set_sampler_level(self.h, l, s.Ne, s.Nf, mr, mc, mv, br, bc, bv, _d(wd),
                                        0 if s.P is None else s.P.shape[1], pr, pc, pv, alpha, g,
                                        1 if lognormal else 0)
            assert rc == 0
            if getattr(s, "T", None) is not None:
                tr, tc, tv, k4 = _csr_args(s.T)
                ts = None if s.Tscale is None else np.ascontiguousarray(s.Tscale, dtype=np.float64)
                assert L.po_set_field_transfer(self.h, l, s.T.shape[0], tr, tc, tv, _d(ts)) == 0
        self.Nout = [s.T.shape[0] if getattr(s, "T", None) is not None else s.Ne for s in sampler_levels]
        self.dNe = [d.Ne for d in (darcy_levels or [])]
        self.dNf = [d.Nf for d in (darcy_levels or [])]
        for l, d in enumerate(darcy_levels or []):
            br, bc, bv, k2 = _csr_args(d.B)
            pr, pc, pv, k3 = _csr_args(d.P_p)
            ep = np.ascontiguousarray(d.elem_ptr, dtype=np.int32)
            ed = np.ascontiguousarray(d.elem_dofs, dtype=np.int32)
            em = np.ascontiguousarray(d.elem_mat, dtype=np.float64)
            eu = np.ascontiguousarray(d.ess_u, dtype=np.int32)
            ev = np.ascontiguousarray(d.ess_data, dtype=np.float64)
            rh = np.ascontiguousarray(d.rhs, dtype=np.float64)
            ob = np.ascontiguousarray(d.obs, dtype=np.float64)
            rc = L.po_set_darcy_level(self.h, l, d.Ne, d.Nf, _i(ep), _i(ed), _d(em), br, bc, bv, _i(eu),
                                      _d(ev), _d(rh), _d(ob), 0 if d.P_p is None else d.P_p.shape[1],
                                      pr, pc, pv)
            assert rc == 0

    def __del__(self):
        try:
            if self.h:
                lib().po_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_tolerances(self, rel: float, abs_: float, maxit: int):
        lib().po_set_tolerances(self.h, rel, abs_, maxit)

    def sampler_eval(self, level: int, xi: np.ndarray, xi_level: Optional[int] = None,
                     embed_s: Optional[np.ndarray] = None, init_level: int = 0, use_init: int = -1):
        """PDESampler::Eval.  Returns (s, embed_s_out, iters)."""
        if xi_level is None:
            xi_level = self.Ne.index(len(xi))
        xi = np.ascontiguousarray(xi, dtype=np.float64)
        s = np.empty(self.Nout[level])
        emb = np.zeros(max(self.Ne[level], self.Ne[init_level] if embed_s is not None else 0))
        if embed_s is not None:
            emb[:len(embed_s)] = embed_s
        it = C.c_int(0)
        rc = lib().po_sampler_eval(self.h, level, xi_level, _d(xi), _d(s), _d(emb), init_level, use_init,
                                   C.byref(it))
        assert rc == 0, rc
        return s, emb[:self.Ne[level]].copy(), it.value

    def darcy_solve(self, level: int, k: np.ndarray, want_sol: bool = False):
        """DarcySolver::SolveFwd.  Returns (Q, C, sol|None, iters)."""
        k = np.ascontiguousarray(k, dtype=np.float64)
        q = C.c_double(0)
        c = C.c_double(0)
        it = C.c_int(0)
        sol = np.empty(self.dNe[level] + self.dNf[level]) if want_sol else None
        rc = lib().po_darcy_solve(self.h, level, _d(k), C.byref(q), C.byref(c), _d(sol), C.byref(it))
        assert rc == 0, rc
        return q.value, c.value, sol, it.value

    def mlmc_level(self, level: int, nsamples: int, pos0: int, mu: float = 0.0, sigma: float = 1.0,
                   nthreads: int = 1, nlevels: Optional[int] = None):
        """One level of MLMC_Manager::InitRun.  Returns (sums[9], rows[nsamples,4], total_iters)."""
        sums = np.zeros(9)
        rows = np.zeros((max(nsamples, 1), 4))
        its = C.c_int64(0)
        rc = lib().po_mlmc_level(self.h, level, nlevels or self.nlevels, nsamples, C.c_uint64(pos0), mu, sigma,
                                 _d(sums), _d(rows), nthreads, C.byref(its))
        assert rc == 0, rc
        return sums, rows[:nsamples], its.value


def _oracle_set_observations(self, level: int, g, G_obs, noise: float):
    g = np.ascontiguousarray(g, dtype=np.float64).reshape(-1, self.dNe[level])
    G_obs = np.ascontiguousarray(G_obs, dtype=np.float64)
    assert lib().po_set_observations(self.h, level, g.shape[0], _d(g), _d(G_obs), float(noise)) == 0


def _oracle_bayes_level(self, level: int, nsamples: int, pos0: int, mu: float = 0.0, sigma: float = 1.0,
                        nthreads: int = 1, nlevels: Optional[int] = None):
    """One level of ML_BayesRatio_Manager::InitRun.  Returns (sums[20], rows[nsamples,5])."""
    sums = np.zeros(20)
    rows = np.zeros((max(nsamples, 1), 5))
    rc = lib().po_bayes_level(self.h, level, nlevels or self.nlevels, nsamples, C.c_uint64(pos0), mu, sigma, _d(sums),
                              _d(rows), nthreads)
    assert rc == 0, rc
    return sums, rows[:nsamples]


OracleProblem.set_observations = _oracle_set_observations
OracleProblem.bayes_level = _oracle_bayes_level


def exp_w_regression(y, x, skip_n_last: int) -> float:
    y = np.ascontiguousarray(y, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    return lib().po_exp_w_regression(_d(y), _d(x), len(y), skip_n_last)


def mlmc_compute(sums, nsamples, M, cost=None, eps2: float = 1e-3, ratio: float = 0.5):
    """MLMC_Manager::computeNSamplesMSE.  Returns a dict of per-level arrays and scalars."""
    sums = np.ascontiguousarray(sums, dtype=np.float64)
    L = sums.shape[0]
    ns = np.ascontiguousarray(nsamples, dtype=np.int32)
    M = np.ascontiguousarray(M, dtype=np.float64)
    cost_a = None if cost is None else np.ascontiguousarray(cost, dtype=np.float64)
    outs = [np.zeros(L) for _ in range(10)]
    missing = np.zeros(L, dtype=np.int32)
    st = _Stats()
    lib().po_mlmc_compute(L, _d(sums), _i(ns), _d(M), _d(cost_a), eps2, ratio, *[_d(o) for o in outs],
                          _i(missing), C.byref(st))
    names = ["eY", "eABSY", "eQ", "eABSQ", "eC", "varY", "varQ", "consistency", "kurtosis", "VC"]
    res = dict(zip(names, outs))
    res["missing"] = missing
    for n, _ in _Stats._fields_:
        res[n] = getattr(st, n)
    return res


def mc_compute(sums, nsamples: int, cost=None, eps2: float = 1e-3, ratio: float = 0.5):
    """MC_Manager::computeNSamplesMSE."""
    sums = np.ascontiguousarray(sums, dtype=np.float64)
    vals = [C.c_double(0) for _ in range(4)]
    missing = C.c_int(0)
    st = _Stats()
    cst = None if cost is None else C.byref(C.c_double(cost))
    lib().po_mc_compute(_d(sums), nsamples, C.cast(cst, _dp) if cst is not None else None, eps2, ratio,
                        *[C.cast(C.byref(v), _dp) for v in vals], C.cast(C.byref(missing), _ip), C.byref(st))
    res = dict(eQ=vals[0].value, eABSQ=vals[1].value, eC=vals[2].value, varQ=vals[3].value,
               missing=missing.value)
    for n, _ in _Stats._fields_:
        res[n] = getattr(st, n)
    return res
