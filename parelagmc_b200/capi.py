"""ctypes binding of the product library `parelagmc_b200/lib/libpmc_b200.so` (C ABI: `include/pmc_b200.h`).

There is no CPU fallback: importing this module is harmless, but `load()` raises if the CUDA library has
not been built, and every compute call raises `PmcError` if no sm_100 GPU is usable.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PMC_B200_LIB") or os.path.join(_HERE, "lib", "libpmc_b200.so")   # env: diagnostic builds
CSRC = os.path.join(_HERE, "csrc")

K_CLASSES = ["saddle_apply", "lanczos_update", "solution_update", "mass_smooth", "schur_smooth", "transfer",
             "setup", "scalar", "rng", "misc"]

# every symbol include/pmc_b200.h declares (checked by tests/test_abi.py against the header text)
SYMBOLS = [
    "pmc_create", "pmc_destroy", "pmc_last_error", "pmc_host_alloc", "pmc_host_free", "pmc_set_stream", "pmc_synchronize", "pmc_set_tolerances",
    "pmc_set_preconditioner", "pmc_set_option", "pmc_set_batch", "pmc_upload_sampler_level", "pmc_upload_darcy_level", "pmc_upload_field_transfer", "pmc_clone", "pmc_prepare",
    "pmc_rng_init", "pmc_rng_fill_int", "pmc_rng_fill", "pmc_rng_map", "pmc_sampler_sample_batch", "pmc_sampler_eval_batch",
    "pmc_darcy_solve_batch", "pmc_darcy_apply_batch", "pmc_mlmc_level_batch", "pmc_mc_level_batch",
    "pmc_upload_observations", "pmc_bayes_level_batch", "pmc_comm_unique_id", "pmc_comm_init", "pmc_allreduce_sums",
    "pmc_comm_destroy", "pmc_reset_stats", "pmc_kernel_stats",
]


class PmcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pmc_b200 error {code}: {msg}")
        self.code = code


class KernelStats(C.Structure):
    _fields_ = [("launches", C.c_int64 * 10), ("algo_bytes", C.c_double * 10), ("ms", C.c_double * 10),
                ("timed_launches", C.c_int64 * 10), ("class_cycle_share", C.c_double * 10),
                ("kernel_launches", C.c_int64), ("other_launches", C.c_int64), ("kernel_ms", C.c_double),
                ("kernel_algo_bytes", C.c_double), ("ops_executed", C.c_int64), ("minres_iterations", C.c_int64)]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(_HERE, "..", "include", "pmc_b200.h")]
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", CSRC], stdout=subprocess.DEVNULL)
    # the C++ host layer (reference-style classes and drivers over the C ABI)
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "host")], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def load():
    """Load the CUDA library; raises (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"or `make -C parelagmc_b200/csrc`.  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.pmc_create.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
    L.pmc_destroy.argtypes = [vp]
    L.pmc_destroy.restype = None
    L.pmc_last_error.argtypes = [vp]
    L.pmc_last_error.restype = C.c_char_p
    L.pmc_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.pmc_host_free.argtypes = [vp]
    L.pmc_host_free.restype = None
    L.pmc_set_stream.argtypes = [vp, vp]
    L.pmc_synchronize.argtypes = [vp]
    L.pmc_set_tolerances.argtypes = [vp, C.c_double, C.c_double, C.c_int]
    L.pmc_set_preconditioner.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double]
    L.pmc_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    L.pmc_set_batch.argtypes = [vp, C.c_int, C.c_int]
    L.pmc_upload_sampler_level.argtypes = [vp, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _ip, _ip, _dp, _dp,
                                           C.c_int, _ip, _ip, _dp, C.c_double, C.c_double, C.c_int]
    L.pmc_upload_darcy_level.argtypes = [vp, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _ip, _ip, _dp, _ip, _dp,
                                         _dp, _dp, C.c_int, _ip, _ip, _dp]
    L.pmc_upload_field_transfer.argtypes = [vp, C.c_int, C.c_int, _ip, _ip, _dp, _dp]
    L.pmc_clone.argtypes = [vp, C.POINTER(vp)]
    L.pmc_prepare.argtypes = [vp]
    L.pmc_rng_init.argtypes = [vp, C.c_double, C.c_double, C.c_int, C.c_int]
    L.pmc_rng_fill_int.argtypes = [vp, C.c_uint64, C.c_int64, C.POINTER(C.c_int32)]
    L.pmc_rng_fill.argtypes = [vp, C.c_uint64, C.c_int64, _dp]
    L.pmc_rng_map.argtypes = [vp, C.c_int64, C.POINTER(C.c_int32), _dp]
    L.pmc_sampler_sample_batch.argtypes = [vp, C.c_int, C.c_int, C.c_uint64, _dp]
    L.pmc_sampler_eval_batch.argtypes = [vp, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int, _dp, _dp, _ip]
    L.pmc_darcy_solve_batch.argtypes = [vp, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _ip]
    L.pmc_darcy_apply_batch.argtypes = [vp, C.c_int, C.c_int, _dp, _dp, _dp]
    L.pmc_mlmc_level_batch.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_uint64, _dp, _dp, C.POINTER(C.c_int64)]
    L.pmc_mc_level_batch.argtypes = [vp, C.c_int, C.c_int, C.c_uint64, _dp, _dp, C.POINTER(C.c_int64)]
    L.pmc_upload_observations.argtypes = [vp, C.c_int, C.c_int, _dp, _dp, C.c_double]
    L.pmc_bayes_level_batch.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_uint64, _dp, _dp, C.POINTER(C.c_int64)]
    L.pmc_comm_unique_id.argtypes = [C.c_char_p]
    L.pmc_comm_init.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
    L.pmc_allreduce_sums.argtypes = [vp, _dp, C.c_int]
    L.pmc_comm_destroy.argtypes = [vp]
    L.pmc_reset_stats.argtypes = [vp]
    L.pmc_kernel_stats.argtypes = [vp, C.POINTER(KernelStats)]
    for name in SYMBOLS:
        f = getattr(L, name)
        if name not in ("pmc_destroy", "pmc_last_error", "pmc_host_free"):
            f.restype = C.c_int
    _lib = L
    return L


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _csr(m):
    if m is None:
        return None, None, None, ()
    rp = np.ascontiguousarray(m.indptr, dtype=np.int32)
    ci = np.ascontiguousarray(m.indices, dtype=np.int32)
    v = np.ascontiguousarray(m.data, dtype=np.float64)
    return _i(rp), _i(ci), _d(v), (rp, ci, v)


class _PinnedBlock:
    def __init__(self, ptr, lib):
        self.ptr, self.lib = ptr, lib

    def __del__(self):
        try:
            self.lib.pmc_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """numpy array over page-locked host memory (`pmc_host_alloc`); freed with the array."""
    L = load()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    rc = L.pmc_host_alloc(n, C.byref(p))
    if rc != 0:
        raise PmcError(rc, (L.pmc_last_error(None) or b"").decode())
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    buf._owner = _PinnedBlock(p, L)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


COMM_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """`pmc_comm_unique_id`: the opaque NCCL id rank 0 creates and ships to the other ranks."""
    L = load()
    buf = C.create_string_buffer(COMM_ID_BYTES)
    rc = L.pmc_comm_unique_id(buf)
    if rc != 0:
        raise PmcError(rc, (L.pmc_last_error(None) or b"").decode())
    return buf.raw


class Context:
    """One `pmc_handle`: a GPU-resident hierarchy plus the batched per-sample path on it."""

    def __init__(self, nlevels: int, device: int = 0, _handle=None):
        self._L = load()
        self._h = C.c_void_p()
        if _handle is not None:
            self._h = _handle
        else:
            rc = self._L.pmc_create(device, nlevels, C.byref(self._h))
            if rc != 0:
                raise PmcError(rc, (self._L.pmc_last_error(None) or b"").decode())
        self.nlevels = nlevels
        self.device = device
        self.Ne = [0] * nlevels      # sampler sizes (noise / Gaussian field)
        self.Nf = [0] * nlevels
        self.Nout = [0] * nlevels    # size of the sampler's output field (= Darcy Ne)
        self.dNe = [0] * nlevels     # Darcy sizes
        self.dNf = [0] * nlevels

    def close(self):
        if getattr(self, "_h", None):
            self._L.pmc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc != 0:
            raise PmcError(rc, (self._L.pmc_last_error(self._h) or b"").decode())

    # -- configuration -----------------------------------------------------------------------------
    def set_stream(self, cuda_stream: int):
        self._ck(self._L.pmc_set_stream(self._h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._ck(self._L.pmc_synchronize(self._h))

    def set_tolerances(self, rel: float, abs_: float, maxit: int):
        self._ck(self._L.pmc_set_tolerances(self._h, rel, abs_, maxit))

    def set_preconditioner(self, mass_degree=0, schur_degree=0, schur_ratio=0.0, coarse_degree=0, coarse_ratio=0.0):
        self._ck(self._L.pmc_set_preconditioner(self._h, mass_degree, schur_degree, schur_ratio, coarse_degree,
                                                coarse_ratio))

    def set_option(self, key: str, value: float):
        self._ck(self._L.pmc_set_option(self._h, key.encode(), float(value)))

    def set_batch(self, max_batch: int = 0, check_every: int = 0):
        self._ck(self._L.pmc_set_batch(self._h, max_batch, check_every))

    # -- uploads -----------------------------------------------------------------------------------
    def upload_sampler_level(self, level: int, s, alpha: float, g: float, lognormal: bool):
        """`s`: a `hierarchy.SamplerLevel`-shaped object (M, B, Wdiag, P in scipy CSR / numpy)."""
        mr, mc, mv, k1 = _csr(s.M)
        br, bc, bv, k2 = _csr(s.B)
        pr, pc, pv, k3 = _csr(s.P)
        wd = np.ascontiguousarray(s.Wdiag, dtype=np.float64)
        self._ck(self._L.pmc_upload_sampler_level(self._h, level, s.Ne, s.Nf, mr, mc, mv, br, bc, bv, _d(wd),
                                                  0 if s.P is None else s.P.shape[1], pr, pc, pv, alpha, g,
                                                  1 if lognormal else 0))
        self.Ne[level], self.Nf[level] = s.Ne, s.Nf
        self.Nout[level] = s.Ne
        if getattr(s, "T", None) is not None:
            tr, tc, tv, k4 = _csr(s.T)
            ts = None if s.Tscale is None else np.ascontiguousarray(s.Tscale, dtype=np.float64)
            self._ck(self._L.pmc_upload_field_transfer(self._h, level, s.T.shape[0], tr, tc, tv, _d(ts)))
            self.Nout[level] = s.T.shape[0]

    def upload_darcy_level(self, level: int, d):
        """`d`: a `hierarchy.DarcyLevel`-shaped object."""
        br, bc, bv, k2 = _csr(d.B)
        pr, pc, pv, k3 = _csr(d.P_p)
        ep = np.ascontiguousarray(d.elem_ptr, dtype=np.int32)
        ed = np.ascontiguousarray(d.elem_dofs, dtype=np.int32)
        em = np.ascontiguousarray(d.elem_mat, dtype=np.float64)
        eu = np.ascontiguousarray(d.ess_u, dtype=np.int32)
        ev = np.ascontiguousarray(d.ess_data, dtype=np.float64)
        rh = np.ascontiguousarray(d.rhs, dtype=np.float64)
        ob = np.ascontiguousarray(d.obs, dtype=np.float64)
        self._ck(self._L.pmc_upload_darcy_level(self._h, level, d.Ne, d.Nf, _i(ep), _i(ed), _d(em), br, bc, bv,
                                                _i(eu), _d(ev), _d(rh), _d(ob),
                                                0 if d.P_p is None else d.P_p.shape[1], pr, pc, pv))
        self.dNe[level], self.dNf[level] = d.Ne, d.Nf
        if self.Ne[level] == 0:
            self.Ne[level], self.Nf[level], self.Nout[level] = d.Ne, d.Nf, d.Ne

    def clone(self) -> "Context":
        """`pmc_clone`: same hierarchy, options and stream of random numbers; own CUDA stream and workspace."""
        h = C.c_void_p()
        self._ck(self._L.pmc_clone(self._h, C.byref(h)))
        c = Context(self.nlevels, self.device, _handle=h)
        c.Ne, c.Nf, c.Nout, c.dNe, c.dNf = list(self.Ne), list(self.Nf), list(self.Nout), list(self.dNe), list(self.dNf)
        return c

    def prepare(self):
        self._ck(self._L.pmc_prepare(self._h))

    # -- RNG ---------------------------------------------------------------------------------------
    def rng_init(self, mu: float = 0.0, sigma: float = 1.0, nparts: int = 1, mypart: int = 0):
        self._ck(self._L.pmc_rng_init(self._h, mu, sigma, nparts, mypart))

    def rng_fill_int(self, pos: int, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.int32)
        self._ck(self._L.pmc_rng_fill_int(self._h, C.c_uint64(pos), n, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def rng_fill(self, pos: int, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.float64)
        self._ck(self._L.pmc_rng_fill(self._h, C.c_uint64(pos), n, _d(out)))
        return out

    def rng_map(self, engine: np.ndarray) -> np.ndarray:
        """Normal deviates of caller-chosen engine outputs (`pmc_rng_map`)."""
        engine = np.ascontiguousarray(engine, dtype=np.int32)
        out = np.empty(engine.shape[0], dtype=np.float64)
        self._ck(self._L.pmc_rng_map(self._h, engine.shape[0], engine.ctypes.data_as(C.POINTER(C.c_int32)), _d(out)))
        return out

    # -- sampler -----------------------------------------------------------------------------------
    def sampler_sample_batch(self, level: int, nsamples: int, pos0: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        out = np.empty((nsamples, self.Ne[level]), dtype=np.float64) if out is None else out
        assert out.shape == (nsamples, self.Ne[level]) and out.flags.c_contiguous
        self._ck(self._L.pmc_sampler_sample_batch(self._h, level, nsamples, C.c_uint64(pos0), _d(out)))
        return out

    def sampler_eval_batch(self, level: int, xi: Optional[np.ndarray], xi_level: Optional[int] = None,
                           init_s: Optional[np.ndarray] = None, init_level: int = 0, use_init: int = -1,
                           want_embed: bool = True, out_s: Optional[np.ndarray] = None,
                           out_embed: Optional[np.ndarray] = None, nsamples: Optional[int] = None):
        """Returns (s [n, Ne], embed_s [n, Ne] | None, iters [n]).  With the option "cache_results" set, xi = None
        (give xi_level and nsamples) takes the noise of the preceding `sampler_sample_batch` from device memory, and
        use_init > 0 with init_s = None the Gaussian field of the preceding `sampler_eval_batch`."""
        if xi is None:
            assert xi_level is not None and nsamples is not None
            n = int(nsamples)
        else:
            xi = np.ascontiguousarray(xi, dtype=np.float64)
            if xi.ndim == 1:
                xi = xi[None, :]
            n = xi.shape[0]
            if xi_level is None:
                xi_level = self.Ne.index(xi.shape[1])
            assert xi.shape[1] == self.Ne[xi_level]
        if init_s is not None:
            init_s = np.ascontiguousarray(init_s, dtype=np.float64).reshape(n, -1)
            assert init_s.shape[1] == self.Ne[init_level]
        s = np.empty((n, self.Nout[level])) if out_s is None else out_s
        emb = (np.empty((n, self.Ne[level])) if out_embed is None else out_embed) if want_embed else None
        assert s.shape == (n, self.Nout[level]) and s.flags.c_contiguous
        it = np.zeros(n, dtype=np.int32)
        self._ck(self._L.pmc_sampler_eval_batch(self._h, level, xi_level, n, _d(xi), _d(init_s), init_level,
                                                use_init, _d(s), _d(emb), _i(it)))
        return s, emb, it

    # -- Darcy -------------------------------------------------------------------------------------
    def darcy_solve_batch(self, level: int, k: Optional[np.ndarray], want_sol: bool = False, nsamples: Optional[int] = None):
        """Returns (Q [n], C [n], sol [n, N] | None, iters [n]).  With the option "cache_results" set, k = None (give
        nsamples) takes the coefficient the preceding `sampler_eval_batch` on this level produced from device memory."""
        if k is None:
            assert nsamples is not None
            n = int(nsamples)
        else:
            k = np.ascontiguousarray(k, dtype=np.float64)
            if k.ndim == 1:
                k = k[None, :]
            n = k.shape[0]
            assert k.shape[1] == self.dNe[level]
        Q = np.empty(n)
        Cc = np.empty(n)
        sol = np.empty((n, self.dNe[level] + self.dNf[level])) if want_sol else None
        it = np.zeros(n, dtype=np.int32)
        self._ck(self._L.pmc_darcy_solve_batch(self._h, level, n, _d(k), _d(Q), _d(Cc), _d(sol), _i(it)))
        return Q, Cc, sol, it

    def darcy_apply_batch(self, level: int, k: np.ndarray, x: np.ndarray) -> np.ndarray:
        k = np.ascontiguousarray(k, dtype=np.float64).reshape(-1, self.dNe[level])
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(k.shape[0], -1)
        y = np.empty_like(x)
        self._ck(self._L.pmc_darcy_apply_batch(self._h, level, k.shape[0], _d(k), _d(x), _d(y)))
        return y

    # -- manager loops -----------------------------------------------------------------------------
    def mlmc_level_batch(self, level: int, nsamples: int, pos0: int, nlevels: Optional[int] = None,
                         want_rows: bool = False, sums: Optional[np.ndarray] = None):
        """Returns (sums[9] (accumulated), rows [n,4] | None, total_iters)."""
        sums = np.zeros(9) if sums is None else sums
        rows = np.zeros((nsamples, 4)) if want_rows else None
        its = C.c_int64(0)
        self._ck(self._L.pmc_mlmc_level_batch(self._h, level, nlevels or self.nlevels, nsamples, C.c_uint64(pos0),
                                              _d(sums), _d(rows), C.byref(its)))
        return sums, rows, its.value

    def mc_level_batch(self, level: int, nsamples: int, pos0: int, want_rows: bool = False,
                       sums: Optional[np.ndarray] = None):
        sums = np.zeros(4) if sums is None else sums
        rows = np.zeros((nsamples, 2)) if want_rows else None
        its = C.c_int64(0)
        self._ck(self._L.pmc_mc_level_batch(self._h, level, nsamples, C.c_uint64(pos0), _d(sums), _d(rows),
                                            C.byref(its)))
        return sums, rows, its.value

    # -- Bayesian inverse problem ------------------------------------------------------------------
    def upload_observations(self, level: int, g: np.ndarray, G_obs: np.ndarray, noise: float):
        g = np.ascontiguousarray(g, dtype=np.float64).reshape(-1, self.dNe[level])
        G_obs = np.ascontiguousarray(G_obs, dtype=np.float64)
        assert g.shape[0] == G_obs.shape[0]
        self._ck(self._L.pmc_upload_observations(self._h, level, g.shape[0], _d(g), _d(G_obs), float(noise)))

    def bayes_level_batch(self, level: int, nsamples: int, pos0: int, nlevels: Optional[int] = None,
                          want_rows: bool = False, sums: Optional[np.ndarray] = None):
        """Returns (sums[20] (accumulated, ML_BayesRatio_Manager enum layout), rows [n,5] | None, total_iters)."""
        sums = np.zeros(20) if sums is None else sums
        rows = np.zeros((nsamples, 5)) if want_rows else None
        its = C.c_int64(0)
        self._ck(self._L.pmc_bayes_level_batch(self._h, level, nlevels or self.nlevels, nsamples, C.c_uint64(pos0),
                                               _d(sums), _d(rows), C.byref(its)))
        return sums, rows, its.value

    # -- instrumentation ---------------------------------------------------------------------------
    # -- ranks sharing a sample budget (NCCL behind the C ABI) --------------------------------------
    def comm_init(self, nranks: int, rank: int, unique_id: Optional[bytes]):
        """`pmc_comm_init` (collective over the ranks).  `unique_id`: the bytes `comm_unique_id()` returned on rank 0."""
        self._ck(self._L.pmc_comm_init(self._h, nranks, rank, unique_id))

    def allreduce_sums(self, sums: np.ndarray) -> np.ndarray:
        """`pmc_allreduce_sums`: in-place sum over the ranks of a contiguous float64 array."""
        assert sums.dtype == np.float64 and sums.flags.c_contiguous
        self._ck(self._L.pmc_allreduce_sums(self._h, _d(sums), int(sums.size)))
        return sums

    def comm_destroy(self):
        self._ck(self._L.pmc_comm_destroy(self._h))

    def kernel_ms(self) -> float:
        """CUDA-event time (ms) of all persistent-kernel launches of this handle since the last `reset_stats`."""
        st = KernelStats()
        self._ck(self._L.pmc_kernel_stats(self._h, C.byref(st)))
        return float(st.kernel_ms)

    def reset_stats(self):
        self._ck(self._L.pmc_reset_stats(self._h))

    def kernel_stats(self) -> dict:
        st = KernelStats()
        self._ck(self._L.pmc_kernel_stats(self._h, C.byref(st)))
        out = {n: dict(launches=st.launches[i], algo_bytes=st.algo_bytes[i], ms=st.ms[i], ops=st.timed_launches[i],
                       cycle_share=st.class_cycle_share[i]) for i, n in enumerate(K_CLASSES)}
        out["kernel"] = dict(launches=st.kernel_launches, ms=st.kernel_ms, algo_bytes=st.kernel_algo_bytes,
                             ops_executed=st.ops_executed, minres_iterations=st.minres_iterations,
                             other_launches=st.other_launches)
        return out
