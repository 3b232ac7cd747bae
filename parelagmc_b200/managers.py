"""Host-side Monte-Carlo managers, mirroring the reference's `MLMC_Manager` / `MC_Manager` over the batched C ABI.

Same names, arguments and behaviour as `/root/reference/src/MLMC_Manager.{hpp,cpp}` and
`/root/reference/src/MC_Manager.{hpp,cpp}` (`Run`, `InitRun`, `ShowMe`, `wallTime`), except that the inner sample
loop (`MLMC_Manager.cpp:110-175`, `MC_Manager.cpp:82-116`) is ONE batched call per level into the GPU library and
that realisations are sharded over the ranks of `comm` (a `torch.distributed` process group or `None`): each rank
processes a contiguous slice of every level's sample budget and the per-level sums are combined with one
all-reduce per `InitRun` (there is no other communication).  The C++ twin of this file, used by the reference-style
drivers, is `parelagmc_b200/host/`.
"""
from __future__ import annotations

import math
import sys
import time
from typing import List, Optional, Sequence

import numpy as np

# enum of /root/reference/src/MLMC_Manager.hpp:65
Y2, Y, ABSY, Q2, Q, ABSQ, C, Y3, Y4, NVAR = range(10)
# enum of /root/reference/src/MC_Manager.hpp:61
SL_Q2, SL_Q, SL_ABSQ, SL_C, SL_NVAR = range(5)  # SL_NVAR == 4 sums


def expWRegression(y: Sequence[float], x: Sequence[float], skip_n_last: int) -> float:
    """`/root/reference/src/Utilities.cpp:257-283`: weighted log-log regression slope, weights 0.5^i."""
    n = len(y) - 1 - skip_n_last
    if n < 1:
        return 0.0
    num = den = 0.0
    for i in range(n):
        logdy = math.log(abs(y[i] / y[i + 1]))
        logdx = math.log(x[i] / x[i + 1])
        w = 0.5 ** i
        num += logdy * w * logdx
        den += logdx * w * logdx
    return num / den


def split_samples(n: int, rank: int, world: int):
    """Contiguous slice [first, first+count) of n realisations owned by `rank`."""
    base, rem = divmod(n, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


class _Comm:
    """The ranks that share a sample budget: a single process, a `torch.distributed` group (any backend; gloo in the CPU
    tests), or -- `_Comm.over_capi` -- the NCCL communicator behind the C ABI (`pmc_comm_init` / `pmc_allreduce_sums`),
    which is what a C/C++ host program uses."""

    def __init__(self, group=None, use_dist: bool = False, device=None):
        self.use_dist = use_dist
        self.group = group
        self.device = device
        self.ctx = None
        if use_dist:
            import torch.distributed as dist
            self.dist = dist
            self.rank = dist.get_rank(group)
            self.size = dist.get_world_size(group)
        else:
            self.rank, self.size = 0, 1

    @classmethod
    def over_capi(cls, ctx, nranks: int, rank: int, unique_id):
        """All-reduce through the library's own NCCL communicator (collective: every rank calls it)."""
        c = cls()
        ctx.comm_init(nranks, rank, unique_id)
        c.ctx, c.rank, c.size = ctx, rank, nranks
        return c

    def allreduce_sum(self, a: np.ndarray) -> np.ndarray:
        if self.size == 1:
            return a
        if self.ctx is not None:
            return self.ctx.allreduce_sums(np.ascontiguousarray(a, dtype=np.float64).copy())
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
        if self.device is not None:
            t = t.to(self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def reduce_sums_and_times(self, local: np.ndarray, times: np.ndarray):
        """One collective for the per-level sums AND the per-level timings of this round: the sample allocation
        (`computeNSamplesMSE`) must see identical inputs on every rank, otherwise the ranks request different sample
        counts, their slices overlap or leave gaps and one of them can leave the adaptive loop alone.  Returns the summed
        sums and the rank-MEAN of the times."""
        if self.size == 1:
            return local, times
        buf = np.concatenate([np.ravel(local), np.ravel(times)]).astype(np.float64)
        buf = self.allreduce_sum(buf)
        return buf[:local.size].reshape(local.shape), buf[local.size:].reshape(times.shape) / self.size


def _check_two_samples(level_nsamples):
    """The unbiased variance n/(n-1) needs two realisations per level; the reference would go on with inf/NaN
    (`/root/reference/src/MLMC_Manager.cpp:319-321`) and report a converged run."""
    bad = [int(i) for i, n in enumerate(np.atleast_1d(level_nsamples)) if n < 2]
    if bad:
        raise ValueError(f"at least 2 samples per level are needed for the variance estimate (levels {bad} have fewer)")


class MLMC_Manager:
    """`parelagmc::MLMC_Manager` (`/root/reference/src/MLMC_Manager.hpp:30-61`).

    backend: object with `mlmc_level_batch(level, nsamples, pos0, nlevels=..., want_rows=..., sums=...)`,
    `Ne[level]` and `Nf[level]` (a `capi.Context`).  params: the "Problem parameters" keys read at
    `MLMC_Manager.cpp:29-36`."""

    def __init__(self, comm: Optional[_Comm], nlevels: int, backend, params: Optional[dict] = None, out=sys.stdout):
        params = params or {}
        self.wallTime = True
        self.comm = comm or _Comm()
        self.nlevels = nlevels
        self.backend = backend
        self.eps2 = float(params.get("Mean square error", 0.001))
        self.auto_eps2 = self.eps2 < 0
        self.ratio = float(params.get("MSE splitting ratio", 0.5))
        self.file_name = params.get("Output filename for MC managers", "MLMC.dat")
        self.init_nsamples = int(params.get("Number of samples", 10))
        self.use_array_samples = bool(params.get("Use array samples", False))
        self.v_init_nsamples = list(params.get("Array number of samples", []))
        if self.use_array_samples and len(self.v_init_nsamples) != nlevels:
            self.use_array_samples = False
        if not self.use_array_samples:
            self.v_init_nsamples = [self.init_nsamples] * nlevels
        self.out = out
        self.pid = self.comm.rank
        dNe, dNf = getattr(backend, "dNe", backend.Ne), getattr(backend, "dNf", backend.Nf)
        self.M = np.array([dNe[i] + dNf[i] for i in range(nlevels)], dtype=np.float64)   # pSolver.GetGlobalNumberOfDofs(i)
        self.logger = open(self.file_name, "w") if (self.pid == 0 and self.file_name) else None
        self.stream_pos = 0            # absolute yarn5 position of the next unused draw (all ranks agree)
        self.concurrent_levels = True  # run the level loops of an InitRun concurrently (one cloned handle per level)
        self._clones = {}
        self._pool = None
        self._reset()
        if self.pid == 0 and out is not None:
            print("\n" + "*" * 50 + "\n*  MLMC_Manager \n*    MSE: %g\n*    MSE splitting ratio: %g\n"
                  "*    Number of Initial Samples: %s \n*    Output filename: %s\n" % (
                      self.eps2, self.ratio, " ".join(map(str, self.v_init_nsamples)), self.file_name) + "*" * 50,
                  file=out)

    def _level_backend(self, ilevel):
        if not (self.concurrent_levels and self.nlevels > 1 and hasattr(self.backend, "clone")):
            return self.backend
        if ilevel == 0:
            return self.backend
        if ilevel not in self._clones:
            self._clones[ilevel] = self.backend.clone()
        return self._clones[ilevel]

    def _reset(self):
        L = self.nlevels
        self.sums = np.zeros((L, NVAR))
        self.level_nsamples = np.zeros(L, dtype=np.int64)
        self.level_nsamples_missing = np.zeros(L, dtype=np.int64)
        self.level_time = np.zeros(L)
        self.ml_estimator_variance = math.inf
        self.expected_discretization_error2 = math.inf
        self.actualMSE = math.inf
        self.total_iters = 0
        for n in ("eY", "eABSY", "eQ", "eABSQ", "eC", "varY", "varQ", "consistency", "kurtosis", "VC"):
            setattr(self, n, np.zeros(L))
        self.alpha = self.alphaABS = self.beta = self.gamma = 0.0

    # --------------------------------------------------------------------------------------------
    def InitRun(self, level_nsamples_init: Sequence[int]):
        """`MLMC_Manager.cpp:103-179`: coarsest level first, then nlevels-2 .. 0; sums accumulate."""
        first_call = self.level_nsamples.max() == 0
        if first_call and self.logger:
            self.logger.write("%" + "level ".rjust(13) + "Y(xi) ".rjust(14) + "Q(xi)".rjust(14) + "Q_c(xi)".rjust(14)
                              + "c \n".rjust(14))
        local = np.zeros((self.nlevels, NVAR))
        want_rows = self.logger is not None and self.comm.size == 1
        # stream positions in the reference's order: coarsest level first, every level after the previous one
        jobs = []
        for ilevel in range(self.nlevels - 1, -1, -1):
            n = int(level_nsamples_init[ilevel])
            Ne = self.backend.Ne[ilevel]
            first, count = split_samples(n, self.comm.rank, self.comm.size)
            jobs.append((ilevel, count, self.stream_pos + first * Ne))
            self.stream_pos += n * Ne
            self.level_nsamples[ilevel] += n

        # handles are created here, on the calling thread, before any worker uses them (pmc_clone prepares its source)
        backends = {ilevel: self._level_backend(ilevel) for ilevel, count, _ in jobs if count > 0}

        def run(job):
            ilevel, count, pos = job
            t0 = time.perf_counter()
            rows, its, dt = None, 0, None
            if count > 0:
                be = backends[ilevel]
                dev0 = be.kernel_ms() if hasattr(be, "kernel_ms") else None
                _, rows, its = be.mlmc_level_batch(ilevel, count, pos, nlevels=self.nlevels, want_rows=want_rows,
                                                   sums=local[ilevel])
                if dev0 is not None:      # device time of this level's launches: free of the host-side queueing behind
                    dt = 1e-3 * (be.kernel_ms() - dev0)     # the other levels' threads (they still share the GPU)
            return ilevel, rows, its, (time.perf_counter() - t0) if dt is None else dt

        if self.concurrent_levels and self.nlevels > 1 and hasattr(self.backend, "clone"):
            # the level loops are independent: one handle (own stream and workspace) and one host thread per level
            from concurrent.futures import ThreadPoolExecutor
            if self._pool is None:
                self._pool = ThreadPoolExecutor(max_workers=self.nlevels)
            results = list(self._pool.map(run, jobs))
        else:
            results = [run(j) for j in jobs]
        round_time = np.zeros(self.nlevels)
        for ilevel, rows, its, dt in results:
            self.total_iters += its
            round_time[ilevel] += dt
            if want_rows and rows is not None:
                for r in rows:
                    self.logger.write(f"{ilevel:14d}{r[0]:14.6g}{r[1]:14.6g}{r[2]:14.6g}{r[3]:14.6g}\n")
        gsums, gtime = self.comm.reduce_sums_and_times(local, round_time)
        self.sums += gsums
        self.level_time += gtime * self.comm.size    # total time of the round over the ranks (n counts all ranks' samples)
        if self.logger:
            self.logger.flush()
        self.computeNSamplesMSE()

    def Run(self):
        """`MLMC_Manager.cpp:181-214`."""
        self._reset()
        self.InitRun(self.v_init_nsamples)
        grain = [0] * self.nlevels
        while self.ml_estimator_variance > self.ratio * self.eps2:
            for i in range(self.nlevels):
                grain[i] = min(int(self.level_nsamples_missing[i]),
                               self.v_init_nsamples[i] + grain[i] + int(self.level_nsamples_missing[i]) // 10)
            if sum(grain) == 0:
                break
            self.InitRun(grain)
        if self.pid == 0 and self.out is not None:
            print("FINAL MLMC ERRORS", file=self.out)
        self.ShowMe()

    # --------------------------------------------------------------------------------------------
    def computeNSamplesMSE(self):
        """`MLMC_Manager.cpp:300-401`."""
        L = self.nlevels
        _check_two_samples(self.level_nsamples)
        n = self.level_nsamples.astype(np.float64)
        e = self.sums / n[:, None]
        self.eY, self.eABSY, self.eQ, self.eABSQ, self.eC = (e[:, Y].copy(), e[:, ABSY].copy(), e[:, Q].copy(),
                                                             e[:, ABSQ].copy(), e[:, C].copy())
        varY, varQ, kurt = e[:, Y2].copy(), e[:, Q2].copy(), e[:, Y4].copy()
        with np.errstate(divide="ignore", invalid="ignore"):
            kurt = kurt / (varY * varY)          # note: before the mean is subtracted (:318-319)
            varY = (varY - self.eY ** 2) * n / (n - 1.0)
            varQ = (varQ - self.eQ ** 2) * n / (n - 1.0)
        self.varY, self.varQ, self.kurtosis = varY, varQ, kurt
        self.consistency = np.zeros(L)
        for l in range(L - 1):
            with np.errstate(divide="ignore", invalid="ignore"):
                self.consistency[l] = abs(self.eQ[l] - self.eQ[l + 1] + self.eY[l]) / (
                    3 * (math.sqrt(max(varQ[l], 0)) + math.sqrt(max(varQ[l + 1], 0)) + math.sqrt(max(varY[l], 0))))
        M = self.M
        self.alpha = expWRegression(self.eY, M, 1)
        self.alphaABS = expWRegression(self.eABSY, M, 1)
        self.beta = expWRegression(varY, M, 1)
        if L == 1:
            bias2 = 0.0
        else:
            m = M[0] / M[1]
            if L > 3:
                bias2 = max(m ** (2.0 * self.alphaABS) * self.eABSY[1] ** 2, self.eABSY[0] ** 2) / (
                    (m ** (-2.0 * self.alphaABS) - 1.0) ** 2)
            elif L == 3:
                bias2 = self.eABSY[0] ** 2 / ((m ** (-self.alphaABS) - 1.0) ** 2)
            else:
                bias2 = self.eABSY[0] ** 2
        self.expected_discretization_error2 = bias2
        if self.auto_eps2:
            self.eps2 = bias2 / (1.0 - self.ratio)
        self.ml_estimator_variance = float(np.sum(varY / n))
        self.actualMSE = bias2 + self.ml_estimator_variance
        cost = (self.level_time / n) if self.wallTime else self.eC
        self.cost = np.asarray(cost, dtype=np.float64)
        self.gamma = expWRegression(self.cost, M, 0)
        prop = float(np.sum(np.sqrt(np.maximum(varY, 0) * self.cost))) / (self.ratio * self.eps2)
        for i in range(L):
            missings = prop * math.sqrt(max(varY[i], 0) / self.cost[i]) - n[i]
            self.level_nsamples_missing[i] = max(int(math.ceil(missings)), 0)
            self.VC[i] = varY[i] * self.cost[i]
        self.ShowMe()

    def ShowMe(self, os=None):
        """`MLMC_Manager.cpp:216-297` (same labels, widths and order)."""
        os = os or self.out
        if self.pid != 0 or os is None:
            return
        W, NW = 79, 42

        def row(name, v):
            os.write(name.ljust(NW) + ("%.8g" % v).ljust(18) + "\n")

        def vec(name, v, fmt="%.8g"):
            os.write(name.ljust(NW) + "  ".join(fmt % x for x in v) + "\n")

        os.write("=" * W + "\nMLMC Manager Errors: \n" + "-" * W + "\n")
        row("Estimate", float(np.sum(self.eY)))
        row("Target MSE", self.eps2)
        row("Actual MSE", self.actualMSE)
        row("ML Estimator Variance", self.ml_estimator_variance)
        row("Estimator Bias", self.expected_discretization_error2)
        row("Alpha", self.alpha)
        row("AlphaAbs", self.alphaABS)
        row("Beta", self.beta)
        row("Gamma", self.gamma)
        os.write("\n")
        vec("DOFS in Forward Problem", self.M)
        vec("C_l ", self.eC)
        os.write("\n")
        vec("NumSamples ", self.level_nsamples, "%d")
        os.write("\n")
        vec("E[Y_l] ", self.eY)
        vec("E[|Y_l|] ", self.eABSY)
        vec("Var[Y_l] ", self.varY)
        vec("E[Q_l] ", self.eQ)
        vec("E[|Q_l|] ", self.eABSQ)
        vec("Var[Q_l] ", self.varQ)
        vec("V[Y_l]*C_l ", self.VC)
        vec("Consistency ", self.consistency)
        vec("Kurtosis", self.kurtosis)
        os.write("=" * W + "\n")
        os.flush()


class MC_Manager:
    """`parelagmc::MC_Manager` (`/root/reference/src/MC_Manager.hpp:28-57`): single level (level 0)."""

    def __init__(self, comm: Optional[_Comm], backend, params: Optional[dict] = None, out=sys.stdout, level: int = 0):
        params = params or {}
        self.wallTime = True
        self.comm = comm or _Comm()
        self.backend = backend
        self.level = level
        self.eps2 = float(params.get("Mean square error", 0.001))
        self.auto_eps2 = self.eps2 < 0
        self.ratio = float(params.get("MSE splitting ratio", 0.5))
        self.file_name = params.get("Output filename for MC managers", "MLMC.dat")
        self.init_nsamples = int(params.get("Number of samples", 10))
        self.out = out
        self.pid = self.comm.rank
        self.M = float(getattr(backend, "dNe", backend.Ne)[level] + getattr(backend, "dNf", backend.Nf)[level])
        self.logger = open(self.file_name, "w") if (self.pid == 0 and self.file_name) else None
        self.stream_pos = 0
        self._reset()

    def _reset(self):
        self.sums = np.zeros(SL_NVAR)[:4].copy()
        self.level_nsamples = 0
        self.level_nsamples_missing = 0
        self.time = 0.0
        self.ml_estimator_variance = math.inf
        self.expected_discretization_error2 = math.inf
        self.actualMSE = math.inf
        self.eQ = self.eABSQ = self.eC = self.varQ = 0.0

    def InitRun(self, nsamples: int):
        """`MC_Manager.cpp:82-116`."""
        Ne = self.backend.Ne[self.level]
        first, count = split_samples(int(nsamples), self.comm.rank, self.comm.size)
        local = np.zeros(4)
        t0 = time.perf_counter()
        if count > 0:
            want_rows = self.logger is not None and self.comm.size == 1
            _, rows, _ = self.backend.mc_level_batch(self.level, count, self.stream_pos + first * Ne,
                                                     want_rows=want_rows, sums=local)
            if want_rows:
                for r in rows:
                    self.logger.write(f"{r[0]:14.6g}{r[1]:14.6g}\n")
        self.stream_pos += int(nsamples) * Ne
        self.level_nsamples += int(nsamples)
        gsums, gtime = self.comm.reduce_sums_and_times(local, np.array([time.perf_counter() - t0]))
        self.sums += gsums
        self.time += float(gtime[0]) * self.comm.size
        self.computeNSamplesMSE()

    def Run(self):
        """`MC_Manager.cpp:118-146`."""
        self._reset()
        grain = self.init_nsamples
        self.InitRun(grain)
        grain = 0
        while self.ml_estimator_variance > self.ratio * self.eps2:
            grain = min(self.level_nsamples_missing, self.init_nsamples + grain + self.level_nsamples_missing // 10)
            if grain == 0:
                break
            self.InitRun(grain)
        if self.pid == 0 and self.out is not None:
            print("FINAL SLMC ERRORS", file=self.out)
        self.ShowMe()

    def computeNSamplesMSE(self):
        """`MC_Manager.cpp:194-239`."""
        _check_two_samples(self.level_nsamples)
        nl = float(self.level_nsamples)
        self.eQ = self.sums[SL_Q] / nl
        self.eABSQ = self.sums[SL_ABSQ] / nl
        self.eC = self.sums[SL_C] / nl
        self.varQ = (self.sums[SL_Q2] / nl - self.eQ ** 2) * nl / (nl - 1.0) if nl > 1 else math.inf
        self.expected_discretization_error2 = 0.0
        if self.auto_eps2:
            self.eps2 = self.expected_discretization_error2 / (1.0 - self.ratio)
        self.ml_estimator_variance = self.varQ / nl
        self.actualMSE = self.expected_discretization_error2 + self.ml_estimator_variance
        cost = (self.time / nl) if self.wallTime else self.eC
        prop = math.sqrt(self.varQ * cost) / (self.ratio * self.eps2)
        missings = prop * math.sqrt(self.varQ / cost) - nl
        self.level_nsamples_missing = max(int(math.ceil(missings)), 0)
        self.ShowMe()

    def ShowMe(self, os=None):
        """`MC_Manager.cpp:148-192`."""
        os = os or self.out
        if self.pid != 0 or os is None:
            return
        W, NW = 79, 42

        def row(name, v):
            os.write(name.ljust(NW) + ("%.8g" % v).ljust(18) + "\n")

        os.write("=" * W + "\nSLMC Manager Errors: \n" + "-" * W + "\n")
        row("Estimate", self.eQ)
        row("Target MSE", self.eps2)
        row("Actual MSE", self.actualMSE)
        row("SL Estimator Variance", self.ml_estimator_variance)
        row("Estimator Bias", self.expected_discretization_error2)
        row("Target Bias Error", math.sqrt(max(self.eps2, 0)) / math.sqrt(2))
        row("DOFS in Forward Problem", self.M)
        row("C_l ", self.eC)
        os.write("\n" + "NumSamples ".ljust(NW) + str(self.level_nsamples) + "\n")
        row("E[Q_l] ", self.eQ)
        row("E[|Q_l|] ", self.eABSQ)
        row("Var[Q_l] ", self.varQ)
        os.write("=" * W + "\n")
        os.flush()


# enum of /root/reference/src/ML_BayesRatio_Manager.hpp:67-70
BR = dict(YZ2=0, YZ=1, ABS_YZ=2, Z2=3, Z=4, ABS_Z=5, YR2=6, YR=7, ABS_YR=8, R2=9, R=10, ABS_R=11, C=18, NVAR=20)


class ML_BayesRatio_Manager:
    """`parelagmc::ML_BayesRatio_Manager` (`/root/reference/src/ML_BayesRatio_Manager.hpp:35-130`): multilevel estimator
    of the posterior expectation E[Q | data] = E[R] / E[Z], R = Q * likelihood, Z = likelihood, with independent prior
    draws for R and Z.  With nlevels = 1 it is the single-level `SL_BayesRatio_Manager`.  The inner loop
    (`InitRun`, hpp:313-424) is one batched device call per level (`pmc_bayes_level_batch`)."""

    def __init__(self, comm: Optional[_Comm], nlevels: int, backend, params: Optional[dict] = None, out=sys.stdout,
                 stream_pos: int = 0):
        params = params or {}
        self.wallTime = True
        self.comm = comm or _Comm()
        self.nlevels = nlevels
        self.backend = backend
        self.eps2 = float(params.get("Mean square error", 0.001))
        self.auto_eps2 = self.eps2 < 0
        self.ratio = float(params.get("MSE splitting ratio", 0.5))
        self.init_nsamples = int(params.get("Number of samples", 10))
        self.v_init_nsamples = list(params.get("Array number of samples", [])) or [self.init_nsamples] * nlevels
        self.out = out
        self.pid = self.comm.rank
        dNe, dNf = getattr(backend, "dNe", backend.Ne), getattr(backend, "dNf", backend.Nf)
        self.M = np.array([dNe[i] + dNf[i] for i in range(nlevels)], dtype=np.float64)
        self.stream_pos = stream_pos    # draws consumed before the run (e.g. by GenerateObservationalData)
        self._reset()

    def _reset(self):
        L = self.nlevels
        self.sums = np.zeros((L, BR["NVAR"]))
        self.level_nsamples = np.zeros(L, dtype=np.int64)
        self.level_nsamples_missing = np.zeros(L, dtype=np.int64)
        self.level_time = np.zeros(L)
        self.ml_estimator_variance = math.inf

    def InitRun(self, level_nsamples_init: Sequence[int]):
        local = np.zeros((self.nlevels, BR["NVAR"]))
        round_time = np.zeros(self.nlevels)
        for ilevel in range(self.nlevels - 1, -1, -1):       # coarsest first (hpp:322,366)
            n = int(level_nsamples_init[ilevel])
            Ne = self.backend.Ne[ilevel]
            first, count = split_samples(n, self.comm.rank, self.comm.size)
            t0 = time.perf_counter()
            if count > 0:                                    # two prior draws per realisation
                self.backend.bayes_level_batch(ilevel, count, self.stream_pos + 2 * first * Ne, nlevels=self.nlevels,
                                               sums=local[ilevel])
            round_time[ilevel] += time.perf_counter() - t0
            self.stream_pos += 2 * n * Ne
            self.level_nsamples[ilevel] += n
        gsums, gtime = self.comm.reduce_sums_and_times(local, round_time)
        self.sums += gsums
        self.level_time += gtime * self.comm.size
        self.computeNSamplesMSE()

    def Run(self):
        self._reset()
        self.InitRun(self.v_init_nsamples)
        grain = [0] * self.nlevels
        while self.ml_estimator_variance > self.ratio * self.eps2:
            for i in range(self.nlevels):
                grain[i] = min(int(self.level_nsamples_missing[i]),
                               self.v_init_nsamples[i] + grain[i] + int(self.level_nsamples_missing[i]) // 10)
            if sum(grain) == 0:
                break
            self.InitRun(grain)
        self.ShowMe()

    def computeNSamplesMSE(self):
        """hpp:572-726."""
        _check_two_samples(self.level_nsamples)
        L, n = self.nlevels, self.level_nsamples.astype(np.float64)
        e = self.sums / n[:, None]
        g = lambda k: e[:, BR[k]].copy()
        self.eR, self.eABS_R, self.eYR, self.eABS_YR = g("R"), g("ABS_R"), g("YR"), g("ABS_YR")
        self.eZ, self.eABS_Z, self.eYZ, self.eABS_YZ, self.eC = g("Z"), g("ABS_Z"), g("YZ"), g("ABS_YZ"), g("C")
        with np.errstate(divide="ignore", invalid="ignore"):
            unb = n / (n - 1.0)
            self.varR = (g("R2") - self.eR ** 2) * unb
            self.varYR = (g("YR2") - self.eYR ** 2) * unb
            self.varZ = (g("Z2") - self.eZ ** 2) * unb
            self.varYZ = (g("YZ2") - self.eYZ ** 2) * unb
        cost = (self.level_time / n) if self.wallTime else self.eC
        self.cost = np.asarray(cost, dtype=np.float64)
        M = self.M
        self.alphaABS_R = expWRegression(self.eABS_YR, M, 1)
        self.alphaABS_Z = expWRegression(self.eABS_YZ, M, 1)

        def bias2(eABS, a):
            if L == 1:
                return 0.0
            m = M[0] / M[1]
            if L > 3:
                return max(m ** (2.0 * a) * eABS[1] ** 2, eABS[0] ** 2) / ((m ** (-2.0 * a) - 1.0) ** 2)
            if L == 3:
                return eABS[0] ** 2 / ((m ** (-a) - 1.0) ** 2)
            return eABS[0] ** 2

        self.expected_discretization_error2 = max(bias2(self.eABS_YR, self.alphaABS_R), bias2(self.eABS_YZ, self.alphaABS_Z))
        if self.auto_eps2:
            self.eps2 = self.expected_discretization_error2 / (1.0 - self.ratio)
        self.ml_estimator_variance_Z = float(np.sum(self.varYZ / n))
        self.ml_estimator_variance_R = float(np.sum(self.varYR / n))
        self.ml_estimator_variance = max(self.ml_estimator_variance_Z, self.ml_estimator_variance_R)
        self.actualMSE = self.expected_discretization_error2 + self.ml_estimator_variance
        prop_R = float(np.sum(np.sqrt(np.maximum(self.varYR, 0) * self.cost))) / (self.ratio * self.eps2)
        prop_Z = float(np.sum(np.sqrt(np.maximum(self.varYZ, 0) * self.cost))) / (self.ratio * self.eps2)
        for i in range(L):
            mR = prop_R * math.sqrt(max(self.varYR[i], 0) / self.cost[i]) - n[i]
            mZ = prop_Z * math.sqrt(max(self.varYZ[i], 0) / self.cost[i]) - n[i]
            self.level_nsamples_missing[i] = max(int(math.ceil(mR)), int(math.ceil(mZ)), 0)

    def estimate(self) -> float:
        """Posterior expectation: sum_l E[Y_R,l] / sum_l E[Y_Z,l]."""
        return float(np.sum(self.eYR) / np.sum(self.eYZ))

    def ShowMe(self, os=None):
        os = os or self.out
        if self.pid != 0 or os is None:
            return
        os.write("=" * 79 + "\nML Bayes Ratio Manager: \n" + "-" * 79 + "\n")
        for name, v in (("E[R]/E[Z]", self.estimate()), ("Target MSE", self.eps2), ("Actual MSE", self.actualMSE),
                        ("ML Estimator Variance", self.ml_estimator_variance)):
            os.write(name.ljust(42) + ("%.8g" % v) + "\n")
        for name, v in (("NumSamples ", self.level_nsamples), ("E[Y_R] ", self.eYR), ("Var[Y_R] ", self.varYR),
                        ("E[Y_Z] ", self.eYZ), ("Var[Y_Z] ", self.varYZ), ("C_l ", self.eC)):
            os.write(name.ljust(42) + "  ".join("%.8g" % x for x in v) + "\n")
        os.write("=" * 79 + "\n")


# ratio columns of /root/reference/src/ML_BayesRatio_Splitting_Manager.hpp:67-70
BR.update(YRatio2=12, YRatio=13, ABS_YRatio=14, Ratio2=15, Ratio=16, ABS_Ratio=17)


class ML_BayesRatio_Splitting_Manager(ML_BayesRatio_Manager):
    """`parelagmc::ML_BayesRatio_Splitting_Manager` (`/root/reference/src/ML_BayesRatio_Splitting_Manager.hpp:35-130`):
    the multilevel estimator of E_post[Q] = E[R / Z] ("divide, then sum": the per-realisation ratio q = R / Z and its
    level difference Y = q_l - q_{l+1}, hpp:381-394).  Same device call per level as `ML_BayesRatio_Manager`
    (`pmc_bayes_level_batch`: the rows carry R, Y_R, Z, Y_Z of every realisation); the six ratio sums (hpp:342-347,
    :417-422) are formed from the rows, and the sample allocation follows the ratio's variance (hpp:603-737)."""

    def InitRun(self, level_nsamples_init: Sequence[int]):
        local = np.zeros((self.nlevels, BR["NVAR"]))
        round_time = np.zeros(self.nlevels)
        for ilevel in range(self.nlevels - 1, -1, -1):
            n = int(level_nsamples_init[ilevel])
            Ne = self.backend.Ne[ilevel]
            first, count = split_samples(n, self.comm.rank, self.comm.size)
            t0 = time.perf_counter()
            if count > 0:
                _, rows, _ = self.backend.bayes_level_batch(ilevel, count, self.stream_pos + 2 * first * Ne,
                                                            nlevels=self.nlevels, want_rows=True, sums=local[ilevel])
                R, YR, Z, YZ = rows[:, 0], rows[:, 1], rows[:, 2], rows[:, 3]
                q = R / Z
                if ilevel == self.nlevels - 1:
                    y = q                                             # hpp:342-347
                else:
                    y = q - (R - YR) / (Z - YZ)                       # q - r_c / z_c (hpp:390-394)
                local[ilevel, BR["Ratio"]] += q.sum()
                local[ilevel, BR["ABS_Ratio"]] += np.abs(q).sum()
                local[ilevel, BR["Ratio2"]] += (q * q).sum()
                local[ilevel, BR["YRatio"]] += y.sum()
                local[ilevel, BR["ABS_YRatio"]] += np.abs(y).sum()
                local[ilevel, BR["YRatio2"]] += (y * y).sum()
            round_time[ilevel] += time.perf_counter() - t0
            self.stream_pos += 2 * n * Ne
            self.level_nsamples[ilevel] += n
        gsums, gtime = self.comm.reduce_sums_and_times(local, round_time)
        self.sums += gsums
        self.level_time += gtime * self.comm.size
        self.computeNSamplesMSE()

    def computeNSamplesMSE(self):
        """hpp:603-737: everything `ML_BayesRatio_Manager` reports, but bias, estimator variance and the sample allocation
        from the ratio's level differences."""
        super().computeNSamplesMSE()
        L, n = self.nlevels, self.level_nsamples.astype(np.float64)
        e = self.sums / n[:, None]
        self.eRatio, self.eABS_Ratio = e[:, BR["Ratio"]].copy(), e[:, BR["ABS_Ratio"]].copy()
        self.eYRatio, self.eABS_YRatio = e[:, BR["YRatio"]].copy(), e[:, BR["ABS_YRatio"]].copy()
        unb = n / (n - 1.0)
        self.varRatio = (e[:, BR["Ratio2"]] - self.eRatio ** 2) * unb
        self.varYRatio = (e[:, BR["YRatio2"]] - self.eYRatio ** 2) * unb
        M = self.M
        self.alpha = expWRegression(self.eYRatio, M, 1)
        self.alphaABS = expWRegression(self.eABS_YRatio, M, 1)
        self.beta = expWRegression(self.varYRatio, M, 1)
        self.gamma = expWRegression(self.cost, M, 0)
        if L == 1:
            bias2 = 0.0
        else:
            m = M[0] / M[1]
            if L > 3:
                bias2 = max(m ** (2.0 * self.alphaABS) * self.eABS_YRatio[1] ** 2, self.eABS_YRatio[0] ** 2) / (
                    (m ** (-2.0 * self.alphaABS) - 1.0) ** 2)
            elif L == 3:
                bias2 = self.eABS_YRatio[0] ** 2 / ((m ** (-self.alphaABS) - 1.0) ** 2)
            else:
                bias2 = self.eABS_YRatio[0] ** 2
        self.expected_discretization_error2 = bias2
        if self.auto_eps2:
            self.eps2 = bias2 / (1.0 - self.ratio)
        self.ml_estimator_variance = float(np.sum(self.varYRatio / n))
        self.actualMSE = bias2 + self.ml_estimator_variance
        prop = float(np.sum(np.sqrt(np.maximum(self.varYRatio, 0) * self.cost))) / (self.ratio * self.eps2)
        for i in range(L):
            missings = prop * math.sqrt(max(self.varYRatio[i], 0) / self.cost[i]) - n[i]
            self.level_nsamples_missing[i] = max(int(math.ceil(missings)), 0)

    def estimate(self) -> float:
        """Posterior expectation: sum_l E[Y_ratio,l]."""
        return float(np.sum(self.eYRatio))


class SL_BayesRatio_Manager(ML_BayesRatio_Manager):
    """`parelagmc::SL_BayesRatio_Manager` (`/root/reference/src/SL_BayesRatio_Manager.hpp:32-150`): single level
    (level 0), E_post[Q] = E[R] / E[Z]; its loop (hpp:240-255) is the coarsest-level loop of the multilevel manager."""

    def __init__(self, comm, backend, params: Optional[dict] = None, out=sys.stdout, stream_pos: int = 0):
        params = dict(params or {})
        params.setdefault("Array number of samples", [int(params.get("Number of samples", 10))])
        super().__init__(comm, 1, backend, params, out=out, stream_pos=stream_pos)


class SL_BayesRatio_Splitting_Manager(ML_BayesRatio_Splitting_Manager):
    """`parelagmc::SL_BayesRatio_Splitting_Manager` (`/root/reference/src/SL_BayesRatio_Splitting_Manager.hpp:32-152`):
    single level, E_post[Q] = E[R / Z]."""

    def __init__(self, comm, backend, params: Optional[dict] = None, out=sys.stdout, stream_pos: int = 0):
        params = dict(params or {})
        params.setdefault("Array number of samples", [int(params.get("Number of samples", 10))])
        super().__init__(comm, 1, backend, params, out=out, stream_pos=stream_pos)
