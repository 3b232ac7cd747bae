"""CPU tests of the host side: the hierarchy provider (stand-in for ParELAG), the managers' logic (sample sharding,
stream positions, statistics, report) and the N>1 path with world_size 2 over gloo."""
import io
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from common import hex_problem, quad_problem, OracleBackend
from parelagmc_b200 import hierarchy as H
from parelagmc_b200 import managers as MG

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_hierarchy_counts_and_identities():
    L = H.build_box_hierarchy([16] * 3, [2.0] * 3, 3)
    assert [(l.Ne, l.Nf) for l in L] == [(4096, 13056), (512, 1728), (64, 240)]
    for l in L:
        assert np.isclose(l.Wdiag.sum(), 8.0)
        M = l.assemble_M()
        assert abs(M - M.T).max() < 1e-15
        # B = W D, rows sum to zero on interior-only... every element has net zero incidence
        assert np.allclose(np.asarray(l.B.sum(axis=1)).ravel(), 0.0)
    for f, c in zip(L[:-1], L[1:]):
        assert np.allclose(np.asarray(f.P_s.sum(axis=1)).ravel(), 1.0)           # aggregation: partition of unity
        # commuting diagram: the divergence of a prolongated coarse flux is the prolongated coarse divergence
        rng = np.random.default_rng(0)
        uc = rng.standard_normal(c.Nf)
        assert np.allclose(f.D @ (f.P_u @ uc), f.P_s @ (c.D @ uc), atol=1e-12)
        # Galerkin mass: P_u^T M_f P_u == M_c (nested RT0 spaces)
        assert abs(f.P_u.T @ f.assemble_M() @ f.P_u - c.assemble_M()).max() < 1e-12


def test_darcy_levels_rhs_obs():
    p = hex_problem(8, 3)
    for lev, d in enumerate(p["darcy"]):
        n = 8 >> lev
        assert d.ess_u.sum() == 4 * n * n                       # four lateral faces
        # flux dofs are total face fluxes: one unit entry per inflow / observation face on every level
        assert np.isclose(np.abs(d.rhs).sum(), n * n) and np.isclose(np.abs(d.obs).sum(), n * n)
    q = quad_problem(4, 2)
    assert q["sampler"][0].ess_u.sum() == 16


def test_split_samples_partition():
    for n in (0, 1, 7, 10, 1000):
        for w in (1, 2, 3, 8):
            parts = [MG.split_samples(n, r, w) for r in range(w)]
            assert sum(c for _, c in parts) == n
            pos = 0
            for f, c in parts:
                assert f == pos
                pos += c


def test_expwregression_matches_oracle():
    from oracle.binding import exp_w_regression
    rng = np.random.default_rng(1)
    x = np.array([17152.0, 2240.0, 304.0, 40.0])
    y = np.abs(rng.standard_normal(4)) + 0.1
    for skip in (0, 1):
        assert MG.expWRegression(y, x, skip) == pytest.approx(exp_w_regression(y, x, skip), rel=1e-13)


def test_mlmc_manager_matches_reference_formulas(tmp_path):
    """InitRun order / stream positions (MLMC_Manager.cpp:110,140) and computeNSamplesMSE against the oracle's
    restatement of :300-401."""
    from oracle.binding import mlmc_compute
    p = hex_problem(4, 3)
    be = OracleBackend(p)
    out = io.StringIO()
    params = {"Use array samples": True, "Array number of samples": [4, 6, 8], "Mean square error": 1e6,
              "Output filename for MC managers": str(tmp_path / "MLMC.dat")}
    m = MG.MLMC_Manager(None, 3, be, params, out=out)
    m.wallTime = False
    m.Run()
    # coarsest level first; every level starts where the previous one stopped
    assert [c[0] for c in be.calls] == [2, 1, 0]
    assert be.calls[0][2] == 0 and be.calls[1][2] == 8 * be.Ne[2] and be.calls[2][2] == 8 * be.Ne[2] + 6 * be.Ne[1]
    r = mlmc_compute(m.sums, m.level_nsamples, m.M, cost=None, eps2=1e6, ratio=0.5)
    assert np.allclose(m.eY, r["eY"]) and np.allclose(m.varY, r["varY"]) and np.allclose(m.varQ, r["varQ"])
    assert np.allclose(m.kurtosis, r["kurtosis"]) and np.allclose(m.consistency, r["consistency"])
    assert m.alpha == pytest.approx(r["alpha"]) and m.alphaABS == pytest.approx(r["alpha_abs"])
    assert m.gamma == pytest.approx(r["gamma"])
    assert m.expected_discretization_error2 == pytest.approx(r["bias2"])
    assert m.ml_estimator_variance == pytest.approx(r["ml_estimator_variance"])
    assert np.array_equal(m.level_nsamples_missing, r["missing"])
    txt = out.getvalue()
    for label in ("MLMC Manager Errors:", "Estimate", "Target MSE", "ML Estimator Variance", "DOFS in Forward Problem",
                  "NumSamples", "E[Y_l]", "Var[Q_l]", "Consistency", "Kurtosis", "FINAL MLMC ERRORS"):
        assert label in txt
    log = open(tmp_path / "MLMC.dat").read().splitlines()
    assert len(log) == 1 + 4 + 6 + 8                              # header + one row per sample (:134-135,:171-172)


def test_mlmc_manager_adaptive_loop(tmp_path):
    """Run() keeps adding samples until the estimator variance meets ratio * eps2 (:202-209)."""
    p = hex_problem(4, 2)
    be = OracleBackend(p)
    params = {"Number of samples": 6, "Mean square error": 4e-3, "Output filename for MC managers": ""}
    m = MG.MLMC_Manager(None, 2, be, params, out=None)
    m.wallTime = False
    m.Run()
    assert m.ml_estimator_variance <= 0.5 * m.eps2
    assert m.level_nsamples.sum() > 12


def test_mc_manager(tmp_path):
    from oracle.binding import mc_compute
    p = hex_problem(4, 2)
    be = OracleBackend(p)
    out = io.StringIO()
    m = MG.MC_Manager(None, be, {"Number of samples": 12, "Mean square error": 1e6,
                                 "Output filename for MC managers": ""}, out=out)
    m.wallTime = False
    m.Run()
    r = mc_compute(m.sums, m.level_nsamples, eps2=1e6, ratio=0.5)
    assert m.eQ == pytest.approx(r["eQ"]) and m.varQ == pytest.approx(r["varQ"])
    assert "FINAL SLMC ERRORS" in out.getvalue() and "SLMC Manager Errors:" in out.getvalue()


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import hex_problem, OracleBackend
    from parelagmc_b200 import managers as MG
    p = hex_problem(4, 2)
    be = OracleBackend(p, threads=1)
    m = MG.MLMC_Manager(MG._Comm(None, True), 2, be, {"Use array samples": True, "Array number of samples": [5, 9],
                                                      "Mean square error": 1e6,
                                                      "Output filename for MC managers": ""}, out=None)
    m.wallTime = False
    m.Run()
    q.put((rank, m.sums.copy(), m.level_nsamples.copy(), list(be.calls), float(np.sum(m.eY))))
    dist.destroy_process_group()


def test_sharded_run_world2_gloo():
    """N > 1 path on CPU: two ranks shard every level's budget; the all-reduced sums equal the 1-rank sums and the
    union of the ranks' stream slices is the 1-rank stream."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = hex_problem(4, 2)
    be = OracleBackend(p, threads=1)
    m = MG.MLMC_Manager(None, 2, be, {"Use array samples": True, "Array number of samples": [5, 9],
                                      "Mean square error": 1e6, "Output filename for MC managers": ""}, out=None)
    m.wallTime = False
    m.Run()
    for rank, sums, ns, calls, est in res:
        assert np.allclose(sums, m.sums, rtol=1e-12, atol=1e-14)
        assert np.array_equal(ns, m.level_nsamples)
        assert est == pytest.approx(float(np.sum(m.eY)), rel=1e-12)
    # rank slices: level 1 (9 samples): 5 + 4, level 0 (5 samples): 3 + 2, contiguous in the stream
    Ne = be.Ne
    c0, c1 = res[0][3], res[1][3]
    assert c0[0] == (1, 5, 0) and c1[0] == (1, 4, 5 * Ne[1])
    assert c0[1] == (0, 3, 9 * Ne[1]) and c1[1] == (0, 2, 9 * Ne[1] + 3 * Ne[0])


def _gloo_adaptive_worker(rank, world, port, q):
    import time
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import hex_problem, OracleBackend
    from parelagmc_b200 import managers as MG

    class Skewed(OracleBackend):          # rank 1 is a 4x slower machine: its local timings differ wildly from rank 0's
        def mlmc_level_batch(self, *a, **k):
            t0 = time.perf_counter()
            r = super().mlmc_level_batch(*a, **k)
            if rank == 1:
                time.sleep(3.0 * (time.perf_counter() - t0))
            return r

    p = hex_problem(4, 2)
    be = Skewed(p, threads=1)
    m = MG.MLMC_Manager(MG._Comm(None, True), 2, be, {"Number of samples": 6, "Mean square error": 4e-3,
                                                      "Output filename for MC managers": ""}, out=None)
    assert m.wallTime            # the default: cost from timers, which differ per rank
    m.concurrent_levels = False
    m.Run()
    q.put((rank, m.sums.copy(), m.level_nsamples.copy(), m.level_nsamples_missing.copy(), m.stream_pos,
           m.ml_estimator_variance, m.eps2, list(be.calls)))
    dist.destroy_process_group()


def test_adaptive_run_world2_wall_time_costs_agree():
    """Run() with a finite MSE target and wallTime = True on two ranks whose timers disagree: the timings are reduced in
    the same collective as the sums, so both ranks request the same sample counts every round, their slices tile the
    stream without gaps or overlap, and both leave the loop together."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_adaptive_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted([q.get(timeout=600) for _ in range(2)], key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    (_, s0, n0, miss0, pos0, var0, eps0, calls0), (_, s1, n1, miss1, pos1, var1, eps1, calls1) = res
    assert np.array_equal(n0, n1) and np.array_equal(miss0, miss1) and pos0 == pos1
    assert np.allclose(s0, s1, rtol=0, atol=0)
    assert var0 == var1 and var0 <= 0.5 * eps0 and n0.sum() > 12
    assert len(calls0) == len(calls1)
    Ne = hex_problem(4, 2)["sampler"]
    for (l0, c0, p0), (l1, c1, p1) in zip(calls0, calls1):       # same level, adjacent slices of the stream
        assert l0 == l1 and p1 == p0 + c0 * Ne[l0].Ne


def test_managers_refuse_single_sample_levels():
    p = hex_problem(4, 2)
    be = OracleBackend(p)
    m = MG.MLMC_Manager(None, 2, be, {"Use array samples": True, "Array number of samples": [1, 5],
                                      "Mean square error": 1e6, "Output filename for MC managers": ""}, out=None)
    with pytest.raises(ValueError, match="at least 2 samples"):
        m.Run()


def test_bench_stream_positions():
    sys.path.insert(0, ROOT)
    import bench
    p = hex_problem(4, 2)
    Ne = [s.Ne for s in p["sampler"]]
    pos0 = bench.stream_positions(p, [10, 20], 0, 2)
    pos1 = bench.stream_positions(p, [10, 20], 1, 2)
    assert pos0[1] == 0 and pos1[1] == 20 * Ne[1]
    assert pos0[0] == 2 * 20 * Ne[1] and pos1[0] == 2 * 20 * Ne[1] + 10 * Ne[0]


def test_bayes_ratio_manager():
    """ML_BayesRatio_Manager mirror: stream positions (two draws per realisation, coarsest level first), statistics
    (hpp:572-726) and the ratio estimate."""
    from common import bayes_problem
    p = bayes_problem()
    be = OracleBackend(p, threads=4)
    for lev in range(p["nlevels"]):
        be.o.set_observations(lev, p["gobs"][lev], p["G_obs"], p["noise"])
    m = MG.ML_BayesRatio_Manager(None, 2, be, {"Array number of samples": [6, 12], "Mean square error": 1e6}, out=None,
                                 stream_pos=p["pos_after_setup"])
    m.wallTime = False
    m.Run()
    n = m.level_nsamples.astype(float)
    assert list(m.level_nsamples) == [6, 12]
    assert np.allclose(m.eYR, m.sums[:, 7] / n) and np.allclose(m.eYZ, m.sums[:, 1] / n)
    assert np.allclose(m.varYZ, (m.sums[:, 0] / n - (m.sums[:, 1] / n) ** 2) * n / (n - 1))
    assert m.ml_estimator_variance == pytest.approx(max(np.sum(m.varYZ / n), np.sum(m.varYR / n)))
    assert m.eC[0] == 2 * p["darcy"][0].N + 2 * p["darcy"][1].N and m.eC[1] == 2 * p["darcy"][1].N
    est = m.estimate()
    assert 1.0 < est < 4.0            # a posterior mean of the effective permeability of the same order as the prior's
    # stream bookkeeping: level 1 first (12 samples x 2 draws), then level 0
    assert m.stream_pos == p["pos_after_setup"] + 2 * 12 * be.Ne[1] + 2 * 6 * be.Ne[0]


def test_ratio_splitting_and_single_level_managers():
    """The remaining ratio managers of the reference (SL_BayesRatio_Manager, SL_/ML_BayesRatio_Splitting_Manager): the
    ratio sums formed from the device call's rows against a direct evaluation, single level = the multilevel manager with
    one level, and the "divide, then sum" estimate."""
    from common import bayes_problem
    p = bayes_problem()
    be = OracleBackend(p, threads=4)
    for lev in range(p["nlevels"]):
        be.o.set_observations(lev, p["gobs"][lev], p["G_obs"], p["noise"])
    params = {"Array number of samples": [5, 9], "Mean square error": 1e6}
    m = MG.ML_BayesRatio_Splitting_Manager(None, 2, be, params, out=None, stream_pos=p["pos_after_setup"])
    m.wallTime = False
    m.Run()
    # direct evaluation of the rows the manager saw (same stream positions: level 1 first)
    pos = p["pos_after_setup"]
    _, r1 = be.o.bayes_level(1, 9, pos, nthreads=2, nlevels=2)
    _, r0 = be.o.bayes_level(0, 5, pos + 2 * 9 * be.Ne[1], nthreads=2, nlevels=2)
    q1 = r1[:, 0] / r1[:, 2]
    q0 = r0[:, 0] / r0[:, 2]
    y0 = q0 - (r0[:, 0] - r0[:, 1]) / (r0[:, 2] - r0[:, 3])
    assert m.sums[1, MG.BR["Ratio"]] == pytest.approx(q1.sum()) and m.sums[1, MG.BR["YRatio2"]] == pytest.approx((q1 * q1).sum())
    assert m.sums[0, MG.BR["YRatio"]] == pytest.approx(y0.sum()) and m.sums[0, MG.BR["ABS_Ratio"]] == pytest.approx(np.abs(q0).sum())
    assert m.estimate() == pytest.approx(y0.mean() + q1.mean())
    assert m.ml_estimator_variance == pytest.approx(np.var(y0, ddof=1) / 5 + np.var(q1, ddof=1) / 9)
    # single-level managers on level 0
    sl = MG.SL_BayesRatio_Manager(None, be, {"Number of samples": 6, "Mean square error": 1e6}, out=None,
                                  stream_pos=p["pos_after_setup"])
    sl.wallTime = False
    sl.Run()
    ml1 = MG.ML_BayesRatio_Manager(None, 1, be, {"Array number of samples": [6], "Mean square error": 1e6}, out=None,
                                   stream_pos=p["pos_after_setup"])
    ml1.wallTime = False
    ml1.Run()
    assert np.array_equal(sl.sums, ml1.sums) and sl.estimate() == ml1.estimate()
    sls = MG.SL_BayesRatio_Splitting_Manager(None, be, {"Number of samples": 6, "Mean square error": 1e6}, out=None,
                                             stream_pos=p["pos_after_setup"])
    sls.wallTime = False
    sls.Run()
    _, r = be.o.bayes_level(0, 6, p["pos_after_setup"], nthreads=2, nlevels=1)
    assert sls.estimate() == pytest.approx((r[:, 0] / r[:, 2]).mean())


def test_lanczos_extremes_of_the_set_up_toolkit(tmp_path):
    """csrc/host_sparse.hpp: lanczos_extremes (the spectrum estimate behind the sampler's a-priori Chebyshev step count):
    the extreme Ritz values of a Jacobi-scaled SPD operator after 80 steps against numpy's eigenvalues."""
    import os, subprocess, textwrap
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n = 20
    src = tmp_path / "lz.cpp"
    src.write_text(textwrap.dedent(f"""
        #include "host_sparse.hpp"
        #include <cstdio>
        using namespace pmc;
        int main() {{
            const int n = {n}, N = n * n * n;   // 3D 7-point operator + mass term, Jacobi scaled
            std::vector<double> d(N);
            auto coef = [&](int i, int j, int k) {{ return 1.0 + 0.9 * std::sin(0.7 * i + 1.3 * j + 2.1 * k); }};
            auto id = [&](int i, int j, int k) {{ return (k * n + j) * n + i; }};
            auto raw = [&](const double *x, double *y) {{
                for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {{
                    const double c = coef(i, j, k);
                    double s = (0.8 + 6.0 * c) * x[id(i, j, k)];
                    if (i > 0) s -= 0.5 * (c + coef(i - 1, j, k)) * x[id(i - 1, j, k)];
                    if (i + 1 < n) s -= 0.5 * (c + coef(i + 1, j, k)) * x[id(i + 1, j, k)];
                    if (j > 0) s -= 0.5 * (c + coef(i, j - 1, k)) * x[id(i, j - 1, k)];
                    if (j + 1 < n) s -= 0.5 * (c + coef(i, j + 1, k)) * x[id(i, j + 1, k)];
                    if (k > 0) s -= 0.5 * (c + coef(i, j, k - 1)) * x[id(i, j, k - 1)];
                    if (k + 1 < n) s -= 0.5 * (c + coef(i, j, k + 1)) * x[id(i, j, k + 1)];
                    y[id(i, j, k)] = s;
                }}
            }};
            for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) d[id(i, j, k)] = 1.0 / std::sqrt(0.8 + 6.0 * coef(i, j, k));
            std::vector<double> t(N);
            auto op = [&](const double *x, double *y) {{
                for (int i = 0; i < N; ++i) t[i] = d[i] * x[i];
                raw(t.data(), y);
                for (int i = 0; i < N; ++i) y[i] *= d[i];
            }};
            double lo, hi;
            lanczos_extremes(N, 80, op, &lo, &hi);
            std::printf("%.12f %.12f ", lo, hi);
        }}
        """))
    exe = tmp_path / "lz"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(root, "parelagmc_b200", "csrc"), "-o", str(exe), str(src)])
    lo, hi = (float(v) for v in subprocess.check_output([str(exe)], text=True).split())
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    idx = np.arange(n ** 3).reshape(n, n, n)          # [k][j][i]
    kk, jj, ii = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    c = 1.0 + 0.9 * np.sin(0.7 * ii + 1.3 * jj + 2.1 * kk)
    rows, cols, vals = [idx.ravel()], [idx.ravel()], [(0.8 + 6.0 * c).ravel()]
    for ax in range(3):
        a = [slice(None)] * 3; b = [slice(None)] * 3
        a[ax] = slice(0, n - 1); b[ax] = slice(1, n)
        w = -0.5 * (c[tuple(a)] + c[tuple(b)]).ravel()
        rows += [idx[tuple(a)].ravel(), idx[tuple(b)].ravel()]
        cols += [idx[tuple(b)].ravel(), idx[tuple(a)].ravel()]
        vals += [w, w]
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))))
    D = sp.diags(1.0 / np.sqrt(A.diagonal()))
    As = D @ A @ D
    e_hi = spl.eigsh(As, k=1, which="LA", return_eigenvectors=False)[0]
    e_lo = spl.eigsh(As, k=1, sigma=0, which="LM", return_eigenvectors=False)[0]
    # Ritz values lie inside the spectrum and have converged to the margins the library widens them by (4 % / 1 %)
    assert e_lo - 1e-10 <= lo <= e_lo * 1.03 and e_hi * 0.995 <= hi <= e_hi + 1e-10, (lo, e_lo, hi, e_hi)


def test_functional_only_minres_recurrence():
    """The algebra behind the functional-only Darcy solve (csrc/program.cuh: sc_beta with a1 = 1): in preconditioned MINRES
    x = sum_k cx_k w_k with w_k = cw0 w_{k-2} + cw1 w_{k-1} + cu z_k, so omega_k = obs . w_k follows the same recurrence
    driven by obs . z_k and Q = obs . x needs neither the direction vectors nor the solution.  numpy restatement of the
    device's scalar recurrences (lazily normalised Lanczos vectors) on a random saddle-point system."""
    import scipy.sparse as sp
    rng = np.random.default_rng(3)
    nu, npr = 60, 25
    M = sp.diags(1.0 + rng.random(nu)) + 0.1 * sp.random(nu, nu, 0.05, random_state=1)
    M = (M + M.T) * 0.5 + sp.identity(nu) * 0.5
    B = sp.random(npr, nu, 0.2, random_state=2) + sp.eye(npr, nu)
    A = sp.bmat([[M, B.T], [B, None]]).tocsr()
    N = nu + npr
    b = rng.standard_normal(N)
    obs = np.where(rng.random(N) < 0.2, rng.standard_normal(N), 0.0)
    dM = M.diagonal()
    S = (B @ sp.diags(1 / dM) @ B.T).toarray()
    Sinv = np.linalg.inv(S)
    P = lambda r: np.concatenate([r[:nu] / dM, Sinv @ r[nu:]])
    # state of the device program: unnormalised Lanczos vectors v0, v1, preconditioned u1 = P v1, scalars as in sc_init / sc_alpha / sc_beta
    v0 = np.zeros(N); v1 = b.copy(); u1 = P(v1)
    beta = np.sqrt(v1 @ u1); ib = 1 / beta; ibprev = 0.0
    g0 = g1 = 1.0; s0 = s1 = 0.0; eta = beta
    w0 = np.zeros(N); w1 = np.zeros(N); x = np.zeros(N)
    om0 = om1 = 0.0; Q = 0.0
    for it in range(200):
        q = A @ u1
        alpha = (u1 @ q) * ib * ib
        zeta = obs @ u1                                    # OP_DOT_SPARSE
        v0 = ib * q - alpha * ib * v1 - beta * ibprev * v0  # OP_LINCOMB3 (cq, cv1, cv0)
        z = P(v0)
        beta_new = np.sqrt(max(v0 @ z, 0.0))
        delta = g1 * alpha - g0 * s1 * beta; rho3 = s0 * beta; rho2 = s1 * alpha + g0 * g1 * beta
        rho1 = np.hypot(delta, beta_new); ir = 1 / rho1
        cw0, cw1, cu = -rho3 * ir, -rho2 * ir, ib * ir
        g0, g1 = g1, delta * ir
        cx = g1 * eta
        s0, s1 = s1, beta_new * ir
        eta = -s1 * eta
        wn = cw0 * w0 + cw1 * w1 + cu * u1                 # OP_SOL_UPDATE (the vectors the functional-only solve never forms)
        x = x + cx * wn
        w0, w1 = w1, wn
        om = cw0 * om0 + cw1 * om1 + cu * zeta             # sc_beta, a1 = 1
        om0, om1 = om1, om
        Q += cx * om
        ibprev, ib, beta = ib, 1 / beta_new, beta_new
        v0, v1 = v1, v0
        u1 = z
        if abs(eta) < 1e-13 * abs(np.sqrt(b @ P(b))):
            break
    xs = np.linalg.solve(A.toarray(), b)
    assert np.linalg.norm(x - xs) / np.linalg.norm(xs) < 1e-9           # the restated MINRES solves the system
    assert abs(Q - obs @ x) <= 1e-12 * max(1.0, abs(obs @ x))           # and the scalar recurrence carries obs . x


def test_bench_clock_sampler_windows_the_timed_region():
    """bench.py's ClockSampler: nvidia-smi lines are time-stamped on arrival and only those inside the timed region count;
    throttle reasons are collected from them; with no sample inside the window the nearest ones are reported as such."""
    import importlib.util, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class _P:
        def terminate(self): pass
        def wait(self, timeout=None): return 0
    cs = bench.ClockSampler(0, 20)
    cs.proc = _P()
    mk = lambda sm, pc="Not Active": f"{sm}, 1965, 700.0, Not Active, Not Active, Not Active, {pc}"
    cs.lines = [(0.5, mk(900)), (1.01, mk(1950, "Active")), (1.05, mk(1965)), (1.09, mk(1960)), (2.0, mk(300))]
    r = cs.stop(1.0, 1.1)
    assert r["samples"] == 3 and r["sm_mhz"] == 1960.0 and r["sm_max_mhz"] == 1965.0
    assert r["reasons"] == ["sw_power_cap"] and r["window"] == "timed region"
    cs.proc = _P()
    r = cs.stop(5.0, 5.1)
    assert r["samples"] >= 1 and r["window"].startswith("nearest")
    cs.proc = None
    assert bench.ClockSampler(0).stop()["sm_mhz"] is None
