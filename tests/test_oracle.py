"""CPU tests of the oracle (the checker): pinned against the reference's own known answers that need no third
party, against an independent sparse direct solve, and against the committed golden fixtures."""
import json
import math
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from common import hex_problem, quad_problem, make_oracle, rel_l2
from parelagmc_b200 import hierarchy as H

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.json")))


def test_matern_scaling_constants():
    """ComputeScalingCoefficientForSPDE (/root/reference/src/Utilities.hpp:188-200) and alpha
    (/root/reference/src/PDESampler.cpp:42); values evaluated from the reference formula."""
    assert H.matern_scaling_coefficient(0.1, 3) == pytest.approx(28.90067818451249, rel=1e-15)
    assert H.matern_scaling_coefficient(0.1, 2) == pytest.approx(50.13256549262001, rel=1e-15)
    assert H.matern_scaling_coefficient(100.0, 3) == pytest.approx(0.9139196898659948, rel=1e-15)
    assert H.spde_alpha(0.1) == 99.99999999999999


def test_darcy_deterministic_known_answer():
    """DarcyDeterministicTest (/root/reference/examples/CMakeLists.txt:62-66): hex 16^3 -> 8^3 -> 4^3, k == 1:
    Q = 2 on every level and 17152 / 2240 / 304 dofs."""
    p = hex_problem(16, 3)
    o = make_oracle(p)
    for lev, dofs in enumerate([17152, 2240, 304]):
        q, c, _, _ = o.darcy_solve(lev, np.ones(p["darcy"][lev].Ne))
        assert c == dofs and p["darcy"][lev].N == dofs
        assert q == pytest.approx(2.0, abs=1e-9)


def test_quad_level_sizes():
    """Config 1 (PDESamplerTest on inline_quad.mesh): 16 / 4 elements, 40 / 12 faces."""
    p = quad_problem(4, 2)
    assert [(s.Ne, s.Nf) for s in p["sampler"]] == [(16, 40), (4, 12)]


def test_tet_hierarchy_known_answers():
    """Tetrahedral levels: the transfer operators commute with the divergence, the coarse spaces are nested (Galerkin
    mass = coarse mass), and with k == 1 the exact solution (pressure linear in z, constant flux) lies in RT0 x P0, so
    Q = 2 on every level as on the hex meshes."""
    from common import tet_problem
    p = tet_problem(4, 2)
    f, c = p["levels"]
    assert (f.Ne, f.Nf, c.Ne, c.Nf) == (384, 864, 48, 120)
    assert abs(f.D @ f.P_u - f.P_s @ c.D).max() < 1e-12
    assert abs(f.P_u.T @ f.assemble_M() @ f.P_u - c.assemble_M()).max() < 1e-12
    assert np.diff(f.assemble_M().indptr).max() == 7        # 2 tetrahedra x 4 faces, one shared
    o = make_oracle(p)
    for lev in range(2):
        d = p["darcy"][lev]
        q, cdofs, sol, _ = o.darcy_solve(lev, np.ones(d.Ne), want_sol=True)
        assert cdofs == d.N and q == pytest.approx(2.0, abs=1e-9)
    rng = np.random.default_rng(4)
    for lev in range(2):                                    # variable k and the SPDE sampler against direct solves
        d, lv, s = p["darcy"][lev], p["levels"][lev], p["sampler"][lev]
        k = np.exp(rng.standard_normal(d.Ne))
        keep = sp.diags((d.ess_u == 0).astype(float))
        A = sp.bmat([[keep @ lv.assemble_M(k) @ keep + sp.diags((d.ess_u != 0).astype(float)), (d.B @ keep).T],
                     [d.B @ keep, None]], format="csc")
        rhs = d.rhs.copy()
        rhs[:d.Nf][d.ess_u != 0] = 0.0
        assert rel_l2(o.darcy_solve(lev, k, want_sol=True)[2], spla.spsolve(A, rhs)) < 1e-8
        xi = rng.standard_normal(s.Ne)
        As = sp.bmat([[s.M, s.B.T], [s.B, -p["alpha"] * sp.diags(s.Wdiag)]], format="csc")
        x = spla.spsolve(As, np.concatenate([np.zeros(s.Nf), -p["g"] * xi * s.w_sqrt]))
        assert rel_l2(o.sampler_eval(lev, xi)[1], x[s.Nf:]) < 1e-9


def test_sampler_against_direct_solve():
    p = hex_problem(8, 3)
    o = make_oracle(p, lognormal=False)
    rng = np.random.default_rng(0)
    for lev in range(3):
        s = p["sampler"][lev]
        xi = rng.standard_normal(s.Ne)
        A = sp.bmat([[s.M, s.B.T], [s.B, -p["alpha"] * sp.diags(s.Wdiag)]], format="csc")
        b = np.concatenate([np.zeros(s.Nf), -p["g"] * xi * s.w_sqrt])
        x = spla.spsolve(A, b)
        got, emb, _ = o.sampler_eval(lev, xi)
        assert rel_l2(got, x[s.Nf:]) < 1e-9
        assert np.array_equal(got, emb)


def test_sampler_restriction_and_lognormal():
    """Noise drawn on level l, evaluated on level l+1 through Ps^T (/root/reference/src/PDESampler.cpp:361-368)."""
    p = hex_problem(8, 3)
    o = make_oracle(p, lognormal=True)
    rng = np.random.default_rng(1)
    s0, s1 = p["sampler"][0], p["sampler"][1]
    xi = rng.standard_normal(s0.Ne)
    A = sp.bmat([[s1.M, s1.B.T], [s1.B, -p["alpha"] * sp.diags(s1.Wdiag)]], format="csc")
    f = s0.P.T @ (-p["g"] * xi * s0.w_sqrt)
    x = spla.spsolve(A, np.concatenate([np.zeros(s1.Nf), f]))
    got, emb, _ = o.sampler_eval(1, xi, xi_level=0, use_init=0)
    assert rel_l2(emb, x[s1.Nf:]) < 1e-9
    assert np.allclose(got, np.exp(emb), rtol=1e-15)


def test_darcy_against_direct_solve():
    p = hex_problem(8, 3)
    o = make_oracle(p)
    rng = np.random.default_rng(2)
    for lev in range(3):
        d, lv = p["darcy"][lev], p["levels"][lev]
        k = np.exp(rng.standard_normal(d.Ne))
        keep = sp.diags((d.ess_u == 0).astype(float))
        Me = keep @ lv.assemble_M(k) @ keep + sp.diags((d.ess_u != 0).astype(float))
        Be = d.B @ keep
        A = sp.bmat([[Me, Be.T], [Be, None]], format="csc")
        rhs = d.rhs.copy()
        rhs[:d.Nf][d.ess_u != 0] = 0.0
        x = spla.spsolve(A, rhs)
        q, c, sol, _ = o.darcy_solve(lev, k, want_sol=True)
        assert rel_l2(sol, x) < 1e-8
        assert q == pytest.approx(float(d.obs @ x), rel=1e-9)


def test_yarn5_jump_split_consistency():
    from oracle.binding import Yarn5
    seq = Yarn5().ints(5000)
    for pos in (1, 15, 16, 17, 1000, 4321):
        assert np.array_equal(Yarn5().jump(pos).ints(50), seq[pos:pos + 50])
    for s, n in ((2, 0), (2, 1), (3, 2), (8, 5)):
        assert np.array_equal(Yarn5().split(s, n).ints(200), seq[n::s][:200])
    assert seq.min() >= 0 and seq.max() < 2**31 - 1


def test_normal_deviates_statistics():
    from oracle.binding import Yarn5, lib
    x = Yarn5().normals(400000, 0.0, 1.0)
    assert abs(x.mean()) < 5e-3 and abs(x.var() - 1.0) < 1e-2
    # exact (log-)normal moments used by PDESamplerTest (/root/reference/examples/PDESamplerTest.cpp:207-209)
    assert abs(np.exp(x).mean() - math.exp(0.5)) < 2e-2
    y = Yarn5().normals(10, 3.0, 2.0)
    assert np.allclose(y, 3.0 + 2.0 * x[:10], rtol=1e-15)
    assert lib().po_uniformoo(0) > 0.0 and lib().po_uniformoo(2**31 - 2) < 1.0
    assert lib().po_inv_Phi(0.5) == 0.0


def test_detmath_accuracy():
    """parelagmc_b200/csrc/detmath.h (the exp/log/erf/erfc both the CUDA build and the oracle evaluate inv_Phi with)
    against mpmath, in ulp of the exact value, on the domain the sampler reaches."""
    import mpmath as mp
    from oracle.binding import det_eval
    mp.mp.dps = 40
    rng = np.random.default_rng(3)

    def max_ulp(which, x, f):
        y = det_eval(which, x)
        worst = 0.0
        for a, b in zip(y, x):
            t = f(mp.mpf(float(b)))
            worst = max(worst, float(abs(mp.mpf(float(a)) - t) / np.spacing(abs(float(t)))))
        return worst

    assert max_ulp("exp", rng.uniform(-30, 30, 3000), mp.exp) < 1.0
    assert max_ulp("log", np.exp(rng.uniform(-25, 3, 3000)), mp.log) < 2.0
    assert max_ulp("erf", rng.uniform(-0.5, 0.5, 3000), mp.erf) < 1.0
    assert max_ulp("erfc", rng.uniform(0.5, 6.8, 6000), mp.erfc) < 4.0
    assert max_ulp("erfc", rng.uniform(6.8, 26.0, 1500), mp.erfc) < 4.0
    assert max_ulp("erfc", rng.uniform(-6.0, 0.5, 1500), mp.erfc) < 1.5
    # special values
    assert det_eval("exp", [0.0, -800.0, 800.0]).tolist() == [1.0, 0.0, np.inf]
    assert det_eval("log", [1.0])[0] == 0.0 and det_eval("erf", [0.0])[0] == 0.0 and det_eval("erfc", [30.0])[0] == 0.0


def test_normals_vs_libm():
    """The deviates computed over detmath.h agree with the same Acklam + Halley sequence over glibc's erf/erfc/exp/log (what
    a TRNG build on this host would call) to a few ulp(1) / phi(y): the Halley step divides the difference of the two
    Phi values (each within a few ulp) by the density.  Most are bit-identical."""
    from oracle.binding import det_eval
    m = 2 ** 31 - 1
    v = np.concatenate([np.arange(0, 20000), np.arange(m - 20000, m), np.random.default_rng(7).integers(0, m, 400000)])
    u = (v.astype(np.float64) + 1.0) / 2147483648.0
    a, b = det_eval("inv_Phi", u), det_eval("inv_Phi_libm", u)
    phi = np.exp(-0.5 * b * b) / np.sqrt(2.0 * np.pi)
    assert np.all(np.abs(a - b) <= 6.0 * np.spacing(1.0) * np.maximum(1.0, 1.0 / phi) * np.maximum(1.0, np.abs(b)))
    assert np.mean(a == b) > 0.85


def test_golden_fixtures():
    """tests/golden/oracle_golden.json (tools/make_golden.py) pins the oracle's stream, fields and QoIs."""
    from oracle.binding import Yarn5
    for pos, v in GOLD["yarn5_ints"].items():
        assert Yarn5().jump(int(pos)).ints(len(v)).tolist() == v
    assert Yarn5().split(4, 3).ints(8).tolist() == GOLD["yarn5_split_4_3"]
    assert [x.hex() for x in Yarn5().normals(16)] == GOLD["normals_pos0"]
    p = hex_problem(4, 2)
    o = make_oracle(p)
    g = GOLD["hex4"]
    assert [d.N for d in p["darcy"]] == g["dofs"]
    for lev in range(2):
        Ne = p["darcy"][lev].Ne
        assert o.darcy_solve(lev, np.ones(Ne))[0] == pytest.approx(g["Q_k1"][lev], rel=1e-10)
        assert o.darcy_solve(lev, np.exp(np.sin(np.arange(Ne, dtype=float))))[0] == pytest.approx(g["Q_ksin"][lev], rel=1e-10)
    xi = Yarn5().normals(p["sampler"][0].Ne)
    assert rel_l2(o.sampler_eval(0, xi)[1], g["field_l0"]) < 1e-10
    assert rel_l2(o.sampler_eval(1, xi, xi_level=0, use_init=0)[1], g["field_l1_from_l0_noise"]) < 1e-10
    for lev, ns in [(1, 4), (0, 3)]:
        sums, rows, _ = o.mlmc_level(lev, ns, 1234)
        assert np.allclose(rows, g[f"mlmc_rows_l{lev}"], rtol=1e-9, atol=1e-12)
        assert np.allclose(sums, g[f"mlmc_sums_l{lev}"], rtol=1e-9, atol=1e-12)


def test_mlmc_level_thread_independence():
    p = hex_problem(4, 2)
    o = make_oracle(p)
    a = o.mlmc_level(0, 6, 77, nthreads=1)
    b = o.mlmc_level(0, 6, 77, nthreads=4)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])


def test_exp_w_regression_and_statistics():
    """expWRegression (/root/reference/src/Utilities.cpp:257-283) and computeNSamplesMSE
    (/root/reference/src/MLMC_Manager.cpp:300-401) against closed forms / an independent numpy restatement."""
    from oracle.binding import exp_w_regression, mlmc_compute, mc_compute
    M = np.array([17152.0, 2240.0, 304.0, 40.0])
    y = 3.0 * M ** (-0.75)
    assert exp_w_regression(y, M, 0) == pytest.approx(-0.75, rel=1e-12)
    assert exp_w_regression(y, M, 1) == pytest.approx(-0.75, rel=1e-12)
    rng = np.random.default_rng(3)
    L = 3
    ns = np.array([50, 200, 800])
    qs = [rng.normal(2.5, 0.3, n) for n in ns]
    ys = [rng.normal(0.1 * 4.0 ** (-l), 0.2 * 2.0 ** (-l), n) if l < L - 1 else qs[l] for l, n in enumerate(ns)]
    sums = np.zeros((L, 9))
    for l in range(L):
        yv, q = ys[l], qs[l]
        sums[l] = [np.sum(yv**2), np.sum(yv), np.sum(np.abs(yv)), np.sum(q**2), np.sum(q), np.sum(np.abs(q)),
                   ns[l] * 100.0 * (l + 1), np.sum(yv**3), np.sum(yv**4)]
    r = mlmc_compute(sums, ns, M[:L], eps2=1e-3, ratio=0.5)
    eY = sums[:, 1] / ns
    varY = (sums[:, 0] / ns - eY**2) * ns / (ns - 1)
    assert np.allclose(r["eY"], eY) and np.allclose(r["varY"], varY)
    assert r["estimate"] == pytest.approx(eY.sum())
    assert r["ml_estimator_variance"] == pytest.approx(np.sum(varY / ns))
    eA = sums[:, 2] / ns
    m = M[0] / M[1]
    assert r["bias2"] == pytest.approx(eA[0]**2 / (m ** (-r["alpha_abs"]) - 1.0) ** 2)
    assert np.allclose(r["kurtosis"], (sums[:, 8] / ns) / (sums[:, 0] / ns) ** 2)
    cost = sums[:, 6] / ns
    prop = np.sum(np.sqrt(varY * cost)) / (0.5 * 1e-3)
    miss = np.maximum(np.ceil(prop * np.sqrt(varY / cost) - ns), 0).astype(int)
    assert np.array_equal(r["missing"], miss)
    s4 = np.array([np.sum(qs[0]**2), np.sum(qs[0]), np.sum(np.abs(qs[0])), 50 * 7.0])
    r1 = mc_compute(s4, 50, eps2=1e-3, ratio=0.5)
    v = (s4[0] / 50 - (s4[1] / 50) ** 2) * 50 / 49
    assert r1["varQ"] == pytest.approx(v) and r1["estimate"] == pytest.approx(s4[1] / 50)
    assert r1["missing"] == max(int(math.ceil(v / (0.5 * 1e-3) - 50)), 0)


@pytest.mark.parametrize("kind", ["embedded", "l2proj", "embedded_tet"])
def test_enlarged_domain_sampler_oracle(kind):
    """EmbeddedPDESampler / L2ProjectionPDESampler apply step: oracle against a direct solve on the
    enlarged mesh followed by the transfer, and basic properties of the transfer matrices."""
    from common import enlarged_problem
    p = enlarged_problem(kind)
    o = make_oracle(p, lognormal=True)
    rng = np.random.default_rng(7)
    for lev in range(p["nlevels"]):
        s = p["sampler"][lev]
        assert s.T.shape == (p["darcy"][lev].Ne, s.Ne)
        if kind in ("embedded", "embedded_tet"):
            assert np.all(s.T.data == 1.0) and np.all(np.diff(s.T.indptr) == 1)   # 0/1 selection
        else:
            assert np.allclose(s.Tscale * np.asarray(s.T.sum(axis=1)).ravel(), 1.0)   # averages: constants preserved
        xi = rng.standard_normal(s.Ne)
        A = sp.bmat([[s.M, s.B.T], [s.B, -p["alpha"] * sp.diags(s.Wdiag)]], format="csc")
        x = spla.spsolve(A, np.concatenate([np.zeros(s.Nf), -p["g"] * xi * s.w_sqrt]))
        ref = s.T @ x[s.Nf:]
        if s.Tscale is not None:
            ref = ref * s.Tscale
        got, emb, _ = o.sampler_eval(lev, xi)
        assert got.shape == (p["darcy"][lev].Ne,)
        assert rel_l2(np.log(got), ref) < 1e-8
        assert rel_l2(emb, x[s.Nf:]) < 1e-8
    sums, rows, _ = o.mlmc_level(0, 3, 0)
    assert np.all(np.isfinite(rows)) and np.all(rows[:, 1] > 0)


def test_bayes_level_oracle():
    """BayesianInverseProblem likelihood / R and one level of ML_BayesRatio_Manager::InitRun: the oracle's level loop
    against a by-hand evaluation with the oracle's building blocks."""
    from common import bayes_problem
    from oracle.binding import Yarn5
    p = bayes_problem()
    o = make_oracle(p)
    for lev in range(p["nlevels"]):
        o.set_observations(lev, p["gobs"][lev], p["G_obs"], p["noise"])
    pos0 = p["pos_after_setup"]
    sums, rows = o.bayes_level(0, 3, pos0, nthreads=2)
    Ne = p["sampler"][0].Ne
    y = Yarn5().jump(pos0)
    for j in range(3):
        zxi, xi = y.normals(Ne), y.normals(Ne)
        vals = {}
        for name, noise_vec in (("z", zxi), ("r", xi)):
            for lev in (0, 1):
                k, _, _ = o.sampler_eval(lev, noise_vec, xi_level=0)
                q, _, sol, _ = o.darcy_solve(lev, k, want_sol=True)
                pr = sol[p["darcy"][lev].Nf:]
                G = np.array([g @ pr / g.sum() for g in p["gobs"][lev]])
                like = np.exp(-np.sum((G - p["G_obs"]) ** 2) / (2 * p["noise"]))
                vals[(name, lev)] = like * (q if name == "r" else 1.0)
        assert rows[j, 0] == pytest.approx(vals[("r", 0)], rel=1e-9)
        assert rows[j, 1] == pytest.approx(vals[("r", 0)] - vals[("r", 1)], rel=1e-7, abs=1e-12)
        assert rows[j, 2] == pytest.approx(vals[("z", 0)], rel=1e-9)
        assert rows[j, 3] == pytest.approx(vals[("z", 0)] - vals[("z", 1)], rel=1e-7, abs=1e-12)
        assert rows[j, 4] == 2 * p["darcy"][0].N + 2 * p["darcy"][1].N
    assert sums[10] == pytest.approx(rows[:, 0].sum()) and sums[4] == pytest.approx(rows[:, 2].sum())
    assert 0.0 < rows[:, 2].min() and rows[:, 2].max() <= 1.0


def test_agglomerated_hierarchy_properties_and_oracle_solves():
    """Unstructured agglomerates: the coarse spaces commute with the divergence and are Galerkin (exact sequence
    D_f P_u = P_s D_c, P_u^T M_f P_u = M_c); the oracle's solves on every level agree with a sparse direct solve."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from common import agglomerated_problem
    p = agglomerated_problem()
    L = p["levels"]
    for l in range(p["nlevels"] - 1):
        f, c = L[l], L[l + 1]
        assert abs(f.D @ f.P_u - f.P_s @ c.D).max() < 1e-10
        assert abs(f.P_u.T @ f.assemble_M() @ f.P_u - c.assemble_M()).max() < 1e-12
    assert max(np.diff(L[1].elem_ptr)) > 8                      # wider than anything the structured meshes produce
    o = make_oracle(p)
    rng = np.random.default_rng(4)
    for lev in range(p["nlevels"]):
        d, lv = p["darcy"][lev], L[lev]
        k = np.exp(rng.standard_normal(d.Ne))
        Q, C, sol, _ = o.darcy_solve(lev, k, want_sol=True)
        M = lv.assemble_M(k)
        ess = d.ess_u != 0
        A = sp.bmat([[M, d.B.T], [d.B, None]]).tolil()
        rhs = d.rhs.copy()
        for i in np.nonzero(ess)[0]:
            A[i, :] = 0
            A[:, i] = 0
            A[i, i] = 1
            rhs[i] = 0
        ref = spla.spsolve(A.tocsc(), rhs)
        assert rel_l2(sol, ref) < 1e-8
        assert Q == pytest.approx(d.obs @ ref, rel=1e-8) and C == d.N
