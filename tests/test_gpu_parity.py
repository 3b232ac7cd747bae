"""GPU parity tests: the CUDA path (through the C ABI, include/pmc_b200.h) against the CPU oracle on the same
seeded inputs.  Bars (BASELINE.json north_star): integer stream bit-exact; per-sample fields and Darcy solutions
within 1e-8 relative L2 at solver tolerance 1e-12; per-level MLMC mean/variance within 1e-6 relative."""
import numpy as np
import pytest
import scipy.sparse as sp

from common import hex_problem, quad_problem, make_oracle, make_context, rel_l2

pytestmark = pytest.mark.gpu

FIELD_TOL = 1e-8      # north_star: per-sample field / solution, relative L2
MOMENT_TOL = 1e-6     # north_star: per-level mean and variance, relative


@pytest.fixture(scope="module")
def prob():
    return hex_problem(8, 3)


@pytest.fixture(scope="module")
def ctx(prob):
    c = make_context(prob)
    yield c
    c.close()


@pytest.fixture(scope="module")
def orc(prob):
    return make_oracle(prob)


def test_yarn5_integer_stream_bit_exact(ctx):
    from oracle.binding import Yarn5
    for pos, n in [(0, 1), (0, 1000), (5, 77), (123456789, 4097), (2**40 + 12345, 513)]:
        ref = Yarn5().jump(pos).ints(n)
        got = ctx.rng_fill_int(pos, n)
        assert np.array_equal(ref, got), (pos, n)


def test_yarn5_split_streams_bit_exact(prob):
    from oracle.binding import Yarn5
    from parelagmc_b200.capi import Context
    c = Context(1)
    try:
        for nparts, mypart in [(2, 0), (2, 1), (8, 5)]:
            c.rng_init(0.0, 1.0, nparts, mypart)
            ref = Yarn5().split(nparts, mypart).ints(300)
            assert np.array_equal(ref, c.rng_fill_int(0, 300))
            # leapfrog property against the parent stream
            parent = Yarn5().ints(300 * nparts)
            assert np.array_equal(ref, parent[mypart::nparts][:300])
    finally:
        c.close()


def test_normal_deviates_bit_exact(ctx):
    """Doubles, bit for bit (north_star: "white-noise vectors ... bit-exact"): the integer draw is exact, and uniformoo,
    Acklam's approximation, the Halley step and erf/erfc/exp/log (csrc/detmath.h) are one IEEE operation sequence on
    both sides.  1.2 M consecutive draws of the stream, both extreme tails and the branch points of inv_Phi through
    `pmc_rng_map`, and a non-trivial (mu, sigma)."""
    from oracle.binding import Yarn5, normal_map
    n = 1200000
    ref = Yarn5().jump(1000).normals(n, 0.0, 1.0)
    got = ctx.rng_fill(1000, n)
    assert np.array_equal(got, ref)
    assert abs(got.mean()) < 0.01 and abs(got.std() - 1.0) < 0.01
    m = 2 ** 31 - 1                                            # engine outputs lie in [0, m - 1]
    lo, hi = int(0.02425 * 2 ** 31), int((1.0 - 0.02425) * 2 ** 31)      # inv_Phi's branch points x_low, x_high
    q1, q3 = int(0.25 * 2 ** 31), int(0.75 * 2 ** 31)                    # Phi's |x| = 0.6745 branch points
    eng = np.concatenate([np.arange(0, 60000), np.arange(m - 60000, m), np.arange(lo - 3000, lo + 3000),
                          np.arange(hi - 3000, hi + 3000), np.arange(q1 - 3000, q1 + 3000), np.arange(q3 - 3000, q3 + 3000),
                          np.arange(2 ** 30 - 3000, 2 ** 30 + 3000),
                          np.random.default_rng(5).integers(0, m, 400000)]).astype(np.int32)
    ref_t = normal_map(eng)
    got_t = ctx.rng_map(eng)
    assert np.array_equal(got_t, ref_t)
    assert ref_t.min() < -6.1 and ref_t.max() > 6.1 and np.all(np.isfinite(ref_t))   # u = 2^-31 and 1 - 2^-31
    ctx.rng_init(3.0, 2.0, 1, 0)
    try:
        assert np.array_equal(ctx.rng_fill(5, 50000), Yarn5().jump(5).normals(50000, 3.0, 2.0))
        assert np.array_equal(ctx.rng_map(eng[:70000]), normal_map(eng[:70000], 3.0, 2.0))
    finally:
        ctx.rng_init(0.0, 1.0, 1, 0)


def test_sample_batch_is_stream_in_order(ctx, prob):
    from oracle.binding import Yarn5
    Ne = prob["sampler"][1].Ne
    xi = ctx.sampler_sample_batch(1, 5, 777)
    ref = Yarn5().jump(777).normals(5 * Ne).reshape(5, Ne)
    assert np.array_equal(xi, ref)


def test_darcy_operator_apply_matches_assembled(ctx, prob):
    """Batched element kernel: A_bc(k) x against scipy assembly + EliminateRowCol."""
    rng = np.random.default_rng(1)
    for lev in range(prob["nlevels"]):
        d, lv = prob["darcy"][lev], prob["levels"][lev]
        n = 5
        k = np.exp(rng.standard_normal((n, d.Ne)))
        x = rng.standard_normal((n, d.N))
        y = ctx.darcy_apply_batch(lev, k, x)
        keep = sp.diags((d.ess_u == 0).astype(float))
        for j in range(n):
            M = lv.assemble_M(k[j])
            Me = keep @ M @ keep + sp.diags((d.ess_u != 0).astype(float))
            Be = d.B @ keep
            A = sp.bmat([[Me, Be.T], [Be, None]], format="csr")
            assert rel_l2(y[j], A @ x[j]) < 1e-13


def test_darcy_known_answer_Q2(ctx, prob):
    """DarcyDeterministicTest (/root/reference/examples/CMakeLists.txt:62-66): k == 1 gives Q = 2 on every level."""
    for lev in range(prob["nlevels"]):
        d = prob["darcy"][lev]
        Q, C, sol, it = ctx.darcy_solve_batch(lev, np.ones((3, d.Ne)), want_sol=True)
        assert np.allclose(Q, 2.0, rtol=0, atol=1e-9), Q
        assert np.all(C == d.N)


def test_darcy_solve_matches_oracle(ctx, orc, prob):
    rng = np.random.default_rng(2)
    for lev in range(prob["nlevels"]):
        d = prob["darcy"][lev]
        n = 7
        k = np.exp(rng.standard_normal((n, d.Ne)))
        Q, C, sol, it = ctx.darcy_solve_batch(lev, k, want_sol=True)
        print(f"darcy level {lev}: iterations {it.min()}..{it.max()}")
        for j in range(n):
            q, c, s, _ = orc.darcy_solve(lev, k[j], want_sol=True)
            assert rel_l2(sol[j], s) < FIELD_TOL, (lev, j, rel_l2(sol[j], s))
            assert abs(Q[j] - q) <= 1e-8 * abs(q)
        assert it.max() < 2000


def test_sampler_eval_matches_oracle(ctx, orc, prob):
    from oracle.binding import Yarn5
    for lev in range(prob["nlevels"]):
        Ne = prob["sampler"][lev].Ne
        n = 6
        xi = Yarn5().jump(31 * lev).normals(n * Ne).reshape(n, Ne)
        s, emb, it = ctx.sampler_eval_batch(lev, xi)
        print(f"sampler level {lev}: iterations {it.min()}..{it.max()}")
        for j in range(n):
            so, eo, _ = orc.sampler_eval(lev, xi[j])
            assert rel_l2(emb[j], eo) < FIELD_TOL
            assert rel_l2(s[j], so) < FIELD_TOL


def test_sampler_methods_agree(orc, prob):
    """The sampler system solved as MINRES on the saddle form ("sampler.method" = 0), as Jacobi-PCG on its SPD form (= 1)
    and by the Chebyshev semi-iteration on the SPD form (= 2; the default for short correlation lengths): same fields,
    equal to the oracle's (which runs MINRES on the saddle form with its own preconditioner) to the field tolerance; also
    through a cluster split."""
    from oracle.binding import Yarn5
    from common import make_context
    ctxs = {m: make_context(prob, True, 1e-12, 1e-30, 2000, options={"sampler.method": m}) for m in (0, 1, 2)}
    ctxs["1c"] = make_context(prob, True, 1e-12, 1e-30, 2000, options={"sampler.method": 1, "cluster_size": 2, "cta_threads": 256})
    ctxs["2c"] = make_context(prob, True, 1e-12, 1e-30, 2000, options={"sampler.method": 2, "cluster_size": 2, "cta_threads": 256})
    try:
        for lev in range(prob["nlevels"]):
            Ne = prob["sampler"][lev].Ne
            n = 5
            xi = Yarn5().jump(977 * lev + 3).normals(n * Ne).reshape(n, Ne)
            res = {m: c.sampler_eval_batch(lev, xi) for m, c in ctxs.items()}
            print(f"sampler level {lev}: MINRES its {res[0][2].max()}, PCG its {res[1][2].max()}, Chebyshev steps {res[2][2].max()}")
            for j in range(n):
                so, eo, _ = orc.sampler_eval(lev, xi[j])
                for m in ctxs:
                    assert rel_l2(res[m][1][j], eo) < FIELD_TOL, (lev, m)
                    assert rel_l2(res[m][0][j], so) < FIELD_TOL, (lev, m)
            assert res[1][2].max() < 2000 and res[0][2].max() < 2000
            # the a-priori step count suffices: no restarted block (every realisation reports the same count, and it is
            # within a few steps of what PCG needs times the Chebyshev / CG ratio)
            assert res[2][2].min() == res[2][2].max() and res[2][2].max() <= 3 * res[1][2].max() + 4
    finally:
        for c in ctxs.values():
            c.close()


def test_chebyshev_sampler_restarts_when_the_spectrum_estimate_is_too_narrow(orc, prob):
    """The Chebyshev semi-iteration runs an a-priori number of steps and then CHECKS the true residual norm per realisation.
    With a deliberately wrong spectrum estimate ("sampler.cheb_lo_scale" = 2.5: the interval misses the lower part of the
    spectrum, so the a-priori count is far too small) the check fails, the tiles run restarted blocks until the stopping
    rule holds, and the fields still equal the oracle's to the field tolerance."""
    from oracle.binding import Yarn5
    from common import make_context
    good = make_context(prob, True, 1e-12, 1e-30, 2000, options={"sampler.method": 2})
    bad = make_context(prob, True, 1e-12, 1e-30, 2000, options={"sampler.method": 2, "sampler.cheb_lo_scale": 2.5})
    try:
        lev, n = 0, 6
        Ne = prob["sampler"][lev].Ne
        xi = Yarn5().jump(31).normals(n * Ne).reshape(n, Ne)
        sg, eg, itg = good.sampler_eval_batch(lev, xi)
        sb, eb, itb = bad.sampler_eval_batch(lev, xi)
        print(f"a-priori steps {itg.max()}, with the narrow interval {itb.min()}..{itb.max()}")
        assert itg.min() == itg.max()                 # no restart with the real estimate
        assert itb.max() < 2000 and itb.max() != itg.max()          # restarted blocks were needed, and they ended
        for j in range(n):
            so, eo, _ = orc.sampler_eval(lev, xi[j])
            assert rel_l2(eb[j], eo) < FIELD_TOL and rel_l2(sb[j], so) < FIELD_TOL
            assert rel_l2(eg[j], eo) < FIELD_TOL
    finally:
        good.close()
        bad.close()


@pytest.mark.parametrize("shape", [{}, {"cluster_size": 2, "cta_threads": 256}, {"group_size": 5}])
def test_functional_only_darcy_solve(prob, shape):
    """DarcySolver::SolveFwd returns Q and C only.  The MINRES that carries Q = obs . x by scalar recurrences and never forms
    the solution ("qoi_only", default) returns the Q of the solve that does form it, to round-off, after the same number of
    iterations -- one CTA per tile, cluster-split and as a grid group; the fused level loop likewise."""
    from common import make_context
    c1 = make_context(prob, True, 1e-12, 1e-30, 2000, options=dict(shape))
    c0 = make_context(prob, True, 1e-12, 1e-30, 2000, options=dict(shape, qoi_only=0))
    try:
        for lev in range(prob["nlevels"]):
            d = prob["darcy"][lev]
            k = np.exp(np.random.default_rng(5 + lev).standard_normal((7, d.Ne)))
            Q1, C1, _, it1 = c1.darcy_solve_batch(lev, k)                   # functional only
            Q0, C0, sol, it0 = c0.darcy_solve_batch(lev, k, want_sol=True)  # forms the solution
            Qs, _, sol1, its = c1.darcy_solve_batch(lev, k, want_sol=True)  # the default handle forms it when it is asked for
            assert np.array_equal(it1, it0) and np.array_equal(C1, C0)
            assert np.allclose(Q1, Q0, rtol=1e-11, atol=1e-13)
            assert np.allclose(Q0, sol @ np.asarray(d.obs), rtol=1e-12, atol=1e-14)
            assert np.array_equal(sol1, sol) and np.array_equal(Qs, Q0)
        for lev in (1, 0):
            _, r1, i1 = c1.mlmc_level_batch(lev, 9, 77, want_rows=True)
            _, r0, i0 = c0.mlmc_level_batch(lev, 9, 77, want_rows=True)
            assert i1 == i0 and np.allclose(r1[:, :3], r0[:, :3], rtol=1e-10, atol=1e-12)
    finally:
        c1.close()
        c0.close()


def test_chained_calls_reuse_device_results(prob):
    """Option "cache_results": Sample -> Eval -> SolveFwd with NULL for the vectors the library itself just produced
    gives bitwise the results of passing the host copies back; a mismatching request fails loudly."""
    from common import make_context
    from parelagmc_b200.capi import PmcError
    c = make_context(prob, True, 1e-6, 1e-12, 300, options={"cache_results": 1})
    try:
        lev, n = 0, 9
        xi = c.sampler_sample_batch(lev, n, 4242)
        sc_h, emb_h, _ = c.sampler_eval_batch(lev + 1, xi, xi_level=lev, use_init=0)
        qc_h = c.darcy_solve_batch(lev + 1, sc_h)[0]
        sf_h, _, _ = c.sampler_eval_batch(lev, xi, xi_level=lev, init_s=emb_h, init_level=lev + 1, use_init=1, want_embed=False)
        q_h = c.darcy_solve_batch(lev, sf_h)[0]
        c.sampler_sample_batch(lev, n, 4242)                                     # same noise again, now chained
        sc_d, emb_d, _ = c.sampler_eval_batch(lev + 1, None, xi_level=lev, use_init=0, nsamples=n)
        qc_d = c.darcy_solve_batch(lev + 1, None, nsamples=n)[0]
        sf_d, _, _ = c.sampler_eval_batch(lev, None, xi_level=lev, init_s=None, init_level=lev + 1, use_init=1, want_embed=False,
                                          nsamples=n)
        q_d = c.darcy_solve_batch(lev, None, nsamples=n)[0]
        for a, b in ((sc_h, sc_d), (emb_h, emb_d), (qc_h, qc_d), (sf_h, sf_d), (q_h, q_d)):
            assert np.array_equal(a, b)
        with pytest.raises(PmcError):
            c.darcy_solve_batch(lev, None, nsamples=n + 1)
        with pytest.raises(PmcError):
            c.sampler_eval_batch(lev + 1, None, xi_level=lev + 1, nsamples=n)
    finally:
        c.close()


def test_sampler_eval_coarse_noise_and_warm_start(ctx, orc, prob):
    """Eval(l+1, xi_l, ., init, false) then Eval(l, xi_l, ., init, true) as in MLMC_Manager.cpp:150-155."""
    from oracle.binding import Yarn5
    lev = 0
    Ne = prob["sampler"][lev].Ne
    n = 4
    xi = Yarn5().normals(n * Ne).reshape(n, Ne)
    sc, embc, _ = ctx.sampler_eval_batch(lev + 1, xi, xi_level=lev, use_init=0)
    sf, embf, itw = ctx.sampler_eval_batch(lev, xi, xi_level=lev, init_s=embc, init_level=lev + 1, use_init=1)
    _, _, itc = ctx.sampler_eval_batch(lev, xi, xi_level=lev)
    print(f"warm start iterations {itw.mean():.1f} vs cold {itc.mean():.1f}")
    for j in range(n):
        so, eo, _ = orc.sampler_eval(lev + 1, xi[j], xi_level=lev, embed_s=None, use_init=0)
        assert rel_l2(embc[j], eo) < FIELD_TOL
        so2, eo2, _ = orc.sampler_eval(lev, xi[j], xi_level=lev, embed_s=eo, init_level=lev + 1, use_init=1)
        assert rel_l2(embf[j], eo2) < FIELD_TOL


def test_mlmc_level_batch_matches_oracle(ctx, orc, prob):
    pos0 = 4242
    for lev, ns in [(2, 24), (1, 12), (0, 6)]:
        sums, rows, its = ctx.mlmc_level_batch(lev, ns, pos0, want_rows=True)
        osums, orows, _ = orc.mlmc_level(lev, ns, pos0, nthreads=8)
        assert np.allclose(rows[:, :3], orows[:, :3], rtol=1e-7, atol=1e-9), (lev, np.abs(rows - orows).max())
        assert np.array_equal(rows[:, 3], orows[:, 3])
        n = float(ns)
        for a, b in [(sums[1] / n, osums[1] / n), (sums[4] / n, osums[4] / n)]:
            assert abs(a - b) <= MOMENT_TOL * abs(b)

        def var(s2, s1):
            return (s2 / n - (s1 / n) ** 2) * n / (n - 1)

        assert abs(var(sums[0], sums[1]) - var(osums[0], osums[1])) <= MOMENT_TOL * abs(var(osums[0], osums[1])) + 1e-15
        assert abs(var(sums[3], sums[4]) - var(osums[3], osums[4])) <= MOMENT_TOL * abs(var(osums[3], osums[4]))
        assert its > 0


def test_batch_split_independence(ctx, prob):
    """Per-sample results must not depend on how realisations are grouped into launches."""
    ns, pos0, lev = 70, 99, 1
    ctx.set_batch(64, 0)
    s1, r1, _ = ctx.mlmc_level_batch(lev, ns, pos0, want_rows=True)
    ctx.set_batch(16, 0)
    s2, r2, _ = ctx.mlmc_level_batch(lev, ns, pos0, want_rows=True)
    ctx.set_batch(0, 0)
    assert np.array_equal(r1, r2)


def test_mc_level_batch(ctx, orc, prob):
    ns, pos0, lev = 10, 5, 1
    sums, rows, _ = ctx.mc_level_batch(lev, ns, pos0, want_rows=True)
    # single-level loop == coarsest-style loop of the oracle on that level
    osums, orows, _ = orc.mlmc_level(lev, ns, pos0, nthreads=8, nlevels=lev + 1)
    assert np.allclose(rows[:, 0], orows[:, 1], rtol=1e-7)
    assert abs(sums[1] - osums[4]) <= 1e-7 * abs(osums[4])


def test_quad_sampler_gaussian():
    """Config 1 (PDESamplerTest on inline_quad.mesh, Gaussian field, 2 levels 16/4 elements)."""
    from oracle.binding import Yarn5
    p = quad_problem(4, 2)
    c = make_context(p, lognormal=False)
    o = make_oracle(p, lognormal=False)
    try:
        for lev in range(2):
            Ne = p["sampler"][lev].Ne
            xi = Yarn5().normals(100 * Ne).reshape(100, Ne)
            s, emb, it = c.sampler_eval_batch(lev, xi)
            for j in range(0, 100, 9):
                so, _, _ = o.sampler_eval(lev, xi[j])
                assert rel_l2(s[j], so) < FIELD_TOL
            assert np.array_equal(s, emb)
    finally:
        c.close()


def test_empty_and_error_paths(ctx, prob):
    from parelagmc_b200.capi import PmcError
    Ne = prob["sampler"][0].Ne
    s, emb, it = ctx.sampler_eval_batch(0, np.zeros((0, Ne)), xi_level=0)
    assert s.shape == (0, Ne)
    sums, rows, its = ctx.mlmc_level_batch(0, 0, 0)
    assert np.all(sums == 0)
    with pytest.raises(PmcError):
        ctx.mlmc_level_batch(7, 1, 0)
    with pytest.raises(PmcError):
        ctx.sampler_eval_batch(0, np.zeros((1, prob["sampler"][1].Ne)), xi_level=1)  # noise coarser than level
    # zero noise -> zero field, converges in 0 iterations
    s, emb, it = ctx.sampler_eval_batch(1, np.zeros((2, prob["sampler"][1].Ne)))
    assert np.all(emb == 0) and np.all(it == 0) and np.all(s == 1.0)


def test_committed_golden_fixtures():
    """CUDA path against tests/golden/oracle_golden.json (tools/make_golden.py)."""
    import json, os
    from oracle.binding import Yarn5
    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.json")))
    p = hex_problem(4, 2)
    c = make_context(p)
    try:
        for pos, v in G["yarn5_ints"].items():
            assert c.rng_fill_int(int(pos), len(v)).tolist() == v
        c.rng_init(0.0, 1.0, 4, 3)
        assert c.rng_fill_int(0, 8).tolist() == G["yarn5_split_4_3"]
        c.rng_init(0.0, 1.0, 1, 0)
        g = G["hex4"]
        for lev in range(2):
            Ne = p["darcy"][lev].Ne
            k = np.stack([np.ones(Ne), np.exp(np.sin(np.arange(Ne, dtype=float)))])
            Q, _, _, _ = c.darcy_solve_batch(lev, k)
            assert abs(Q[0] - g["Q_k1"][lev]) <= 1e-9 * abs(g["Q_k1"][lev])
            assert abs(Q[1] - g["Q_ksin"][lev]) <= 1e-9 * abs(g["Q_ksin"][lev])
        xi = c.sampler_sample_batch(0, 1, 0)
        _, e0, _ = c.sampler_eval_batch(0, xi)
        _, e1, _ = c.sampler_eval_batch(1, xi, xi_level=0, use_init=0)
        assert rel_l2(e0[0], g["field_l0"]) < FIELD_TOL
        assert rel_l2(e1[0], g["field_l1_from_l0_noise"]) < FIELD_TOL
        for lev, ns in [(1, 4), (0, 3)]:
            sums, rows, _ = c.mlmc_level_batch(lev, ns, 1234, want_rows=True)
            assert np.allclose(rows, g[f"mlmc_rows_l{lev}"], rtol=1e-7, atol=1e-9)
            assert np.allclose(sums, g[f"mlmc_sums_l{lev}"], rtol=1e-6, atol=1e-9)
    finally:
        c.close()


def test_managers_on_gpu_match_oracle_backend(tmp_path):
    """MLMC_Manager / MC_Manager (host mirror of the reference managers) driving the CUDA library, against the same
    managers driving the oracle: estimates and variances within 1e-6 relative."""
    from common import OracleBackend
    from parelagmc_b200 import managers as MG
    p = hex_problem(4, 3)
    c = make_context(p, rel=1e-10)
    try:
        params = {"Use array samples": True, "Array number of samples": [6, 10, 16], "Mean square error": 1e6,
                  "Output filename for MC managers": ""}
        mg = MG.MLMC_Manager(None, 3, c, params, out=None)
        mg.wallTime = False
        mg.Run()
        mo = MG.MLMC_Manager(None, 3, OracleBackend(p, rel=1e-10), params, out=None)
        mo.wallTime = False
        mo.Run()
        assert np.allclose(mg.eY, mo.eY, rtol=MOMENT_TOL, atol=1e-12)
        assert np.allclose(mg.varY, mo.varY, rtol=MOMENT_TOL, atol=1e-14)
        assert np.allclose(mg.eQ, mo.eQ, rtol=MOMENT_TOL) and np.allclose(mg.varQ, mo.varQ, rtol=MOMENT_TOL)
        assert np.array_equal(mg.level_nsamples, mo.level_nsamples)
        sl = MG.MC_Manager(None, c, {"Number of samples": 8, "Mean square error": 1e6,
                                     "Output filename for MC managers": ""}, out=None)
        sl.wallTime = False
        sl.Run()
        so = MG.MC_Manager(None, OracleBackend(p, rel=1e-10), {"Number of samples": 8, "Mean square error": 1e6,
                                                                 "Output filename for MC managers": ""}, out=None)
        so.wallTime = False
        so.Run()
        assert abs(sl.eQ - so.eQ) <= MOMENT_TOL * abs(so.eQ) and abs(sl.varQ - so.varQ) <= MOMENT_TOL * abs(so.varQ)
    finally:
        c.close()


def _custom_problem(n, lengths, nlevels, bc, ess_value=0.0, corlen=0.2):
    """Anisotropic box, other boundary conditions, optionally non-zero essential data."""
    from parelagmc_b200 import hierarchy as H
    L = H.build_box_hierarchy(n, lengths, nlevels)
    SL = H.build_sampler_levels(L)
    DL = H.build_darcy_levels(L, **bc)
    if ess_value != 0.0:
        # prescribed outward flux ess_value per essential face (non-zero ess_data exercises the right-hand-side
        # fix-up of BlockMatrix::EliminateRowCol, /root/reference/src/DarcySolver.cpp:498)
        for lv, d in zip(L, DL):
            on_ess = np.asarray(bc["ess_attr"])[lv.bdr_attr - 1] != 0
            d.ess_data[lv.bdr_face[on_ess]] = ess_value * lv.bdr_sign[on_ess]
    return dict(levels=L, sampler=SL, darcy=DL, alpha=H.spde_alpha(corlen),
                g=H.matern_scaling_coefficient(corlen, len(n)), nlevels=nlevels)


@pytest.mark.parametrize("case", ["spe10_bc", "anisotropic", "ess_data"])
def test_other_problems_match_oracle(case):
    from parelagmc_b200 import hierarchy as H
    if case == "spe10_bc":
        p = _custom_problem([8, 8, 8], [2.0, 2.0, 2.0], 3, H.SPE10_BC)
    elif case == "anisotropic":
        p = _custom_problem([8, 12, 4], [2.0, 1.0, 0.5], 2, H.MLMC_DEFAULT_BC)
    else:
        p = _custom_problem([8, 8, 8], [2.0, 2.0, 2.0], 2, H.MLMC_DEFAULT_BC, ess_value=0.01)
    c = make_context(p)
    o = make_oracle(p)
    try:
        rng = np.random.default_rng(5)
        for lev in range(p["nlevels"]):
            d = p["darcy"][lev]
            k = np.exp(0.7 * rng.standard_normal((5, d.Ne)))
            Q, C, sol, it = c.darcy_solve_batch(lev, k, want_sol=True)
            for j in range(5):
                q, _, s, _ = o.darcy_solve(lev, k[j], want_sol=True)
                assert rel_l2(sol[j], s) < FIELD_TOL, (case, lev, j, rel_l2(sol[j], s))
                assert abs(Q[j] - q) <= 1e-8 * max(abs(q), 1e-3)
        lev = 0
        sums, rows, _ = c.mlmc_level_batch(lev, 6, 17, want_rows=True)
        osums, orows, _ = o.mlmc_level(lev, 6, 17, nthreads=4)
        assert np.allclose(rows[:, :3], orows[:, :3], rtol=1e-7, atol=1e-9)
    finally:
        c.close()


def test_tetrahedral_hierarchy_matches_oracle():
    """Unstructured-style input (RT0 on tetrahedra: 4 dofs per element, 7-entry mass rows, 4-entry prolongator rows):
    Darcy solutions, sampler fields and MLMC rows against the oracle, and Q = 2 for k == 1."""
    from common import tet_problem
    p = tet_problem(4, 2)
    c = make_context(p)
    o = make_oracle(p)
    try:
        rng = np.random.default_rng(11)
        for lev in range(2):
            d, s = p["darcy"][lev], p["sampler"][lev]
            Q, C, _, _ = c.darcy_solve_batch(lev, np.ones((3, d.Ne)))
            assert np.allclose(Q, 2.0, atol=1e-9) and np.all(C == d.N)
            k = np.exp(0.8 * rng.standard_normal((5, d.Ne)))
            Q, C, sol, it = c.darcy_solve_batch(lev, k, want_sol=True)
            for j in range(5):
                q, _, so, _ = o.darcy_solve(lev, k[j], want_sol=True)
                assert rel_l2(sol[j], so) < FIELD_TOL, (lev, j, rel_l2(sol[j], so))
                assert abs(Q[j] - q) <= 1e-8 * max(abs(q), 1e-3)
            xi = rng.standard_normal((4, s.Ne))
            kk, emb, _ = c.sampler_eval_batch(lev, xi)
            for j in range(4):
                ko, eo, _ = o.sampler_eval(lev, xi[j])
                assert rel_l2(emb[j], eo) < FIELD_TOL and rel_l2(kk[j], ko) < FIELD_TOL
        sums, rows, _ = c.mlmc_level_batch(0, 7, 31, want_rows=True)
        osums, orows, _ = o.mlmc_level(0, 7, 31, nthreads=4)
        assert np.allclose(rows[:, :3], orows[:, :3], rtol=1e-7, atol=1e-9)
        assert np.allclose(sums, osums, rtol=1e-6, atol=1e-9)
    finally:
        c.close()


def test_clone_runs_levels_concurrently_with_identical_results(ctx, prob):
    """pmc_clone: same hierarchy and stream of random numbers, own CUDA stream; results are bitwise those of the
    original handle, also when the level batches run from concurrent host threads."""
    from concurrent.futures import ThreadPoolExecutor
    c2 = ctx.clone()
    try:
        ref = [ctx.mlmc_level_batch(lev, 20, 321, want_rows=True)[1] for lev in (1, 0)]
        with ThreadPoolExecutor(max_workers=2) as pool:
            f1 = pool.submit(ctx.mlmc_level_batch, 1, 20, 321, None, True)
            f0 = pool.submit(c2.mlmc_level_batch, 0, 20, 321, None, True)
            got = [f1.result()[1], f0.result()[1]]
        assert np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1])
        # solutions cross the ABI in the caller's numbering on a clone as well (the library renumbers the RT dofs inside)
        d = prob["darcy"][0]
        k = np.exp(np.random.default_rng(21).standard_normal((3, d.Ne)))
        a, b = ctx.darcy_solve_batch(0, k, want_sol=True), c2.darcy_solve_batch(0, k, want_sol=True)
        assert np.array_equal(a[2], b[2]) and np.array_equal(a[0], b[0])
    finally:
        c2.close()


@pytest.mark.parametrize("kind", ["embedded", "l2proj", "embedded_tet"])
def test_enlarged_domain_samplers_match_oracle(kind):
    """SURVEY 8f-1 (EmbeddedPDESampler, matching enlarged mesh + meshP selection) and 8f-2 (L2ProjectionPDESampler apply,
    non-matching enlarged mesh + W^-1 G^T): sampler outputs on the forward mesh, and the fused MLMC level loop."""
    from common import enlarged_problem
    from oracle.binding import Yarn5
    p = enlarged_problem(kind)
    c = make_context(p)
    o = make_oracle(p)
    try:
        for lev in range(p["nlevels"]):
            Ne = p["sampler"][lev].Ne
            xi = Yarn5().jump(5 * lev).normals(5 * Ne).reshape(5, Ne)
            s, emb, it = c.sampler_eval_batch(lev, xi)
            assert s.shape == (5, p["darcy"][lev].Ne) and emb.shape == (5, Ne)
            for j in range(5):
                so, eo, _ = o.sampler_eval(lev, xi[j])
                assert rel_l2(emb[j], eo) < FIELD_TOL and rel_l2(s[j], so) < FIELD_TOL
        for lev, ns in [(1, 9), (0, 6)]:
            sums, rows, _ = c.mlmc_level_batch(lev, ns, 99, want_rows=True)
            osums, orows, _ = o.mlmc_level(lev, ns, 99, nthreads=4)
            assert np.allclose(rows[:, :3], orows[:, :3], rtol=1e-7, atol=1e-9), (kind, lev)
            assert np.array_equal(rows[:, 3], orows[:, 3])
    finally:
        c.close()


def test_bayes_level_batch_matches_oracle():
    """BayesianInverseProblem likelihood / R and the level loop of ML_BayesRatio_Manager::InitRun."""
    from common import bayes_problem
    p = bayes_problem()
    c = make_context(p)
    o = make_oracle(p)
    try:
        for lev in range(p["nlevels"]):
            c.upload_observations(lev, p["gobs"][lev], p["G_obs"], p["noise"])
            o.set_observations(lev, p["gobs"][lev], p["G_obs"], p["noise"])
        pos0 = p["pos_after_setup"]
        for lev, ns in [(1, 10), (0, 6)]:
            sums, rows, its = c.bayes_level_batch(lev, ns, pos0, want_rows=True)
            osums, orows = o.bayes_level(lev, ns, pos0, nthreads=4)
            assert np.allclose(rows, orows, rtol=1e-7, atol=1e-10), (lev, np.abs(rows - orows).max())
            assert np.allclose(sums, osums, rtol=1e-7, atol=1e-10)
            assert its > 0
        # the ratio estimate E[R]/E[Z] of the two paths agrees
        est = sums[10] / sums[4]
        assert est == pytest.approx(osums[10] / osums[4], rel=1e-6)
        c2 = c.clone()                       # observations travel with pmc_clone
        s2, r2, _ = c2.bayes_level_batch(0, 6, pos0, want_rows=True)
        c2.close()
        assert np.array_equal(r2, rows)
    finally:
        c.close()


@pytest.mark.parametrize("cs", [2, 4, 8])
def test_cluster_split_matches_single_cta(prob, cs):
    """A tile owned by a thread-block cluster (rows split over CS CTAs, cluster barriers, DSMEM dot products) gives the
    results of the one-CTA-per-tile kernel to round-off, and matches the oracle."""
    from parelagmc_b200.capi import Context
    def ctx_with(csize):
        c = Context(prob["nlevels"], 0)
        c.set_option("cta_threads", 256)
        c.set_option("cluster_size", csize)
        for l, s in enumerate(prob["sampler"]):
            c.upload_sampler_level(l, s, prob["alpha"], prob["g"], True)
        for l, d in enumerate(prob["darcy"]):
            c.upload_darcy_level(l, d)
        c.set_tolerances(1e-12, 1e-30, 2000)
        c.rng_init(0.0, 1.0, 1, 0)
        return c
    c1, ck = ctx_with(1), ctx_with(cs)
    try:
        for lev, ns in [(0, 10), (1, 7)]:
            _, r1, _ = c1.mlmc_level_batch(lev, ns, 55, want_rows=True)
            _, rk, _ = ck.mlmc_level_batch(lev, ns, 55, want_rows=True)
            assert np.allclose(r1, rk, rtol=1e-9, atol=1e-12), (cs, lev, np.abs(r1 - rk).max())
        d = prob["darcy"][0]
        k = np.exp(np.random.default_rng(3).standard_normal((5, d.Ne)))
        Q1, _, s1, _ = c1.darcy_solve_batch(0, k, want_sol=True)
        Qk, _, sk, _ = ck.darcy_solve_batch(0, k, want_sol=True)
        assert rel_l2(sk, s1) < 1e-9
    finally:
        c1.close()
        ck.close()


@pytest.mark.parametrize("group,solo", [(3, 0), (5, 64), (16, 4096), (37, 100)])
def test_grid_group_matches_single_cta(prob, group, solo):
    """A tile owned by a group of G CTAs of a cooperative launch (rows split G ways, barrier and dot-product shares
    through global memory, small operations on the group's first CTA alone) gives the results of the
    one-CTA-per-tile kernel to round-off."""
    from parelagmc_b200.capi import Context
    def ctx_with(g):
        c = Context(prob["nlevels"], 0)
        if g:
            c.set_option("group_size", g)
            c.set_option("solo_rows", solo)
        else:
            c.set_option("group_size", -1)
            c.set_option("cluster_size", 1)
        for l, s in enumerate(prob["sampler"]):
            c.upload_sampler_level(l, s, prob["alpha"], prob["g"], True)
        for l, d in enumerate(prob["darcy"]):
            c.upload_darcy_level(l, d)
        c.set_tolerances(1e-12, 1e-30, 2000)
        c.rng_init(0.0, 1.0, 1, 0)
        return c
    c1, cg = ctx_with(0), ctx_with(group)
    try:
        for lev, ns in [(0, 6), (1, 7), (2, 3)]:
            _, r1, _ = c1.mlmc_level_batch(lev, ns, 55, want_rows=True)
            for rep in range(2):                                  # the barrier words are reused launch after launch
                _, rg, _ = cg.mlmc_level_batch(lev, ns, 55, want_rows=True)
                assert np.allclose(r1, rg, rtol=1e-9, atol=1e-12), (group, lev, rep, np.abs(r1 - rg).max())
        d = prob["darcy"][0]
        k = np.exp(np.random.default_rng(3).standard_normal((5, d.Ne)))
        Q1, _, s1, it1 = c1.darcy_solve_batch(0, k, want_sol=True)
        Qg, _, sg, itg = cg.darcy_solve_batch(0, k, want_sol=True)
        assert rel_l2(sg, s1) < 1e-9 and np.array_equal(it1, itg)
    finally:
        c1.close()
        cg.close()


@pytest.mark.parametrize("opt", ["stage_operators", "defer_x", "fuse_coarse", "renumber", "single_wave", "cheb_three_term"])
def test_kernel_variants_agree(prob, opt):
    """The performance switches do not change what is computed: operator entries staged by TMA or read from L2, the
    MINRES solution update deferred within an iteration pair or not, the coarsest Chebyshev iteration as one
    shared-memory operation or step by step (all bitwise), the library's internal renumbering
    of the RT dofs on or off (round-off: rows are summed in a different order; solutions cross the ABI in the caller's
    numbering either way), one wave of smaller CTAs, the sampler's Chebyshev steps with or without a separate update
    vector (round-off)."""
    from parelagmc_b200.capi import Context
    def run(value):
        c = Context(prob["nlevels"], 0)
        c.set_option(opt, value)
        for l, s in enumerate(prob["sampler"]):
            c.upload_sampler_level(l, s, prob["alpha"], prob["g"], True)
        for l, d in enumerate(prob["darcy"]):
            c.upload_darcy_level(l, d)
        c.set_tolerances(1e-12, 1e-30, 2000)
        c.rng_init(0.0, 1.0, 1, 0)
        try:
            _, rows, its = c.mlmc_level_batch(0, 9, 77, want_rows=True)
            d = prob["darcy"][0]
            k = np.exp(np.random.default_rng(8).standard_normal((5, d.Ne)))
            Q, _, sol, it = c.darcy_solve_batch(0, k, want_sol=True)
            y = c.darcy_apply_batch(0, k, sol)
            return rows, its, Q, sol, it, y
        finally:
            c.close()
    a, b = run(0), run(1)
    if opt in ("stage_operators", "defer_x", "fuse_coarse", "single_wave"):
        if opt == "fuse_coarse":     # 9 realisations run cluster-split: the fused operation stays on one CTA -> round-off
            assert np.allclose(a[0], b[0], rtol=1e-13, atol=1e-14) and a[1] == b[1]
        else:
            assert np.array_equal(a[0], b[0]) and a[1] == b[1]
        assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4]) and np.array_equal(a[5], b[5])
    else:
        assert np.allclose(a[0][:, :3], b[0][:, :3], rtol=1e-9, atol=1e-12)
        assert rel_l2(a[3], b[3]) < 1e-9 and rel_l2(a[5], b[5]) < 1e-9 and np.allclose(a[2], b[2], rtol=1e-9)


def test_iteration_cap_and_options(prob):
    """Iteration cap (the reference's 'Maximum iterations'): realisations stop after max_iter iterations with finite
    output; options are validated and frozen once the preconditioner exists."""
    from parelagmc_b200.capi import Context, PmcError
    c = Context(prob["nlevels"], 0)
    try:
        with pytest.raises(PmcError):
            c.set_option("darcy.no_such_key", 1)
        with pytest.raises(PmcError):
            c.set_option("sampler.schur_ratio", 0.5)
        c.set_option("darcy.schur_degree", 3)
        for l, s in enumerate(prob["sampler"]):
            c.upload_sampler_level(l, s, prob["alpha"], prob["g"], True)
        for l, d in enumerate(prob["darcy"]):
            c.upload_darcy_level(l, d)
        with pytest.raises(PmcError):
            c.upload_darcy_level(0, prob["darcy"][0])          # uploaded twice
        c.set_tolerances(1e-14, 1e-300, 5)
        c.rng_init(0.0, 1.0, 1, 0)
        with pytest.raises(PmcError):
            c.rng_init(0.0, 1.0, 4, 7)                          # mypart out of range
        k = np.exp(np.random.default_rng(0).standard_normal((6, prob["darcy"][0].Ne)))
        Q, C, sol, it = c.darcy_solve_batch(0, k, want_sol=True)
        assert np.all(it == 5) and np.all(np.isfinite(sol)) and np.all(np.isfinite(Q))
        with pytest.raises(PmcError):
            c.set_option("darcy.schur_degree", 2)               # preconditioner already built
        c.set_tolerances(1e-10, 1e-300, 2000)
        Q2, _, _, it2 = c.darcy_solve_batch(0, k)
        assert it2.max() < 2000 and np.all(np.abs(Q2 - Q) < 0.5)
    finally:
        c.close()


def test_full_size_workload_properties():
    """BASELINE configs[1] at full size (16^3 / 8^3 / 4^3, N = 17152 / 2240 / 304; 1000 / 3000 / 6000 realisations, the
    bench's stopping rule): size-independent properties of the CUDA path -- the deterministic known answer Q = 2 on
    every level, rows that telescope exactly (Y = Q_l - Q_{l+1}), sums that are the sums of the rows, results that do
    not depend on how the batch is cut (bitwise), and agreement with the oracle on a bounded subsample."""
    p = hex_problem(16, 3)
    c = make_context(p, True, 1e-6, 1e-12, 300)
    c1 = make_context(p, True, 1e-6, 1e-12, 300, options={"cta_threads": 256, "cluster_size": 1, "group_size": -1})
    o = make_oracle(p, True, 1e-6, 1e-12, 300)
    try:
        for lev, dofs in enumerate([17152, 2240, 304]):
            Q, C, _, it = c.darcy_solve_batch(lev, np.ones((4, p["darcy"][lev].Ne)))
            assert np.allclose(Q, 2.0, atol=1e-5) and np.all(C == dofs) and np.all(it < 300)
        pos = 0
        for lev, ns in [(2, 6000), (1, 3000), (0, 1000)]:
            sums, rows, its = c.mlmc_level_batch(lev, ns, pos, want_rows=True)
            assert rows.shape == (ns, 4) and np.all(np.isfinite(rows)) and its > 0
            if lev < 2:
                assert np.array_equal(rows[:, 0], rows[:, 1] - rows[:, 2])
            y, q = rows[:, 0], rows[:, 1]
            ref = [np.sum(y * y), np.sum(y), np.sum(np.abs(y)), np.sum(q * q), np.sum(q), np.sum(np.abs(q)), np.sum(rows[:, 3]),
                   np.sum(y ** 3), np.sum(y ** 4)]
            assert np.allclose(sums, ref, rtol=1e-10)
            # cutting the batch: bitwise the same rows for one launch shape (the dot products are reduced in an order that
            # depends on the CTA and cluster size only), equal to round-off when the library picks another shape for the halves
            half = ns // 2
            _, ra, _ = c.mlmc_level_batch(lev, half, pos, want_rows=True)
            _, rb, _ = c.mlmc_level_batch(lev, ns - half, pos + half * p["sampler"][lev].Ne, want_rows=True)
            assert np.allclose(np.vstack([ra, rb]), rows, rtol=1e-6, atol=1e-9)
            _, f1, _ = c1.mlmc_level_batch(lev, ns, pos, want_rows=True)
            _, fa, _ = c1.mlmc_level_batch(lev, half, pos, want_rows=True)
            _, fb, _ = c1.mlmc_level_batch(lev, ns - half, pos + half * p["sampler"][lev].Ne, want_rows=True)
            assert np.array_equal(np.vstack([fa, fb]), f1)
            _, orows, _ = o.mlmc_level(lev, 16, pos, nthreads=8)
            assert np.allclose(rows[:16, :3], orows[:, :3], rtol=1e-4, atol=2e-5)      # both stop at rel 1e-6, different preconditioners
            pos += ns * p["sampler"][lev].Ne
        assert 2.0 < sums[4] / 1000 < 3.2                                               # E[Q_0] of the lognormal problem
    finally:
        c.close()
        c1.close()


@pytest.mark.parametrize("tets", [False, True])
def test_unstructured_agglomerates_match_oracle(tets):
    """Irregular agglomerates with ParELAG-style order-0 coarse spaces (what "Unstructured coarsening" hands over): dense
    agglomerate mass blocks of up to ~18 x 18, operator rows of up to ~35 entries (beyond the 8-entry staging buffers:
    the entries of those slices are read from L2).  Darcy solutions, sampler fields and MLMC rows on every level against
    the oracle."""
    from common import agglomerated_problem, make_context, make_oracle
    p = agglomerated_problem(n=4 if tets else 8, nlevels=3, tets=tets)
    o = make_oracle(p)
    c = make_context(p, True, 1e-12, 1e-30, 3000)
    try:
        rng = np.random.default_rng(8)
        for lev in range(p["nlevels"]):
            d, sl = p["darcy"][lev], p["sampler"][lev]
            k = np.exp(rng.standard_normal((3, d.Ne)))
            Q, C, sol, it = c.darcy_solve_batch(lev, k, want_sol=True)
            xi = rng.standard_normal((3, sl.Ne))
            s, emb, _ = c.sampler_eval_batch(lev, xi)
            for j in range(3):
                Qo, _, so, _ = o.darcy_solve(lev, k[j], want_sol=True)
                assert rel_l2(sol[j], so) < FIELD_TOL and Q[j] == pytest.approx(Qo, rel=1e-8)
                fo, eo, _ = o.sampler_eval(lev, xi[j])
                assert rel_l2(emb[j], eo) < FIELD_TOL and rel_l2(s[j], fo) < FIELD_TOL
        c_l2 = make_context(p, True, 1e-12, 1e-30, 3000, options={"stage_wide": 0})   # wide slices read from L2 instead
        pos = 0
        for lev, ns in [(2, 5), (1, 6), (0, 5)]:
            sums, rows, _ = c.mlmc_level_batch(lev, ns, pos, want_rows=True)
            osums, orows, _ = o.mlmc_level(lev, ns, pos, nthreads=2)
            assert np.allclose(rows[:, :3], orows[:, :3], rtol=1e-7, atol=1e-9)
            assert np.allclose(sums, osums, rtol=1e-6)
            _, rows2, _ = c_l2.mlmc_level_batch(lev, ns, pos, want_rows=True)
            assert np.array_equal(rows, rows2)       # chunk-staged and L2-read entries: the same arithmetic in the same order
            pos += ns * p["sampler"][lev].Ne
        c_l2.close()
    finally:
        c.close()
