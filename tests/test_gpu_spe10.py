"""What the SPE10-scale configuration (BASELINE.json configs[4]) depends on, each against the oracle or against the
one-CTA-per-tile kernel: strongly anisotropic cells (the library builds its own aggregation coarse spaces for the Schur
V-cycle, `darcy.amg` / `sampler.amg`), the long correlation length (the sampler stays on MINRES with a V-cycle), levels
too large for one CTA per tile (thread-block clusters, grid groups), and the 4-level MLMC loop on the SPE10 geometry."""
import numpy as np
import pytest

from common import hex_problem, make_context, make_oracle, rel_l2, spe10_problem

pytestmark = pytest.mark.gpu


def _ctx(p, rel=1e-10, options=None):
    return make_context(p, True, rel, 1e-30, 5000, options=options)


def test_aggregation_coarse_spaces_match_oracle():
    """`amg` = 1 (strength-aware pairwise aggregation) against `amg` = 0 (the hierarchy's own L2 prolongators) and the
    oracle on SPE10-shaped cells: same solutions, fewer iterations, and the automatic choice (-1) is the aggregation."""
    p = spe10_problem(0.125, 3)        # 8 x 28 x 11 cells of 150 x 78.6 x 15.5 ft
    o = make_oracle(p, True, 1e-10, 1e-30, 5000)
    ctxs = {a: _ctx(p, options={"darcy.amg": a, "sampler.amg": a}) for a in (0, 1, -1)}
    try:
        rng = np.random.default_rng(11)
        for lev in range(2):
            d = p["darcy"][lev]
            k = np.exp(rng.standard_normal((3, d.Ne)))
            res = {a: c.darcy_solve_batch(lev, k, want_sol=True) for a, c in ctxs.items()}
            for j in range(3):
                Qo, _, so, _ = o.darcy_solve(lev, k[j], want_sol=True)
                for a in ctxs:
                    assert rel_l2(res[a][2][j], so) < 1e-7, (lev, a)
                    assert res[a][0][j] == pytest.approx(Qo, rel=1e-8)
            it = {a: int(res[a][3].sum()) for a in ctxs}
            print(f"level {lev}: Darcy iterations amg=0 {it[0]}, amg=1 {it[1]}, auto {it[-1]}")
            assert it[-1] == it[1]
            if lev == 0:
                assert it[1] < it[0]
        # the sampler at correlation length 100: MINRES + V-cycle (alpha W does not dominate), all three variants
        s0 = p["sampler"][0]
        xi = rng.standard_normal((2, s0.Ne))
        for a, c in ctxs.items():
            s, emb, its = c.sampler_eval_batch(0, xi)
            for j in range(2):
                so, eo, _ = o.sampler_eval(0, xi[j])
                assert rel_l2(emb[j], eo) < 1e-7, a
    finally:
        for c in ctxs.values():
            c.close()


def test_clusters_and_grid_groups_on_a_large_level():
    """A 64^3 level (N = 1.06 M rows, 128 MB per tile): a handful of realisations run as one tile per thread-block cluster
    of 8 and as one tile per group of 37 / 148 co-resident CTAs; both give the one-CTA-per-tile results."""
    p = hex_problem(64, 2)
    opts = {"single": {"cluster_size": 1, "group_size": -1}, "cluster8": {"cluster_size": 8, "group_size": -1},
            "group37": {"group_size": 37}, "group148": {"group_size": 148}, "auto": {}}
    rows, its = {}, {}
    for name, o in opts.items():
        c = make_context(p, True, 1e-10, 1e-30, 2000, options=o)
        try:
            _, rows[name], its[name] = c.mlmc_level_batch(0, 4, 99, want_rows=True)
        finally:
            c.close()
    for name in opts:
        assert np.allclose(rows[name], rows["single"], rtol=1e-8, atol=1e-11), (name, np.abs(rows[name] - rows["single"]).max())
        assert its[name] == its["single"], name


def test_scaled_spe10_mlmc_rows_match_oracle():
    """SPE10_MLMC at quarter scale (15 x 55 x 21 cells, 4 levels, N = 58 k / 8.2 k / 1.3 k / 0.2 k): the rows of every level
    of `MLMC_Manager::InitRun` against the oracle's."""
    p = spe10_problem(0.25, 4)
    o = make_oracle(p, True, 1e-10, 1e-30, 5000)
    c = _ctx(p)
    try:
        pos = 0
        for lev, ns in [(3, 6), (2, 4), (1, 3), (0, 2)]:
            sums, rows, its = c.mlmc_level_batch(lev, ns, pos, want_rows=True)
            osums, orows, _ = o.mlmc_level(lev, ns, pos, nthreads=4)
            assert np.allclose(rows[:, :3], orows[:, :3], rtol=1e-6, atol=1e-8), (lev, np.abs(rows[:, :3] - orows[:, :3]).max())
            assert np.allclose(sums, osums, rtol=1e-6)
            print(f"level {lev}: {its / ns:.0f} MINRES iterations per realisation")
            pos += ns * p["sampler"][lev].Ne
    finally:
        c.close()
