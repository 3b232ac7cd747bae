"""Shared problem builders for the tests: the hierarchy (host once), the oracle problem (checker) and the
GPU context (product), all fed with the same arrays."""
from __future__ import annotations

import functools

import numpy as np

from parelagmc_b200 import hierarchy as H


@functools.lru_cache(maxsize=None)
def hex_problem(n=16, nlevels=3, corlen=0.1, length=2.0, bc="mlmc"):
    """The reference's default MLMC problem (examples/MLMC.cpp with CreateMLMCParameterList.hpp):
    hex n^3 on [0,2]^3 coarsened by 2 per direction."""
    L = H.build_box_hierarchy([n] * 3, [length] * 3, nlevels)
    SL = H.build_sampler_levels(L)
    DL = H.build_darcy_levels(L, **(H.MLMC_DEFAULT_BC if bc == "mlmc" else H.SPE10_BC))
    return dict(levels=L, sampler=SL, darcy=DL, alpha=H.spde_alpha(corlen),
                g=H.matern_scaling_coefficient(corlen, 3), nlevels=nlevels)


@functools.lru_cache(maxsize=None)
def tet_problem(n=4, nlevels=2, corlen=0.3, length=2.0):
    """Tetrahedral meshes (BASELINE configs[2] runs on a tet mesh): Kuhn triangulation of an n^3 grid on [0,2]^3,
    nested coarsening, RT0 / P0 on every level, the MLMC drivers' boundary conditions."""
    L = H.build_tet_hierarchy(n, length, nlevels)
    SL = H.build_sampler_levels(L)
    DL = H.build_darcy_levels(L, **H.MLMC_DEFAULT_BC)
    return dict(levels=L, sampler=SL, darcy=DL, alpha=H.spde_alpha(corlen),
                g=H.matern_scaling_coefficient(corlen, 3), nlevels=nlevels)


@functools.lru_cache(maxsize=None)
def quad_problem(n=4, nlevels=2, corlen=0.1):
    """PDESamplerTest on inline_quad.mesh: unit square, 2x2 quads refined to n x n."""
    L = H.build_box_hierarchy([n] * 2, [1.0] * 2, nlevels)
    SL = H.build_sampler_levels(L)
    return dict(levels=L, sampler=SL, darcy=None, alpha=H.spde_alpha(corlen),
                g=H.matern_scaling_coefficient(corlen, 2), nlevels=nlevels)


@functools.lru_cache(maxsize=None)
def enlarged_problem(kind="embedded", n=8, nlevels=2, corlen=0.3):
    """Enlarged-domain samplers: the forward problem lives on the n^3 box [0,2]^3, the SPDE on a
    larger box.  kind = "embedded": matching mesh, 2 cells of padding per side, meshP selection
    (EmbeddedPDESampler); kind = "l2proj": NON-matching enlarged mesh (different h), mortar transfer W^-1 G^T
    (L2ProjectionPDESampler)."""
    import dataclasses
    h = 2.0 / n
    if kind == "embedded_tet":      # BASELINE configs[2]: tetrahedral forward mesh inside a tetrahedral enlarged mesh
        n, pad = 4, 2
        h = 2.0 / n
        orig = H.build_tet_hierarchy(n, 2.0, nlevels)
        emb = H.build_tet_hierarchy(n + 2 * pad, 2.0 + 2 * pad * h, nlevels)
        T = [(m, None) for m in H.tet_embedded_selection(n, pad, nlevels)]
        SL = [dataclasses.replace(s, T=t[0], Tscale=t[1]) for s, t in zip(H.build_sampler_levels(emb), T)]
        DL = H.build_darcy_levels(orig, **H.MLMC_DEFAULT_BC)
        return dict(levels=orig, embed_levels=emb, sampler=SL, darcy=DL, alpha=H.spde_alpha(corlen),
                    g=H.matern_scaling_coefficient(corlen, 3), nlevels=nlevels)
    orig = H.build_box_hierarchy([n] * 3, [2.0] * 3, nlevels)
    if kind == "embedded":
        pad = 2
        emb = H.build_box_hierarchy([n + 2 * pad] * 3, [2.0 + 2 * pad * h] * 3, nlevels)
        T = [(m, None) for m in H.embedded_selection(orig, emb, pad)]
    else:
        emb = H.build_box_hierarchy([n + 2] * 3, [2.0] * 3, nlevels)   # same box, non-matching cells (h' = 2/(n+2))
        T = H.l2_projection_transfers(orig, emb)
    SL = [dataclasses.replace(s, T=t[0], Tscale=t[1]) for s, t in zip(H.build_sampler_levels(emb), T)]
    DL = H.build_darcy_levels(orig, **H.MLMC_DEFAULT_BC)
    return dict(levels=orig, embed_levels=emb, sampler=SL, darcy=DL, alpha=H.spde_alpha(corlen),
                g=H.matern_scaling_coefficient(corlen, 3), nlevels=nlevels)


def make_oracle(p, lognormal=True, rel=1e-12, abs_=1e-30, maxit=2000):
    from oracle.binding import OracleProblem
    op = OracleProblem(p["sampler"], p["darcy"], p["alpha"], p["g"], lognormal)
    op.set_tolerances(rel, abs_, maxit)
    return op


def make_context(p, lognormal=True, rel=1e-12, abs_=1e-30, maxit=2000, device=0, options=None):
    import os
    from parelagmc_b200.capi import Context
    ctx = Context(p["nlevels"], device)
    options = dict(options or {})
    for kv in filter(None, os.environ.get("PMC_OPTS", "").split(",")):    # diagnostic: PMC_OPTS="key=value,..."
        k, v = kv.split("=")
        options[k] = float(v)
    for k, v in options.items():
        ctx.set_option(k, v)
    for l, s in enumerate(p["sampler"]):
        ctx.upload_sampler_level(l, s, p["alpha"], p["g"], lognormal)
    if p["darcy"] is not None:
        for l, d in enumerate(p["darcy"]):
            ctx.upload_darcy_level(l, d)
    ctx.set_tolerances(rel, abs_, maxit)
    ctx.rng_init(0.0, 1.0, 1, 0)
    ctx.prepare()
    return ctx


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


class OracleBackend:
    """Test double with the `capi.Context` manager-facing surface, computed by the CPU oracle.  Lets the host-side
    manager logic (sample sharding, stream positions, all-reduce, statistics) be tested without a GPU."""

    def __init__(self, p, rel=1e-8, abs_=1e-30, maxit=1000, threads=4):
        self.o = make_oracle(p, True, rel, abs_, maxit)
        self.nlevels = p["nlevels"]
        self.Ne = [s.Ne for s in p["sampler"]]
        self.Nf = [s.Nf for s in p["sampler"]]
        self.dNe = [d.Ne for d in p["darcy"]]
        self.dNf = [d.Nf for d in p["darcy"]]
        self.threads = threads
        self.calls = []

    def mlmc_level_batch(self, level, nsamples, pos0, nlevels=None, want_rows=False, sums=None):
        s, rows, its = self.o.mlmc_level(level, nsamples, pos0, nthreads=self.threads, nlevels=nlevels)
        self.calls.append((level, nsamples, pos0))
        if sums is None:
            sums = np.zeros(9)
        sums += s
        return sums, (rows if want_rows else None), its

    def bayes_level_batch(self, level, nsamples, pos0, nlevels=None, want_rows=False, sums=None):
        s, rows = self.o.bayes_level(level, nsamples, pos0, nthreads=self.threads, nlevels=nlevels)
        if sums is None:
            sums = np.zeros(20)
        sums += s
        return sums, (rows if want_rows else None), 0

    def mc_level_batch(self, level, nsamples, pos0, want_rows=False, sums=None):
        s, rows, its = self.o.mlmc_level(level, nsamples, pos0, nthreads=self.threads, nlevels=level + 1)
        if sums is None:
            sums = np.zeros(4)
        sums += np.array([s[3], s[4], s[5], s[6]])
        return sums, (rows[:, [1, 3]] if want_rows else None), its


@functools.lru_cache(maxsize=None)
def bayes_problem(n=8, nlevels=2, noise=0.05):
    """RatioEstimator_MLMC-style set-up (BASELINE configs[3]) on the hex box: local-average-pressure observations at
    three interior points, synthetic data G_obs = G(k(xi_0)) + eta generated by the oracle from stream position 0 (the
    noise eta from a second, default-seeded stream, as in
    /root/reference/src/BayesianInverseProblem.cpp:159-175)."""
    from oracle.binding import Yarn5
    p = dict(hex_problem(n, nlevels))
    coords = [(0.5, 0.5, 0.5), (1.0, 1.25, 0.75), (1.5, 0.75, 1.5)]
    p["gobs"] = H.observation_functionals(p["levels"], coords, eps=0.3)
    o = make_oracle(p)
    Ne0 = p["sampler"][0].Ne
    xi0 = Yarn5().normals(Ne0)                           # prior.Sample(0, xi)
    k0, _, _ = o.sampler_eval(0, xi0)
    _, _, sol, _ = o.darcy_solve(0, k0, want_sol=True)
    pr = sol[p["darcy"][0].Nf:]
    G = np.array([g @ pr / g.sum() for g in p["gobs"][0]])
    eta = Yarn5().normals(len(G), 0.0, np.sqrt(noise))   # NormalDistributionSampler noise_dist(0, noise)
    p["G_obs"] = G + eta
    p["noise"] = noise
    p["pos_after_setup"] = Ne0                            # the prior draw of the set-up consumed Ne0 positions
    return p


# ------------------------------------------------------------------------------------------------------------------
# The reference's own ctest problems (examples/CMakeLists.txt:55-118), in the reference's element numbering
# ------------------------------------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def reference_sampler_problem():
    """PDESamplerTest.exe defaults (`/root/reference/examples/example_helpers/CreateSamplerParameterList.hpp:26-33`):
    Build3DHexMesh 4^3 on [0,2]^3, 2 parallel refinements (16^3 / 8^3 / 4^3), Gaussian, variance 1, corlen 0.1, elements
    numbered as MFEM numbers them under UniformRefinement."""
    p = dict(hex_problem(16, 3, 0.1))
    num = H.mfem_refined_box_numbering([4] * 3, 2)
    p["sampler"] = H.renumber_sampler_levels(p["sampler"], num)
    p["darcy"] = H.renumber_darcy_levels(p["darcy"], num)
    p["numbering"] = num
    return p


@functools.lru_cache(maxsize=None)
def reference_enlarged_problem(with_transfer=True):
    """The matching enlarged-mesh pair of EmbeddedPDESamplerTest / DarcyTest_RandomInput / LikelihoodExample /
    RatioEstimator_MC (`/root/reference/examples/example_helpers/Build3DMesh.hpp:23-40`): forward mesh 4^3 on [0,2]^3,
    enlarged mesh 6^3 on [-0.5,2.5]^3, both refined twice (16^3 in 24^3, 4 fine cells of padding), MFEM numbering on
    both.  On matching meshes the L2 projection W^-1 G^T of L2ProjectionPDESampler is the 0/1 selection meshP."""
    import dataclasses
    orig = H.build_box_hierarchy([16] * 3, [2.0] * 3, 3)
    emb = H.build_box_hierarchy([24] * 3, [3.0] * 3, 3)
    num_o = H.mfem_refined_box_numbering([4] * 3, 2)
    num_e = H.mfem_refined_box_numbering([6] * 3, 2)
    SL = H.build_sampler_levels(emb)
    T = H.embedded_selection(orig, emb, 4)
    inside = [np.asarray(t.sum(axis=0)).ravel()[n] > 0 for t, n in zip(T, num_e)]   # enlarged elements inside [0,2]^3
    if with_transfer:
        SL = [dataclasses.replace(s, T=t, Tscale=None) for s, t in zip(SL, T)]
    SL = H.renumber_sampler_levels(SL, num_e, num_o)
    DL = H.renumber_darcy_levels(H.build_darcy_levels(orig, **H.MLMC_DEFAULT_BC), num_o)
    # BayesianInverseProblem defaults (CreateBayesianParameterList.hpp:48-51): one point (1,1,1), noise 0.1
    gobs = [g[:, n] for g, n in zip(H.observation_functionals(orig, [(1.0, 1.0, 1.0)], 0.01, rule="bbox"), num_o)]
    return dict(levels=orig, embed_levels=emb, sampler=SL, darcy=DL if with_transfer else None,
                alpha=H.spde_alpha(0.1), g=H.matern_scaling_coefficient(0.1, 3), nlevels=3, inside=inside, gobs=gobs,
                noise=0.1)


@functools.lru_cache(maxsize=None)
def spe10_problem(scale=0.25, nlevels=4):
    """BASELINE configs[4] at reduced size: `Create_SPE10_Mesh(3, {60,220,85}, {20,10,2})` cells
    (`/root/reference/src/MeshUtilities.cpp:21-37`: 1200 x 2200 x 170 ft), correlation length 100
    (`examples/SPE10/spe10_3D_parameters.xml:20`), the SPE10 boundary attributes (`:45-49`), unit mass coefficient
    (`examples/SPE10/SPE10_MLMC.cpp:209`).  The cells keep their 10 : 5 : 1 aspect ratio whatever the scale."""
    n = [max(8, int(round(x * scale))) for x in (60, 220, 85)]
    L = H.build_box_hierarchy(n, [1200.0, 2200.0, 170.0], nlevels)
    return dict(levels=L, sampler=H.build_sampler_levels(L), darcy=H.build_darcy_levels(L, **H.SPE10_BC),
                alpha=H.spde_alpha(100.0), g=H.matern_scaling_coefficient(100.0, 3), nlevels=nlevels, grid=n)


@functools.lru_cache(maxsize=None)
def agglomerated_problem(n=8, nlevels=3, corlen=0.3, target=8, seed=1, tets=False):
    """Unstructured coarsening ("Unstructured coarsening" = true in the reference's drivers: METIS agglomerates +
    ParELAG's order-0 coarse spaces, /root/reference/src/Utilities.cpp:125-155): hex n^3 (or Kuhn tetrahedra) on [0,2]^3
    with irregular agglomerates of ~`target` elements per level; coarse agglomerates have 4-18 faces, so the operators
    have rows of up to ~35 entries (the structured meshes stop at 7)."""
    fine = (H.build_tet_hierarchy(n, 2.0, 1) if tets else H.build_box_hierarchy([n] * 3, [2.0] * 3, 1))[0]
    L = H.build_agglomerated_hierarchy(fine, nlevels, target=target, seed=seed)
    return dict(levels=L, sampler=H.build_sampler_levels(L), darcy=H.build_darcy_levels(L, **H.MLMC_DEFAULT_BC),
                alpha=H.spde_alpha(corlen), g=H.matern_scaling_coefficient(corlen, 3), nlevels=nlevels)
