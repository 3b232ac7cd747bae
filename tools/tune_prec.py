"""Times the fused level-0 batch of the bench for preconditioner variants (diagnostic)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import hex_problem
from parelagmc_b200.capi import Context
p = hex_problem(16, 3)
S = [1000, 3000, 6000]
variants = [{}, {"darcy.mass_degree": 1}, {"darcy.omega": 2.5}, {"darcy.mass_degree": 1, "darcy.omega": 2.5},
            {"darcy.schur_degree": 1}, {"darcy.mass_degree": 1, "darcy.omega": 2.5, "sampler.mass_degree": 1},
            {"darcy.mass_degree": 1, "darcy.omega": 2.5, "darcy.schur_degree": 3, "darcy.schur_ratio": 6},
            {"darcy.mass_degree": 3, "darcy.omega": 2.5}]
for opts in variants:
    c = Context(3, 0)
    for k, v in opts.items():
        c.set_option(k, float(v))
    for l, s in enumerate(p["sampler"]):
        c.upload_sampler_level(l, s, p["alpha"], p["g"], True)
    for l, d in enumerate(p["darcy"]):
        c.upload_darcy_level(l, d)
    c.set_tolerances(1e-6, 1e-12, 300); c.rng_init(0.0, 1.0, 1, 0); c.prepare()
    out = []
    for lev in (2, 1, 0):
        c.mlmc_level_batch(lev, S[lev], 0)
        c.reset_stats()
        sums, _, its = c.mlmc_level_batch(lev, S[lev], 0)
        k = c.kernel_stats()["kernel"]
        out.append(f"L{lev}: {k['ms']:6.2f} ms its {its/S[lev]:6.1f} E[Y]={sums[1]/S[lev]:.5f}")
    print(opts, " | ".join(out))
    c.close()
