import os, sys
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from common import hex_problem
from parelagmc_b200.capi import Context
prob = hex_problem(int(os.environ.get("PMC_N", "8")), 3)
def ctx_with(csize, stage, nt=256):
    c = Context(prob["nlevels"], 0)
    c.set_option("cta_threads", nt)
    c.set_option("cluster_size", csize)
    c.set_option("stage_operators", stage)
    for l, s in enumerate(prob["sampler"]):
        c.upload_sampler_level(l, s, prob["alpha"], prob["g"], True)
    for l, d in enumerate(prob["darcy"]):
        c.upload_darcy_level(l, d)
    c.set_tolerances(1e-12, 1e-30, 2000)
    c.rng_init(0.0, 1.0, 1, 0)
    return c
ref = ctx_with(1, 0)
R = {}
for lev, ns in [(0, 10), (1, 7), (2, 7)]:
    R[lev] = ref.mlmc_level_batch(lev, ns, 55, want_rows=True)[1]
for cs in (1, 2, 4, 8):
    for stage in (0, 1):
        for nt in (256, 512):
            c = ctx_with(cs, stage, nt)
            out = []
            for lev, ns in [(0, 10), (1, 7), (2, 7)]:
                for rep in range(2):
                    r = c.mlmc_level_batch(lev, ns, 55, want_rows=True)[1]
                    out.append(f"{np.abs(r - R[lev]).max():.2e}")
            print(f"cs={cs} stage={stage} nt={nt}: ", " ".join(out), flush=True)
            c.close()
