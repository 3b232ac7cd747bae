"""Iteration-count distribution of the batched solves at the bench tolerances (diagnostic)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import hex_problem, make_context
p = hex_problem(16, 3)
rel = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-6
opts = dict(kv.split("=") for kv in sys.argv[2:])
import common
def _ctx():
    from parelagmc_b200.capi import Context
    c = Context(p["nlevels"], 0)
    for k, v in opts.items():
        c.set_option(k, float(v))
    for l, s in enumerate(p["sampler"]):
        c.upload_sampler_level(l, s, p["alpha"], p["g"], True)
    for l, d in enumerate(p["darcy"]):
        c.upload_darcy_level(l, d)
    c.set_tolerances(rel, 1e-12, 300); c.rng_init(0.0, 1.0, 1, 0); c.prepare()
    return c
ctx = _ctx()
print("options", opts, "rel", rel)
n = 1024
for lev in range(3):
    xi = ctx.sampler_sample_batch(lev, n, 0)
    t = time.time(); s, emb, it = ctx.sampler_eval_batch(lev, xi); t1 = time.time() - t
    t = time.time(); Q, C, _, itd = ctx.darcy_solve_batch(lev, s); t2 = time.time() - t
    print(f"level {lev}: sampler its min/mean/max {it.min()}/{it.mean():.1f}/{it.max()} ({t1*1e3:.1f} ms)  "
          f"darcy {itd.min()}/{itd.mean():.1f}/{itd.max()} ({t2*1e3:.1f} ms)  Q mean {Q.mean():.4f}")
    if lev < 2:
        sc, embc, itc = ctx.sampler_eval_batch(lev + 1, xi, xi_level=lev, use_init=0)
        sf, _, itw = ctx.sampler_eval_batch(lev, xi, xi_level=lev, init_s=embc, init_level=lev + 1, use_init=1)
        print(f"   warm-started sampler its {itw.min()}/{itw.mean():.1f}/{itw.max()}")
ctx.close()
