"""Iterations per solve of the Darcy and the sampler systems on the SPE10 geometry for a list of option sets.
  python tools/spe10_prec_sweep.py [--scale 0.5] [--level 0] "k=v,k=v" "k=v" ...      ('-' = defaults)"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import spe10_problem, make_context
ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=0.5)
ap.add_argument("--level", type=int, default=0)
ap.add_argument("--ns", type=int, default=8)
ap.add_argument("cfgs", nargs="*", default=["-"])
a = ap.parse_args()
p = spe10_problem(a.scale, 4)
lev = a.level
print("grid", p["grid"], "N", p["darcy"][lev].N, flush=True)
for cfg in a.cfgs:
    opts = {} if cfg == "-" else {kv.split("=")[0]: float(kv.split("=")[1]) for kv in cfg.split(",")}
    c = make_context(p, True, 1e-6, 1e-14, 3000, options=opts)
    xi = c.sampler_sample_batch(lev, a.ns, 0)
    c.sampler_eval_batch(lev, xi[:4], want_embed=False)
    c.reset_stats(); t0 = time.perf_counter()
    k, _, sits = c.sampler_eval_batch(lev, xi, want_embed=False)
    ts = c.kernel_stats()["kernel"]["ms"]
    c.darcy_solve_batch(lev, k[:4])
    c.reset_stats()
    Q, _, _, dits = c.darcy_solve_batch(lev, k)
    td = c.kernel_stats()["kernel"]["ms"]
    print(f"{cfg:50s} sampler its {sits.mean():6.1f} ({ts:7.1f} ms)   darcy its {dits.mean():6.1f} ({td:7.1f} ms)  Q {Q.mean():.6f}", flush=True)
    c.close()
