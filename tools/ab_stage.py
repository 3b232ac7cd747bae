"""A/B of an option on the bench's level batches (diagnostic): python tools/ab_stage.py key=v1,v2 [level ...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import hex_problem
from parelagmc_b200.capi import Context
p = hex_problem(16, 3)
key, vals = sys.argv[1].split("=")
levels = [int(a) for a in sys.argv[2:]] or [0, 1, 2]
S = [1000, 3000, 6000]
for v in vals.split(","):
    c = Context(p["nlevels"], 0)
    c.set_option(key, float(v))
    for l, s in enumerate(p["sampler"]):
        c.upload_sampler_level(l, s, p["alpha"], p["g"], True)
    for l, d in enumerate(p["darcy"]):
        c.upload_darcy_level(l, d)
    c.set_tolerances(1e-6, 1e-12, 300); c.rng_init(0.0, 1.0, 1, 0); c.prepare()
    for lev in levels:
        c.mlmc_level_batch(lev, S[lev], 0)
        c.reset_stats()
        sums, _, its = c.mlmc_level_batch(lev, S[lev], 0)
        st = c.kernel_stats()
        k = st["kernel"]
        shares = {n: round(x["cycle_share"], 3) for n, x in st.items() if n != "kernel" and x["ops"] and x["cycle_share"] > 0.01}
        print(f"{key}={v} level {lev}: {k['ms']:8.2f} ms  {k['algo_bytes']/(k['ms']*1e-3)/1e9:7.0f} GB/s  its {its}  meanQ {sums[4]/S[lev]:.10f}  {shares}", flush=True)
    c.close()
