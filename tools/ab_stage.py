"""A/B of library options on the bench's level batches (diagnostic).
  python tools/ab_stage.py "k1=v,k2=v" "k1=w" ...     (PMC_LEVELS=0,1,2 selects the levels; '-' = defaults)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import hex_problem
from parelagmc_b200.capi import Context
p = hex_problem(16, 3)
levels = [int(a) for a in os.environ.get("PMC_LEVELS", "0").split(",")]
S = [1000, 3000, 6000]
for cfg in sys.argv[1:]:
    c = Context(p["nlevels"], 0)
    if cfg != "-":
        for kv in cfg.split(","):
            k, v = kv.split("=")
            c.set_option(k, float(v))
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
        shares = {n[:6]: round(x["cycle_share"], 2) for n, x in st.items() if n != "kernel" and x["ops"] and x["cycle_share"] > 0.02}
        print(f"{cfg:40s} L{lev}: {k['ms']:7.2f} ms {k['algo_bytes']/(k['ms']*1e-3)/1e9:6.0f} GB/s its {its} Q {sums[4]/S[lev]:.10f} {shares}", flush=True)
    c.close()
