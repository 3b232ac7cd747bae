"""Times one fused level batch of the bench workload for every CTA size of the persistent kernel (diagnostic)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import hex_problem, make_context
p = hex_problem(16, 3)
ctx = make_context(p, True, 1e-6, 1e-12, 300)
S = [1000, 3000, 6000]
for lev in (2, 1, 0):
    for nt in (0, 64, 128, 256, 512):
        ctx.set_option("cta_threads", nt)
        ctx.mlmc_level_batch(lev, S[lev], 0)
        ctx.reset_stats()
        t = time.perf_counter()
        for _ in range(2):
            ctx.mlmc_level_batch(lev, S[lev], 0)
        dt = (time.perf_counter() - t) / 2
        k = ctx.kernel_stats()["kernel"]
        print(f"level {lev} cta {nt:4d}: wall {dt*1e3:7.2f} ms  kernel {k['ms']/2:7.2f} ms  {k['algo_bytes']/2/ (k['ms']/2*1e-3)/1e9:7.0f} GB/s")
ctx.close()
