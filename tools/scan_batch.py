"""Bandwidth of the level-0 fused batch as a function of the number of realisations (diagnostic)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import hex_problem, make_context
p = hex_problem(16, 3)
ctx = make_context(p, True, 1e-6, 1e-12, 300)
for lev, sizes in ((0, (500, 1000, 1184, 2000, 2368, 4736)), (1, (3000, 4736, 9472))):
    for S in sizes:
        ctx.mlmc_level_batch(lev, S, 0)
        ctx.reset_stats()
        ctx.mlmc_level_batch(lev, S, 0)
        k = ctx.kernel_stats()["kernel"]
        print(f"level {lev} S={S:5d}: kernel {k['ms']:7.2f} ms  {k['algo_bytes']/(k['ms']*1e-3)/1e9:6.0f} GB/s  {S/(k['ms']*1e-3):9.0f} samples/s")
ctx.close()
