import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import hex_problem, make_context
p = hex_problem(16, 3)
ctx = make_context(p, True, 1e-6, 1e-12, 300)
for lev, S in ((0, 4736), (0, 2368), (1, 9472)):
    for nt in (128, 256, 512):
        ctx.set_option("cta_threads", nt)
        ctx.mlmc_level_batch(lev, S, 0)
        ctx.reset_stats()
        ctx.mlmc_level_batch(lev, S, 0)
        k = ctx.kernel_stats()["kernel"]
        print(f"level {lev} S={S:5d} NT={nt}: kernel {k['ms']:7.2f} ms  {k['algo_bytes']/(k['ms']*1e-3)/1e9:6.0f} GB/s")
ctx.close()
