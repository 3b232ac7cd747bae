import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from common import hex_problem, make_context
p = hex_problem(16, 3)
ctx = make_context(p, True, 1e-6, 1e-12, 300)
lev = int(sys.argv[1]); ns = int(sys.argv[2])
ctx.mlmc_level_batch(lev, ns, 0)
ctx.reset_stats()
ctx.mlmc_level_batch(lev, ns, 0)
print(ctx.kernel_stats()["kernel"])
