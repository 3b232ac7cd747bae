"""Effect of the cluster split on small fine-level batches of the bench hierarchy (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import hex_problem, make_context
p = hex_problem(16, 3)
ctx = make_context(p, True, 1e-6, 1e-12, 300)
for S in (16, 64, 128, 256, 500):
    for cs in (1, 0):
        ctx.set_option("cluster_size", cs)
        ctx.mlmc_level_batch(0, S, 0)
        ctx.reset_stats()
        ctx.mlmc_level_batch(0, S, 0)
        k = ctx.kernel_stats()["kernel"]
        print(f"level 0 S={S:4d} cluster={'auto' if cs == 0 else 1}: kernel {k['ms']:7.2f} ms  {k['algo_bytes']/(k['ms']*1e-3)/1e9:6.0f} GB/s  {S/(k['ms']*1e-3):8.0f} samples/s")
ctx.close()
