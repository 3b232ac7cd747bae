"""Level-0 batch of the bench hierarchy for several batch sizes / CTA sizes (diagnostic: how the achieved bandwidth
depends on how full the machine is).  python tools/scan_fill.py "nt:samples" ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import hex_problem, make_context
p = hex_problem(16, 3)
ctx = make_context(p, True, 1e-6, 1e-12, 300)
for a in sys.argv[1:]:
    nt, ns = (int(x) for x in a.split(":"))
    ctx.set_option("cta_threads", nt)
    ctx.mlmc_level_batch(0, ns, 0)
    ctx.reset_stats()
    ctx.mlmc_level_batch(0, ns, 0)
    k = ctx.kernel_stats()["kernel"]
    print(f"nt {nt:4d} samples {ns:5d} tiles {(ns+3)//4:4d}: {k['ms']:8.2f} ms  {ns/k['ms']*1e3:9.0f} samples/s  {k['algo_bytes']/(k['ms']*1e-3)/1e9:6.0f} GB/s", flush=True)
ctx.close()
