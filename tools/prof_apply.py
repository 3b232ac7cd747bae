"""One launch of the persistent kernel that contains only the Darcy block-operator apply (pmc_darcy_apply_batch): lets ncu
see the hardware counters of that operation alone.   python tools/prof_apply.py [nsamples] [reps]"""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from common import hex_problem, make_context
p = hex_problem(16, 3)
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
opts = {}
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    opts[k] = float(v)
ctx = make_context(p, True, 1e-6, 1e-12, 300, options=opts)
d = p["darcy"][0]
rng = np.random.default_rng(0)
k = np.exp(rng.standard_normal((ns, d.Ne)))
x = rng.standard_normal((ns, d.N))
for _ in range(reps):
    ctx.reset_stats()
    y = ctx.darcy_apply_batch(0, k, x)
    st = ctx.kernel_stats()["kernel"]
    print(st["ms"], "ms", st["algo_bytes"] / max(st["ms"], 1e-9) / 1e6, "GB/s (credited)")
