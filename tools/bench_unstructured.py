"""Throughput of the level batches on unstructured agglomerates (wide operator rows) next to the structured hierarchy of
the same fine mesh.   python tools/bench_unstructured.py [n] [samples_l1] [samples_l2]"""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from common import hex_problem, agglomerated_problem, make_context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = {0: 200, 1: int(sys.argv[2]) if len(sys.argv) > 2 else 3000, 2: int(sys.argv[3]) if len(sys.argv) > 3 else 6000}
for name, p in (("structured", hex_problem(n, 3, 0.1)), ("agglomerated", agglomerated_problem(n=n, nlevels=3, corlen=0.1))):
    c = make_context(p, True, 1e-6, 1e-12, 300)
    for lev in (2, 1, 0):
        d = p["darcy"][lev]
        ne = np.diff(d.elem_ptr)
        c.mlmc_level_batch(lev, S[lev], 0)
        c.reset_stats()
        sums, _, its = c.mlmc_level_batch(lev, S[lev], 0)
        k = c.kernel_stats()["kernel"]
        print(f"{name:13s} level {lev}: N={d.N:6d} faces/element {ne.min()}-{ne.max()} samples {S[lev]:5d} {k['ms']:8.2f} ms "
              f"{S[lev]/k['ms']*1e3:10.0f} samples/s {k['algo_bytes']/max(k['ms'],1e-9)/1e6:6.0f} GB/s its/sample {its/S[lev]:6.1f} "
              f"E[Q]={sums[4]/S[lev]:.5f}", flush=True)
    c.close()
