"""Few realisations of a large level: one CTA per tile / cluster split / grid groups (diagnostic).
  python tools/scan_group.py [n] [samples ...]      hex n^3, 3 levels, fused level-0 batch at the bench tolerances"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import hex_problem, make_context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sizes = [int(a) for a in sys.argv[2:]] or [4, 8, 16]
p = hex_problem(n, 3)
print("hex", n, "N per level", [d.N for d in p["darcy"]], flush=True)
for name, opts in [("one CTA per tile", {"cluster_size": 1, "group_size": -1}), ("clusters", {"group_size": -1}),
                   ("grid groups", {}), ("grid groups, solo 0", {"solo_rows": 0}), ("grid groups, solo 16384", {"solo_rows": 16384})]:
    ctx = make_context(p, True, 1e-6, 1e-12, 300, options=opts)
    for ns in sizes:
        ctx.mlmc_level_batch(0, ns, 0)
        ctx.reset_stats()
        sums, _, its = ctx.mlmc_level_batch(0, ns, 0)
        k = ctx.kernel_stats()["kernel"]
        print(f"{name:24s} samples {ns:3d}: {k['ms']:9.2f} ms  {ns/k['ms']*1e3:8.1f} samples/s  {k['algo_bytes']/(k['ms']*1e-3)/1e9:6.0f} GB/s  its {its}  E[Q] {sums[4]/ns:.9f}", flush=True)
    ctx.close()
