"""Small fixed case for ncu: one fused MLMC level batch on the bench hierarchy with a capped iteration count,
so that a handful of launches of every hot kernel (at bench sizes) is captured quickly.
  python tools/profile_case.py [--level 0] [--samples 1024] [--maxit 6]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import hex_problem, make_context

ap = argparse.ArgumentParser()
ap.add_argument("--level", type=int, default=0)
ap.add_argument("--samples", type=int, default=1024)
ap.add_argument("--maxit", type=int, default=6)
ap.add_argument("--n", type=int, default=16)
ap.add_argument("--mc", action="store_true", help="single-level loop (sampler + Darcy at --level only)")
a = ap.parse_args()
p = hex_problem(a.n, 3)
ctx = make_context(p, True, 1e-6, 1e-12, a.maxit)
sums, rows, its = (ctx.mc_level_batch if a.mc else ctx.mlmc_level_batch)(a.level, a.samples, 0)
st = ctx.kernel_stats()
print("kernel", st["kernel"], "iters", its)
print({k: (round(v["cycle_share"], 3), round(v["algo_bytes"] / 1e9, 2)) for k, v in st.items() if k != "kernel"})
ctx.close()
