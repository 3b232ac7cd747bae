"""Host-buffer path of one level, call by call (diagnostic): where the end-to-end time of bench.py's e2e leg goes."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import hex_problem, make_context
from parelagmc_b200.capi import pinned_empty
p = hex_problem(16, 3)
reuse = os.environ.get("PMC_E2E_REUSE", "1") != "0"
ctx = make_context(p, True, 1e-6, 1e-12, 300, options={"cache_results": 1} if reuse else None)
lev, n = int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 1000
Ne, Nec = p["sampler"][lev].Ne, p["sampler"][lev + 1].Ne
b = {"xi": pinned_empty((n, Ne)), "s": pinned_empty((n, Ne)), "sc": pinned_empty((n, Nec)), "emb": pinned_empty((n, Nec))}
def T(f):
    t = time.perf_counter(); r = f(); return r, (time.perf_counter() - t) * 1e3
for rep in range(3):
    ts = {}
    xi, ts["sample"] = T(lambda: ctx.sampler_sample_batch(lev, n, 0, out=b["xi"]))
    R = (lambda v: None) if reuse else (lambda v: v)
    (sc, emb, _), ts["eval coarse"] = T(lambda: ctx.sampler_eval_batch(lev + 1, R(xi), xi_level=lev, use_init=0, out_s=b["sc"], out_embed=b["emb"], nsamples=n))
    _, ts["darcy coarse"] = T(lambda: ctx.darcy_solve_batch(lev + 1, R(sc), nsamples=n))
    (sf, _, _), ts["eval fine"] = T(lambda: ctx.sampler_eval_batch(lev, R(xi), xi_level=lev, init_s=R(emb), init_level=lev + 1, use_init=1, want_embed=False, out_s=b["s"], nsamples=n))
    _, ts["darcy fine"] = T(lambda: ctx.darcy_solve_batch(lev, R(sf), nsamples=n))
    ctx.reset_stats()
    _, ts["fused level batch"] = T(lambda: ctx.mlmc_level_batch(lev, n, 0))
    k = ctx.kernel_stats()["kernel"]
    print(" | ".join(f"{k_}: {v:6.2f}" for k_, v in ts.items()), f"| sum of calls {sum(v for k_, v in ts.items() if k_ != 'fused level batch'):.2f} ms, fused kernel {k['ms']:.2f} ms", flush=True)
ctx.close()
