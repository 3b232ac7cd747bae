import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from common import hex_problem, make_context
p = hex_problem(16, 3)
ctx = make_context(p, True, 1e-6, 1e-12, 300)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
S=[1000,3000,6000]
for lev in (2,1,0):
    ctx.mlmc_level_batch(lev,S[lev],0)
for lev in (2,1):
    ts=[]
    for rep in range(30):
        t=time.perf_counter(); ctx.mlmc_level_batch(lev,S[lev],0); ts.append((time.perf_counter()-t)*1e3)
    print("level",lev,"wall ms:", " ".join("%.1f"%x for x in ts))
# interleaved like the bench
for rep in range(6):
    out=[]
    for lev in (2,1,0):
        t=time.perf_counter(); ctx.mlmc_level_batch(lev,S[lev],0); out.append((time.perf_counter()-t)*1e3)
    print("interleaved", ["%.1f"%x for x in out])
