import os, sys, time
import numpy as np
ROOT="/root/repo"; sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from common import hex_problem, make_context
p = hex_problem(16, 3)
ctx = make_context(p, True, 1e-6, 1e-12, 300)
S=[1000,3000,6000]
def run(tag):
    for rep in range(3):
        out=[]
        for lev in (2,1,0):
            t=time.perf_counter(); ctx.mlmc_level_batch(lev,S[lev],0); out.append((time.perf_counter()-t)*1e3)
        print(tag, ["%.1f"%x for x in out])
run("own stream")
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
run("torch stream0")
s2=torch.cuda.Stream()
ctx.set_stream(s2.cuda_stream)
run("torch side stream")
