"""BASELINE.json configs[4]: SPE10-scale 3D Darcy MLMC (60 x 220 x 85 hex cells on 1200 x 2200 x 170 ft, 4 levels,
correlation length 100, the SPE10 boundary conditions, unit mass coefficient -- /root/reference/examples/SPE10/
SPE10_MLMC.cpp with spe10_3D_parameters.xml) through MLMC_Manager with the realisations of every level sharded over
the ranks (one rank per GPU) and one all-reduce of the per-level sums.

  python tools/spe10_mlmc.py --samples 8,32,128,512
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/spe10_mlmc.py --samples 64,256,1024,4096
"""
import argparse, io, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parelagmc_b200 import hierarchy as H
from parelagmc_b200 import managers as MG
from parelagmc_b200.capi import Context

ap = argparse.ArgumentParser()
ap.add_argument("--samples", default="8,32,128,512", help="realisations per level (level 0 = finest first), whole job")
ap.add_argument("--scale", type=float, default=1.0, help="shrink the grid (1.0 = 60x220x85)")
ap.add_argument("--rel", type=float, default=1e-6)
ap.add_argument("--concurrent-levels", action="store_true")
a = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
import torch
torch.cuda.set_device(local)
comm = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = MG._Comm(use_dist=True, device=torch.device("cuda", local))
samples = [int(x) for x in a.samples.split(",")]
nl = len(samples)
n = [max(8, int(round(x * a.scale))) for x in (60, 220, 85)]
t0 = time.time()
L = H.build_box_hierarchy(n, [1200.0, 2200.0, 170.0], nl)
SL = H.build_sampler_levels(L)
DL = H.build_darcy_levels(L, **H.SPE10_BC)
corlen = 100.0
c = Context(nl, local)
for l, s in enumerate(SL):
    c.upload_sampler_level(l, s, H.spde_alpha(corlen), H.matern_scaling_coefficient(corlen, 3), True)
for l, d in enumerate(DL):
    c.upload_darcy_level(l, d)
c.set_tolerances(a.rel, 1e-14, 3000)
c.rng_init(0.0, 1.0, 1, 0)
c.prepare()
t_setup = time.time() - t0
params = {"Use array samples": True, "Array number of samples": samples, "Mean square error": 1e30,
          "Output filename for MC managers": ""}
buf = io.StringIO()
mg = MG.MLMC_Manager(comm, nl, c, params, out=buf if rank == 0 else None)
mg.concurrent_levels = a.concurrent_levels
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.time()
mg.Run()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.time() - t0
if rank == 0:
    print(buf.getvalue())
    per_level = {f"level{l}": {"N": int(DL[l].N), "samples": samples[l], "s": round(float(mg.level_time[l]), 3),
                               "samples_per_s": round(samples[l] / max(float(mg.level_time[l]), 1e-9), 2)} for l in range(nl)}
    print(json.dumps({"config": "SPE10-scale MLMC", "grid": n, "n_gpus": world, "samples": samples, "wall_s": round(dt, 3),
                      "setup_s": round(t_setup, 1), "samples_per_s": round(sum(samples) / dt, 2), "estimate": float(np.sum(mg.eY)),
                      "minres_iterations": int(mg.total_iters), "per_level": per_level}))
c.close()
if world > 1:
    dist.destroy_process_group()
