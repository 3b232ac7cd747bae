"""SPE10-scale run (BASELINE.json configs[4]): 60 x 220 x 85 hex cells on 1200 x 2200 x 170 ft, 4 levels, correlation
length 100, the SPE10 boundary conditions, unit mass coefficient -- built by the Cartesian hierarchy provider (the
grid is not 8-divisible; remainder cells close each axis).  Runs a few realisations per level through the fused level
loop and checks level 3 against the oracle.
  python tools/spe10_scale.py [--samples 16] [--levels 4] [--scale 1.0]"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from parelagmc_b200 import hierarchy as H
from parelagmc_b200.capi import Context

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=16)
ap.add_argument("--levels", type=int, default=4)
ap.add_argument("--scale", type=float, default=1.0, help="shrink the grid (1.0 = 60x220x85)")
ap.add_argument("--check", action="store_true", help="compare the coarsest level with the CPU oracle")
ap.add_argument("--rel", type=float, default=1e-6)
ap.add_argument("--opt", action="append", default=[], help="key=value for pmc_set_option")
ap.add_argument("--min-level", type=int, default=0)
ap.add_argument("--max-level", type=int, default=-1, help="coarsest level to run (default: all)")
ap.add_argument("--repeat", type=int, default=1)
a = ap.parse_args()
n = [max(8, int(round(x * a.scale))) for x in (60, 220, 85)]
t0 = time.time()
L = H.build_box_hierarchy(n, [1200.0, 2200.0, 170.0], a.levels)
SL = H.build_sampler_levels(L)
DL = H.build_darcy_levels(L, **H.SPE10_BC)
print("grid", n, "levels", [(l.Ne, l.Nf) for l in L], f"host hierarchy {time.time()-t0:.1f} s", flush=True)
corlen = 100.0
p = dict(sampler=SL, darcy=DL, alpha=H.spde_alpha(corlen), g=H.matern_scaling_coefficient(corlen, 3), nlevels=a.levels)
c = Context(a.levels, 0)
for kv in a.opt:
    k, v = kv.split("=")
    c.set_option(k, float(v))
t0 = time.time()
for l, s in enumerate(SL):
    c.upload_sampler_level(l, s, p["alpha"], p["g"], True)
for l, d in enumerate(DL):
    c.upload_darcy_level(l, d)
c.set_tolerances(a.rel, 1e-14, 3000)
c.rng_init(0.0, 1.0, 1, 0)
c.prepare()
print(f"upload + device set-up {time.time()-t0:.1f} s", flush=True)
for lev in range(a.levels - 1 if a.max_level < 0 else a.max_level, a.min_level - 1, -1):
  for rep in range(a.repeat):
    ns = a.samples * (4 ** min(lev, 2))
    c.reset_stats()
    t0 = time.time()
    sums, rows, its = c.mlmc_level_batch(lev, ns, 0, want_rows=True)
    dt = time.time() - t0
    k = c.kernel_stats()["kernel"]
    print(f"level {lev}: N={DL[lev].N:8d} samples {ns:5d}  {dt*1e3:9.1f} ms  {ns/dt:9.1f} samples/s  its/sample {its/ns:7.1f}  "
          f"E[Y]={sums[1]/ns:.6g} E[Q]={sums[4]/ns:.6g}  kernel {k['algo_bytes']/max(k['ms'],1e-9)/1e6:6.0f} GB/s", flush=True)
    if a.check and lev == a.levels - 1:
        from oracle.binding import OracleProblem
        o = OracleProblem(SL, DL, p["alpha"], p["g"], True)
        o.set_tolerances(a.rel, 1e-14, 3000)
        osums, orows, _ = o.mlmc_level(lev, min(ns, 8), 0, nthreads=8)
        print("   oracle rows max |diff|:", np.abs(orows[:, :3] - rows[:min(ns, 8), :3]).max(), flush=True)
c.close()
