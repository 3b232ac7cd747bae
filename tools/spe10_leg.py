"""The SPE10 leg of bench.py on its own, for library options given as PMC_OPTS="k=v,k=v" (read by tests/common.make_context).
  PMC_OPTS="darcy.amg_smooth=0.9,darcy.omega=1.25" python tools/spe10_leg.py [--scale 1.0]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
a = ap.parse_args()
args = argparse.Namespace(spe10_scale=a.scale)
r = bench.run_spe10(args, 0, 1, 0, None, torch.device("cuda", 0))
print(os.environ.get("PMC_OPTS", "-"), "| InitRun %.1f ms, %.1f samples/s | " % (r["ms_per_initrun"], r["value"]) +
      " ".join("%s: %.1f ms its %.1f" % (k, v["ms"], v["darcy_its_per_solve"]) for k, v in r["per_level"].items()), flush=True)
