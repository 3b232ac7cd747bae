"""Times the level batches of the bench for option variants given on the command line (diagnostic).
   python tools/scan_opts.py "darcy.mass_scale=0.7" "darcy.mass_scale=1.4,darcy.omega=2.0" ...   ("" = defaults)
Every variant runs in the same process on the same box, so the times compare."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import hex_problem
from parelagmc_b200.capi import Context
p = hex_problem(16, 3)
S = [1000, 3000, 6000]
levels = [int(x) for x in os.environ.get("SCAN_LEVELS", "0").split(",")]
for spec in sys.argv[1:] or [""]:
    opts = dict((kv.split("=")[0], float(kv.split("=")[1])) for kv in spec.split(",") if kv)
    c = Context(3, 0)
    for k, v in opts.items():
        c.set_option(k, v)
    for l, s in enumerate(p["sampler"]):
        c.upload_sampler_level(l, s, p["alpha"], p["g"], True)
    for l, d in enumerate(p["darcy"]):
        c.upload_darcy_level(l, d)
    c.set_tolerances(1e-6, 1e-12, 300); c.rng_init(0.0, 1.0, 1, 0); c.prepare()
    out = []
    for lev in levels:
        c.mlmc_level_batch(lev, S[lev], 0)
        c.reset_stats()
        sums, _, its = c.mlmc_level_batch(lev, S[lev], 0)
        k = c.kernel_stats()["kernel"]
        out.append(f"L{lev}: {k['ms']:6.2f} ms its/sample {its/S[lev]:6.1f} E[Y]={sums[1]/S[lev]:.6f}")
    print(f"{spec or 'defaults':50s}", " | ".join(out), flush=True)
    c.close()
