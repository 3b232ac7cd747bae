"""Generates tests/golden/oracle_golden.json from the CPU oracle (oracle/pmc_oracle.c) on tiny seeded cases.

The reference itself cannot be built or imported here (C++ over MPI/MFEM/hypre/ParELAG/TRNG, none present), so
these vectors pin the ORACLE (and, through the GPU tests, the CUDA path) against regressions; the reference's own
known answers that need no third party (Q = 2, dof counts, Matern constants) are asserted separately in
tests/test_oracle.py.   Run:  python tools/make_golden.py
"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import hex_problem, quad_problem, make_oracle
from oracle.binding import Yarn5
from parelagmc_b200 import hierarchy as H

g = {}
g["yarn5_ints"] = {str(pos): Yarn5().jump(pos).ints(n).tolist() for pos, n in [(0, 16), (10**6, 8), (2**40, 8)]}
g["yarn5_split_4_3"] = Yarn5().split(4, 3).ints(8).tolist()
g["normals_pos0"] = [x.hex() for x in Yarn5().normals(16)]
g["matern"] = {"g(0.1,3)": H.matern_scaling_coefficient(0.1, 3), "g(0.1,2)": H.matern_scaling_coefficient(0.1, 2),
               "g(100,3)": H.matern_scaling_coefficient(100.0, 3), "alpha(0.1)": H.spde_alpha(0.1)}
p = hex_problem(4, 2)
o = make_oracle(p)
g["hex4"] = {"dofs": [d.N for d in p["darcy"]]}
g["hex4"]["Q_k1"] = [o.darcy_solve(l, np.ones(p["darcy"][l].Ne))[0] for l in range(2)]
ks = [np.exp(np.sin(np.arange(p["darcy"][l].Ne, dtype=np.float64))) for l in range(2)]
g["hex4"]["Q_ksin"] = [o.darcy_solve(l, ks[l])[0] for l in range(2)]
xi = Yarn5().normals(p["sampler"][0].Ne)
s0, e0, _ = o.sampler_eval(0, xi)
s1, e1, _ = o.sampler_eval(1, xi, xi_level=0, use_init=0)
g["hex4"]["field_l0"] = e0.tolist()
g["hex4"]["field_l1_from_l0_noise"] = e1.tolist()
for lev, ns in [(1, 4), (0, 3)]:
    sums, rows, _ = o.mlmc_level(lev, ns, 1234)
    g["hex4"][f"mlmc_rows_l{lev}"] = rows.tolist()
    g["hex4"][f"mlmc_sums_l{lev}"] = sums.tolist()
q = quad_problem(4, 2)
oq = make_oracle(q, lognormal=False)
xi = Yarn5().jump(99).normals(q["sampler"][0].Ne)
g["quad4"] = {"field_l0": oq.sampler_eval(0, xi)[0].tolist(),
              "field_l1": oq.sampler_eval(1, xi, xi_level=0, use_init=0)[0].tolist()}
os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
with open(os.path.join(ROOT, "tests", "golden", "oracle_golden.json"), "w") as f:
    json.dump(g, f, indent=1)
print("wrote tests/golden/oracle_golden.json")
