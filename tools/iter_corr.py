"""How well do the coarse-level Darcy iteration counts of a realisation predict the fine-level ones?  Wasted lane-iterations
of 4-wide tiles when realisations are grouped (a) in stream order, (b) sorted by the coarse count, (c) sorted by the fine
count itself (the bound).   python tools/iter_corr.py [level] [nsamples]"""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from common import hex_problem, make_context
lev = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
p = hex_problem(16, 3)
c = make_context(p, True, 1e-6, 1e-12, 300)
xi = c.sampler_sample_batch(lev, n, 0)
kc, emb, sit_c = c.sampler_eval_batch(lev + 1, xi, xi_level=lev, use_init=0)
kf, _, sit_f = c.sampler_eval_batch(lev, xi, xi_level=lev, want_embed=False)
_, _, _, itc = c.darcy_solve_batch(lev + 1, kc)
_, _, _, itf = c.darcy_solve_batch(lev, kf)
def waste(order):
    it = itf[order]
    pad = (-len(it)) % 4
    t = np.concatenate([it, np.full(pad, it[-1])]).reshape(-1, 4)
    return float((t.max(axis=1)[:, None] - t).sum() / (t.max(axis=1).sum() * 4)), int(t.max(axis=1).max())
print("fine its mean %.1f min %d max %d; coarse its mean %.1f; corr %.3f" % (itf.mean(), itf.min(), itf.max(), itc.mean(), np.corrcoef(itf, itc)[0, 1]))
print("sampler its fine %.1f..%d coarse %.1f" % (sit_f.mean(), sit_f.max(), sit_c.mean()))
s = np.log(kf); rng_ = s.max(axis=1) - s.min(axis=1)
print("corr with log-contrast of k: %.3f, with std of log k: %.3f" % (np.corrcoef(itf, rng_)[0, 1], np.corrcoef(itf, s.std(axis=1))[0, 1]))
print("wasted lane-iterations: stream order %.3f, sorted by coarse its %.3f, by (coarse its, contrast) %.3f, by fine its %.3f" % (
    waste(np.arange(n))[0], waste(np.argsort(itc, kind="stable"))[0], waste(np.lexsort((rng_, itc)))[0], waste(np.argsort(itf))[0]))
