"""numpy prototype: Darcy preconditioner variants vs MINRES iteration count at the bench tolerance (1e-6)."""
import sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, '.')
from parelagmc_b200.hierarchy import *
from tools.prototype_solver import cheb_apply, minres

class MG:
    def __init__(self, S, Ps, deg=2, ratio=4.0, coarse_deg=8, coarse_ratio=30.0, omega=1.0, gamma=1):
        self.S=[S.tocsr()]; self.P=[]
        for P in Ps:
            self.P.append(P.tocsr()); self.S.append((P.T@self.S[-1]@P).tocsr())
        self.invD=[1.0/np.asarray(abs(s).sum(axis=1)).ravel() for s in self.S]
        self.deg=deg; self.ratio=ratio; self.coarse_deg=coarse_deg; self.coarse_ratio=coarse_ratio
        self.omega=omega; self.gamma=gamma
    def smooth(self, m, r, x0=None, deg=None, ratio=None):
        deg = deg or self.deg; ratio = ratio or self.ratio
        return cheb_apply(lambda v: self.S[m]@v, self.invD[m], r, 1.0/ratio, 1.0, deg, x0)
    def vcycle(self, m, r):
        if m == len(self.S)-1:
            return self.smooth(m, r, None, self.coarse_deg, self.coarse_ratio)
        x = self.smooth(m, r)
        for g in range(self.gamma if m>0 else 1):
            res = r - self.S[m]@x
            xc = self.vcycle(m+1, self.P[m].T@res)
            x = x + self.omega*(self.P[m]@xc)
        return self.smooth(m, r, x)

n=16; nl=3
L=build_box_hierarchy([n]*3,[2,2,2],nl)
SL=build_sampler_levels(L); DL=build_darcy_levels(L,**MLMC_DEFAULT_BC)
alpha=spde_alpha(0.1); g=matern_scaling_coefficient(0.1,3)
rng=np.random.default_rng(0)
lev=0
s=SL[lev]
A=sp.bmat([[s.M,s.B.T],[s.B,-alpha*sp.diags(s.Wdiag)]],format='csc')
ks=[]
for j in range(3):
    xi=rng.standard_normal(s.Ne); b=np.concatenate([np.zeros(s.Nf),-g*xi*s.w_sqrt])
    ks.append(np.exp(spla.spsolve(A,b)[s.Nf:]))
d=DL[lev]
keep=sp.diags((d.ess_u==0).astype(float))
def run(tag, mdeg, mgkw, rel=1e-6, minterval=(0.5,1.5)):
    its=[]
    for k in ks:
        M=L[lev].assemble_M(k)
        Me=(keep@M@keep+sp.diags((d.ess_u!=0).astype(float))).tocsr(); Be=(d.B@keep).tocsr()
        Aop=sp.bmat([[Me,Be.T],[Be,None]],format='csr')
        rhs=d.rhs.copy(); rhs[:d.Nf][d.ess_u!=0]=0
        Md=Me.diagonal()
        Sm=(Be@sp.diags(1/Md)@Be.T).tocsr()
        mg=MG(Sm,[DL[m].P_p for m in range(lev,nl-1)],**mgkw)
        def Pop(r):
            zu=cheb_apply(lambda v:Me@v,1/Md,r[:d.Nf],minterval[0],minterval[1],mdeg)
            return np.concatenate([zu,mg.vcycle(0,r[d.Nf:])])
        x,it=minres(lambda v:Aop@v,Pop,rhs,rel=rel,abs_=1e-12)
        its.append(it)
    print(f"{tag:50s} its={its}")
run("base m2 s2/4 c8/30", 2, dict(deg=2,ratio=4.0,coarse_deg=8,coarse_ratio=30.0))
run("m1", 1, dict(deg=2,ratio=4.0))
run("m3", 3, dict(deg=2,ratio=4.0))
run("s1", 2, dict(deg=1,ratio=4.0))
run("s3/6", 2, dict(deg=3,ratio=6.0))
run("s2/8", 2, dict(deg=2,ratio=8.0))
run("s3/10", 2, dict(deg=3,ratio=10.0))
run("omega1.5", 2, dict(deg=2,ratio=4.0,omega=1.5))
run("omega2", 2, dict(deg=2,ratio=4.0,omega=2.0))
run("gamma2 (W)", 2, dict(deg=2,ratio=4.0,gamma=2))
run("c16/100", 2, dict(deg=2,ratio=4.0,coarse_deg=16,coarse_ratio=100.))
run("s3/6 omega1.5 c16/100", 2, dict(deg=3,ratio=6.0,omega=1.5,coarse_deg=16,coarse_ratio=100.))
run("s4/10 omega1.5 c16/100", 2, dict(deg=4,ratio=10.0,omega=1.5,coarse_deg=16,coarse_ratio=100.))
print("---- omega sweep")
for om in (1.8, 2.0, 2.2, 2.5, 3.0):
    run(f"omega{om}", 2, dict(deg=2,ratio=4.0,omega=om))
run("m1 omega2", 1, dict(deg=2,ratio=4.0,omega=2.0))
run("m1 s1 omega2", 1, dict(deg=1,ratio=4.0,omega=2.0))
run("m2 s1 omega2", 2, dict(deg=1,ratio=4.0,omega=2.0))
run("m2 s3/6 omega2", 2, dict(deg=3,ratio=6.0,omega=2.0))
run("m3 s3/6 omega2", 3, dict(deg=3,ratio=6.0,omega=2.0))
run("m2 s2/4 omega2 rel1e-12", 2, dict(deg=2,ratio=4.0,omega=2.0), rel=1e-12)
run("base rel1e-12", 2, dict(deg=2,ratio=4.0), rel=1e-12)
