"""Throw-away numpy/scipy prototype used to pick the preconditioner parameters (Chebyshev degrees,
V-cycle shape) before writing the CUDA kernels.  Not part of the product path."""
import sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, '.')
from parelagmc_b200.hierarchy import *

def cheb_apply(Aop, invD, r, a, b, deg, x0=None):
    """deg steps of Chebyshev-accelerated Jacobi for A z = r with spectrum(D^-1 A) in [a,b]."""
    theta=(b+a)/2; delta=(b-a)/2; sigma=theta/delta; rho=1/sigma
    if x0 is None:
        z = invD*r/theta; d = z.copy(); start=1
    else:
        z = x0.copy(); res = r - Aop(z); d = invD*res/theta; z = z + d; start=1
    for j in range(start,deg):
        rho_new = 1/(2*sigma-rho)
        res = r - Aop(z)
        d = rho_new*rho*d + (2*rho_new/delta)*(invD*res)
        z = z + d
        rho = rho_new
    return z

class MG:
    def __init__(self, S, Ps, deg=2, ratio=4.0, coarse_deg=8, coarse_ratio=30.0, l1=True):
        self.S=[S.tocsr()]; self.P=[]
        for P in Ps:
            self.P.append(P.tocsr()); self.S.append((P.T@self.S[-1]@P).tocsr())
        self.l1=l1
        self.invD=[]
        for s in self.S:
            dg = np.asarray(abs(s).sum(axis=1)).ravel() if l1 else s.diagonal()
            self.invD.append(1.0/dg)
        self.lmax = 1.0 if l1 else 2.0
        self.deg=deg; self.ratio=ratio; self.coarse_deg=coarse_deg; self.coarse_ratio=coarse_ratio
    def smooth(self, m, r, x0=None, deg=None, ratio=None):
        deg = deg or self.deg; ratio = ratio or self.ratio
        return cheb_apply(lambda v: self.S[m]@v, self.invD[m], r, self.lmax/ratio, self.lmax*1.0, deg, x0)
    def vcycle(self, m, r):
        if m == len(self.S)-1:
            return self.smooth(m, r, None, self.coarse_deg, self.coarse_ratio)
        x = self.smooth(m, r)
        res = r - self.S[m]@x
        xc = self.vcycle(m+1, self.P[m].T@res)
        x = x + self.P[m]@xc
        return self.smooth(m, r, x)

def minres(Aop, Pop, b, x0=None, rel=1e-12, abs_=1e-300, maxit=2000):
    """MFEM-style preconditioned MINRES."""
    n=len(b)
    x = np.zeros(n) if x0 is None else x0.copy()
    v1 = b - Aop(x) if x0 is not None else b.copy()
    u1 = Pop(v1)
    eta = beta = np.sqrt(u1@v1)
    gamma0=gamma1=1.0; sigma0=sigma1=0.0
    goal = max(rel*eta, abs_)
    v0=np.zeros(n); w0=np.zeros(n); w1=np.zeros(n)
    if eta<=goal: return x,0
    for it in range(1,maxit+1):
        v1/=beta; u1/=beta
        q=Aop(u1); alpha=u1@q
        if it>1: q-=beta*v0
        v0=q-alpha*v1
        delta=gamma1*alpha-gamma0*sigma1*beta
        rho3=sigma0*beta
        rho2=sigma1*alpha+gamma0*gamma1*beta
        q=Pop(v0); beta=np.sqrt(max(v0@q,0))
        rho1=np.hypot(delta,beta)
        if it==1: w0=u1/rho1
        elif it==2: w0=u1/rho1-(rho2/rho1)*w1
        else: w0=(-rho3/rho1)*w0-(rho2/rho1)*w1+u1/rho1
        gamma0=gamma1; gamma1=delta/rho1
        x+=gamma1*eta*w0
        sigma0=sigma1; sigma1=beta/rho1
        eta=-sigma1*eta
        if abs(eta)<=goal: return x,it
        u1,q=q,u1
        v0,v1=v1,v0
        w0,w1=w1,w0
    return x,maxit

if __name__=='__main__':
    n=int(sys.argv[1]) if len(sys.argv)>1 else 16
    nl=int(sys.argv[2]) if len(sys.argv)>2 else 3
    L=build_box_hierarchy([n]*3,[2,2,2],nl)
    SL=build_sampler_levels(L); DL=build_darcy_levels(L,**MLMC_DEFAULT_BC)
    corlen=0.1; alpha=spde_alpha(corlen); g=matern_scaling_coefficient(corlen,3)
    rng=np.random.default_rng(0)
    for lev in range(nl):
        s=SL[lev]
        A=sp.bmat([[s.M,s.B.T],[s.B,-alpha*sp.diags(s.Wdiag)]],format='csr')
        Md=s.M.diagonal()
        Sm=(alpha*sp.diags(s.Wdiag)+s.B@sp.diags(1/Md)@s.B.T).tocsr()
        Ps=[SL[m].P for m in range(lev,nl-1)]
        for (mdeg, sdeg, sratio) in [(1,2,4.0),(2,2,4.0),(3,2,4.0),(3,3,6.0)]:
            mg=MG(Sm,Ps,deg=sdeg,ratio=sratio)
            def Pop(r):
                zu=cheb_apply(lambda v:s.M@v,1/Md,r[:s.Nf],0.5,1.5,mdeg)
                return np.concatenate([zu,mg.vcycle(0,r[s.Nf:])])
            xi=rng.standard_normal(s.Ne)
            b=np.concatenate([np.zeros(s.Nf),-g*xi*s.w_sqrt])
            x,it=minres(lambda v:A@v,Pop,b)
            xd=spla.spsolve(A.tocsc(),b)
            print(f'sampler lev{lev} mdeg{mdeg} sdeg{sdeg}: it={it} relerr={np.linalg.norm(x-xd)/np.linalg.norm(xd):.2e} s.std={x[s.Nf:].std():.3f}')
        # Darcy with k = exp(s)
        sfield=xd[s.Nf:]; k=np.exp(sfield)
        d=DL[lev]
        M=L[lev].assemble_M(k)
        keep=sp.diags((d.ess_u==0).astype(float))
        Me=(keep@M@keep+sp.diags((d.ess_u!=0).astype(float))).tocsr(); Be=(d.B@keep).tocsr()
        A=sp.bmat([[Me,Be.T],[Be,None]],format='csr')
        rhs=d.rhs.copy(); rhs[:d.Nf][d.ess_u!=0]=0
        Md=Me.diagonal()
        Sm=(Be@sp.diags(1/Md)@Be.T).tocsr()
        xd=spla.spsolve(A.tocsc(),rhs)
        # eigen-bounds of D^-1 M
        for (mdeg, sdeg, sratio, cdeg) in [(1,2,4.0,8),(2,2,4.0,8),(3,2,4.0,8),(3,3,6.0,12),(2,1,4.0,8),(2,3,6.0,12)]:
            mg=MG(Sm,[DL[m].P_p for m in range(lev,nl-1)],deg=sdeg,ratio=sratio,coarse_deg=cdeg)
            def Pop(r):
                zu=cheb_apply(lambda v:Me@v,1/Md,r[:d.Nf],0.5,1.5,mdeg)
                return np.concatenate([zu,mg.vcycle(0,r[d.Nf:])])
            x,it=minres(lambda v:A@v,Pop,rhs)
            print(f'darcy   lev{lev} mdeg{mdeg} sdeg{sdeg} cdeg{cdeg}: it={it} relerr={np.linalg.norm(x-xd)/np.linalg.norm(xd):.2e} Q={d.obs@x:.10f} Qd={d.obs@xd:.10f}')
