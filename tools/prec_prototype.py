"""CPU prototype of the Darcy solve's preconditioner on the bench hierarchy (diagnostic; numpy / scipy, no GPU).

Reproduces the iteration counts of the device path (block-diagonal MINRES preconditioner: Jacobi on the RT mass block,
V-cycle with l1-Jacobi Chebyshev smoothing on the lumped Schur complement B diag(M(k))^-1 B^T over the hierarchy's
piecewise-constant prolongators, over-correction 2.5) and varies one ingredient at a time:

  python tools/prec_prototype.py            # realisations 0..2 of the bench's level-0 stream

Findings recorded in DESIGN.md section 5 (iterations at rel 1e-6 on the 16^3 level):
  * device default (Chebyshev degree 2 / 3 / coarsest 8)            33
  * exact solve of the lumped Schur complement instead of the V-cycle 28-29   (ceiling of any better V-cycle)
  * red-black symmetric Gauss-Seidel on the finest V-level (2 operator applications per cycle instead of 4)   37-39
  * mass block: Chebyshev degree 2 / 3 / exact M^-1 (Schur part unchanged)    29-30 / 28-29 / 25
  * fine solve started from the prolongated coarse solution of the same realisation: initial preconditioned residual
    0.11-0.21 of the zero start's, 24-27 instead of 27-29 iterations (exact Schur block)
  * python tools/prec_prototype.py --consistent : the two blocks must approximate the SAME inverse of M.  With the Jacobi
    mass block even the exact Schur complement B M^-1 B^T gives 27 iterations (lumped: 28); with a degree-2 Chebyshev mass
    block the lumped Schur complement gives 26-27, but the consistent one, B p2(D^-1 M) D^-1 B^T (first-order Neumann
    correction, a 13-point pattern), gives 16-17 (exact Schur complement: 16) -- each with an exact solve of the Schur block.
    With the device's V-cycle on the consistent Schur complement instead of the exact solve: 27-29 (over-correction 2.5-3).
"""
import os, sys
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spl
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import hex_problem, make_oracle
from oracle.binding import Yarn5

p = hex_problem(16, 3)
orc = make_oracle(p, True, 1e-12, 1e-30, 2000)
d0, d1 = p["darcy"][0], p["darcy"][1]
P01, P12 = sp.csr_matrix(d0.P_p), sp.csr_matrix(d1.P_p)
Pu = sp.csr_matrix(d0.P_u)


def darcy_mats(d, k):
    """[M(k) B^T; B 0] with the essential RT dofs eliminated (identity rows), and the right-hand side."""
    Nf, Ne = d.Nf, d.Ne
    rows, cols, vals = [], [], []
    ptr, dofs, mp, em = (np.asarray(a) for a in (d.elem_ptr, d.elem_dofs, d.elem_mat_ptr, d.elem_mat))
    for e in range(Ne):
        n = ptr[e + 1] - ptr[e]
        dd = dofs[ptr[e]:ptr[e + 1]]
        Me = em[mp[e]:mp[e] + n * n].reshape(n, n) * k[e]
        rows += list(np.repeat(dd, n)); cols += list(np.tile(dd, n)); vals += list(Me.ravel())
    M = sp.csr_matrix((vals, (rows, cols)), shape=(Nf, Nf))
    B = sp.csr_matrix(d.B)
    ess = np.asarray(d.ess_u).astype(bool)
    rhs = np.asarray(d.rhs).copy(); ed = np.asarray(d.ess_data)[:Nf]
    xe = np.where(ess, ed, 0.0)
    bu = rhs[:Nf] - M @ xe; bp = rhs[Nf:] - B @ xe
    Dm = sp.diags((~ess).astype(float)); Ie = sp.diags(ess.astype(float))
    Mf = Dm @ M @ Dm + Ie; Bf = B @ Dm
    bu = np.where(ess, ed, bu)
    return sp.bmat([[Mf, Bf.T], [Bf, None]]).tocsr(), Mf, Bf, np.concatenate([bu, bp])


def minres(A, b, P, x0, goal, maxit=500):
    """Preconditioned MINRES; stops when the preconditioned residual norm estimate |eta| <= goal.  Returns (x, iterations)."""
    x = x0.copy(); r = b - A @ x; z = P(r); beta = np.sqrt(r @ z); eta = beta
    vk, zk, betak, vkm, beta_prev = r.copy(), z.copy(), beta, np.zeros_like(b), 0.0
    g0 = g1 = 1.0; s0 = s1 = 0.0
    w0 = np.zeros_like(b); w1 = np.zeros_like(b); it = 0
    while abs(eta) > goal and it < maxit:
        zn_ = zk / betak
        q = A @ zn_
        alpha = q @ zn_
        vn = q - (alpha / betak) * vk - ((betak / beta_prev) * vkm if beta_prev != 0 else 0.0)
        zn = P(vn); beta_new = np.sqrt(max(vn @ zn, 0))
        delta = g1 * alpha - g0 * s1 * betak; rho1 = np.hypot(delta, beta_new); rho2 = s1 * alpha + g0 * g1 * betak; rho3 = s0 * betak
        wn = (zn_ - rho3 * w0 - rho2 * w1) / rho1
        g0 = g1; g1 = delta / rho1; s0 = s1; s1 = beta_new / rho1
        x = x + g1 * eta * wn; eta = -s1 * eta
        w0, w1 = w1, wn
        vkm, vk, zk, beta_prev, betak = vk, vn, zn, betak, beta_new
        it += 1
    return x, it


def cheb(S, dinv, r, z, deg, lo, hi):
    theta = 0.5 * (hi + lo); delta = 0.5 * (hi - lo); sigma = theta / delta; rho = 1 / sigma
    d = np.zeros_like(r)
    for j in range(deg):
        if j == 0: ca, cb = 0.0, 1 / theta
        else:
            rn = 1 / (2 * sigma - rho); ca = rn * rho; cb = 2 * rn / delta; rho = rn
        d = ca * d + cb * dinv * (r - S @ z); z = z + d
    return z


def make_prec(Mf, Bf, schur="cheb", mass="jacobi", omega=2.5, deg_mid=3):
    dM = Mf.diagonal(); Nf = Mf.shape[0]
    S0 = (Bf @ sp.diags(1 / dM) @ Bf.T).tocsr(); S1 = (P01.T @ S0 @ P01).tocsr(); S2 = (P12.T @ S1 @ P12).tocsr()
    l1 = [1 / np.asarray(abs(S).sum(axis=1)).ravel() for S in (S0, S1, S2)]
    n0 = S0.shape[0]
    col = -np.ones(n0, dtype=int)       # two-colouring of the 7-point pattern
    for s in range(n0):
        if col[s] >= 0: continue
        col[s] = 0; st = [s]
        while st:
            i = st.pop()
            for j in S0.indices[S0.indptr[i]:S0.indptr[i + 1]]:
                if j != i and col[j] < 0: col[j] = 1 - col[i]; st.append(j)
    c1, c2 = np.where(col == 0)[0], np.where(col == 1)[0]
    dg = 1 / S0.diagonal(); S0c1, S0c2 = S0[c1], S0[c2]
    luS = spl.splu(S0.tocsc()) if schur == "exact" else None
    luM = spl.splu(Mf.tocsc()) if mass == "exact" else None

    def coarse(r1):
        z = cheb(S1, l1[1], r1, np.zeros_like(r1), deg_mid, 0.25, 1.0)
        z2 = cheb(S2, l1[2], P12.T @ (r1 - S1 @ z), np.zeros(S2.shape[0]), 8, 1 / 30, 1.0)
        return cheb(S1, l1[1], r1, z + omega * (P12 @ z2), deg_mid, 0.25, 1.0)

    def vc(r):
        if schur == "exact": return luS.solve(r)
        if schur == "cheb":
            z = cheb(S0, l1[0], r, np.zeros_like(r), 2, 0.25, 1.0)
            z = z + omega * (P01 @ coarse(P01.T @ (r - S0 @ z)))
            return cheb(S0, l1[0], r, z, 2, 0.25, 1.0)
        z = np.zeros_like(r)                         # red-black symmetric Gauss-Seidel
        z[c1] = dg[c1] * r[c1]
        z[c2] = z[c2] + dg[c2] * (r[c2] - S0c2 @ z)
        z = z + omega * (P01 @ coarse(P01.T @ (r - S0 @ z)))
        z[c2] = z[c2] + dg[c2] * (r[c2] - S0c2 @ z)
        z[c1] = z[c1] + dg[c1] * (r[c1] - S0c1 @ z)
        return z

    def apply(r):
        ru = r[:Nf]
        if mass == "exact": zu = luM.solve(ru)
        elif mass.startswith("cheb"): zu = cheb(Mf, 1 / dM, ru, np.zeros(Nf), int(mass[4:]), 0.5, 1.5)
        else: zu = ru / dM
        return np.concatenate([zu, vc(r[Nf:])])
    return apply


def consistent_blocks():
    for j in range(2):
        xi = Yarn5().jump(4096 * j).normals(4096)
        kf = orc.sampler_eval(0, xi)[0]
        A0, M0, B0, b0 = darcy_mats(d0, kf)
        dM = M0.diagonal(); Nf = M0.shape[0]
        Dinv = sp.diags(1 / dM)
        S1 = (B0 @ Dinv @ B0.T).tocsc()
        S2 = (B0 @ (Dinv - Dinv @ (M0 - sp.diags(dM)) @ Dinv) @ B0.T).tocsc()
        Sx = sp.csc_matrix(B0 @ spl.splu(M0.tocsc()).solve(B0.T.toarray()))
        out = {}
        for name, S in (("lumped", S1), ("Neumann-1", S2), ("exact", Sx)):
            lu = spl.splu(S)
            for mass in ("jacobi", "cheb2"):
                def P(r, lu=lu, mass=mass):
                    ru = r[:Nf]
                    zu = ru / dM if mass == "jacobi" else cheb(M0, 1 / dM, ru, np.zeros(Nf), 2, 0.5, 1.5)
                    return np.concatenate([zu, lu.solve(r[Nf:])])
                out[f"S {name} / M {mass}"] = minres(A0, b0, P, np.zeros_like(b0), 1e-6 * np.sqrt(b0 @ P(b0)))[1]
        print(f"realisation {j}: {out}; entries per row: lumped {S1.nnz / S1.shape[0]:.1f}, Neumann-1 {S2.nnz / S2.shape[0]:.1f}", flush=True)


if __name__ == "__main__":
    if "--consistent" in sys.argv:
        consistent_blocks()
        sys.exit(0)
    for j in range(3):
        xi = Yarn5().jump(4096 * j).normals(4096)
        kf = orc.sampler_eval(0, xi)[0]
        kc = orc.sampler_eval(1, xi, xi_level=0)[0]
        A0, M0, B0, b0 = darcy_mats(d0, kf)
        out = {}
        for name, kw in [("default", {}), ("exact Schur", {"schur": "exact"}), ("RB-SGS", {"schur": "gs"}),
                         ("RB-SGS omega 2", {"schur": "gs", "omega": 2.0}), ("mass cheb2", {"mass": "cheb2"}),
                         ("mass cheb3", {"mass": "cheb3"}), ("mass exact", {"mass": "exact"})]:
            P = make_prec(M0, B0, **kw)
            goal = 1e-6 * np.sqrt(b0 @ P(b0))
            out[name] = minres(A0, b0, P, np.zeros_like(b0), goal)[1]
        # warm start from the prolongated coarse solution (exact Schur block)
        A1, M1, B1, b1 = darcy_mats(d1, kc)
        xc = spl.spsolve(A1.tocsc(), b1)
        x0 = np.concatenate([Pu @ xc[:d1.Nf], P01 @ xc[d1.Nf:]])
        ess = np.asarray(d0.ess_u).astype(bool)
        x0[:d0.Nf][ess] = np.asarray(d0.ess_data)[:d0.Nf][ess]
        P = make_prec(M0, B0, schur="exact")
        etab = np.sqrt(b0 @ P(b0)); r0 = b0 - A0 @ x0
        out["warm start: residual ratio"] = round(float(np.sqrt(r0 @ P(r0)) / etab), 3)
        out["warm start: its (cold)"] = (minres(A0, b0, P, x0, 1e-6 * etab)[1], minres(A0, b0, P, np.zeros_like(b0), 1e-6 * etab)[1])
        print(f"realisation {j}: {out}", flush=True)
