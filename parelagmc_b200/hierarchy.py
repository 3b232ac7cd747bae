"""Structured RT0/P0 multilevel hierarchy provider (the stand-in for ParELAG in this repo).

In the reference the agglomerated hierarchy (prolongators P, the Raviart-Thomas / L2 spaces,
the B and W operators, per-agglomerate mass blocks) is built once on the host by ParELAG
(`/root/reference/src/DarcySolver.cpp:60-244`, `/root/reference/src/PDESampler.cpp:177-334`) and is
NOT part of the per-sample hot path.  ParELAG/MFEM are not available here, so this module produces the
same *kind* of data for Cartesian quad/hex meshes with geometric (tensor-product) agglomeration:

  * lowest-order Raviart-Thomas dofs = total flux through a face in the +axis direction,
  * piecewise-constant (VALUE-type) L2 dofs, so W = diag(|e|), D = signed incidence / |e|, B = W D,
  * prolongators: P_u = coarse RT0 functions expressed in fine flux dofs, P_s = 0/1 aggregation,
  * per-element dense local mass blocks for the unit coefficient (`VectorFEMassIntegrator(1)`,
    `/root/reference/examples/MLMC.cpp:225`).

On nested Cartesian meshes the ParELAG order-0 coarse spaces coincide with the coarse RT0/P0 spaces
(the coarse RT0 function is the minimum-energy extension of a uniform face flux), so this is the same
discretisation, not an approximation of it.  Everything downstream (oracle, C-ABI upload) consumes plain
CSR arrays and does not care where they came from.

Mesh conventions follow MFEM's Cartesian mesh generator as used by the reference's inline meshes
(`/root/reference/meshes/cube_hex.mesh`, `/root/reference/meshes/inline_quad.mesh`):
boundary attributes 3D: bottom(z=0)=1, front(y=0)=2, right(x=L)=3, back(y=L)=4, left(x=0)=5, top(z=L)=6;
2D: bottom(y=0)=1, right(x=L)=2, top(y=L)=3, left(x=0)=4.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import scipy.sparse as sp


# --------------------------------------------------------------------------------------
# closed-form constants of the path
# --------------------------------------------------------------------------------------
def matern_scaling_coefficient(corlen: float, ndim: int) -> float:
    """g of `ComputeScalingCoefficientForSPDE` (`/root/reference/src/Utilities.hpp:188-200`).

    Uses tgamma(nu + d) exactly as the code does (not the doc comment's nu + d/2).
    """
    d = float(ndim)
    nu = 2.0 - d / 2.0
    gnu = math.gamma(nu)
    gnudim = math.gamma(nu + d)
    c = math.pow(16.0 * math.atan(1.0), 0.5 * d)
    k = math.pow(1.0 / corlen, 2.0 * nu)
    return math.sqrt(c * gnudim * k / gnu)


def spde_alpha(corlen: float) -> float:
    """alpha = kappa^2 = 1/corlen^2 (`/root/reference/src/PDESampler.cpp:42`)."""
    return 1.0 / (corlen * corlen)


# --------------------------------------------------------------------------------------
# one level of a tensor-product box mesh
# --------------------------------------------------------------------------------------
@dataclass
class BoxLevel:
    """One level: a tensor-product grid given by its node coordinates per axis."""

    nodes: List[np.ndarray]            # per axis, increasing coordinates, len n_a + 1
    dim: int = field(init=False)
    n: np.ndarray = field(init=False)  # elements per axis
    Ne: int = field(init=False)
    Nf: int = field(init=False)
    face_off: np.ndarray = field(init=False)  # offset of each axis' face block

    def __post_init__(self):
        self.dim = len(self.nodes)
        self.n = np.array([len(x) - 1 for x in self.nodes], dtype=np.int64)
        self.Ne = int(np.prod(self.n))
        cnt = []
        for a in range(self.dim):
            m = self.n.copy()
            m[a] += 1
            cnt.append(int(np.prod(m)))
        self.face_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        self.Nf = int(self.face_off[-1])

    # -- index helpers (x fastest) --------------------------------------------------------
    def elem_index(self, idx: Sequence[np.ndarray]) -> np.ndarray:
        e = np.zeros_like(idx[0], dtype=np.int64)
        stride = 1
        for a in range(self.dim):
            e = e + stride * idx[a]
            stride *= int(self.n[a])
        return e

    def face_index(self, axis: int, idx: Sequence[np.ndarray]) -> np.ndarray:
        """Face normal to `axis`; idx[axis] in [0, n_axis], the others in [0, n_b)."""
        f = np.zeros_like(idx[0], dtype=np.int64)
        stride = 1
        for a in range(self.dim):
            f = f + stride * idx[a]
            stride *= int(self.n[a]) + (1 if a == axis else 0)
        return f + int(self.face_off[axis])

    def elem_grid(self) -> List[np.ndarray]:
        """Per-axis integer index arrays of all elements, in element numbering order."""
        grids = np.meshgrid(*[np.arange(int(m)) for m in self.n], indexing="ij")
        # element numbering is x fastest -> order 'F'
        return [g.ravel(order="F") for g in grids]

    def h(self, axis: int) -> np.ndarray:
        return np.diff(self.nodes[axis])

    def elem_sizes(self) -> List[np.ndarray]:
        idx = self.elem_grid()
        return [self.h(a)[idx[a]] for a in range(self.dim)]

    def volumes(self) -> np.ndarray:
        v = np.ones(self.Ne)
        for s in self.elem_sizes():
            v = v * s
        return v


def _csr(mat) -> sp.csr_matrix:
    m = sp.csr_matrix(mat)
    m.sum_duplicates()
    m.sort_indices()
    m.indices = m.indices.astype(np.int32)
    m.indptr = m.indptr.astype(np.int32)
    return m


@dataclass
class LevelData:
    """Everything ParELAG would hand over for one level (host arrays, uploaded once)."""

    dim: int
    Ne: int
    Nf: int
    # element -> RT dofs and dense unit-coefficient local mass blocks
    elem_ptr: np.ndarray      # int32 [Ne+1]
    elem_dofs: np.ndarray     # int32 [sum n_e]
    elem_mat_ptr: np.ndarray  # int64 [Ne+1] offsets into elem_mat (n_e^2 each)
    elem_mat: np.ndarray      # float64
    Wdiag: np.ndarray         # float64 [Ne]  (L2 mass, diagonal)
    D: sp.csr_matrix          # Ne x Nf  discrete divergence
    B: sp.csr_matrix          # Ne x Nf  = W D
    bdr_face: np.ndarray      # int32 [nb] boundary RT dofs
    bdr_attr: np.ndarray      # int32 [nb] attribute (1-based)
    bdr_sign: np.ndarray      # float64 [nb] +1 if dof orientation is outward, else -1
    P_u: Optional[sp.csr_matrix] = None   # Nf x Nf_coarse   (to the next coarser level)
    P_s: Optional[sp.csr_matrix] = None   # Ne x Ne_coarse
    grid: Optional[BoxLevel] = None

    @property
    def N(self) -> int:
        return self.Nf + self.Ne

    def assemble_M(self, k: Optional[np.ndarray] = None) -> sp.csr_matrix:
        """M(k) = sum_e k_e R_e^T M_e R_e  (`DeRhamSequence::ComputeMassOperator(uform[,k])`,
        call sites `/root/reference/src/DarcySolver.cpp:479`, `/root/reference/src/PDESampler.cpp:232`)."""
        ne = np.diff(self.elem_ptr)
        rows = np.repeat(self.elem_dofs, np.repeat(ne, ne))
        # column index: for each element, tile dofs n_e times
        cols = np.concatenate([np.tile(self.elem_dofs[self.elem_ptr[e]:self.elem_ptr[e + 1]], ne[e])
                               for e in range(self.Ne)]) if not _uniform(ne) else _tile_uniform(
            self.elem_dofs, int(ne[0]))
        vals = self.elem_mat
        if k is not None:
            vals = vals * np.repeat(np.asarray(k, dtype=np.float64), ne * ne)
        M = sp.coo_matrix((vals, (rows, cols)), shape=(self.Nf, self.Nf))
        return _csr(M)


def _uniform(ne: np.ndarray) -> bool:
    return bool(np.all(ne == ne[0]))


def _tile_uniform(elem_dofs: np.ndarray, n: int) -> np.ndarray:
    d = elem_dofs.reshape(-1, n)
    return np.repeat(d[:, None, :], n, axis=1).reshape(-1)


def _build_level(grid: BoxLevel) -> LevelData:
    dim = grid.dim
    idx = grid.elem_grid()
    hs = grid.elem_sizes()
    vol = grid.volumes()
    nloc = 2 * dim
    Ne = grid.Ne
    # element dofs: [x-, x+, y-, y+, (z-, z+)]
    dofs = np.empty((Ne, nloc), dtype=np.int64)
    for a in range(dim):
        lo = [i.copy() for i in idx]
        hi = [i.copy() for i in idx]
        hi[a] = hi[a] + 1
        dofs[:, 2 * a] = grid.face_index(a, lo)
        dofs[:, 2 * a + 1] = grid.face_index(a, hi)
    # local mass: direction a block = h_a^2 / |e| * [[1/3, 1/6], [1/6, 1/3]]
    # (flux dofs: u_a = (F0 (1-t) + F1 t) / (|e| / h_a)  =>  int u_a^2 = h_a^2/|e| * (F0^2/3 + F0 F1/3 + F1^2/3))
    mats = np.zeros((Ne, nloc, nloc))
    for a in range(dim):
        c = hs[a] * hs[a] / vol
        mats[:, 2 * a, 2 * a] = c / 3.0
        mats[:, 2 * a + 1, 2 * a + 1] = c / 3.0
        mats[:, 2 * a, 2 * a + 1] = c / 6.0
        mats[:, 2 * a + 1, 2 * a] = c / 6.0
    # divergence in VALUE-type P0 dofs: (sum of outward fluxes) / |e|
    rows = np.repeat(np.arange(Ne, dtype=np.int64), nloc)
    sgn = np.tile(np.array([-1.0, 1.0] * dim), Ne)
    Binc = _csr(sp.coo_matrix((sgn, (rows, dofs.ravel())), shape=(Ne, grid.Nf)))
    D = _csr(sp.diags(1.0 / vol) @ Binc)
    B = _csr(sp.diags(vol) @ D)
    # boundary faces and attributes
    if dim == 3:
        attr = {(2, 0): 1, (1, 0): 2, (0, 1): 3, (1, 1): 4, (0, 0): 5, (2, 1): 6}
    elif dim == 2:
        attr = {(1, 0): 1, (0, 1): 2, (1, 1): 3, (0, 0): 4}
    else:
        raise ValueError("dim must be 2 or 3")
    bf, ba, bs = [], [], []
    for a in range(dim):
        others = [b for b in range(dim) if b != a]
        grids = np.meshgrid(*[np.arange(int(grid.n[b])) for b in others], indexing="ij")
        for side in (0, 1):
            ii = [None] * dim
            for b, g in zip(others, grids):
                ii[b] = g.ravel(order="F")
            ii[a] = np.full_like(ii[others[0]], 0 if side == 0 else int(grid.n[a]))
            f = grid.face_index(a, ii)
            bf.append(f)
            ba.append(np.full(f.shape, attr[(a, side)], dtype=np.int32))
            bs.append(np.full(f.shape, -1.0 if side == 0 else 1.0))
    return LevelData(
        dim=dim, Ne=Ne, Nf=grid.Nf,
        elem_ptr=(np.arange(Ne + 1, dtype=np.int64) * nloc).astype(np.int32),
        elem_dofs=dofs.ravel().astype(np.int32),
        elem_mat_ptr=np.arange(Ne + 1, dtype=np.int64) * nloc * nloc,
        elem_mat=mats.ravel(),
        Wdiag=vol.copy(),
        D=D, B=B,
        bdr_face=np.concatenate(bf).astype(np.int32),
        bdr_attr=np.concatenate(ba).astype(np.int32),
        bdr_sign=np.concatenate(bs),
        grid=grid,
    )


def _coarsen_axis(nodes: np.ndarray, factor: int = 2):
    """Group consecutive fine cells by `factor`; a short remainder group closes the axis."""
    n = len(nodes) - 1
    starts = np.arange(0, n, factor)
    cnodes = np.concatenate([nodes[starts], nodes[-1:]])
    parent = np.minimum(np.arange(n) // factor, len(starts) - 1)
    return cnodes, parent.astype(np.int64), np.concatenate([starts, [n]]).astype(np.int64)


def _prolongators(fine: BoxLevel, coarse: BoxLevel, parents, starts):
    dim = fine.dim
    # --- P_s : 0/1 aggregation ---
    fidx = fine.elem_grid()
    cidx = [parents[a][fidx[a]] for a in range(dim)]
    ce = coarse.elem_index(cidx)
    P_s = _csr(sp.coo_matrix((np.ones(fine.Ne), (np.arange(fine.Ne), ce)), shape=(fine.Ne, coarse.Ne)))
    # --- P_u : coarse RT0 flux functions sampled on fine faces ---
    rows, cols, vals = [], [], []
    for a in range(dim):
        m = fine.n.copy()
        m[a] += 1
        grids = np.meshgrid(*[np.arange(int(x)) for x in m], indexing="ij")
        ii = [g.ravel(order="F") for g in grids]
        f = fine.face_index(a, ii)
        # transverse share = fine face area / coarse face area
        share = np.ones(f.shape)
        cpar = [None] * dim
        for b in range(dim):
            if b == a:
                continue
            cpar[b] = parents[b][ii[b]]
            share = share * fine.h(b)[ii[b]] / coarse.h(b)[cpar[b]]
        # position along the axis: fine node ii[a] lies in coarse cell I (or on a coarse node)
        xa = fine.nodes[a][ii[a]]
        # coarse cell containing the node (right-closed at the end)
        I = np.minimum(np.searchsorted(coarse.nodes[a], xa, side="right") - 1, int(coarse.n[a]) - 1)
        t = (xa - coarse.nodes[a][I]) / coarse.h(a)[I]
        on_left = ii[a] == starts[a][I]
        on_right = ii[a] == starts[a][I + 1]
        t = np.where(on_left, 0.0, np.where(on_right, 1.0, t))
        for side, w in ((0, (1.0 - t) * share), (1, t * share)):
            jj = [c for c in cpar]
            jj[a] = I + side
            F = coarse.face_index(a, jj)
            keep = w != 0.0
            rows.append(f[keep])
            cols.append(F[keep])
            vals.append(w[keep])
    P_u = _csr(sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                             shape=(fine.Nf, coarse.Nf)))
    return P_u, P_s


def build_box_hierarchy(n_fine: Sequence[int], lengths: Sequence[float], nlevels: int,
                        factor: int = 2) -> List[LevelData]:
    """Levels 0 (finest) .. nlevels-1 (coarsest) of an n_fine box mesh on [0,lengths]."""
    dim = len(n_fine)
    nodes = [np.linspace(0.0, float(lengths[a]), int(n_fine[a]) + 1) for a in range(dim)]
    grids = [BoxLevel(nodes)]
    maps = []
    for _ in range(nlevels - 1):
        cn, par, st = [], [], []
        for a in range(dim):
            c, p, s = _coarsen_axis(grids[-1].nodes[a], factor)
            cn.append(c)
            par.append(p)
            st.append(s)
        grids.append(BoxLevel(cn))
        maps.append((par, st))
    levels = [_build_level(g) for g in grids]
    for l in range(nlevels - 1):
        P_u, P_s = _prolongators(grids[l], grids[l + 1], *maps[l])
        levels[l].P_u = P_u
        levels[l].P_s = P_s
    return levels


# --------------------------------------------------------------------------------------
# tetrahedral meshes (BASELINE configs[2] is a tet mesh): Kuhn triangulation of a cube grid, nested under refinement
# --------------------------------------------------------------------------------------
_KUHN = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]


def _kuhn_locate(x: np.ndarray, n: int, h: float) -> np.ndarray:
    """Index of a Kuhn simplex (cell * 6 + permutation) whose closure contains each point of x [m, 3]."""
    c = np.clip(np.floor(x / h + 1e-12).astype(np.int64), 0, n - 1)
    loc = x / h - c
    order = np.argsort(-loc, axis=1, kind="stable")          # a_{pi0} >= a_{pi1} >= a_{pi2}
    code = order[:, 0] * 9 + order[:, 1] * 3 + order[:, 2]
    lut = np.full(27, -1, dtype=np.int64)
    for i, pm in enumerate(_KUHN):
        lut[pm[0] * 9 + pm[1] * 3 + pm[2]] = i
    cell = (c[:, 2] * n + c[:, 1]) * n + c[:, 0]
    return cell * 6 + lut[code]


@dataclass
class _TetMesh:
    n: int
    h: float
    verts: np.ndarray      # [nv, 3]
    tets: np.ndarray       # [Ne, 4] vertex ids; local face a is opposite vertex a
    tet_face: np.ndarray   # [Ne, 4] global face ids
    tet_sign: np.ndarray   # [Ne, 4] +1 if the tet's outward normal agrees with the face's global orientation
    face_verts: np.ndarray  # [Nf, 3] sorted vertex ids (global orientation = (b-a) x (c-a))
    vol: np.ndarray        # [Ne]


def _kuhn_mesh(n: int, length: float) -> _TetMesh:
    h = length / n
    g = np.arange(n + 1) * h
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    vid = lambda i, j, k: (k * (n + 1) + j) * (n + 1) + i
    verts = np.empty(((n + 1) ** 3, 3))
    I, J, K = np.meshgrid(np.arange(n + 1), np.arange(n + 1), np.arange(n + 1), indexing="ij")
    verts[vid(I, J, K).ravel()] = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    ci, cj, ck = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    order = np.argsort(((ck * n + cj) * n + ci).ravel(), kind="stable")
    ci, cj, ck = ci.ravel()[order], cj.ravel()[order], ck.ravel()[order]
    tets = np.empty((n ** 3, 6, 4), dtype=np.int64)
    for t, pm in enumerate(_KUHN):
        off = np.zeros((4, 3), dtype=np.int64)
        for s_ in range(3):
            off[s_ + 1] = off[s_]
            off[s_ + 1, pm[s_]] += 1
        for v in range(4):
            tets[:, t, v] = vid(ci + off[v, 0], cj + off[v, 1], ck + off[v, 2])
    tets = tets.reshape(-1, 4)
    Ne = tets.shape[0]
    # faces: local face a = the three vertices other than a; numbered in order of first appearance (locality)
    loc = np.stack([np.sort(np.delete(tets, a, axis=1), axis=1) for a in range(4)], axis=1).reshape(-1, 3)   # [Ne*4, 3]
    key = (loc[:, 0] * verts.shape[0] + loc[:, 1]) * verts.shape[0] + loc[:, 2]
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    rank = np.empty_like(first)
    rank[np.argsort(first, kind="stable")] = np.arange(first.size)
    tet_face = rank[inv].reshape(Ne, 4)
    face_verts = np.empty((first.size, 3), dtype=np.int64)
    face_verts[rank] = loc[first]
    Pv = verts[tets]                                                  # [Ne, 4, 3]
    vol = np.abs(np.einsum("ex,ex->e", np.cross(Pv[:, 1] - Pv[:, 0], Pv[:, 2] - Pv[:, 0]), Pv[:, 3] - Pv[:, 0])) / 6.0
    fa, fb, fc = (verts[face_verts[:, i]] for i in range(3))
    area_vec = 0.5 * np.cross(fb - fa, fc - fa)                       # global orientation
    cen = (fa + fb + fc) / 3.0
    sign = np.empty((Ne, 4))
    for a in range(4):
        f = tet_face[:, a]
        sign[:, a] = np.sign(np.einsum("ex,ex->e", area_vec[f], cen[f] - Pv[:, a]))
    return _TetMesh(n=n, h=h, verts=verts, tets=tets, tet_face=tet_face, tet_sign=sign, face_verts=face_verts, vol=vol)


def _tet_level(m: _TetMesh, length: float) -> LevelData:
    """RT0 / P0 on tetrahedra with flux dofs: phi_a = sigma_a (x - p_a) / (3 |T|), so that the flux of phi_a through
    face a in the face's global orientation is 1 (the dofs ParELAG's finest level inherits from MFEM's RT0)."""
    Ne, Nf = m.tets.shape[0], m.face_verts.shape[0]
    Pv = m.verts[m.tets]
    A = Pv[:, None, :, :] - Pv[:, :, None, :]                          # A[e, a, c] = p_c - p_a
    G = np.einsum("eacx,ebdx->eabcd", A, A)
    I = (G.sum(axis=(3, 4)) + np.einsum("eabcc->eab", G)) * (m.vol / 20.0)[:, None, None]
    mats = m.tet_sign[:, :, None] * m.tet_sign[:, None, :] * I / (9.0 * m.vol * m.vol)[:, None, None]
    rows = np.repeat(np.arange(Ne, dtype=np.int64), 4)
    Binc = _csr(sp.coo_matrix((m.tet_sign.ravel(), (rows, m.tet_face.ravel())), shape=(Ne, Nf)))
    D = _csr(sp.diags(1.0 / m.vol) @ Binc)
    B = _csr(sp.diags(m.vol) @ D)
    # boundary faces: exactly one adjacent tet; attribute from the cube side (same numbering as the hex meshes)
    cnt = np.bincount(m.tet_face.ravel(), minlength=Nf)
    bf = np.nonzero(cnt == 1)[0]
    fv = m.verts[m.face_verts[bf]]                                    # [nb, 3, 3]
    attr_of = {(2, 0): 1, (1, 0): 2, (0, 1): 3, (1, 1): 4, (0, 0): 5, (2, 1): 6}
    ba = np.zeros(bf.size, dtype=np.int32)
    for (axis, side), at in attr_of.items():
        on = np.all(np.abs(fv[:, :, axis] - (0.0 if side == 0 else length)) < 1e-9 * max(length, 1.0), axis=1)
        ba[on] = at
    assert np.all(ba > 0)
    sgn_of_face = np.zeros(Nf)
    sgn_of_face[m.tet_face.ravel()] = m.tet_sign.ravel()              # boundary faces have one contributor
    return LevelData(dim=3, Ne=Ne, Nf=Nf, elem_ptr=(np.arange(Ne + 1, dtype=np.int64) * 4).astype(np.int32),
                     elem_dofs=m.tet_face.ravel().astype(np.int32), elem_mat_ptr=np.arange(Ne + 1, dtype=np.int64) * 16,
                     elem_mat=mats.ravel(), Wdiag=m.vol.copy(), D=D, B=B, bdr_face=bf.astype(np.int32), bdr_attr=ba,
                     bdr_sign=sgn_of_face[bf].copy())


def _tet_prolongators(fine: _TetMesh, coarse: _TetMesh):
    """Nested Kuhn meshes: P_s = injection of piecewise constants; P_u = RT0 interpolation, i.e. the flux of every
    coarse basis function through every fine face."""
    cen_t = fine.verts[fine.tets].mean(axis=1)
    parent = _kuhn_locate(cen_t, coarse.n, coarse.h)
    P_s = _csr(sp.coo_matrix((np.ones(parent.size), (np.arange(parent.size), parent)),
                             shape=(fine.tets.shape[0], coarse.tets.shape[0])))
    fa, fb, fc = (fine.verts[fine.face_verts[:, i]] for i in range(3))
    S = 0.5 * np.cross(fb - fa, fc - fa)
    xc = (fa + fb + fc) / 3.0
    T = _kuhn_locate(xc, coarse.n, coarse.h)
    Pv = coarse.verts[coarse.tets[T]]                                  # [Nf_f, 4, 3]
    val = coarse.tet_sign[T] * np.einsum("fax,fx->fa", xc[:, None, :] - Pv, S) / (3.0 * coarse.vol[T])[:, None]
    rows = np.repeat(np.arange(xc.shape[0]), 4)
    cols = coarse.tet_face[T].ravel()
    v = val.ravel()
    keep = np.abs(v) > 1e-12
    P_u = _csr(sp.coo_matrix((v[keep], (rows[keep], cols[keep])), shape=(xc.shape[0], coarse.face_verts.shape[0])))
    return P_u, P_s


def build_tet_hierarchy(n_fine: int, length: float, nlevels: int) -> List[LevelData]:
    """Levels 0 (finest) .. nlevels-1 of the Kuhn triangulation (6 tetrahedra per cube) of an n_fine^3 grid on
    [0, length]^3, coarsened by halving the grid; the coarse spaces are RT0 / P0 on the coarse tetrahedra, which the
    nested refinement makes subspaces of the fine ones."""
    assert n_fine % (2 ** (nlevels - 1)) == 0
    meshes = [_kuhn_mesh(n_fine >> l, length) for l in range(nlevels)]
    levels = [_tet_level(m, length) for m in meshes]
    for l in range(nlevels - 1):
        levels[l].P_u, levels[l].P_s = _tet_prolongators(meshes[l], meshes[l + 1])
    return levels


# --------------------------------------------------------------------------------------
# unstructured agglomerates (what `BuildTopologyAlgebraic` + ParELAG's order-0 coarsening yield)
# --------------------------------------------------------------------------------------
def greedy_partition(level: LevelData, target: int, seed: int = 0) -> np.ndarray:
    """Connected, irregular agglomerates of about `target` elements each, grown breadth-first from random seeds over the
    element-face-element graph (a METIS stand-in: `/root/reference/src/Utilities.cpp:125-155` asks METIS for contiguous
    k-way parts of about `coarsening_factor` elements)."""
    inc = sp.csr_matrix((np.ones(level.B.nnz), level.B.indices, level.B.indptr), shape=level.B.shape)
    adj = sp.csr_matrix(inc @ inc.T)
    adj.setdiag(0)
    adj.eliminate_zeros()
    rng = np.random.default_rng(seed)
    part = np.full(level.Ne, -1, dtype=np.int64)
    npart = 0
    for e0 in rng.permutation(level.Ne):
        if part[e0] >= 0:
            continue
        size = int(rng.integers(max(2, target // 2), target + target // 2 + 1))
        front, members = [int(e0)], [int(e0)]
        part[e0] = npart
        while front and len(members) < size:
            e = front.pop(0)
            nb = adj.indices[adj.indptr[e]:adj.indptr[e + 1]]
            for j in rng.permutation(nb):
                if part[j] < 0 and len(members) < size:
                    part[j] = npart
                    members.append(int(j))
                    front.append(int(j))
        npart += 1
    # merge singletons into a neighbour so that every agglomerate has interior faces or at least two elements
    sizes = np.bincount(part, minlength=npart)
    for e in np.nonzero(sizes[part] == 1)[0]:
        nb = adj.indices[adj.indptr[e]:adj.indptr[e + 1]]
        if nb.size:
            part[e] = part[nb[0]]
    _, part = np.unique(part, return_inverse=True)
    return part.astype(np.int64)


def coarsen_by_agglomeration(fine: LevelData, part: np.ndarray) -> LevelData:
    """The order-0 coarse spaces ParELAG builds on arbitrary agglomerates (`DeRhamSequence::Coarsen`, reached from
    `/root/reference/src/PDESampler.cpp:166-171`, `src/DarcySolver.cpp:160-170`): one L2 dof per agglomerate (piecewise
    constants), one RT dof per agglomerated face (the set of fine faces shared by the same two agglomerates, or lying on
    the boundary with the same attribute), whose basis function has a flux trace proportional to the fine faces' share and
    is extended into the two agglomerates by the minimum-energy flux with constant divergence (local mixed problems).
    Fills `fine.P_u`, `fine.P_s` and returns the coarse `LevelData`: dof lists of the agglomerates, their dense local mass
    blocks `P_A^T M_A P_A` (up to ~20 x 20: rows far wider than the 6-7 entries of the structured meshes), `D`, `W`,
    boundary data.  `D_f P_u = P_s D_c` holds by construction."""
    Ne, Nf = fine.Ne, fine.Nf
    nA = int(part.max()) + 1
    Binc = sp.csr_matrix(fine.B)                        # signed incidence (+1: the face's + direction leaves the element)
    Bt = Binc.T.tocsr()
    vol = fine.Wdiag
    volA = np.bincount(part, weights=vol, minlength=nA)
    battr = np.zeros(Nf, dtype=np.int64)
    battr[fine.bdr_face] = fine.bdr_attr
    # ---- coarse faces ----
    key_of, faces_of, F_low, F_high, F_attr, interior_of = {}, [], [], [], [], [[] for _ in range(nA)]
    face_sign = np.zeros(Nf)                             # +1 if the fine face's + direction agrees with its coarse face's
    cface = np.full(Nf, -1, dtype=np.int64)
    for f in range(Nf):
        els = Bt.indices[Bt.indptr[f]:Bt.indptr[f + 1]]
        sg = Bt.data[Bt.indptr[f]:Bt.indptr[f + 1]]
        if els.size == 2 and part[els[0]] == part[els[1]]:
            interior_of[part[els[0]]].append(f)
            continue
        if els.size == 2:
            a, b = (0, 1) if part[els[0]] < part[els[1]] else (1, 0)
            key = (int(part[els[a]]), int(part[els[b]]), 0)
            sign = sg[a]                                 # + direction of the coarse face: out of the lower agglomerate
        else:
            key = (int(part[els[0]]), -1, int(battr[f]))
            sign = sg[0]                                 # boundary: outward
        if key not in key_of:
            key_of[key] = len(faces_of)
            faces_of.append([])
            F_low.append(key[0]); F_high.append(key[1]); F_attr.append(key[2])
        F = key_of[key]
        faces_of[F].append(f)
        cface[f] = F
        face_sign[f] = sign
    nF = len(faces_of)
    # ---- agglomerate-local fine mass matrices (unit coefficient) ----
    ne = np.diff(fine.elem_ptr)
    el_of = [np.nonzero(part == A)[0] for A in range(nA)]
    cfaces_of = [[] for _ in range(nA)]
    for F in range(nF):
        cfaces_of[F_low[F]].append(F)
        if F_high[F] >= 0:
            cfaces_of[F_high[F]].append(F)
    rows_u, cols_u, vals_u = [], [], []
    c_ptr, c_dofs, c_mats, c_mptr = [0], [], [], [0]
    for A in range(nA):
        els = el_of[A]
        bnd = [f for F in cfaces_of[A] for f in faces_of[F]]
        inte = interior_of[A]
        loc = {f: i for i, f in enumerate(inte + bnd)}
        nI, nL = len(inte), len(inte) + len(bnd)
        MA = np.zeros((nL, nL))
        BA = np.zeros((els.size, nL))
        for r, e in enumerate(els):
            d = fine.elem_dofs[fine.elem_ptr[e]:fine.elem_ptr[e + 1]]
            Me = fine.elem_mat[fine.elem_mat_ptr[e]:fine.elem_mat_ptr[e + 1]].reshape(ne[e], ne[e])
            idx = [loc[int(f)] for f in d]
            MA[np.ix_(idx, idx)] += Me
            for p_ in range(Binc.indptr[e], Binc.indptr[e + 1]):
                BA[r, loc[int(Binc.indices[p_])]] = Binc.data[p_]
        PA = np.zeros((nL, len(cfaces_of[A])))
        for c, F in enumerate(cfaces_of[A]):
            g = np.zeros(nL)
            ff = faces_of[F]
            for f in ff:
                g[loc[f]] = face_sign[f] / len(ff)        # total flux 1 through F in its + direction
            sigma = 1.0 if F_low[F] == A else -1.0       # ... which leaves A (lower agglomerate) or enters it
            if nI > 0:
                K = np.zeros((nI + els.size, nI + els.size))
                K[:nI, :nI] = MA[:nI, :nI]
                K[:nI, nI:] = BA[:, :nI].T
                K[nI:, :nI] = BA[:, :nI]
                rhs = np.concatenate([-MA[:nI, nI:] @ g[nI:], sigma * vol[els] / volA[A] - BA[:, nI:] @ g[nI:]])
                g[:nI] = np.linalg.lstsq(K, rhs, rcond=None)[0][:nI]
            PA[:, c] = g
        for f, i in loc.items():
            for c, F in enumerate(cfaces_of[A]):
                if PA[i, c] != 0.0 and (i < nI or cface[f] == F):
                    # a boundary fine face belongs to one coarse face only; its value is written by the lower agglomerate
                    if i >= nI and F_low[F] != A:
                        continue
                    rows_u.append(f); cols_u.append(F); vals_u.append(PA[i, c])
        Mc = PA.T @ MA @ PA
        c_dofs.extend(cfaces_of[A])
        c_ptr.append(len(c_dofs))
        c_mats.append(0.5 * (Mc + Mc.T).ravel())
        c_mptr.append(c_mptr[-1] + Mc.size)
    fine.P_u = _csr(sp.coo_matrix((vals_u, (rows_u, cols_u)), shape=(Nf, nF)))
    fine.P_s = _csr(sp.coo_matrix((np.ones(Ne), (np.arange(Ne), part)), shape=(Ne, nA)))
    # ---- coarse D, B, boundary ----
    r, c, v = [], [], []
    for F in range(nF):
        r.append(F_low[F]); c.append(F); v.append(1.0)
        if F_high[F] >= 0:
            r.append(F_high[F]); c.append(F); v.append(-1.0)
    Binc_c = _csr(sp.coo_matrix((v, (r, c)), shape=(nA, nF)))
    D_c = _csr(sp.diags(1.0 / volA) @ Binc_c)
    bsel = [F for F in range(nF) if F_high[F] < 0]
    return LevelData(dim=fine.dim, Ne=nA, Nf=nF, elem_ptr=np.array(c_ptr, dtype=np.int32),
                     elem_dofs=np.array(c_dofs, dtype=np.int32), elem_mat_ptr=np.array(c_mptr, dtype=np.int64),
                     elem_mat=np.concatenate(c_mats), Wdiag=volA.copy(), D=D_c, B=Binc_c,
                     bdr_face=np.array(bsel, dtype=np.int32), bdr_attr=np.array([F_attr[F] for F in bsel], dtype=np.int32),
                     bdr_sign=np.ones(len(bsel)))


def build_agglomerated_hierarchy(fine: LevelData, nlevels: int, target: int = 8, seed: int = 0) -> List[LevelData]:
    """`fine` plus nlevels - 1 levels of irregular agglomerates of about `target` elements (unstructured coarsening)."""
    levels = [fine]
    for l in range(nlevels - 1):
        part = greedy_partition(levels[-1], target, seed + l)
        levels.append(coarsen_by_agglomeration(levels[-1], part))
    return levels


# --------------------------------------------------------------------------------------
# host-once setup that the reference performs in BuildHierarchy / Build*Functional
# --------------------------------------------------------------------------------------
@dataclass
class SamplerLevel:
    """Operators of `[M B^T; B -alpha W]` for one level, as `PDESampler::BuildHierarchy` leaves them
    (`/root/reference/src/PDESampler.cpp:218-284`)."""

    Ne: int
    Nf: int
    M: sp.csr_matrix        # eliminated: essential rows/cols zeroed, unit diagonal
    B: sp.csr_matrix        # W * D with essential columns zeroed
    Wdiag: np.ndarray       # diag(W) (positive, before the -alpha scaling)
    w_sqrt: np.ndarray      # sqrt(diag W)
    ess_u: np.ndarray       # int32 0/1 mask over RT dofs
    P: Optional[sp.csr_matrix]  # Ne x Ne_coarse  (Ps[i], `PDESampler.cpp:189-193`)
    nnz: int
    # optional transfer of the sampled field to the forward problem's mesh (applied before exp):
    # s = Tscale .* (T field);  T is n_out x Ne
    T: Optional[sp.csr_matrix] = None
    Tscale: Optional[np.ndarray] = None


def _eliminate_rowcol(M: sp.csr_matrix, ess: np.ndarray) -> sp.csr_matrix:
    """mfem::SparseMatrix::EliminateRowCol(rc) with the default DIAG_ONE policy."""
    keep = sp.diags((ess == 0).astype(np.float64))
    return _csr(keep @ M @ keep + sp.diags((ess != 0).astype(np.float64)))


def build_sampler_levels(levels: List[LevelData]) -> List[SamplerLevel]:
    out = []
    for lv in levels:
        ess = np.zeros(lv.Nf, dtype=np.int32)
        ess[lv.bdr_face] = 1                          # u.n = 0 on the whole boundary (`:210-214`)
        M = _eliminate_rowcol(lv.assemble_M(), ess)   # `:232,236-241`
        Dt = _csr(lv.D @ sp.diags((ess == 0).astype(np.float64)))   # EliminateCols `:243`
        Dt.eliminate_zeros()
        B = _csr(sp.diags(lv.Wdiag) @ Dt)             # `:245`
        B.eliminate_zeros()
        nnz = M.nnz + 2 * B.nnz + lv.Ne               # `:265`
        out.append(SamplerLevel(Ne=lv.Ne, Nf=lv.Nf, M=M, B=B, Wdiag=lv.Wdiag.copy(),
                                w_sqrt=np.sqrt(lv.Wdiag), ess_u=ess, P=lv.P_s, nnz=int(nnz)))
    return out


@dataclass
class DarcyLevel:
    """Per-level data of `DarcySolver` after BuildHierachySpaces / Build*ObservationFunctional /
    SetEssBdrConditions / BuildForcingTerms (`/root/reference/src/DarcySolver.cpp:60-414`)."""

    Ne: int
    Nf: int
    elem_ptr: np.ndarray
    elem_dofs: np.ndarray
    elem_mat_ptr: np.ndarray
    elem_mat: np.ndarray
    B: sp.csr_matrix        # un-eliminated W*D (`:203-207`)
    ess_u: np.ndarray       # int32 0/1 mask of essential RT dofs
    ess_data: np.ndarray    # float64 [N]
    rhs: np.ndarray         # float64 [N]
    obs: np.ndarray         # float64 [N]
    P_u: Optional[sp.csr_matrix]
    P_p: Optional[sp.csr_matrix]

    @property
    def N(self):
        return self.Nf + self.Ne


def build_darcy_levels(levels: List[LevelData], ess_attr: Sequence[int], obs_attr: Sequence[int],
                       inflow_attr: Sequence[int], p_inflow: float = -1.0,
                       qoi: str = "eff_perm") -> List[DarcyLevel]:
    """The MLMC drivers' deterministic problem (`/root/reference/examples/MLMC.cpp:214-239`):
    f = 0, q = 0, p_bdr = p_inflow on inflow attributes, u.n = 0 on essential attributes,
    QoI = int_{Gamma_obs} u.n  (eff_perm) or int p (p_int)."""
    ess_attr = np.asarray(ess_attr)
    obs_attr = np.asarray(obs_attr)
    inflow_attr = np.asarray(inflow_attr)
    out = []
    rhs = obs = None
    for l, lv in enumerate(levels):
        N = lv.N
        if l == 0:
            rhs = np.zeros(N)
            obs = np.zeros(N)
            a = lv.bdr_attr - 1
            # VectorFEBoundaryFluxLFIntegrator(coef): int coef * (phi . n_outward); phi.n integrates to +-1
            inflow = inflow_attr[a] != 0
            np.add.at(rhs, lv.bdr_face[inflow], p_inflow * lv.bdr_sign[inflow])
            if qoi == "eff_perm":
                ob = obs_attr[a] != 0
                np.add.at(obs, lv.bdr_face[ob], lv.bdr_sign[ob])
            elif qoi == "p_int":
                obs[lv.Nf:] = lv.Wdiag
            else:
                raise ValueError(qoi)
        else:
            prev = levels[l - 1]
            Pblk = sp.block_diag([prev.P_u, prev.P_s], format="csr")
            rhs = Pblk.T @ rhs          # `DarcySolver.cpp:410-411`
            obs = Pblk.T @ obs          # `:314-315`
        ess = np.zeros(lv.Nf, dtype=np.int32)
        on_ess = ess_attr[lv.bdr_attr - 1] != 0
        ess[lv.bdr_face[on_ess]] = 1
        out.append(DarcyLevel(Ne=lv.Ne, Nf=lv.Nf, elem_ptr=lv.elem_ptr, elem_dofs=lv.elem_dofs,
                              elem_mat_ptr=lv.elem_mat_ptr, elem_mat=lv.elem_mat, B=lv.B,
                              ess_u=ess, ess_data=np.zeros(N), rhs=np.asarray(rhs).copy(),
                              obs=np.asarray(obs).copy(), P_u=lv.P_u, P_p=lv.P_s))
    return out


# --------------------------------------------------------------------------------------
# element numbering of the reference's meshes (what maps stream position -> element)
# --------------------------------------------------------------------------------------
# offsets of the children of a quad / hex in the order of the parent's local vertices (MFEM's reference elements)
_CHILD_OFFSETS = {2: [(0, 0), (1, 0), (1, 1), (0, 1)],
                  3: [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]}


def mfem_refined_box_numbering(n_coarse: Sequence[int], nref: int) -> List[np.ndarray]:
    """Element numbering of an MFEM Cartesian mesh (`mfem::Mesh(nx, ny[, nz], ...)`, elements x fastest: what
    `/root/reference/examples/example_helpers/Build3DMesh.hpp:23-27` and the inline meshes under `/root/reference/meshes`
    produce) after `nref` calls of `UniformRefinement()` as the drivers make them
    (`/root/reference/examples/PDESamplerTest.cpp:151-156`): the child at the parent's local vertex 0 keeps the parent's
    index i, the children at local vertices 1 .. 2^d - 1 are appended at `NE + (2^d - 1) i + (j - 1)`.  This is the
    numbering ParELAG's `MFEMRefinedMeshPartitioner` relies on (`/root/reference/src/Utilities.cpp:28-38`: the first
    `level_nElements[l+1]` elements are the children 0, the rest come in groups of 2^d - 1), so the agglomerate of fine
    element i is `i` if `i < NE_coarse` else `(i - NE_coarse) / (2^d - 1)`, and level l's elements are numbered like the
    mesh refined `nref - l` times.

    Returns, for levels 0 (finest) .. nref (the unrefined mesh), `to_cart[l][e]` = x-fastest Cartesian index of element e
    in this numbering.  Pinned by the reference's own ctest goldens (`tests/test_reference_goldens.py`): the noise vector
    is consumed in this order (`/root/reference/src/NormalDistributionSampler.cpp:31-37`)."""
    dim = len(n_coarse)
    off = np.array(_CHILD_OFFSETS[dim], dtype=np.int64)
    nch = 1 << dim
    n = np.array(n_coarse, dtype=np.int64)
    grids = np.meshgrid(*[np.arange(int(m)) for m in n], indexing="ij")
    ijk = np.stack([g.ravel(order="F") for g in grids], axis=1)          # x fastest
    out = []
    for r in range(nref + 1):
        stride = np.concatenate([[1], np.cumprod(n[:-1])])
        out.append((ijk * stride[None, :]).sum(axis=1))
        if r == nref:
            break
        ne = ijk.shape[0]
        new = np.empty((nch * ne, dim), dtype=np.int64)
        new[:ne] = 2 * ijk + off[0][None, :]
        rest = (2 * ijk[:, None, :] + off[None, 1:, :]).reshape(-1, dim)     # parent-major, children 1..2^d-1
        new[ne:] = rest
        ijk = new
        n = 2 * n
    return out[::-1]


def _perm_rows(m, p):
    return None if m is None else _csr(sp.csr_matrix(m)[p, :])


def _perm_cols(m, p):
    return None if m is None else _csr(sp.csr_matrix(m)[:, p])


def renumber_sampler_levels(sampler: List["SamplerLevel"], to_old: Sequence[np.ndarray],
                            out_to_old: Optional[Sequence[np.ndarray]] = None) -> List["SamplerLevel"]:
    """The same sampler hierarchy with element (L2 dof) e of level l being the old element `to_old[l][e]`; RT dofs keep
    their numbers.  `out_to_old` renumbers the rows of the field transfer T (the forward mesh's elements)."""
    import dataclasses
    out = []
    for l, s in enumerate(sampler):
        p = np.asarray(to_old[l])
        pc = np.asarray(to_old[l + 1]) if s.P is not None else None
        T, Ts = s.T, s.Tscale
        if T is not None:
            T = _perm_cols(T, p)
            if out_to_old is not None:
                T = _perm_rows(T, np.asarray(out_to_old[l]))
                Ts = None if Ts is None else Ts[np.asarray(out_to_old[l])]
        out.append(dataclasses.replace(
            s, B=_perm_rows(s.B, p), Wdiag=s.Wdiag[p].copy(), w_sqrt=s.w_sqrt[p].copy(),
            P=None if s.P is None else _perm_cols(_perm_rows(s.P, p), pc), T=T, Tscale=Ts))
    return out


def renumber_darcy_levels(darcy: List["DarcyLevel"], to_old: Sequence[np.ndarray]) -> List["DarcyLevel"]:
    """The same Darcy hierarchy with element e of level l being the old element `to_old[l][e]`."""
    import dataclasses
    out = []
    for l, d in enumerate(darcy):
        p = np.asarray(to_old[l])
        pc = np.asarray(to_old[l + 1]) if d.P_p is not None else None
        ne = np.diff(d.elem_ptr)
        ptr = np.concatenate([[0], np.cumsum(ne[p])]).astype(np.int32)
        dofs = np.concatenate([d.elem_dofs[d.elem_ptr[e]:d.elem_ptr[e + 1]] for e in p]).astype(np.int32)
        mptr = np.concatenate([[0], np.cumsum((ne[p].astype(np.int64)) ** 2)])
        mats = np.concatenate([d.elem_mat[d.elem_mat_ptr[e]:d.elem_mat_ptr[e + 1]] for e in p])
        perm_all = np.concatenate([np.arange(d.Nf), d.Nf + p])
        out.append(dataclasses.replace(
            d, elem_ptr=ptr, elem_dofs=dofs, elem_mat_ptr=mptr, elem_mat=mats, B=_perm_rows(d.B, p),
            ess_data=d.ess_data[perm_all].copy(), rhs=d.rhs[perm_all].copy(), obs=d.obs[perm_all].copy(),
            P_p=None if d.P_p is None else _perm_cols(_perm_rows(d.P_p, p), pc)))
    return out


# --------------------------------------------------------------------------------------
# enlarged-domain samplers
# --------------------------------------------------------------------------------------
def embedded_selection(orig: List[LevelData], embed: List[LevelData], pad_cells: int) -> List[sp.csr_matrix]:
    """meshP of `EmbeddedPDESampler` (`/root/reference/src/EmbeddedPDESampler.cpp:63-89`, applied at `:426-435`) for a
    MATCHING enlarged box: the original n^d box sits `pad_cells` fine cells inside the enlarged one on every side; on
    level l the offset is pad_cells / 2^l.  Row e of level l selects the embedded element that coincides with e."""
    out = []
    for l, (lo, le) in enumerate(zip(orig, embed)):
        pad = pad_cells >> l
        assert pad << l == pad_cells, "padding must survive the coarsening"
        idx = lo.grid.elem_grid()
        j = le.grid.elem_index([i + pad for i in idx])
        out.append(_csr(sp.coo_matrix((np.ones(lo.Ne), (np.arange(lo.Ne), j)), shape=(lo.Ne, le.Ne))))
    return out


def tet_embedded_selection(n_orig: int, pad_cells: int, nlevels: int) -> List[sp.csr_matrix]:
    """meshP for the tetrahedral pair of `MLMC_EmbeddedPDESampler` (BASELINE configs[2]): the Kuhn mesh of the n^3 box
    sits `pad_cells` fine cells inside the Kuhn mesh of the (n + 2 pad)^3 box; tetrahedron (cell, permutation) of the
    original mesh coincides with (shifted cell, same permutation) of the enlarged one on every level."""
    out = []
    for l in range(nlevels):
        n, pad = n_orig >> l, pad_cells >> l
        assert pad << l == pad_cells and n << l == n_orig
        ne = n + 2 * pad
        i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
        co = ((k * n + j) * n + i).ravel()
        ce = (((k + pad) * ne + (j + pad)) * ne + (i + pad)).ravel()
        rows = (co[:, None] * 6 + np.arange(6)[None, :]).ravel()
        cols = (ce[:, None] * 6 + np.arange(6)[None, :]).ravel()
        out.append(_csr(sp.coo_matrix((np.ones(rows.size), (rows, cols)), shape=(6 * n ** 3, 6 * ne ** 3))))
    return out


def _overlap_1d(a: np.ndarray, b: np.ndarray) -> sp.csr_matrix:
    """|[a_i, a_i+1] n [b_j, b_j+1]| as a sparse (len(a)-1) x (len(b)-1) matrix."""
    lo = np.maximum(a[:-1, None], b[None, :-1])
    hi = np.minimum(a[1:, None], b[None, 1:])
    return sp.csr_matrix(np.maximum(hi - lo, 0.0))


def l2_projection_transfers(orig: List[LevelData], embed: List[LevelData]):
    """Mortar matrices of `L2ProjectionPDESampler` for two (possibly non-matching) box hierarchies: on level 0
    Gt[i, j] = |e_i n ebar_j| (piecewise constants; what ParMortarAssembler + L2MortarIntegrator assemble,
    `/root/reference/src/L2ProjectionPDESampler.cpp:488-505`), on coarser levels the Galerkin product
    Gt_{l+1} = P_orig^T Gt_l P_embed (`:507-514`); the apply is s = W_orig^-1 Gt s_bar (`:595-611`).
    Returns [(Gt_l, 1/diag W_orig_l)]."""
    go, ge = orig[0].grid, embed[0].grid
    G = None
    for a in range(go.dim):          # element numbering is x fastest => kron(last axis, ..., first axis)
        O = _overlap_1d(go.nodes[a], ge.nodes[a])
        G = O if G is None else sp.kron(O, G, format="csr")
    out = []
    for l in range(len(orig)):
        out.append((_csr(G), 1.0 / orig[l].Wdiag))
        if l + 1 < len(orig):
            G = (orig[l].P_s.T @ G @ embed[l].P_s).tocsr()
    return out


# --------------------------------------------------------------------------------------
# Bayesian inverse problem set-up
# --------------------------------------------------------------------------------------
def observation_functionals(levels: List[LevelData], coords: Sequence[Sequence[float]], eps: float,
                            rule: str = "center") -> List[np.ndarray]:
    """g_obs_func[i][level] of `BayesianInverseProblem` (`/root/reference/src/BayesianInverseProblem.cpp:46-104`): on
    level 0 the domain integral of 1 (`DomainLFIntegrator`: the element volumes) over the elements marked for observation
    point i, on coarser levels P^T of the finer one.  rule = "bbox" marks exactly what `ChangeMeshAttributes` does
    (`/root/reference/src/MeshUtilities.cpp:268-335`, default eps 0.01 at `MeshUtilities.hpp:61-62`): the elements whose
    bounding box, enlarged by eps, contains the point (`min - eps <= x < max + eps` per axis); rule = "center" marks the
    elements whose centre lies within `eps` (max-norm) of the point.  With no coordinates: one functional, the integral
    of p over the domain.  Returns [level] -> array [m, Ne]."""
    lv = levels[0]
    if len(coords) == 0:
        g0 = lv.Wdiag[None, :].copy()
    else:
        idx = lv.grid.elem_grid()
        ctr = [0.5 * (lv.grid.nodes[a][:-1] + lv.grid.nodes[a][1:])[idx[a]] for a in range(lv.dim)]
        g0 = np.zeros((len(coords), lv.Ne))
        for i, pt in enumerate(coords):
            near = np.ones(lv.Ne, dtype=bool)
            for a in range(lv.dim):
                if rule == "bbox":
                    lo, hi = lv.grid.nodes[a][:-1][idx[a]], lv.grid.nodes[a][1:][idx[a]]
                    near &= (lo - eps <= pt[a]) & (pt[a] < hi + eps)
                else:
                    near &= np.abs(ctr[a] - pt[a]) <= eps
            assert near.any(), f"observation point {pt} marks no element"
            g0[i, near] = lv.Wdiag[near]
    out = [g0]
    for l in range(len(levels) - 1):
        out.append((levels[l].P_s.T @ out[-1].T).T)
    return out


# The reference's default MLMC problem (`examples/example_helpers/CreateMLMCParameterList.hpp:27-41`)
MLMC_DEFAULT_BC = dict(ess_attr=[0, 1, 1, 1, 1, 0], obs_attr=[1, 0, 0, 0, 0, 0], inflow_attr=[0, 0, 0, 0, 0, 1])
# SPE10 XML (`examples/SPE10/spe10_3D_parameters.xml:45-49`)
SPE10_BC = dict(ess_attr=[1, 0, 1, 0, 1, 1], obs_attr=[0, 1, 0, 0, 0, 0], inflow_attr=[0, 0, 0, 1, 0, 0])


# --------------------------------------------------------------------------------------
# binary dump read by the C++ host layer (parelagmc_b200/host/HierarchyData.cpp)
# --------------------------------------------------------------------------------------
def dump_problem(path: str, sampler_levels: List[SamplerLevel], darcy_levels: List[DarcyLevel], dim: int,
                 corlen: float, gobs: Optional[Sequence[np.ndarray]] = None) -> None:
    """Write the host-once hierarchy data in the "PMCH2" layout of `HierarchyData::Load` ("PMCH3" when the observation
    functionals of a BayesianInverseProblem, `gobs[level]` = array [m, Ne], travel with it)."""
    import struct

    def ivec(f, a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        f.write(struct.pack("<q", a.size))
        f.write(a.tobytes())

    def dvec(f, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        f.write(struct.pack("<q", a.size))
        f.write(a.tobytes())

    def csr(f, m):
        if m is None:
            f.write(struct.pack("<i", 0))
            return
        m = _csr(m)
        f.write(struct.pack("<iii", 1, m.shape[0], m.shape[1]))
        ivec(f, m.indptr)
        ivec(f, m.indices)
        dvec(f, m.data)

    with open(path, "wb") as f:
        f.write(b"PMCH3\0\0\0" if gobs is not None else b"PMCH2\0\0\0")
        f.write(struct.pack("<iid", len(sampler_levels), dim, corlen))
        for s, d in zip(sampler_levels, darcy_levels):
            f.write(struct.pack("<ii", s.Ne, s.Nf))
            csr(f, s.M)
            csr(f, s.B)
            csr(f, s.P)
            dvec(f, s.Wdiag)
            csr(f, getattr(s, "T", None))                       # enlarged-domain samplers: field transfer (or absent)
            dvec(f, s.Tscale if getattr(s, "Tscale", None) is not None else np.zeros(0))
            f.write(struct.pack("<ii", d.Ne, d.Nf))
            ivec(f, d.elem_ptr)
            ivec(f, d.elem_dofs)
            dvec(f, d.elem_mat)
            csr(f, d.B)
            csr(f, d.P_p)
            ivec(f, d.ess_u)
            dvec(f, d.ess_data)
            dvec(f, d.rhs)
            dvec(f, d.obs)
        if gobs is not None:
            f.write(struct.pack("<i", int(np.asarray(gobs[0]).shape[0])))
            for g in gobs:
                dvec(f, np.asarray(g).ravel())
