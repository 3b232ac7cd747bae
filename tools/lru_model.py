"""LRU model of the Darcy block apply's gathers under different element orders (diagnostic, no GPU).

The library numbers the RT dofs by first touch in the caller's element order (csrc/pmc_b200.cu: first_touch_order), so
the element order decides how far apart the uses of a row of the gathered vector are.  This walks the rows of
[M(k) | B^T] in their packed order, touches the 32-byte rows each one gathers (u, k, p) and counts misses of an LRU
cache of the given size per CTA, relative to the compulsory reads.  Result on the bench's 16^3 level (DESIGN.md
section 5): the MFEM refinement order the mesh arrives in is the best of the four tried.
"""
import os, sys
from collections import OrderedDict
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import hex_problem
from parelagmc_b200 import hierarchy as H

p = hex_problem(16, 3)
d = p["darcy"][0]
Ne, Nf = d.Ne, d.Nf
ptr, dofs = np.asarray(d.elem_ptr), np.asarray(d.elem_dofs)
n = 16
to_cart = H.mfem_refined_box_numbering([4, 4, 4], 2)[0]
ci, cj, ck = to_cart % n, (to_cart // n) % n, to_cart // (n * n)


def morton(i, j, k):
    r = 0
    for b in range(4):
        r |= ((i >> b) & 1) << (3 * b) | ((j >> b) & 1) << (3 * b + 1) | ((k >> b) & 1) << (3 * b + 2)
    return r


orders = {
    "caller (MFEM refinement order)": np.arange(Ne),
    "lexicographic": np.argsort(to_cart, kind="stable"),
    "4x4x4 blocks": np.argsort(((ck // 4) * 16 + (cj // 4) * 4 + (ci // 4)) * 64 + (ck % 4) * 16 + (cj % 4) * 4 + (ci % 4), kind="stable"),
    "Morton": np.argsort(np.array([morton(int(a), int(b), int(c)) for a, b, c in zip(ci, cj, ck)]), kind="stable"),
}


def first_touch(eorder, blk=64):
    perm = -np.ones(Nf, dtype=np.int64); nxt = 0
    for e0 in range(0, Ne, blk):
        es = eorder[e0:e0 + blk]
        for sl in range(6):
            for e in es:
                f = dofs[ptr[e] + sl]
                if perm[f] < 0: perm[f] = nxt; nxt += 1
    return perm


def trace(eorder):
    enew = np.empty(Ne, dtype=np.int64); enew[eorder] = np.arange(Ne)
    perm = first_touch(eorder)
    f2e = [[] for _ in range(Nf)]
    for e in range(Ne):
        for s in range(6): f2e[dofs[ptr[e] + s]].append((e, s))
    rows = [None] * Nf
    for f in range(Nf):
        acc = []
        for (e, s) in f2e[f]:
            opp = dofs[ptr[e] + (s ^ 1)]
            acc += [("x", perm[f]), ("x", perm[opp]), ("k", enew[e]), ("x", Nf + enew[e])]
        rows[perm[f]] = acc
    return rows


def misses(rows, cap_kb):
    cap = cap_kb * 1024 // 32
    lru = OrderedDict(); miss = 0
    for r, ac in enumerate(rows):
        for key in ac + [("y", r)]:
            if key in lru: lru.move_to_end(key)
            else:
                if key[0] != "y": miss += 1      # the written row allocates but is not a read miss
                lru[key] = 1
                if len(lru) > cap: lru.popitem(last=False)
    return miss


compulsory = Nf + 2 * Ne
for name, eo in orders.items():
    rows = trace(eo)
    print(f"{name:32s}", {c: round(misses(rows, c) / compulsory, 2) for c in (32, 64, 128, 256, 512)}, flush=True)
