"""TEST INFRASTRUCTURE: numpy model of the *schedule* the CUDA kernels use (half-edge rows keyed by
vertex, edge constants recomputed from per-vertex S and T = Ga*S, each half-edge updating only its own
edge copy from ping buffers).  It lets the CPU suite check that this reformulation reproduces the
reference iteration (cp_pfdr_graph_d1_b200/csrc/quad.cu header comment) without a GPU."""
import numpy as np


def half_edges(V, Eu, Ev):
    """half-edge h = 2e+side owned by (Eu|Ev)[e]; returns arrays in (owner, h) order + twin positions"""
    E = Eu.shape[0]
    h = np.arange(2 * E)
    own = np.where(h & 1, Ev[h >> 1], Eu[h >> 1])
    nbr = np.where(h & 1, Eu[h >> 1], Ev[h >> 1])
    order = np.lexsort((h, own))                 # stable: by owner, then half-edge id
    pos = np.empty(2 * E, np.int64)
    pos[order] = np.arange(2 * E)
    return dict(own=own[order], nbr=nbr[order], edge=(h >> 1)[order], twin=pos[h[order] ^ 1])


def quad_n0(V, X, Y, A, Eu, Ev, La_d1, La_l1, positivity, box, Ltype, L, rho, condMin, difTol, itMax):
    """N == 0 quadratic family without reconditioning, as quad.cu schedules it."""
    dt = X.dtype.type
    g = half_edges(V, Eu, Ev)
    own, nbr, twin = g["own"], g["nbr"], g["twin"]
    ga0 = np.ones(V, X.dtype) if A is None else A.copy()
    aux = np.where(ga0 > 0, Y / ga0, 0)
    nz = aux != 0
    c = dt(dt(nz.sum()) / dt(np.abs(aux[nz]).sum()))
    w = (c * La_d1[g["edge"]]).astype(X.dtype)
    S = np.zeros(V, X.dtype)
    np.add.at(S, own, w)
    ga = ga0 + S
    if La_l1 is not None:
        ga = ga + c * La_l1
    ga = dt(1) / ga
    bound = dt(1.9) * (dt(2) - dt(rho))
    if Ltype == 0 or L is None:
        ga = np.minimum(ga, bound / (L[0] if L is not None else dt(1)))
    else:
        ga = np.where(L > 0, np.minimum(ga, bound / np.where(L > 0, L, 1)), ga)
    T, Sinv, thc = ga * S, dt(1) / S, dt(1) / c
    grad = lambda x: (x if A is None else A * x) - Y
    x = X.copy()
    P = 2 * x - ga * grad(x)
    Z = x[own].copy()
    eps = difTol if 0 < difTol < np.finfo(X.dtype).eps else np.finfo(X.dtype).eps
    it, dif = 0, difTol ** 2
    while not (it == itMax or dif < difTol ** 2):
        tsum = T[own] + T[nbr]
        th = thc * tsum
        wo, wt = T[nbr] / tsum, T[own] / tsum
        po, pt = P[own] - Z, P[nbr] - Z[twin]
        a = wo * po + wt * pt
        b = po - pt
        b = np.where(b > th, b - th, np.where(b < -th, b + th, 0))
        Z = Z + dt(rho) * (a + wt * b - x[own])
        xn = np.zeros(V, X.dtype)
        np.add.at(xn, own, (w * Sinv[own]) * Z)
        if box is not None:
            xn = np.clip(xn, box[0], box[1])
        elif La_l1 is not None:
            th1 = ga * La_l1
            xn = np.where(xn > th1, xn - th1, np.where((not positivity) & (xn < -th1), xn + th1, 0))
        elif positivity:
            xn = np.maximum(xn, 0)
        d, cc = np.sum((x - xn) ** 2), np.sum(xn ** 2)
        dif = d / cc if cc > eps else d / eps
        x = xn
        P = 2 * x - ga * grad(x)
        it += 1
    return x, it
