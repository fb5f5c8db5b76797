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


# ---------------------------------------------------------------------------------------------------
# sharded schedule (cp_pfdr_graph_d1_b200/shard.py): what each rank of the multi-GPU path computes and
# exchanges, in numpy, with the communication abstracted (gloo in tests/test_sharding_gloo.py).
# ---------------------------------------------------------------------------------------------------
class LocalComm:
    """single-rank stand-in"""
    def allreduce(self, a):
        return a

    def exchange(self, sends, recv_counts):
        return {}


def _pos_of_keys(g, E):
    """position (in the (owner, key)-sorted half-edge list) of every half-edge key 2e+side"""
    order_key = 2 * g["edge"] + (g["own"] != g["_eu"][g["edge"]]).astype(np.int64)
    pos = np.empty(2 * E, np.int64)
    pos[order_key] = np.arange(2 * E)
    return pos


def quad_n0_sharded(plan, comm, X0, Y, A, La_d1, rho, difTol, itMax, box=None):
    """N == 0, no l1, no reconditioning; arrays are GLOBAL, every rank slices its own part through the plan.
    Returns the rank's owned part of the iterate and the iteration count."""
    dt = X0.dtype.type
    Vown, Vloc = plan.Vown, plan.Vloc
    x = plan.vertex_array(X0).copy()
    y = plan.vertex_array(Y)
    a = None if A is None else plan.vertex_array(A)
    lad1 = plan.edge_array(La_d1)
    E = plan.Eu.shape[0]
    g = half_edges(Vloc, plan.Eu.astype(np.int64), plan.Ev.astype(np.int64))
    # self-loops have own == Eu == Ev: the side cannot be recovered from the owner, recompute keys explicitly
    h = np.arange(2 * E)
    own_all = np.where(h & 1, plan.Ev[h >> 1], plan.Eu[h >> 1]).astype(np.int64)
    order = np.lexsort((h, own_all))
    pos = np.empty(2 * E, np.int64)
    pos[order] = np.arange(2 * E)
    own, nbr, twin, edge = g["own"], g["nbr"], g["twin"], g["edge"]
    mine = own < Vown                                   # half-edges this rank computes
    ga0 = np.ones(Vloc, x.dtype) if a is None else a.copy()
    aux = np.where(ga0 > 0, y / ga0, 0)[:Vown]
    nz = aux != 0
    s = comm.allreduce(np.array([np.abs(aux[nz]).sum(dtype=np.float64), float(nz.sum())]))
    c = dt(dt(s[1]) / dt(s[0]))
    w = (c * lad1[edge]).astype(x.dtype)
    S = np.zeros(Vloc, x.dtype)
    np.add.at(S, own[mine], w[mine])
    ga = np.zeros(Vloc, x.dtype)
    ga[:Vown] = dt(1) / (ga0[:Vown] + S[:Vown])
    bound = dt(1.9) * (dt(2) - dt(rho))
    ga[:Vown] = np.minimum(ga[:Vown], bound) if a is None else np.where(
        a[:Vown] > 0, np.minimum(ga[:Vown], bound / np.where(a[:Vown] > 0, a[:Vown], 1)), ga[:Vown])
    T = ga * S
    Sinv = np.zeros(Vloc, x.dtype)
    Sinv[:Vown] = dt(1) / S[:Vown]
    grad = lambda xx: ((xx if a is None else a[:Vown] * xx) - y[:Vown])
    P = np.zeros(Vloc, x.dtype)
    P[:Vown] = 2 * x[:Vown] - ga[:Vown] * grad(x[:Vown])

    def halo(vertex_arrays, Z):
        sends = {nb.rank: np.concatenate([va[nb.send_vertices] for va in vertex_arrays] +
                                         ([Z[pos[nb.send_halfedges]]] if Z is not None else []))
                 for nb in plan.neighbors}
        nva = len(vertex_arrays)
        counts = {nb.rank: nva * nb.recv_vertices.shape[0] + (nb.recv_halfedges.shape[0] if Z is not None else 0)
                  for nb in plan.neighbors}
        recvs = comm.exchange(sends, counts)
        for nb in plan.neighbors:
            buf, o = recvs[nb.rank], 0
            for va in vertex_arrays:
                va[nb.recv_vertices] = buf[o:o + nb.recv_vertices.shape[0]]
                o += nb.recv_vertices.shape[0]
            if Z is not None:
                Z[pos[nb.recv_halfedges]] = buf[o:]

    halo([T, P], None)
    Z = x[own].copy()
    eps = difTol if 0 < difTol < np.finfo(x.dtype).eps else np.finfo(x.dtype).eps
    it, dif = 0, difTol ** 2
    thc = dt(1) / c
    while not (it == itMax or dif < difTol ** 2):
        tsum = T[own] + T[nbr]
        th = thc * tsum
        wo, wt = T[nbr] / tsum, T[own] / tsum
        po, pt = P[own] - Z, P[nbr] - Z[twin]
        aa = wo * po + wt * pt
        b = po - pt
        b = np.where(b > th, b - th, np.where(b < -th, b + th, 0))
        Zn = Z + dt(rho) * (aa + wt * b - x[own])
        Zn[~mine] = 0                                   # not ours: arrives with the halo
        xn = np.zeros(Vloc, x.dtype)
        np.add.at(xn, own[mine], (w * Sinv[own] * Zn)[mine])
        if box is not None:
            xn[:Vown] = np.clip(xn[:Vown], box[0], box[1])
        sums = comm.allreduce(np.array([np.sum((x[:Vown] - xn[:Vown]).astype(np.float64) ** 2),
                                        np.sum(xn[:Vown].astype(np.float64) ** 2)]))
        dif = sums[0] / sums[1] if sums[1] > eps else sums[0] / eps
        x[:Vown] = xn[:Vown]
        P[:Vown] = 2 * x[:Vown] - ga[:Vown] * grad(x[:Vown])
        Z = Zn
        halo([P], Z)
        it += 1
    return x[:Vown], it
