"""TEST INFRASTRUCTURE: numpy model of the slab-sharded stencil schedule (cp_pfdr_graph_d1_b200/csrc/slab.cuh, dia.cuh):
what every rank holds (vertex arrays over [lo, hi), 2.ND planes of edge copies indexed by vertex), what it computes
(owned vertices only) and what it writes into a neighbour's ghost zone after each iteration (forward point of the
boundary vertices, its copies on the cut edges) - with the transport abstracted (gloo in tests/test_sharding_gloo.py).
Lets the CPU suite check the host logic of shard.SlabPlan and the exchange pattern without a GPU."""
import numpy as np


def stencil(plan):
    """offsets (ascending) and, per offset, the boolean mask over local vertices of `edge (v, v + d) exists` + its local edge id"""
    Eu, Ev = plan.Eu.astype(np.int64), plan.Ev.astype(np.int64)
    offs = np.unique(Ev - Eu)
    out, eid = [], []
    for d in offs:
        m = np.zeros(plan.Vloc, bool)
        e = np.full(plan.Vloc, -1, np.int64)
        sel = np.nonzero(Ev - Eu == d)[0]
        m[Eu[sel]] = True
        e[Eu[sel]] = sel
        out.append(m); eid.append(e)
    return [int(d) for d in offs], out, eid


def quad_n0_slab(plan, comm, X0, Y, A, La_d1, rho, difTol, itMax, box=None):
    """N == 0, no l1, no reconditioning; arrays are GLOBAL, the plan slices them.  comm: allreduce(array) and
    push(dict rank -> list of (name, plane, local_start_at_receiver, values)) delivering into the receiver's arrays."""
    dt = X0.dtype.type
    Vloc, b, e = plan.Vloc, plan.v0 - plan.lo, plan.v1 - plan.lo               # owned local range [b, e)
    own = np.zeros(Vloc, bool); own[b:e] = True
    x = plan.vertex_array(X0).astype(X0.dtype).copy()
    y = plan.vertex_array(Y)
    a = None if A is None else plan.vertex_array(A)
    lad1 = plan.edge_array(La_d1) if plan.edge_ids is not None else La_d1
    offs, mout, eid = stencil(plan)
    ND = len(offs)
    # half-edges of vertex v in the reference's summation order: in_{ND-1} .. in_0, out_0 .. out_{ND-1}
    halves = [(i, False) for i in range(ND - 1, -1, -1)] + [(i, True) for i in range(ND)]

    def exists(i, is_out):
        if is_out:
            return mout[i] & own
        m = np.zeros(Vloc, bool)
        m[offs[i]:] = mout[i][:Vloc - offs[i]]
        return m & own

    def nbr(i, is_out):
        v = np.arange(Vloc)
        return v + offs[i] if is_out else v - offs[i]

    def edge_of(i, is_out):
        v = np.arange(Vloc)
        return eid[i] if is_out else np.where(v - offs[i] >= 0, eid[i][np.maximum(v - offs[i], 0)], -1)

    ga0 = np.ones(Vloc, x.dtype) if a is None else a.copy()
    aux = np.where(ga0 > 0, y / ga0, 0)[b:e]
    nz = aux != 0
    s = comm.allreduce(np.array([np.abs(aux[nz]).sum(dtype=np.float64), float(nz.sum())]))
    c = dt(dt(s[1]) / dt(s[0]))
    S = np.zeros(Vloc, x.dtype)
    w = {}
    for h in halves:
        ex = exists(*h)
        wh = np.zeros(Vloc, x.dtype)
        wh[ex] = c * lad1[edge_of(*h)[ex]]
        w[h] = wh
        S = S + wh
    ga = np.zeros(Vloc, x.dtype)
    ga[b:e] = dt(1) / (ga0[b:e] + S[b:e])
    bound = dt(1.9) * (dt(2) - dt(rho))
    ga[b:e] = np.minimum(ga[b:e], bound) if a is None else np.where(a[b:e] > 0, np.minimum(ga[b:e], bound / np.where(a[b:e] > 0, a[b:e], 1)), ga[b:e])
    T = ga * S
    Sinv = np.zeros(Vloc, x.dtype); Sinv[b:e] = dt(1) / S[b:e]
    grad = lambda xx: ((xx if a is None else a[b:e] * xx) - y[b:e])
    P = np.zeros(Vloc, x.dtype)
    P[b:e] = 2 * x[b:e] - ga[b:e] * grad(x[b:e])
    # planes: Z[i] = copy of v on out_i(v), Z[ND + i] = copy of v on in_i(v)
    Z = np.zeros((2 * ND, Vloc), x.dtype)
    for (i, is_out) in halves:
        ex = exists(i, is_out)
        Z[i if is_out else ND + i][ex] = x[ex]
    arrays = {"T": T, "P": P, "Z": Z}

    def halo(names):
        """what slab_refresh / the iteration kernel's peer stores move: to the rank above the last `depth` owned vertices,
        to the rank below the first ones; copies on cut edges of offset d: only the d vertices next to the cut"""
        msgs = {}
        for side, r in ((0, plan.rank - 1), (1, plan.rank + 1)):
            if r < 0 or r >= plan.world:
                continue
            lo_r, hi_r, v0_r, v1_r = comm.ranges[r]
            depth = (v0_r - lo_r) if side else (hi_r - v1_r)
            first = (e - depth) if side else b
            shift = plan.lo - lo_r
            out = []
            for nme in names:
                if nme == "Z":
                    for i in range(ND):
                        d = min(offs[i], depth)
                        f = (e - d) if side else b
                        plane = i if side else ND + i
                        out.append(("Z", plane, f + shift, Z[plane][f:f + d].copy()))
                else:
                    out.append((nme, -1, first + shift, arrays[nme][first:first + depth].copy()))
            msgs[r] = out
        for nme, plane, start, vals in comm.push(msgs):
            (Z[plane] if nme == "Z" else arrays[nme])[start:start + vals.shape[0]] = vals

    halo(["T", "P", "Z"])
    eps = difTol if 0 < difTol < np.finfo(x.dtype).eps else np.finfo(x.dtype).eps
    it, dif = 0, difTol ** 2
    thc = dt(1) / c
    while not (it == itMax or dif < difTol ** 2):
        Zn = Z.copy()
        acc = np.zeros(Vloc, x.dtype)
        for (i, is_out) in halves:                                  # the reference's summation order
            ex = np.nonzero(exists(i, is_out))[0]
            n = nbr(i, is_out)[ex]
            po_pl, tw_pl = (i, ND + i) if is_out else (ND + i, i)
            tsum = T[ex] + T[n]
            th = thc * tsum
            wo, wt = T[n] / tsum, T[ex] / tsum
            po, pt = P[ex] - Z[po_pl][ex], P[n] - Z[tw_pl][n]
            d0 = po - pt
            bb = d0 - np.clip(d0, -th, th)
            zn = Z[po_pl][ex] + dt(rho) * ((pt + wo * d0) + wt * bb - x[ex])
            Zn[po_pl][ex] = zn
            acc[ex] = acc[ex] + (w[(i, is_out)][ex] * Sinv[ex]) * zn
        xn = acc
        if box is not None:
            xn[b:e] = np.clip(xn[b:e], box[0], box[1])
        sums = comm.allreduce(np.array([np.sum((x[b:e] - xn[b:e]).astype(np.float64) ** 2), np.sum(xn[b:e].astype(np.float64) ** 2)]))
        dif = sums[0] / sums[1] if sums[1] > eps else sums[0] / eps
        x[b:e] = xn[b:e]
        P[b:e] = 2 * x[b:e] - ga[b:e] * grad(x[b:e])
        Z[:] = Zn
        halo(["P", "Z"])
        it += 1
    return x[b:e], it
