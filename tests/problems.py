"""Seeded problem instances shared by the parity tests (CPU and GPU)."""
import zlib

import numpy as np

from cp_pfdr_graph_d1_b200 import synth


def random_graph(V, E, seed, self_loops=3, dups=5, hub=True):
    """irregular multigraph: random edges + duplicated edges + self-loops (as the cut-pursuit reduced
    problems contain, reference src/CP_PFDR_graph_quadratic_d1_l1.cpp:646-649) + one hub vertex;
    every vertex is covered by a ring so that no vertex is isolated."""
    rng = np.random.default_rng(seed)
    ring_u = np.arange(V)
    ring_v = (ring_u + 1) % V
    u = rng.integers(0, V, E)
    v = rng.integers(0, V, E)
    keep = u != v
    u, v = u[keep], v[keep]
    Eu = [ring_u, u, u[:dups]]
    Ev = [ring_v, v, v[:dups]]
    if self_loops:
        s = rng.integers(0, V, self_loops)
        Eu.append(s), Ev.append(s)
    if hub:
        t = rng.choice(np.arange(1, V), min(V - 1, 70), replace=False)
        Eu.append(np.zeros_like(t)), Ev.append(t)
    Eu, Ev = np.concatenate(Eu).astype(np.int32), np.concatenate(Ev).astype(np.int32)
    perm = rng.permutation(Eu.shape[0])
    return Eu[perm], Ev[perm]


def quad_case(name, dtype):
    """returns (family, args, kwargs) for Binder.PFDR_graph_quadratic_d1_{l1,bounds}; X starts at zero
    unless stated."""
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    kw = dict(rho=1.5, condMin=1e-3, difRcd=0.0, difTol=1e-5, itMax=300, want_obj=True, want_dif=True)
    if name.startswith("grid"):
        n = 13
        w = synth.c2_denoise(n, dtype, diag=("diag" in name))
        V, A = w["V"], w["La_l2"]
        Y = w["Y"] if A is None else (A * w["Y"]).astype(dtype)
        Ltype, L = (1, A) if A is not None else (0, None)
        X = np.zeros(V, dtype)
        if "warm" in name:
            X = rng.uniform(0, 3, V).astype(dtype)
        if "box" in name:
            lo, hi = (0.5, 2.5) if "box2" in name else ((-np.inf, 2.0) if "boxhi" in name else (1.0, np.inf))
            return "bounds", (X, Y, A, w["Eu"], w["Ev"], w["La_d1"], lo, hi, Ltype, L), dict(kw, N=0)
        La_l1 = np.full(V, 0.05, dtype) if "l1" in name else None
        pos = int("pos" in name)
        if "l1" in name and not pos:
            kw["want_obj"] = False          # reference Obj reads out of bounds for x<0
        if "recond" in name:
            kw.update(difRcd=1e-3, difTol=1e-6, itMax=500)
        if "notrack" in name:
            kw.update(difTol=0.0, itMax=37, want_dif=False)
        if "it0" in name:
            kw.update(itMax=0)
        if "rho1" in name:
            kw.update(rho=1.0)
        return "l1", (X, Y, A, w["Eu"], w["Ev"], w["La_d1"], La_l1, pos, Ltype, L), dict(kw, N=0)
    if name.startswith("rand"):
        V = 1000 if "big" in name else 77
        Eu, Ev = random_graph(V, 3 * V, 5)
        E = Eu.shape[0]
        A = rng.uniform(0.5, 1.5, V).astype(dtype)
        Y = (A * (np.round(rng.uniform(0, 3, V)) + rng.normal(0, 0.2, V))).astype(dtype)
        La_d1 = rng.uniform(0.05, 0.5, E).astype(dtype)
        if "recond" in name:
            kw.update(difRcd=1e-3, difTol=1e-6, itMax=500)
        return "l1", (np.zeros(V, dtype), Y, A, Eu, Ev, La_d1, np.full(V, 0.02, dtype), 1, 1, A), dict(kw, N=0)
    if name.startswith("dense"):
        n, N = (9, 40) if "odd" not in name else (7, 23)       # N=40: vector loads; N=23: scalar path
        w = synth.dense_problem(n, N, dtype)
        V = w["V"]
        A64 = w["A"].astype(np.float64)
        lam = np.linalg.eigvalsh(A64.T @ A64).max()
        if "ata" in name:
            A, Y, Nn = np.asfortranarray((A64.T @ A64).astype(dtype)), (A64.T @ w["Y"]).astype(dtype), -V
        else:
            A, Y, Nn = w["A"], w["Y"], N
        if "diagL" in name:
            Ltype, L = 1, np.full(V, lam, dtype)
        elif "noL" in name:
            Ltype, L = 0, None
            kw.update(itMax=60)
        else:
            Ltype, L = 0, np.array([lam], dtype)
        if "recond" in name:
            kw.update(difRcd=1e-3, difTol=1e-6, itMax=500)
        X = np.zeros(V, dtype)
        if "box" in name:
            return "bounds", (X, Y, A, w["Eu"], w["Ev"], w["La_d1"], 0.0, 1.5, Ltype, L), dict(kw, N=Nn)
        return "l1", (X, Y, A, w["Eu"], w["Ev"], w["La_d1"], np.full(V, 0.01, dtype), 1, Ltype, L), dict(kw, N=Nn)
    raise KeyError(name)


QUAD_CASES = ["grid_identity", "grid_diag", "grid_diag_l1_pos", "grid_diag_l1", "grid_diag_pos", "grid_diag_box2",
              "grid_diag_boxhi", "grid_diag_boxlo", "grid_diag_recond", "grid_diag_l1_pos_recond", "grid_diag_warm",
              "grid_diag_notrack", "grid_diag_it0", "grid_identity_rho1", "rand_small", "rand_big", "rand_big_recond",
              "dense_direct", "dense_direct_odd", "dense_direct_box", "dense_direct_diagL", "dense_direct_noL",
              "dense_direct_recond", "dense_ata", "dense_ata_odd", "dense_ata_box", "dense_ata_recond"]


def irregular_case(V, kind, dt=np.float64):
    """multigraph with self-loops, duplicate edges and (V > 8) a hub; diagonal, dense N x V or Gram operator; l1 + positivity"""
    rng = np.random.default_rng(V * 7 + len(kind))
    Eu, Ev = random_graph(V, 2 * V, V, self_loops=min(3, V), dups=min(5, V), hub=V > 8)
    E = Eu.shape[0]
    if kind == "diag":
        A = rng.uniform(0.5, 1.5, V); Y = A * np.round(rng.uniform(0, 3, V)); N, Lt, L = 0, 1, A
    else:
        M = max(3, V // 3)
        B = rng.normal(0, 1, (M, V))
        lam = float(np.linalg.eigvalsh(B.T @ B).max())
        x0 = np.round(rng.uniform(0, 2, V))
        if kind == "dense":
            A, Y, N = np.asfortranarray(B), B @ x0, M
        else:
            A, Y, N = np.asfortranarray(B.T @ B), B.T @ (B @ x0), -V
        Lt, L = 0, np.array([lam])
    La_d1 = rng.uniform(0.05, 0.5, E)
    args = (np.zeros(V, dt), Y.astype(dt), A.astype(dt), Eu, Ev, La_d1.astype(dt), np.full(V, 0.01, dt), 1, Lt, L.astype(dt))
    kw = dict(N=N, rho=1.5, condMin=1e-3, difRcd=0.0, difTol=1e-7, itMax=400, want_obj=False, want_dif=True)
    return "l1", args, kw


def run_quad(binder, name, dtype, **override):
    fam, args, kw = quad_case(name, dtype)
    kw.update(override)
    args = tuple(a.copy() if isinstance(a, np.ndarray) else a for a in args)
    f = binder.PFDR_graph_quadratic_d1_l1 if fam == "l1" else binder.PFDR_graph_quadratic_d1_bounds
    return f(*args, **kw)


SIMPLEX_CASES = ["kla", "kla_weighted", "kla_recond", "kla_labels", "kla_k16", "kla_k3", "kla_k1", "kla_k32",
                 "kla_k12", "kla_k64", "kla_recond_k16", "kla_labels_k16", "kla_weighted_k16", "lin_k16", "lin_recond_k16",
                 "quad_k16", "quad_recond_k16", "quad_weighted_k16", "kla_rand_k16",
                 "lin", "lin_recond", "quad", "quad_weighted", "quad_recond", "kla_rand", "kla_it0", "kla_notrack",
                 # more than 32 labels, not a multiple of 4: rows padded to 4 labels per lane on the device
                 "kla_k33", "kla_recond_k33", "kla_labels_k33", "lin_k33", "quad_weighted_k50", "kla_rand_k50", "kla_k127",
                 # more than 128 labels: 8 / 16 labels per lane
                 "kla_k200", "kla_recond_k200", "quad_weighted_k300", "lin_k511"]


def simplex_case(name, dtype):
    K = 5
    for tag, k in (("k16", 16), ("k3", 3), ("k1", 1), ("k32", 32), ("k12", 12), ("k64", 64), ("k33", 33), ("k50", 50), ("k127", 127),
                   ("k200", 200), ("k300", 300), ("k511", 511)):
        if name.endswith("_" + tag) or ("_" + tag + "_") in name:
            K = k
    al = 0.0 if name.startswith("lin") else (1.0 if name.startswith("quad") else 0.05)
    kw = dict(rho=1.5, condMin=1e-2, difRcd=0.0, difTol=1e-4, itMax=150, want_obj=True, want_dif=True)
    if "rand" in name:
        V = 300
        Eu, Ev = random_graph(V, 3 * V, 8)
        rng = np.random.default_rng(21)
        Q = rng.uniform(0, 1, (V, K))
        Q[np.arange(V), rng.integers(0, K, V)] += 1.5
        Q = (Q / Q.sum(axis=1, keepdims=True)).astype(dtype)
        La_d1 = rng.uniform(0.05, 0.3, Eu.shape[0]).astype(dtype)
    else:
        w = synth.c4_simplex(9 if K <= 5 else (7 if K <= 16 else 5), K, dtype)
        V, Eu, Ev, Q, La_d1 = w["V"], w["Eu"], w["Ev"], w["Q"], w["La_d1"]
    if "weighted" in name:
        kw["La_f"] = np.random.default_rng(9).uniform(0.5, 2.0, V).astype(dtype)
    if "recond" in name:
        kw.update(difRcd=1e-2, difTol=1e-5, itMax=300)
    if "labels" in name:
        kw.update(difTol=1.0, itMax=60)
    if "it0" in name:
        kw.update(itMax=0)
    if "notrack" in name:
        kw.update(difTol=0.0, itMax=25, want_dif=False)
    return (Q.copy(), Q, al, Eu, Ev, La_d1), kw


def run_simplex(binder, name, dtype):
    args, kw = simplex_case(name, dtype)
    return binder.PFDR_graph_loss_d1_simplex(*args, **kw)
