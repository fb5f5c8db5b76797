"""Seeded synthetic workloads of BASELINE.json / SURVEY.md section 8(d) (host generation, numpy)."""
import numpy as np


def grid_graph(n, nz=None):
    """3D 6-connected grid, v = (z*n + y)*n + x; edges emitted vertex-major in the order +x, +y, +z."""
    nz = n if nz is None else nz
    idx = np.arange(n * n * nz, dtype=np.int64).reshape(nz, n, n)
    cand_u = np.stack([idx, idx, idx], axis=-1)                                   # (z, y, x, dir)
    cand_v = np.stack([idx + 1, idx + n, idx + n * n], axis=-1)
    ok = np.zeros(cand_u.shape, bool)
    ok[:, :, :-1, 0] = True
    ok[:, :-1, :, 1] = True
    ok[:-1, :, :, 2] = True
    Eu = cand_u[ok].astype(np.int32)
    Ev = cand_v[ok].astype(np.int32)
    return n * n * nz, Eu, Ev


def c2_denoise(n=100, dtype=np.float32, diag=True, seed=2, const_weights=False):
    """C2: PFDR_graph_l22_d1_l1 denoising on the n^3 grid (4 constant z-slabs, levels 0,1,3,2, + U(-.5,.5))."""
    V, Eu, Ev = grid_graph(n)
    z = np.arange(V) // (n * n)
    levels = np.array([0.0, 1.0, 3.0, 2.0])
    Y = levels[np.minimum(4 * z // n, 3)] + np.random.default_rng(seed).uniform(-0.5, 0.5, V)
    La_d1 = np.full(Eu.shape[0], 0.5) if const_weights else np.random.default_rng(seed + 1).uniform(0.25, 0.75, Eu.shape[0])
    La_l2 = np.random.default_rng(seed + 2).uniform(0.5, 1.5, V) if diag else None
    return dict(V=V, E=Eu.shape[0], Y=Y.astype(dtype), La_l2=None if La_l2 is None else La_l2.astype(dtype),
                Eu=Eu, Ev=Ev, La_d1=La_d1.astype(dtype))


def c4_simplex(n=64, K=16, dtype=np.float32, seed=6, nz=None):
    """C4: K-label smoothed-KL problem on the n^3 grid; truth label = 4*floor(4z/n) + floor(4y/n).
    nz < n: only the first nz z-planes of that problem (a bounded sample for the CPU legs of bench.py)."""
    V, Eu, Ev = grid_graph(n, nz)
    v = np.arange(V)
    z, y = v // (n * n), (v // n) % n
    lab = (4 * np.minimum(4 * z // n, 3) + np.minimum(4 * y // n, 3)) % K
    Q = np.random.default_rng(seed).uniform(0.0, 1.0, (V, K))
    Q[v, lab] += 2.0
    Q /= Q.sum(axis=1, keepdims=True)
    return dict(V=V, E=Eu.shape[0], K=K, Q=Q.astype(dtype), Eu=Eu, Ev=Ev,
                La_d1=np.full(Eu.shape[0], 0.2, dtype), al=0.05, labels=lab)


def knn_graph(V=32768, k=10, seed=4):
    """C3 graph: symmetrised, de-duplicated k-NN edges (u < v) of V uniform points in the unit cube."""
    from scipy.spatial import cKDTree
    pts = np.random.default_rng(seed).uniform(0.0, 1.0, (V, 3))
    _, nn = cKDTree(pts).query(pts, k=k + 1)
    u = np.repeat(np.arange(V), k)
    w = nn[:, 1:].reshape(-1)
    a, b = np.minimum(u, w), np.maximum(u, w)
    key = np.unique(a.astype(np.int64) * V + b)
    return pts, (key // V).astype(np.int32), (key % V).astype(np.int32)


def c3_bounds_ata(V=32768, k=10, m=2048, dtype=np.float32, seed=4):
    """C3: box-constrained problem with dense precomputed A^tA = G^tG/m + 0.1 I on a k-NN graph."""
    pts, Eu, Ev = knn_graph(V, k, seed)
    G = np.random.default_rng(seed + 1).standard_normal((m, V)).astype(np.float32)
    AtA = (G.T @ G) / np.float32(m)
    AtA[np.arange(V), np.arange(V)] += np.float32(0.1)
    x0 = np.round(4 * pts[:, 0]) / 4
    AtY = AtA.astype(np.float64) @ x0 + 0.01 * np.random.default_rng(seed + 2).standard_normal(V)
    return dict(V=V, E=Eu.shape[0], AtA=np.asfortranarray(AtA.astype(dtype)), AtY=AtY.astype(dtype), Eu=Eu, Ev=Ev,
                La_d1=np.full(Eu.shape[0], 0.02, dtype), x0=x0)


def dense_problem(n=16, N=64, dtype=np.float32, seed=7):
    """C5 in miniature: dense N-by-V lead field on the n^3 grid, two boxes of value 2, Y = A x0."""
    V, Eu, Ev = grid_graph(n)
    rng = np.random.default_rng(seed)
    A = np.asfortranarray(rng.uniform(-1 / 16, 1 / 16, (N, V)).astype(dtype))
    v = np.arange(V)
    z, y, x = v // (n * n), (v // n) % n, v % n
    x0 = np.zeros(V)
    x0[(z < n // 3) & (y < n // 2) & (x < n // 2)] = 2.0
    x0[(z >= 2 * n // 3) & (y >= n // 2)] = 2.0
    Y = (A.astype(np.float64) @ x0).astype(dtype)
    return dict(V=V, E=Eu.shape[0], N=N, A=A, Y=Y, Eu=Eu, Ev=Ev, La_d1=np.full(Eu.shape[0], 0.05, dtype), x0=x0)


def icosphere(level=5):
    """triangulated unit sphere: vertices (V, 3) and undirected mesh edges (u < v); V = 10*4^level + 2"""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    verts = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
             (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
             (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
             (8, 6, 7), (9, 8, 1)]
    verts = [np.array(v, float) / np.linalg.norm(v) for v in verts]
    for _ in range(level):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = verts[a] + verts[b]
                verts.append(m / np.linalg.norm(m))
                cache[key] = len(verts) - 1
            return cache[key]
        for a, b, c in faces:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        faces = nf
    F = np.array(faces)
    e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]])
    e = np.unique(np.sort(e, axis=1), axis=0)
    return np.array(verts), e[:, 0].astype(np.int32), e[:, 1].astype(np.int32)


def eeg_like(level=5, N=91, dtype=np.float64, seed=1):
    """C1 stand-in (data/EEG.mat is absent from the reference checkout): icosphere cortex,
    N sensors on a concentric sphere with a smooth dipole-like lead field, sparse positive piecewise-constant
    sources, noisy observations; parameters of octave/example_EEG_CP.m:20-27 apply."""
    rng = np.random.default_rng(seed)
    P, Eu, Ev = icosphere(level)
    V = P.shape[0]
    S = rng.standard_normal((N, 3))
    S = 1.25 * S / np.linalg.norm(S, axis=1, keepdims=True)
    D = S[:, None, :] - P[None, :, :]
    A = np.einsum("nvk,vk->nv", D, P) / np.linalg.norm(D, axis=2) ** 3          # radial dipole potential
    A /= np.linalg.norm(A, axis=0, keepdims=True).mean()
    x0 = np.zeros(V)
    for c, val in zip(rng.choice(V, 4, replace=False), (1.0, 1.5, 2.0, 1.0)):
        x0[np.arccos(np.clip(P @ P[c], -1, 1)) < 0.25] = val
    y = A @ x0
    y = y + 0.02 * np.abs(y).max() * rng.standard_normal(N)
    return dict(V=V, E=Eu.shape[0], N=N, A=np.asfortranarray(A.astype(dtype)), Y=y.astype(dtype), Eu=Eu, Ev=Ev,
                La_d1=np.full(Eu.shape[0], 0.02, dtype), La_l1=np.full(V, 0.005, dtype), x0=x0, points=P)
