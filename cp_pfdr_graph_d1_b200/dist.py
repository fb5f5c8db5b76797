"""Multi-GPU (one process per GPU) front end of the vertex-sharded solvers.

torch.distributed is the plumbing (rendezvous, broadcasting the NCCL id, gathering results); the data path -
halo exchange of boundary forward points / cut-edge copies and the all-reduce of the stopping scalars - runs
inside libpfdr_b200.so on its own NCCL communicator (csrc/comm.cu).  shard.py builds each rank's plan."""
import ctypes as C

import numpy as np

from . import _lib, shard
from ._abi import _CT, PfdrError
from .api import _Solver, last_stats


class _Shard(C.Structure):
    _fields_ = [("Vown", C.c_int), ("Vglobal", C.c_int), ("nnb", C.c_int)] + \
               [(n, C.POINTER(C.c_int)) for n in ("nb_rank", "sv_off", "sv", "sh_off", "sh", "rv_off", "rv", "rh_off", "rh")]


def init_comm(device_index=None):
    """creates the library's NCCL communicator over the ranks of the default torch.distributed group"""
    import torch
    import torch.distributed as dist
    lib = _lib.load()
    world, rank = dist.get_world_size(), dist.get_rank()
    if device_index is not None:
        lib.pfdr_b200_set_device(int(device_index))
    ident = (C.c_ubyte * 128)()
    if rank == 0:
        if lib.pfdr_b200_comm_unique_id(ident) != 0:
            raise PfdrError(lib.pfdr_b200_last_error().decode())
    box = [bytes(ident)]
    dist.broadcast_object_list(box, src=0)
    buf = (C.c_ubyte * 128).from_buffer_copy(box[0])
    if lib.pfdr_b200_comm_init(buf, world, rank) != 0:
        raise PfdrError(lib.pfdr_b200_last_error().decode())
    torch.cuda.synchronize()
    return world, rank


def destroy_comm():
    _lib.load().pfdr_b200_comm_destroy()


def _shard_struct(plan, keep):
    def cat(parts):
        off = np.zeros(len(parts) + 1, np.int32)
        off[1:] = np.cumsum([p.shape[0] for p in parts])
        data = np.ascontiguousarray(np.concatenate(parts) if parts else np.zeros(0), dtype=np.int32)
        if data.shape[0] == 0:
            data = np.zeros(1, np.int32)
        keep.extend([off, data])
        return off.ctypes.data_as(C.POINTER(C.c_int)), data.ctypes.data_as(C.POINTER(C.c_int))
    nb = plan.neighbors
    ranks = np.array([n.rank for n in nb] or [0], np.int32)
    keep.append(ranks)
    s = _Shard(plan.Vown, plan.V_global, len(nb), ranks.ctypes.data_as(C.POINTER(C.c_int)))
    s.sv_off, s.sv = cat([n.send_vertices for n in nb])
    s.sh_off, s.sh = cat([n.send_halfedges for n in nb])
    s.rv_off, s.rv = cat([n.recv_vertices for n in nb])
    s.rh_off, s.rh = cat([n.recv_halfedges for n in nb])
    return s


class ShardedQuadraticSolver(_Solver):
    """PFDR_graph_quadratic_d1_{l1,bounds} on this rank's shard: A is None / 1-D (N = 0) or a dense 2-D
    N-by-V matrix (N > 0, this rank keeps its owned columns; Y then has length N and is not sliced).  Arrays are the
    GLOBAL host arrays and the plan slices them; pass `local=True` when they are already local (dense A: owned
    columns only; arrays may then also be torch CUDA tensors)."""

    def __init__(self, plan, X0, Y, A, La_d1, La_l1=None, positivity=0, min=-np.inf, max=np.inf, prox="l1", Ltype=0,
                 L=None, rho=1.5, condMin=1e-3, difTol=1e-5, itMax=1000, want_dif=False, local=False, N=None):
        super().__init__()
        from ._abi import _real_dtype, _ptr
        dt = _real_dtype(X0)
        R, suf = _CT[dt]
        P, PI = C.POINTER(R), C.POINTER(C.c_int)
        dense = A is not None and len(A.shape) == 2
        if N is None:
            N = int(A.shape[0]) if dense else 0
        if local:
            X0l, Yl, Al, La_l1l, Ll, La_d1l = X0, Y, A, La_l1, L, La_d1
        else:
            va, ea = plan.vertex_array, plan.edge_array
            X0l, La_l1l = (None if a is None else np.ascontiguousarray(va(a), dtype=dt) for a in (X0, La_l1))
            if dense:
                Yl = np.ascontiguousarray(Y, dtype=dt)
                Al = np.asfortranarray(np.asarray(A)[:, plan.v0:plan.v1], dtype=dt)
            else:
                Yl = np.ascontiguousarray(va(Y), dtype=dt)
                Al = None if A is None else np.ascontiguousarray(va(A), dtype=dt)
            Ll = None if L is None else (np.ascontiguousarray(L, dtype=dt) if Ltype == 0 else np.ascontiguousarray(va(L), dtype=dt))
            La_d1l = np.ascontiguousarray(ea(La_d1), dtype=dt)
        Eu, Ev = np.ascontiguousarray(plan.Eu, np.int32), np.ascontiguousarray(plan.Ev, np.int32)
        self._keep = [X0l, Yl, Al, La_l1l, Ll, La_d1l, Eu, Ev]
        self._dt, self._shape, self._itMax, self.plan = dt, (plan.Vown,), itMax, plan
        sh = _shard_struct(plan, self._keep)
        f = getattr(self._lib, f"pfdr_b200_solver_create_quadratic_sharded_{suf}")
        f.argtypes = [C.POINTER(C.c_void_p), C.POINTER(_Shard), C.c_int, C.c_int, C.c_int, C.c_int, P, P, P, PI, PI, P, P,
                      C.c_int, R, R, C.c_int, P, R, R, R, R, C.c_int, C.c_int, C.c_int]
        self._chk(f(C.byref(self._h), C.byref(sh), 0 if prox == "l1" else 1, plan.Vloc, int(Eu.shape[0]), int(N),
                    _ptr(X0l, R), _ptr(Yl, R), _ptr(Al, R), _ptr(Eu, C.c_int), _ptr(Ev, C.c_int), _ptr(La_d1l, R),
                    _ptr(La_l1l, R), int(positivity), float(min), float(max), int(Ltype), _ptr(Ll, R), rho, condMin, 0.0,
                    difTol, itMax, 0, int(want_dif)))
        self.setup_stats = last_stats()


class ShardedSimplexSolver(_Solver):
    """PFDR_graph_loss_d1_simplex on this rank's shard; P0, Q are GLOBAL (V, K) host arrays unless local=True."""

    def __init__(self, plan, P0, Q, al, La_d1, La_f=None, rho=1.5, condMin=1e-2, difTol=1e-4, itMax=500,
                 want_dif=False, local=False):
        super().__init__()
        dt = np.dtype(P0.dtype)
        R, suf = _CT[dt]
        P, PI = C.POINTER(R), C.POINTER(C.c_int)
        va = (lambda a: a) if local else plan.vertex_array
        ea = (lambda a: a) if local else plan.edge_array
        P0l, Ql, La_fl = (None if a is None else np.ascontiguousarray(va(a), dtype=dt) for a in (P0, Q, La_f))
        La_d1l = np.ascontiguousarray(ea(La_d1), dtype=dt)
        Eu, Ev = np.ascontiguousarray(plan.Eu, np.int32), np.ascontiguousarray(plan.Ev, np.int32)
        K = int(P0l.shape[1])
        self._keep = [P0l, Ql, La_fl, La_d1l, Eu, Ev]
        self._dt, self._shape, self._itMax, self.plan = dt, (plan.Vown, K), itMax, plan
        sh = _shard_struct(plan, self._keep)
        ptr = lambda a, t=R: None if a is None else a.ctypes.data_as(C.POINTER(t))
        f = getattr(self._lib, f"pfdr_b200_solver_create_simplex_sharded_{suf}")
        f.argtypes = [C.POINTER(C.c_void_p), C.POINTER(_Shard), C.c_int, C.c_int, C.c_int, R, P, P, P, PI, PI, P, R, R, R,
                      R, C.c_int, C.c_int, C.c_int]
        self._chk(f(C.byref(self._h), C.byref(sh), K, plan.Vloc, int(Eu.shape[0]), al, ptr(La_fl), ptr(P0l), ptr(Ql),
                    ptr(Eu, C.c_int), ptr(Ev, C.c_int), ptr(La_d1l), rho, condMin, 0.0, difTol, itMax, 0, int(want_dif)))
        self.setup_stats = last_stats()


class _Slab(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("world", "rank", "Vglobal", "v0", "v1", "lo", "hi")]


SLAB_BLOB = 256


def exchange_blobs(mine, world, rank):
    """all ranks' blobs, in rank order, over the default torch.distributed group (any backend)"""
    import torch
    import torch.distributed as dist
    box = [None] * world
    dist.all_gather_object(box, bytes(mine))
    return b"".join(box)


class _SlabSolver(_Solver):
    def _connect(self, plan, same_process=False, blobs=None):
        mine = (C.c_ubyte * SLAB_BLOB)()
        self._lib.pfdr_b200_slab_export.argtypes = [C.c_void_p, C.c_void_p]
        self._chk(self._lib.pfdr_b200_slab_export(self._h, mine))
        self.blob = bytes(mine)
        if blobs is None and plan.world > 1:
            blobs = exchange_blobs(mine, plan.world, plan.rank)
        if blobs is not None or plan.world == 1:
            self.connect(blobs if blobs is not None else self.blob, same_process)

    def connect(self, blobs, same_process=False):
        buf = (C.c_ubyte * len(blobs)).from_buffer_copy(blobs)
        self._lib.pfdr_b200_slab_connect.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        self._chk(self._lib.pfdr_b200_slab_connect(self._h, buf, int(same_process)))


class SlabSimplexSolver(_SlabSolver):
    """PFDR_graph_loss_d1_simplex on this rank's slab of a grid graph (shard.SlabPlan): halo exchange by direct stores
    into the neighbours' memory, no collective library.  P0, Q are GLOBAL (V, K) host arrays unless local=True (then
    rows [lo, hi), and La_d1 per local edge).  `connect=False` defers the (collective) attach: export `.blob`,
    gather all ranks' blobs yourself and call .connect(blobs)."""

    def __init__(self, plan, P0, Q, al, La_d1, La_f=None, rho=1.5, condMin=1e-2, difTol=1e-4, itMax=500,
                 want_dif=False, local=False, connect=True, same_process=False, difRcd=0.0, want_obj=False):
        super().__init__()
        dt = np.dtype(P0.dtype)
        R, suf = _CT[dt]
        P, PI = C.POINTER(R), C.POINTER(C.c_int)
        va = (lambda a: a) if local else plan.vertex_array
        ea = (lambda a: a) if local else plan.edge_array
        P0l, Ql, La_fl = (None if a is None else np.ascontiguousarray(va(a), dtype=dt) for a in (P0, Q, La_f))
        La_d1l = np.ascontiguousarray(ea(La_d1), dtype=dt)
        Eu, Ev = np.ascontiguousarray(plan.Eu, np.int32), np.ascontiguousarray(plan.Ev, np.int32)
        K = int(P0l.shape[1])
        assert P0l.shape[0] == plan.Vloc and La_d1l.shape[0] == Eu.shape[0]
        self._keep = [P0l, Ql, La_fl, La_d1l, Eu, Ev]
        self._dt, self._shape, self._itMax, self.plan = dt, (plan.Vown, K), itMax, plan
        sl = _Slab(plan.world, plan.rank, plan.V_global, plan.v0, plan.v1, plan.lo, plan.hi)
        ptr = lambda a, t=R: None if a is None else a.ctypes.data_as(C.POINTER(t))
        f = getattr(self._lib, f"pfdr_b200_solver_create_simplex_slab_{suf}")
        f.argtypes = [C.POINTER(C.c_void_p), C.POINTER(_Slab), C.c_int, C.c_int, C.c_int, R, P, P, P, PI, PI, P, R, R, R,
                      R, C.c_int, C.c_int, C.c_int]
        self._chk(f(C.byref(self._h), C.byref(sl), K, plan.Vloc, int(Eu.shape[0]), al, ptr(La_fl), ptr(P0l), ptr(Ql),
                    ptr(Eu, C.c_int), ptr(Ev, C.c_int), ptr(La_d1l), rho, condMin, difRcd, difTol, itMax, int(want_obj), int(want_dif)))
        self.setup_stats = last_stats()
        if connect:
            self._connect(plan, same_process)


class SlabQuadraticSolver(_SlabSolver):
    """PFDR_graph_quadratic_d1_{l1,bounds} on this rank's slab of a grid graph: A None / 1-D (N = 0) or dense N-by-V
    (N > 0: this rank keeps the columns of its owned vertices, Y has length N).  Arrays are GLOBAL host arrays unless
    local=True (vertex arrays then cover [lo, hi), a dense A the owned columns)."""

    def __init__(self, plan, X0, Y, A, La_d1, La_l1=None, positivity=0, min=-np.inf, max=np.inf, prox="l1", Ltype=0,
                 L=None, rho=1.5, condMin=1e-3, difTol=1e-5, itMax=1000, want_dif=False, local=False, N=None,
                 connect=True, same_process=False, difRcd=0.0, want_obj=False):
        super().__init__()
        from ._abi import _real_dtype, _ptr
        dt = _real_dtype(X0)
        R, suf = _CT[dt]
        P, PI = C.POINTER(R), C.POINTER(C.c_int)
        dense = A is not None and len(A.shape) == 2
        if N is None:
            N = int(A.shape[0]) if dense else 0
        if local:
            X0l, Yl, Al, La_l1l, Ll, La_d1l = X0, Y, A, La_l1, L, La_d1
        else:
            va, ea = plan.vertex_array, plan.edge_array
            X0l, La_l1l = (None if a is None else np.ascontiguousarray(va(a), dtype=dt) for a in (X0, La_l1))
            if dense:
                Yl = np.ascontiguousarray(Y, dtype=dt)
                Al = np.asfortranarray(np.asarray(A)[:, plan.v0:plan.v1], dtype=dt)
            else:
                Yl = np.ascontiguousarray(va(Y), dtype=dt)
                Al = None if A is None else np.ascontiguousarray(va(A), dtype=dt)
            Ll = None if L is None else (np.ascontiguousarray(L, dtype=dt) if Ltype == 0 else np.ascontiguousarray(va(L), dtype=dt))
            La_d1l = np.ascontiguousarray(ea(La_d1), dtype=dt)
        Eu, Ev = np.ascontiguousarray(plan.Eu, np.int32), np.ascontiguousarray(plan.Ev, np.int32)
        self._keep = [X0l, Yl, Al, La_l1l, Ll, La_d1l, Eu, Ev]
        self._dt, self._shape, self._itMax, self.plan = dt, (plan.Vown,), itMax, plan
        sl = _Slab(plan.world, plan.rank, plan.V_global, plan.v0, plan.v1, plan.lo, plan.hi)
        f = getattr(self._lib, f"pfdr_b200_solver_create_quadratic_slab_{suf}")
        f.argtypes = [C.POINTER(C.c_void_p), C.POINTER(_Slab), C.c_int, C.c_int, C.c_int, C.c_int, P, P, P, PI, PI, P, P,
                      C.c_int, R, R, C.c_int, P, R, R, R, R, C.c_int, C.c_int, C.c_int]
        self._chk(f(C.byref(self._h), C.byref(sl), 0 if prox == "l1" else 1, plan.Vloc, int(Eu.shape[0]), int(N),
                    _ptr(X0l, R), _ptr(Yl, R), _ptr(Al, R), _ptr(Eu, C.c_int), _ptr(Ev, C.c_int), _ptr(La_d1l, R),
                    _ptr(La_l1l, R), int(positivity), float(min), float(max), int(Ltype), _ptr(Ll, R), rho, condMin, difRcd,
                    difTol, itMax, int(want_obj), int(want_dif)))
        self.setup_stats = last_stats()
        if connect:
            self._connect(plan, same_process)


def gather_owned(local, plan):
    """all ranks' owned parts -> the global array (on every rank)"""
    import torch.distributed as dist
    parts = [None] * plan.world
    dist.all_gather_object(parts, (plan.v0, local))
    return np.concatenate([p[1] for p in sorted(parts, key=lambda t: t[0])], axis=0)
