#!/usr/bin/env python
"""bench.py - PFDR edge-updates/s on the BASELINE.json workloads (see DESIGN.md "Measurement").

Default workload = the configuration the north-star target is quoted on: C4, PFDR_graph_loss_d1_simplex with the
smoothed-KL loss, K = 16 labels, 252^3 grid (16.0M vertices / 47.8M edges), float32.

A *step* is one complete PFDR solve of `iters` iterations (difTol = 0, so exactly itMax iterations run)
on the resident synthetic problem: solver.reset() (initial iterate + first preconditioning, device only)
followed by solver.run().  `value` = E * iters * steps / time with every input already in HBM;
`e2e` = the same metric through the one-shot C-ABI call with HOST buffers (upload, layout, preconditioning,
iterations, download of the iterate all inside the timed region): pinned buffers (the headline) and pageable ones
(what the MEX / cut-pursuit callers hand over).  `extra` = short resident runs of the other BASELINE configs,
`cp_solve` = BASELINE's second figure (cut-pursuit wall time, reference CP code on libpfdr_b200 vs all-reference),
`cp_outer` = the dense and graph passes of one cut-pursuit outer iteration on the device.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4_f32|c4_f64|c2_f32|c2_f64|...] [--impl reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def algorithmic_bytes(kind, V, E, r, K=1, diag=True, l1=False, N=0):
    """minimal bytes that must cross HBM once per iteration."""
    if kind == "quad":
        return E * (8 + 5 * r) + V * 5 * r + (V * r if diag else 0) + (V * r if l1 else 0)
    if kind == "ata":
        return E * (8 + 5 * r) + V * 5 * r + V * V * r + 2 * r * V
    if kind == "dense":
        return E * (8 + 5 * r) + V * 5 * r + 2 * N * V * r + 2 * r * V
    return E * (8 + r + 4 * r * K) + V * 5 * r * K


WORKLOADS = {
    # name: (kind, grid n, dtype, K, iterations per step)
    "c2_f32": ("quad", 100, np.float32, 1, 500),
    "c2_f64": ("quad", 100, np.float64, 1, 500),
    "c4_f32": ("simplex", 252, np.float32, 16, 100),
    "c4_f64": ("simplex", 252, np.float64, 16, 100),
    "c4s_f32": ("simplex", 64, np.float32, 16, 100),      # parity-sized C4 (also what the CPU legs of C4 run)
    "c4s_f64": ("simplex", 64, np.float64, 16, 100),
    # dense operators (device-generated): C3 direct PFDR with precomputed A^tA (V = 32768, k-NN graph),
    # and a single-GPU slice of C5 (dense N = 256 lead field on a 160^3 grid: A is 4.2 GB)
    "c3_f32": ("ata", 32768, np.float32, 1, 50),
    "c5s_f32": ("dense", 160, np.float32, 1, 50),
    # C5 at full size (sharded runs only: the 256 x 128M matrix is 131 GB) and a mid-size version for 2-GPU boxes
    "c5_f32": ("dense", 504, np.float32, 1, 20),
    "c5m_f32": ("dense", 252, np.float32, 1, 20),
}


def describe(name):
    kind, n, dt, K, iters = WORKLOADS[name]
    if kind == "quad":
        return (f"C2 PFDR_graph_l22_d1_l1 (diagonal A) denoising, 3D 6-connected grid {n}^3, "
                f"{np.dtype(dt).name}, rho=1.5, difTol=0, itMax={iters}")
    if kind == "ata":
        return (f"C3 PFDR_graph_quadratic_d1_bounds, dense precomputed AtA V={n}, box [0,1], 10-NN graph, "
                f"{np.dtype(dt).name}, SCAL L, rho=1.5, difTol=0, itMax={iters}")
    if kind == "dense":
        which = "C5" if n == 504 else "C5 (reduced grid)"
        return (f"{which} PFDR_graph_quadratic_d1_l1, dense A N=256 on a 3D grid {n}^3, {np.dtype(dt).name}, SCAL L, "
                f"rho=1, difTol=0, itMax={iters}")
    return (f"C4 PFDR_graph_loss_d1_simplex smoothed-KL (al=0.05), K={K}, 3D grid {n}^3, {np.dtype(dt).name}, "
            f"rho=1.5, condMin=1e-2, difTol=0, itMax={iters}")


def static_config(name, iters):
    """the `config` object: a function of the workload alone, so that both arms print the same one"""
    kind, n, dt, K, _ = WORKLOADS[name]
    if kind == "ata":
        V, E = n, None
    else:
        V, E = n ** 3, 3 * n * n * (n - 1)
    B = algorithmic_bytes(kind, V, E or 0, np.dtype(dt).itemsize, K, N=256 if kind == "dense" else 0)
    return {"workload": describe(name).replace(f"itMax={WORKLOADS[name][4]}", f"itMax={iters}"), "name": name, "V": V, "E": E, "K": K,
            "iterations_per_step": iters,
            "l2": f"no flush: one iteration streams {B / 1e6:.0f} MB algorithmic ({'>' if B > 126e6 else '<'} 126 MB L2); "
                  "steps re-run the same resident problem"}


# ------------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(name):
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(name)
    return None


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons every 50 ms; the summary is over the samples that arrived inside the
    timed region (mark_start() .. mark_end()).  nvidia-smi needs a few hundred ms to start: start() is called ahead of the
    warm-up and waits for the first sample, so that even a sub-second timed region is covered."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index, self.t0, self.t1 = [], None, index, None, None

    def start(self, wait_s=3.0):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t_end = time.perf_counter() + wait_s
            while not self.rows and time.perf_counter() < t_end:
                time.sleep(0.01)
        except OSError:
            self.proc = None
        return self

    def __enter__(self):                      # with-block = the timed region
        if self.proc is None:
            self.start()
        self.mark_start()
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def __exit__(self, *a):
        if self.t1 is None:
            self.mark_end()
        self.stop()

    def stop(self):
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
            self.t.join(timeout=2)
            self.proc = None

    def summary(self):
        t0, t1 = self.t0 or 0.0, self.t1 or float("inf")
        inside = [r for t, r in self.rows if t0 <= t <= t1 + 0.06]     # (+ one sampling period: the row reporting the last 50 ms)
        rows, where = (inside, "timed region") if inside else ([r for t, r in self.rows if t0 - 0.5 <= t <= t1 + 0.5], "within 0.5 s of the timed region")
        sm, mx, reasons = [], 0.0, set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm), "window": where, "timed_region_s": round(t1 - t0, 4) if self.t1 else None}


# ------------------------------------------------------------------------------------------------
def host_inputs(name):
    from cp_pfdr_graph_d1_b200 import synth
    kind, n, dt, K, iters = WORKLOADS[name]
    if kind in ("ata", "dense"):
        return None
    if kind == "quad":
        w = synth.c2_denoise(n, dt, diag=True)
        w["Ys"] = (w["La_l2"] * w["Y"]).astype(dt)          # the l22 gateway's pre-scaling (mex :54-60)
        return w
    return synth.c4_simplex(n, K, dt)


def cpu_sample_inputs(name):
    """bounded sample of the workload for the CPU legs: C2 whole; full-size C4: the first 16 z-planes of the same grid"""
    from cp_pfdr_graph_d1_b200 import synth
    kind, n, dt, K, iters = WORKLOADS[name]
    if kind == "simplex" and n > 128:
        return synth.c4_simplex(n, K, dt, nz=16), f"first 16 of {n} z-planes of the same problem ({n}x{n}x16 grid, {n * n * 16} vertices)"
    return host_inputs(name), "the whole problem"


def grid_slab_simplex(n, K, dt, world, rank):
    """rank's slab of the C4 grid (shard.SlabPlan: vertices [lo, hi) = owned planes + one ghost plane each side),
    generated directly, no global arrays: same recipe as synth.c4_simplex with a counter-based RNG per (vertex, label),
    so that the ranks agree on the rows they share"""
    from cp_pfdr_graph_d1_b200 import shard
    plan = shard.grid_slab_plan(n, world, rank)
    gl = np.arange(plan.lo, plan.hi, dtype=np.int64)
    zz, yy = gl // (n * n), (gl // n) % n
    lab = (4 * np.minimum(4 * zz // n, 3) + np.minimum(4 * yy // n, 3)) % K
    # counter-based uniform(0,1): splitmix64 of (vertex, label)
    ctr = (gl[:, None].astype(np.uint64) * np.uint64(K) + np.arange(K, dtype=np.uint64)[None, :]) + np.uint64(0x9E3779B97F4A7C15)
    zq = (ctr ^ (ctr >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    zq = (zq ^ (zq >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    zq = zq ^ (zq >> np.uint64(31))
    Q = (zq >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
    Q[np.arange(gl.shape[0]), lab] += 2.0
    Q /= Q.sum(axis=1, keepdims=True)
    E_global = 3 * n * n * (n - 1)
    return dict(plan=plan, V=n ** 3, E=E_global, Q=Q.astype(dt), La_d1=np.full(plan.Eu.shape[0], 0.2, dt))


def device_grid_edges(n, dev):
    """edge list of synth.grid_graph(n) built on the device (the full C4 graph has 48M edges)"""
    import torch
    V = n ** 3
    v = torch.arange(V, device=dev)
    z, y, x = v // (n * n), (v // n) % n, v % n
    eu, ev = [], []
    for ok, d in ((x < n - 1, 1), (y < n - 1, n), (z < n - 1, n * n)):
        eu.append(torch.where(ok, v, -1)); ev.append(torch.where(ok, v + d, -1))
    Eu = torch.stack(eu, 1).reshape(-1); Ev = torch.stack(ev, 1).reshape(-1)
    keep = Eu >= 0
    return Eu[keep].to(torch.int32).contiguous(), Ev[keep].to(torch.int32).contiguous()


OPNORM_C3_S = None


def torch_lambda_max(matvec, V, dev, tdt, iters=30):
    """largest eigenvalue of A^tA by power iteration in torch (bench input preparation only)"""
    import torch
    x = torch.randn(V, device=dev, dtype=tdt, generator=torch.Generator(device=dev).manual_seed(11))
    lam = 1.0
    for _ in range(iters):
        x = x / x.norm()
        x = matvec(x)
        lam = float(x.norm())
    return lam * 1.01


def device_inputs_dense(name, dev):
    """C3 / C5-slice inputs generated on the device (the matrices are several GB)"""
    import torch
    from cp_pfdr_graph_d1_b200 import synth
    kind, n, dt, K, iters = WORKLOADS[name]
    tdt = torch.float32 if dt is np.float32 else torch.float64
    g = torch.Generator(device=dev).manual_seed(5)
    if kind == "ata":
        V, m = n, 2048
        pts, Eu, Ev = synth.knn_graph(V, 10, 4)
        G = torch.randn((m, V), generator=g, device=dev, dtype=tdt)
        AtA = (G.t() @ G) / m
        AtA.diagonal().add_(0.1)
        del G
        x0 = torch.from_numpy(np.round(4 * pts[:, 0]) / 4).to(dev, tdt)
        AtY = AtA @ x0 + 0.01 * torch.randn(V, generator=g, device=dev, dtype=tdt)
        # SCAL Lipschitz constant from the library's own operator_norm_matrix (the routine cut-pursuit calls for it,
        # reference src/CP_PFDR_graph_quadratic_d1_bounds.cpp:769): the symmetric V x V system takes its batched path
        from cp_pfdr_graph_d1_b200 import api
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lam = 1.01 * api.operator_norm_matrix(AtA, 0, V, 1e-3, 100, 10)
        global OPNORM_C3_S
        OPNORM_C3_S = time.perf_counter() - t0
        return dict(V=V, E=int(Eu.shape[0]), N=-V, A=AtA.contiguous(), Y=AtY, Eu=to_dev(Eu, dev), Ev=to_dev(Ev, dev),
                    La_d1=torch.full((Eu.shape[0],), 0.02, device=dev, dtype=tdt), L=torch.tensor([lam], device=dev, dtype=tdt))
    V, Eu, Ev = synth.grid_graph(n)
    N = 256
    At = (torch.rand((V, N), generator=g, device=dev, dtype=tdt) - 0.5) / 8       # (V, N) row-major == N-by-V column-major
    v = torch.arange(V, device=dev)
    z, y, x = v // (n * n), (v // n) % n, v % n
    x0 = torch.zeros(V, device=dev, dtype=tdt)
    x0[(z < n // 3) & (y < n // 2) & (x < n // 2)] = 2.0
    x0[(z >= 2 * n // 3) & (y >= n // 2)] = 2.0
    Y = At.t() @ x0
    lam = torch_lambda_max(lambda t: At @ (At.t() @ t), V, dev, tdt, iters=15)
    return dict(V=V, E=int(Eu.shape[0]), N=N, A=At, Y=Y, Eu=to_dev(Eu, dev), Ev=to_dev(Ev, dev),
                La_d1=torch.full((Eu.shape[0],), 0.05, device=dev, dtype=tdt), L=torch.tensor([lam], device=dev, dtype=tdt))


def grid_slab_plan(n, world, rank):
    """shard plan of the n^3 grid for this rank, built from the edges touching its z-slab only"""
    from cp_pfdr_graph_d1_b200 import shard
    V = n ** 3
    b = shard.slab_bounds(V, world)
    v0, v1 = int(b[rank]), int(b[rank + 1])
    lo, hi = max(v0 - n * n, 0), min(v1 + n * n, V)
    g = np.arange(lo, hi, dtype=np.int64)
    z, y, x = g // (n * n), (g // n) % n, g % n
    eu, ev = [], []
    for ok, d in ((x < n - 1, 1), (y < n - 1, n), (z < n - 1, n * n)):
        eu.append(np.where(ok, g, -1)); ev.append(np.where(ok, g + d, -1))
    Eu, Ev = np.stack(eu, 1).reshape(-1), np.stack(ev, 1).reshape(-1)
    keep = (Eu >= 0) & (Ev < hi) & (((Eu >= v0) & (Eu < v1)) | ((Ev >= v0) & (Ev < v1)))
    return shard.make_plan_local(V, Eu[keep], Ev[keep], world, rank, b)


def device_slab_dense(n, N, dt, world, rank, dev):
    """this rank's part of C5: owned columns of the hashed lead field (generated on the device), Y = A x0 summed
    over the ranks, SCAL L from the Marchenko-Pastur edge of a matrix with iid U(-1/16, 1/16) entries"""
    import torch
    import torch.distributed as dist
    tdt = torch.float32 if dt is np.float32 else torch.float64
    plan = grid_slab_plan(n, world, rank)
    V, Vown = n ** 3, plan.Vown
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    At = (torch.rand((Vown, N), generator=g, device=dev, dtype=tdt) - 0.5) / 8     # (Vown, N) row-major == N x Vown col-major
    v = torch.arange(plan.v0, plan.v1, device=dev)
    z, y, x = v // (n * n), (v // n) % n, v % n
    x0 = torch.zeros(Vown, device=dev, dtype=tdt)
    x0[(z < n // 3) & (y < n // 2) & (x < n // 2)] = 2.0
    x0[(z >= 2 * n // 3) & (y >= n // 2)] = 2.0
    Y = (At.t() @ x0).contiguous()
    if world > 1:
        dist.all_reduce(Y)
    lam = 1.05 * (V ** 0.5 + N ** 0.5) ** 2 / 768.0
    E = 3 * n * n * (n - 1)
    return dict(plan=plan, V=V, E=E, N=N, A=At, Y=Y, X0=torch.zeros(plan.Vloc, device=dev, dtype=tdt),
                La_d1=torch.full((plan.Eu.shape[0],), 0.05, device=dev, dtype=tdt),
                L=torch.tensor([lam], device=dev, dtype=tdt))


def to_dev(a, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def make_solver(name, w, dev, itMax):
    """solver handle over device-resident inputs"""
    from cp_pfdr_graph_d1_b200 import api, DIAG
    kind, n, dt, K, iters = WORKLOADS[name]
    import torch
    if kind in ("ata", "dense"):
        d = dict(w)
        d["X0"] = torch.zeros(w["V"], device=dev, dtype=w["Y"].dtype)
        torch.cuda.synchronize()
        if kind == "ata":
            return api.QuadraticSolver(d["X0"], d["Y"], d["A"], d["Eu"], d["Ev"], d["La_d1"], min=0.0, max=1.0,
                                       prox="bounds", Ltype=0, L=d["L"], rho=1.5, condMin=1e-3, difRcd=0.0, difTol=0.0,
                                       itMax=itMax, N=w["N"]), d
        return api.QuadraticSolver(d["X0"], d["Y"], d["A"], d["Eu"], d["Ev"], d["La_d1"], None, 0, Ltype=0, L=d["L"],
                                   rho=1.0, condMin=1e-3, difRcd=0.0, difTol=0.0, itMax=itMax, N=w["N"]), d
    if kind == "quad":
        d = {k: to_dev(w[k], dev) for k in ("Ys", "La_l2", "Eu", "Ev", "La_d1")}
        d["X0"] = torch.zeros(w["V"], device=dev, dtype=d["Ys"].dtype)
        torch.cuda.synchronize()
        return api.QuadraticSolver(d["X0"], d["Ys"], d["La_l2"], d["Eu"], d["Ev"], d["La_d1"], None, 0, Ltype=DIAG,
                                   L=d["La_l2"], rho=1.5, condMin=1e-3, difRcd=0.0, difTol=0.0, itMax=itMax, N=0), d
    if isinstance(w["Q"], np.ndarray):
        d = {k: to_dev(w[k], dev) for k in ("Q", "Eu", "Ev", "La_d1")}
    else:
        d = {k: w[k] for k in ("Q", "Eu", "Ev", "La_d1")}
    torch.cuda.synchronize()
    return api.SimplexSolver(d["Q"], d["Q"], 0.05, d["Eu"], d["Ev"], d["La_d1"], None, rho=1.5, condMin=1e-2,
                             difRcd=0.0, difTol=0.0, itMax=itMax), d


def prep_host(name, bufs):
    """initial iterate of the in/out argument (X = 0; P = Q as octave/mex/PFDR_graph_loss_d1_simplex_mex.cpp:33)"""
    if WORKLOADS[name][0] == "quad":
        bufs["X"][:] = 0
    else:
        bufs["P"][:] = bufs["Q"]


def one_shot_host(name, w, pinned, itMax, backend=None, prepare=True):
    """the reference-facing call with HOST buffers (e2e leg and CPU legs); returns iterations done"""
    from cp_pfdr_graph_d1_b200 import api, DIAG
    b = backend or api._b()
    kind = WORKLOADS[name][0]
    if prepare:
        prep_host(name, pinned)
    if kind == "quad":
        r = b.PFDR_graph_quadratic_d1_l1(pinned["X"], pinned["Ys"], pinned["La_l2"], pinned["Eu"], pinned["Ev"],
                                         pinned["La_d1"], None, 0, DIAG, pinned["La_l2"], 1.5, 1e-3, 0.0, 0.0,
                                         itMax, 0)
    else:
        r = b.PFDR_graph_loss_d1_simplex(pinned["P"], pinned["Q"], 0.05, pinned["Eu"], pinned["Ev"], pinned["La_d1"],
                                         None, 1.5, 1e-2, 0.0, 0.0, itMax)
    return r.it


def pin(w, name, use_cuda=True):
    import torch
    kind = WORKLOADS[name][0]
    keys = ("Ys", "La_l2", "Eu", "Ev", "La_d1") if kind == "quad" else ("Q", "Eu", "Ev", "La_d1")
    out = {}
    for k in keys:
        t = torch.from_numpy(np.ascontiguousarray(w[k]))
        out[k] = (t.pin_memory() if use_cuda else t).numpy()
    xk, like = ("X", out["Ys"]) if kind == "quad" else ("P", out["Q"])
    t = torch.zeros(like.shape, dtype=torch.from_numpy(like).dtype)
    out[xk] = (t.pin_memory() if use_cuda else t).numpy()
    return out


def cpu_reference_rate(name, w, note, seconds_target=12.0):
    """edge-updates/s of the reference's own CPU implementation (oracle/_ref, all host cores) on a bounded
    sample of the workload: two-point timing (itMax = a, b) removes the setup.  Falls back to the serial port."""
    from oracle import oracle as O
    ref = O.reference()
    kind_label, be, cores = ("reference", ref, ref.lib.ref_num_procs()) if ref is not None else ("port", O.oracle(), 1)
    pinned = pin(w, name, use_cuda=False)
    E = int(w["E"])

    def run(itmax):
        t0 = time.perf_counter()
        one_shot_host(name, w, pinned, itmax, backend=be)
        return time.perf_counter() - t0

    a = 2
    ta = run(a)
    per_it = max(ta / (a + 1), 1e-4)
    b = int(max(a + 4, min(400, seconds_target / 2 / per_it)))
    tb = min(run(b), run(b))
    ta = min(ta, run(a))
    rate = E * (b - a) / max(tb - ta, 1e-9)
    return {"value": rate, "unit": "edge_updates/s", "cores": int(cores), "kind": kind_label,
            "sample": f"{describe(name).split(',')[0]}: {note}, iterations {a}..{b} (two-point timing, "
                      f"setup removed), {1e3 * (tb - ta) / (b - a):.2f} ms/iteration"}


# ------------------------------------------------------------------------------------------------
def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    kind, n, dt, K, iters = WORKLOADS[name]
    iters = args.iters or iters
    if kind in ("ata", "dense"):
        print(json.dumps({"impl": "reference", "unavailable": "device-generated multi-GB workload: CPU arm runs c2/c4 only"}), flush=True)
        return
    w, note = cpu_sample_inputs(name)
    per_step = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb = cpu_reference_rate(name, w, note, seconds_target=6.0)
        if i >= args.warmup:
            per_step.append(cb["value"])
    val = float(np.mean(per_step))
    cfg = static_config(name, iters)
    cb["value"] = val
    line = {"impl": "reference", "metric": "pfdr_edge_updates_per_s", "value": val, "unit": "edge_updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * cfg["E"] * iters / val, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32" if dt is np.float32 else "f64", "data": "synthetic",
            "config": cfg, "cpu_baseline": cb,
            "e2e": {"value": val, "unit": "edge_updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def resident_run(name, dev, stream, steps, warmup, iters=0, clock_index=None):
    """builds the workload in HBM and times `steps` solves; returns (line fragment, host inputs or None)"""
    import torch
    from cp_pfdr_graph_d1_b200 import api
    kind, n, dt, K, it0 = WORKLOADS[name]
    iters = iters or it0
    r = np.dtype(dt).itemsize
    w = device_inputs_dense(name, dev) if kind in ("ata", "dense") else host_inputs(name)
    V, E = int(w["V"]), int(w["E"])
    B_iter = algorithmic_bytes(kind, V, E, r, K, N=abs(int(w.get("N", 0))) if kind == "dense" else 0)
    solver, keep = make_solver(name, w, dev, iters)
    setup = dict(solver.setup_stats)

    def step():
        solver.reset()
        return solver.run()

    clk = ClockSampler(clock_index).start() if clock_index is not None else None
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    loop_ms = launches = its = 0
    if clk:
        clk.mark_start()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(steps):
            st = step()
            loop_ms += st["loop_ms"]; launches += st["launches"]; its += st["iterations"]
        ev1.record(stream)
    torch.cuda.synchronize()
    if clk:
        clk.mark_end()
        clk.stop()
    ms = ev0.elapsed_time(ev1)
    assert its == iters * steps, (its, iters, steps)
    solver.close()
    del keep
    torch.cuda.empty_cache()
    peak, peak_src = peaks()
    achieved = B_iter * its / (loop_ms * 1e-3) / 1e9
    return dict(V=V, E=E, K=K, iters=iters, ms=ms, loop_ms=loop_ms, launches=launches, its=its, value=E * its / (ms * 1e-3),
                B_iter=B_iter, achieved=achieved, peak=peak, peak_src=peak_src, setup=setup,
                clocks=clk.summary() if clk else None, kind=kind, dt=dt), w


KERNEL_NAMES = {"quad": "quad_dia_iter_kernel (grid graphs) / quad_iter_pipe_kernel", "simplex": "simplex_dia_iter_kernel (grid graphs) / simplex_iter_kernel",
                "ata": "ata_gemv_kernel", "dense": "ax_partial_kernel + atr_kernel"}


def cp_solve_time():
    """BASELINE.json's second figure: wall time of the reference's UNMODIFIED cut-pursuit (oracle/_ref, built from the
    reference's sources) with its PFDR / operator-norm calls served by libpfdr_b200 through the C++ shim, beside the
    all-reference build on the host cores; EEG-like stand-in of config 1 (data/EEG.mat is absent), float64"""
    from oracle import oracle as O
    from cp_pfdr_graph_d1_b200 import synth
    ref, b200 = O.cp_reference(), O.cp_b200()
    if ref is None or b200 is None:
        return None
    w = synth.eeg_like(5, 91, np.float64)
    kw = dict(CP_difTol=1e-4, CP_itMax=15, rho=1.5, condMin=1e-3, difRcd=0.0, difTol=1e-7, itMax=10000)
    a = (w["Y"], w["A"], w["Eu"], w["Ev"], w["La_d1"], w["La_l1"], 1, w["N"])
    # the reference seeds operator_norm_matrix's power method with time(NULL) + thread id; oracle/_ref runs with its clock
    # pinned and libpfdr_b200 (which draws the same starts, csrc/opnorm.cu) is given the same clock, so that the two
    # cut-pursuit runs take the same metric and their components can be compared one by one
    os.environ["PFDR_B200_OPNORM_TIME"] = "42"
    b200.quadratic_d1_l1(*a, **kw)                                 # warm-up: context, memory pool, module load
    t0 = time.perf_counter(); r1 = b200.quadratic_d1_l1(*a, **kw); t1 = time.perf_counter() - t0
    t0 = time.perf_counter(); r0 = ref.quadratic_d1_l1(*a, **kw); t_ref = time.perf_counter() - t0
    same = bool(r1["rV"] == r0["rV"] and np.array_equal(r1["Cv"], r0["Cv"]))
    err = float(np.linalg.norm(r1["rX"] - r0["rX"]) / max(np.linalg.norm(r0["rX"]), 1e-300)) if same else None
    return {"problem": f"CP_PFDR_graph_quadratic_d1_l1, EEG-like icosphere level 5 (V={w['V']}, N=91), float64, settings of "
                       "octave/example_EEG_CP.m", "cp_b200_s": round(t1, 4), "cp_reference_s": round(t_ref, 4),
            "rV": int(r0["rV"]), "rV_b200": int(r1["rV"]), "Cv_identical": same, "rX_rel_err": err,
            "host_cores": os.cpu_count()}


def cp_outer_times(dev, with_cpu=True):
    """the passes of one cut-pursuit outer iteration around the PFDR call, operator resident in HBM,
    device in / device out.  Config 3 size for the dense passes (A^tA, V = 32768, float32: 4.3 GB read once per pass),
    config 4's graph (252^3 grid, 48M edges) for the graph passes; the CPU figure is the oracle restatement of the
    reference's serial sweep on the same graph (kind "port", 1 core)."""
    import torch
    from cp_pfdr_graph_d1_b200 import api
    peak, _ = peaks()

    def timed(f, reps=5):
        f(); torch.cuda.synchronize()
        t = []
        for _ in range(reps):
            torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); t.append(time.perf_counter() - t0)
        return float(np.median(t))

    out = {}
    V, rV = 32768, 300
    g = torch.Generator(device=dev).manual_seed(5)
    G = torch.randn(2048, V, device=dev, dtype=torch.float32, generator=g)
    M = (G.t() @ G) / 2048 + 0.1 * torch.eye(V, device=dev)
    del G
    rng = np.random.default_rng(12)
    Cv = rng.integers(0, rV, V).astype(np.int32); Cv[:rV] = np.arange(rV)
    AtY = torch.randn(V, device=dev, dtype=torch.float32, generator=g)
    rX = torch.rand(rV, device=dev, dtype=torch.float32, generator=g)
    op = api.CpOperator(M, V, -V)
    t_set = timed(lambda: op.set_components(Cv, rV))
    t_red = timed(lambda: op.reduce(AtY, on_device=True))
    t_grad = timed(lambda: op.gradient(rX, AtY, on_device=True))
    B = V * V * 4
    out["dense_ata_f32"] = {"V": V, "rV": rV, "operator_bytes": B, "set_components_ms": 1e3 * t_set,
                            "reduce_ms": 1e3 * t_red, "reduce_gbs": B / t_red / 1e9, "reduce_frac": B / t_red / 1e9 / peak,
                            "gradient_ms": 1e3 * t_grad, "gradient_gbs": B / t_grad / 1e9, "gradient_frac": B / t_grad / 1e9 / peak,
                            "what": "wall clock around the C-ABI call (launches + one sync), outputs left on the device"}
    op.close(); del M, op
    torch.cuda.empty_cache()

    n = 252
    Vg = n ** 3
    Eu, Ev = device_grid_edges(n, dev)
    E = int(Eu.shape[0])
    v = torch.arange(Vg, device=dev)
    lab = 4 * (4 * (v // (n * n)) // n) + 4 * ((v // n) % n) // n            # config 4's ground-truth labels: 16 bars
    del v
    active = ((lab[Eu.long()] != lab[Ev.long()]) | (torch.rand(E, device=dev, generator=g) < 0.02)).to(torch.uint8)
    del lab
    res = {}
    t_cc = timed(lambda: res.__setitem__("cc", api.cp_components(Vg, Eu, Ev, active, on_device=True)))
    Cvd, rVg = res["cc"]
    La = torch.full((E,), 0.2, device=dev, dtype=torch.float32)
    t_re = timed(lambda: res.__setitem__("re", api.cp_reduced_edges(Vg, Eu, Ev, active, La, Cvd, rVg)))
    rXg = torch.rand(rVg, device=dev, dtype=torch.float32, generator=g)
    act2 = active.clone()
    t_mg = timed(lambda: api.cp_merge(Eu, Ev, Cvd, rXg, 1e-3, act2))
    out["graph_252"] = {"V": Vg, "E": E, "rV": int(rVg), "rE": int(res["re"][0].shape[0]), "components_ms": 1e3 * t_cc,
                        "reduced_edges_ms": 1e3 * t_re, "merge_ms": 1e3 * t_mg}
    if with_cpu:
        from oracle import oracle as O
        hEu, hEv, hact = Eu.cpu().numpy(), Ev.cpu().numpy(), active.cpu().numpy()
        t0 = time.perf_counter()
        oCv, orV, _, _ = O.CpGraphOracle().components(Vg, hEu, hEv, hact)
        out["graph_252"]["components_cpu_port_ms"] = 1e3 * (time.perf_counter() - t0)
        out["graph_252"]["Cv_identical"] = bool(orV == rVg and np.array_equal(oCv, Cvd.cpu().numpy()))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c4_f32", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads and the cut-pursuit timing")
    ap.add_argument("--iters", type=int, default=0, help="override iterations per step")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from cp_pfdr_graph_d1_b200 import dist_bench
        return dist_bench.main(args)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    from cp_pfdr_graph_d1_b200 import api
    stream = torch.cuda.Stream(dev)
    api.set_stream(stream.cuda_stream)

    name = args.workload
    kind, n, dt, K, iters = WORKLOADS[name]
    iters = args.iters or iters
    if kind in ("ata", "dense"):
        args.no_cpu = args.no_e2e = True                 # several-GB device-generated matrices: resident leg only
    R, w = resident_run(name, dev, stream, args.steps, args.warmup, iters, clock_index=local)
    E = R["E"]
    line = {
        "metric": "pfdr_edge_updates_per_s", "value": R["value"], "unit": "edge_updates/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": R["ms"] / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32" if dt is np.float32 else "f64", "data": "synthetic",
        "config": static_config(name, iters),
        "detail": {"slots": R["setup"]["slots"], "us_per_iteration": 1e3 * R["loop_ms"] / R["its"],
                   "edge_label_updates_per_s": R["value"] * K, "setup_ms_device_resident": R["setup"]["setup_ms"]},
        "gpu_launches": int(R["launches"]),
        "roofline": {"bound": "hbm", "achieved": R["achieved"], "peak": R["peak"], "unit": "GB/s", "frac": R["achieved"] / R["peak"],
                     "traffic": measured_traffic(name), "peak_source": R["peak_src"], "kernel": KERNEL_NAMES[kind],
                     "algorithmic_bytes_per_launch": R["B_iter"], "avg_launch_us": 1e3 * R["loop_ms"] / R["its"]},
        "clocks": R["clocks"],
    }

    if not args.no_e2e:
        def e2e_leg(bufs, reps):
            for _ in range(2):                                   # warm-up: memory pool growth, module load
                one_shot_host(name, w, bufs, iters)
            torch.cuda.synchronize()
            t, h2d, d2h = 0.0, 0, 0
            for _ in range(reps):
                prep_host(name, bufs)                            # the caller's initial iterate: not part of the call
                t0 = time.perf_counter()
                done = one_shot_host(name, w, bufs, iters, prepare=False)
                torch.cuda.synchronize()
                t += time.perf_counter() - t0
                s = api.last_stats()
                h2d, d2h = s["h2d_bytes"], s["d2h_bytes"]
                assert done == iters
            return {"value": E * iters * reps / t, "unit": "edge_updates/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * t / reps}
        reps = max(2, min(args.steps, 3 if n > 128 else 5))
        e = e2e_leg(pin(w, name), reps)
        e["what"] = "one-shot C-ABI call (upload, layout, preconditioning, iterations, download): pinned host buffers in, host iterate out"
        p = e2e_leg(pin(w, name, use_cuda=False), reps)
        e["pageable"] = {"value": p["value"], "ms_per_step": p["ms_per_step"],
                         "what": "same call with ordinary (pageable) host arrays, as the MEX / cut-pursuit callers pass"}
        line["e2e"] = e
    else:
        line["e2e"] = None
    del w
    if not args.no_cpu:
        ws, note = cpu_sample_inputs(name)
        line["cpu_baseline"] = cpu_reference_rate(name, ws, note)
        del ws
    if not args.no_extra:
        extra = {}
        for xn in ("c2_f32", "c2_f64", "c3_f32", "c5s_f32"):
            if xn == name:
                continue
            try:
                X, _ = resident_run(xn, dev, stream, 2, 2)
                extra[xn] = {"workload": describe(xn), "value": X["value"], "unit": "edge_updates/s",
                             "us_per_iteration": 1e3 * X["loop_ms"] / X["its"], "roofline_frac": X["achieved"] / X["peak"],
                             "achieved_gbs": X["achieved"], "algorithmic_bytes_per_launch": X["B_iter"]}
                if xn == "c3_f32" and OPNORM_C3_S is not None:
                    extra[xn]["operator_norm_matrix_s"] = round(OPNORM_C3_S, 4)
            except Exception as ex:                               # a secondary line must not lose the headline
                extra[xn] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
        line["extra"] = extra
        try:
            line["cp_solve"] = cp_solve_time()
        except Exception as ex:
            line["cp_solve"] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
        try:
            line["cp_outer"] = cp_outer_times(dev, with_cpu=not args.no_cpu)
        except Exception as ex:
            line["cp_outer"] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
