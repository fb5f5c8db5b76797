"""bench.py under torchrun (WORLD_SIZE > 1): the same workload, vertex-sharded over the ranks (strong scaling).

Grid workloads (C2, C4 - every BASELINE config that shards but the dense one) run slab-sharded (csrc/slab.cuh): the
iteration kernel of a rank stores its boundary forward points and cut-edge copies straight into the neighbours' memory
over NVLink, one flag per rank pair and iteration; torch.distributed only carries the 256-byte arena descriptors at
set-up.  The dense-operator workloads (C5) use the NCCL halo engine (csrc/comm.cu).  Time = max over ranks of the
CUDA-event time of the timed region.  During warm-up rank 0 also solves the SAME problem unsharded on its own GPU and
the line carries `parity_vs_1gpu`: relative error of the gathered sharded iterate against it (0 expected: the sharded
arithmetic is the single-GPU arithmetic; only the stopping sums - unused here, difTol = 0 - are associated differently)."""
import json
import os
import sys
import time

import numpy as np


def main(args):
    import torch
    import torch.distributed as dist
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    import bench as B
    from cp_pfdr_graph_d1_b200 import api, dist as pd, shard

    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream()
    api.set_stream(stream.cuda_stream)
    api._lib.load().pfdr_b200_set_device(local)

    name = args.workload
    kind, n, dt, K, iters = B.WORKLOADS[name]
    iters = args.iters or iters
    r = np.dtype(dt).itemsize
    tdt = torch.float32 if dt is np.float32 else torch.float64
    Ndense, slab = 0, kind in ("quad", "simplex")
    if kind == "ata":
        raise SystemExit("AtA problems do not shard (DESIGN.md section 6): run this workload with --gpus 1")
    if kind == "dense":
        pd.init_comm(local)
        w = B.device_slab_dense(n, 256, dt, world, rank, dev)
        plan, V, E, Ndense = w["plan"], w["V"], w["E"], w["N"]
        torch.cuda.synchronize()
        mk = lambda: pd.ShardedQuadraticSolver(plan, w["X0"], w["Y"], w["A"], w["La_d1"], None, 0, Ltype=0, L=w["L"], rho=1.0,
                                               condMin=1e-3, difTol=0.0, itMax=iters, local=True, N=Ndense)
    elif kind == "simplex":
        w = B.grid_slab_simplex(n, K, dt, world, rank)            # this rank's planes only (no global arrays)
        plan, V, E = w["plan"], w["V"], w["E"]
        mk = lambda: pd.SlabSimplexSolver(plan, w["Q"], w["Q"], 0.05, w["La_d1"], rho=1.5, condMin=1e-2, difTol=0.0,
                                          itMax=iters, local=True)
    else:
        g = B.host_inputs(name)
        V, E = int(g["V"]), int(g["E"])
        plan = shard.make_slab_plan(V, g["Eu"], g["Ev"], world, rank)
        w = dict(X0=np.zeros(plan.Vloc, dt), Ys=plan.vertex_array(g["Ys"]), La_l2=plan.vertex_array(g["La_l2"]),
                 La_d1=plan.edge_array(g["La_d1"]))
        mk = lambda: pd.SlabQuadraticSolver(plan, w["X0"], w["Ys"], w["La_l2"], w["La_d1"], None, 0, Ltype=1, L=w["La_l2"],
                                            rho=1.5, condMin=1e-3, difTol=0.0, itMax=iters, local=True, N=0)
    solver = mk()
    B_iter = B.algorithmic_bytes(kind, V, E, r, K, N=Ndense)

    def step():
        solver.reset()
        return solver.run()

    clk = B.ClockSampler(local).start() if rank == 0 else None        # nvidia-smi is up and sampling before the timed region
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()

    # ---- correctness of the sharded solve: against the unsharded solve of the same problem on rank 0's GPU ----
    parity = None
    if slab:
        bounds = shard.slab_bounds_aligned(V, plan.band, world)

        def gather_to0(t):                                   # rows of every rank's owned range -> rank 0 (sizes differ)
            if rank != 0:
                dist.send(t.contiguous(), 0)
                return None
            out = [t]
            for q in range(1, world):
                buf = torch.empty((int(bounds[q + 1] - bounds[q]),) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
                dist.recv(buf, q)
                out.append(buf)
            return out

        parts = gather_to0(torch.from_numpy(np.ascontiguousarray(solver.fetch().X)).to(dev))
        if kind == "simplex":
            qparts = gather_to0(torch.from_numpy(np.ascontiguousarray(plan.owned(w["Q"]))).to(dev))
        if rank == 0:
            got = torch.cat(parts)
            if kind == "simplex":
                Qg = torch.cat(qparts)
                del qparts
                Eu, Ev = B.device_grid_edges(n, dev)
                La = torch.full((Eu.shape[0],), 0.2, device=dev, dtype=tdt)
                torch.cuda.synchronize()
                one = api.SimplexSolver(Qg, Qg, 0.05, Eu, Ev, La, None, rho=1.5, condMin=1e-2, difRcd=0.0, difTol=0.0, itMax=iters)
            else:
                one = api.QuadraticSolver(np.zeros(V, dt), g["Ys"], g["La_l2"], g["Eu"], g["Ev"], g["La_d1"], None, 0, Ltype=1,
                                          L=g["La_l2"], rho=1.5, condMin=1e-3, difRcd=0.0, difTol=0.0, itMax=iters, N=0)
            one.run()
            ref = torch.from_numpy(one.fetch().X).to(dev)
            one.close()
            parity = float((got - ref).double().norm() / ref.double().norm())
            del got, ref, parts
            torch.cuda.empty_cache()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    loop_ms = launches = its = 0
    if clk:
        clk.mark_start()
    ev0.record(stream)
    for _ in range(args.steps):
        st = step()
        loop_ms += st["loop_ms"]; launches += st["launches"]; its += st["iterations"]
    ev1.record(stream)
    torch.cuda.synchronize()
    if clk:
        clk.mark_end()
        clk.stop()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1), loop_ms, float(launches)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, loop_ms_max = float(t[0]), float(t[1])

    # end to end at N GPUs: every rank builds its sharded solver from HOST (pinned) arrays of its slab - upload, layout,
    # preconditioning, attach - iterates and reads its part of the iterate back; max over ranks of the wall time
    e2e = None
    if slab and not args.no_e2e:
        solver.close()
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        for k in list(w):
            if isinstance(w[k], np.ndarray):
                w[k] = pin(w[k])
        times, h2d, d2h = [], 0, 0
        for i in range(4):                                   # first two: warm-up (memory pool growth)
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            s2 = mk()
            h2d = s2.setup_stats["h2d_bytes"]
            s2.run()
            r2 = s2.fetch()
            d2h = api.last_stats()["d2h_bytes"]
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
            assert r2.it == iters
            s2.close()
        te = torch.tensor([sum(times[2:]) / 2, float(h2d), float(d2h)], device="cuda", dtype=torch.float64)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": E * iters / float(te[0]), "unit": "edge_updates/s", "h2d_bytes_per_step": int(te[1]),
               "d2h_bytes_per_step": int(te[2]), "ms_per_step": 1e3 * float(te[0]),
               "what": "per rank: slab solver built from pinned host arrays of its slab, attached, run, owned iterate read back; "
                       "max over ranks (h2d/d2h: largest rank)"}
        solver = None
    if rank == 0:
        assert its == iters * args.steps
        value = E * its / (ms * 1e-3)
        peak, peak_src = B.peaks()
        achieved = B_iter * its / (loop_ms_max * 1e-3) / 1e9
        par = (f"vertex slabs x{world}: halo (boundary forward points + cut-edge copies) stored by the iteration kernel into the "
               "neighbours' memory over NVLink (CUDA IPC), one flag per rank pair and iteration, no collective on the data path"
               if slab else f"vertex shards x{world}, dense A by columns: NCCL halo exchange + all-reduce of A x and of the stopping scalars")
        line = {
            "metric": "pfdr_edge_updates_per_s", "value": value, "unit": "edge_updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32" if dt is np.float32 else "f64", "data": "synthetic",
            "config": B.static_config(name, iters),
            "detail": {"us_per_iteration": 1e3 * loop_ms_max / its, "parallelism": par, "parity_vs_1gpu": parity,
                       "edge_label_updates_per_s": value * K},
            "parity_vs_1gpu": parity,
            "gpu_launches": int(t[2]),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s",
                         "frac": achieved / (peak * world), "traffic": None, "peak_source": peak_src + f" x {world} GPUs",
                         "kernel": B.KERNEL_NAMES[kind],
                         "algorithmic_bytes_per_launch": B_iter / world, "avg_launch_us": 1e3 * loop_ms_max / its},
            "e2e": e2e, "clocks": clk.summary(),
        }
        print(json.dumps(line), flush=True)
    if solver is not None:
        solver.close()
    if kind == "dense":
        pd.destroy_comm()
    dist.destroy_process_group()
