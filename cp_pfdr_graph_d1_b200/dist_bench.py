"""bench.py under torchrun (WORLD_SIZE > 1): the same workload, vertex-sharded over the ranks (strong scaling).
Each rank builds the seeded global problem on the host, keeps its slab (shard.make_plan), creates a sharded
solver and runs the same steps; time is the max over ranks of the CUDA-event time of the timed region."""
import json
import os
import sys

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
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    api.set_stream(stream.cuda_stream)
    pd.init_comm(local)

    name = args.workload
    kind, n, dt, K, iters = B.WORKLOADS[name]
    iters = args.iters or iters
    r = np.dtype(dt).itemsize
    Ndense = 0
    if kind == "dense":
        w = B.device_slab_dense(n, 256, dt, world, rank, torch.device("cuda", local))
        plan, V, E, Ndense = w["plan"], w["V"], w["E"], w["N"]
        torch.cuda.synchronize()
        solver = pd.ShardedQuadraticSolver(plan, w["X0"], w["Y"], w["A"], w["La_d1"], None, 0, Ltype=0, L=w["L"], rho=1.0,
                                           condMin=1e-3, difTol=0.0, itMax=iters, local=True, N=Ndense)
    elif kind == "ata":
        raise SystemExit("AtA problems do not shard (DESIGN.md section 6): run this workload with --gpus 1")
    elif n > 128 and kind == "simplex":
        w = B.grid_slab_simplex(n, K, dt, world, rank)            # slab generated directly (no global arrays)
        plan, V, E = w["plan"], w["V"], w["E"]
        solver = pd.ShardedSimplexSolver(plan, w["Q"], w["Q"], 0.05, w["La_d1"], rho=1.5, condMin=1e-2, difTol=0.0,
                                         itMax=iters, local=True)
    else:
        w = B.host_inputs(name)
        V, E = int(w["V"]), int(w["E"])
        plan = shard.make_plan(V, w["Eu"], w["Ev"], world, rank)
        if kind == "quad":
            solver = pd.ShardedQuadraticSolver(plan, np.zeros(V, dt), w["Ys"], w["La_l2"], w["La_d1"], None, 0, Ltype=1,
                                               L=w["La_l2"], rho=1.5, condMin=1e-3, difTol=0.0, itMax=iters)
        else:
            solver = pd.ShardedSimplexSolver(plan, w["Q"], w["Q"], 0.05, w["La_d1"], rho=1.5, condMin=1e-2, difTol=0.0,
                                             itMax=iters)
    B_iter = B.algorithmic_bytes(kind, V, E, r, K, N=Ndense)

    def step():
        solver.reset()
        return solver.run()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    loop_ms = launches = its = 0
    with B.ClockSampler(local) as clk:
        ev0.record(stream)
        for _ in range(args.steps):
            st = step()
            loop_ms += st["loop_ms"]; launches += st["launches"]; its += st["iterations"]
        ev1.record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1), loop_ms, float(launches)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, loop_ms_max = float(t[0]), float(t[1])

    # end to end at N GPUs: every rank builds its sharded solver from HOST (pinned) local arrays - upload, half-edge
    # layout, preconditioning - iterates and reads its part of the iterate back; max over ranks of the wall time
    e2e = None
    if kind in ("quad", "simplex") and not (n > 128 and kind == "simplex" and world < 2):
        import time
        solver.close()
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        if kind == "quad":
            loc = [pin(plan.vertex_array(a)) for a in (np.zeros(V, dt), w["Ys"], w["La_l2"])] + [pin(plan.edge_array(w["La_d1"]))]
            mk = lambda: pd.ShardedQuadraticSolver(plan, loc[0], loc[1], loc[2], loc[3], None, 0, Ltype=1, L=loc[2], rho=1.5,
                                                   condMin=1e-3, difTol=0.0, itMax=iters, local=True, N=0)
        else:
            Ql = pin(w["Q"] if n > 128 else plan.vertex_array(w["Q"]))
            Ld = pin(w["La_d1"] if n > 128 else plan.edge_array(w["La_d1"]))
            mk = lambda: pd.ShardedSimplexSolver(plan, Ql, Ql, 0.05, Ld, rho=1.5, condMin=1e-2, difTol=0.0, itMax=iters, local=True)
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
               "what": "per rank: sharded solver built from pinned host arrays of its shard, run, owned iterate read back; "
                       "max over ranks (h2d/d2h: largest rank)"}
        solver = None
    if rank == 0:
        assert its == iters * args.steps
        value = E * its / (ms * 1e-3)
        peak, peak_src = B.peaks()
        achieved = B_iter * its / (loop_ms_max * 1e-3) / 1e9
        halo_bytes = sum((nb.send_vertices.shape[0] + nb.send_halfedges.shape[0]) for nb in plan.neighbors) * K * r
        line = {
            "metric": "pfdr_edge_updates_per_s", "value": value, "unit": "edge_updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32" if dt is np.float32 else "f64", "data": "synthetic",
            "config": {"workload": B.describe(name), "V": V, "E": E, "K": K, "iterations_per_step": iters,
                       "us_per_iteration": 1e3 * loop_ms_max / its, "parallelism": f"vertex slabs x{world}, NCCL halo "
                       "exchange (boundary forward points + cut-edge copies) + all-reduce of the stopping scalars",
                       "halo_bytes_per_iteration_rank0": int(halo_bytes),
                       "l2": "no flush: steps re-run the same resident problem"},
            "gpu_launches": int(t[2]),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s",
                         "frac": achieved / (peak * world), "traffic": None, "peak_source": peak_src + f" x {world} GPUs",
                         "kernel": {"quad": "quad_iter_pipe_kernel", "simplex": "simplex_iter_kernel",
                                    "dense": "ax_partial_kernel + atr_kernel"}[kind],
                         "algorithmic_bytes_per_launch": B_iter / world, "avg_launch_us": 1e3 * loop_ms_max / its},
            "e2e": e2e, "clocks": clk.summary(),
        }
        print(json.dumps(line), flush=True)
    if solver is not None:
        solver.close()
    pd.destroy_comm()
    dist.destroy_process_group()
