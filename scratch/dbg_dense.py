import os, sys, socket
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, torch.distributed as dist, torch.multiprocessing as mp

def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cp_pfdr_graph_d1_b200 import dist as pd, shard, synth
    pd.init_comm(rank)
    w = synth.dense_problem(12, 40, np.float64)
    lam = np.array([float(np.linalg.norm(w["A"], 2) ** 2)])
    plan = shard.make_plan(w["V"], w["Eu"], w["Ev"], world, rank)
    for itmax in (1, 2, 5):
        s = pd.ShardedQuadraticSolver(plan, np.zeros(w["V"]), w["Y"], w["A"], w["La_d1"], None, 0, Ltype=0, L=lam, rho=1.5, condMin=1e-3, difTol=0.0, itMax=itmax, want_dif=True)
        s.run(); r = s.fetch(want_dif=True)
        full = pd.gather_owned(r.X, plan)
        if rank == 0: np.save(f"{out}_{world}_{itmax}.npy", full); print("world", world, "itmax", itmax, "Dif", r.Dif[-3:])
        s.close()
    pd.destroy_comm(); dist.destroy_process_group()

def port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); return s.getsockname()[1]

if __name__ == "__main__":
    out = "/tmp/dbg"
    for world in (1, 2):
        mp.spawn(worker, args=(world, port(), out), nprocs=world, join=True)
    from cp_pfdr_graph_d1_b200 import api, synth
    w = synth.dense_problem(12, 40, np.float64)
    lam = np.array([float(np.linalg.norm(w["A"], 2) ** 2)])
    for itmax in (1, 2, 5):
        one = api.PFDR_graph_quadratic_d1_l1(np.zeros(w["V"]), w["Y"], w["A"], w["Eu"], w["Ev"], w["La_d1"], None, 0, 0, lam, 1.5, 1e-3, 0.0, 0.0, itmax, w["N"], want_dif=True)
        for world in (1, 2):
            x = np.load(f"{out}_{world}_{itmax}.npy")
            d = np.abs(x - one.X)
            print("itmax", itmax, "world", world, "max abs diff", d.max(), "n bad", (d > 1e-9).sum(), "first bad", np.nonzero(d > 1e-9)[0][:8], "V", x.shape)
        print("one Dif", one.Dif[-3:])
