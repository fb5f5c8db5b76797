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
    w = synth.c2_denoise(20, np.float64, diag=True)
    A = w["La_l2"]; Y = A * w["Y"]
    plan = shard.make_plan(w["V"], w["Eu"], w["Ev"], world, rank)
    for itmax in (1, 5, 17, 40, 150):
        s = pd.ShardedQuadraticSolver(plan, np.zeros(w["V"]), Y, A, w["La_d1"], np.full(w["V"], 0.05), 1, Ltype=1, L=A, rho=1.5, condMin=1e-3, difTol=0.0, itMax=itmax, want_dif=True)
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
    w = synth.c2_denoise(20, np.float64, diag=True)
    A = w["La_l2"]; Y = A * w["Y"]
    for itmax in (1, 5, 17, 40, 150):
        one = api.PFDR_graph_quadratic_d1_l1(np.zeros(w["V"]), Y, A, w["Eu"], w["Ev"], w["La_d1"], np.full(w["V"], 0.05), 1, 1, A, 1.5, 1e-3, 0.0, 0.0, itmax, 0, want_dif=True)
        for world in (1, 2):
            x = np.load(f"{out}_{world}_{itmax}.npy")
            d = np.abs(x - one.X)
            print("itmax", itmax, "world", world, "max abs diff", d.max(), "argmax", d.argmax(), "n bad", (d > 1e-12).sum(), "first bad", np.nonzero(d > 1e-12)[0][:8])
        print("one Dif", one.Dif[-3:])
