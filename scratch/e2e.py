import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench as B
from cp_pfdr_graph_d1_b200 import api
for name in ("c2_f32", "c4s_f32"):
    w = B.host_inputs(name); pinned = B.pin(w, name)
    it = B.WORKLOADS[name][4]
    for rep in range(4):
        t0 = time.perf_counter(); B.one_shot_host(name, w, pinned, it); torch.cuda.synchronize(); t = time.perf_counter() - t0
        s = api.last_stats()
        print(name, "rep", rep, "wall %.1f ms" % (1e3 * t), "total %.1f setup %.1f loop %.1f ms" % (s["total_ms"], s["setup_ms"], s["loop_ms"]), "h2d %.1f MB" % (s["h2d_bytes"] / 1e6))
