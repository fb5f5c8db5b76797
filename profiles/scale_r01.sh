#!/bin/bash
# strong-scaling sweep on one 8-GPU box: JSON lines -> gpurun_out/scale_r01.jsonl
OUT=gpurun_out/scale_r01.jsonl
: > $OUT
run1() { timeout 600 python bench.py --workload $1 --steps 2 --warmup 2 --no-cpu --no-e2e ${2:+--iters $2} 2>&1 | grep -E '^\{' >> $OUT; }
runN() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $1 --steps 2 --warmup 2 --workload $2 ${3:+--iters $3} 2>&1 | grep -E '^\{|Error|error' | tail -3 >> $OUT; }
nvidia-smi -L | wc -l
run1 c2_f32; run1 c4_f32 50
for n in 2 4 8; do runN $n c2_f32; runN $n c4_f32 50; runN $n c5_f32 10; done
python - <<'PY'
import json
for l in open("gpurun_out/scale_r01.jsonl"):
    try:
        d = json.loads(l)
        print(d["n_gpus"], d["config"]["workload"][:34], "| us/it", round(d["config"]["us_per_iteration"], 1), "| eu/s %.3e" % d["value"], "| frac", round(d["roofline"]["frac"], 3))
    except Exception:
        print(l[:200])
PY
