#!/bin/bash
# round 2, session 23 (4 GPUs): bench N=4 with the final kernels (sanity check of the scaling run's path)
mkdir -p gpurun_out
SECONDS=0
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench N=4 rc=$? wall=${SECONDS}s"
tail -c 800 gpurun_out/bench_n4.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_n4.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=4 value", d["value"], "us/step", d["roofline"]["step_kernel"]["us_per_scheduler_step"], "e2e", d["e2e"]["value"], "tp", d["e2e"]["tp"])
    print(" scatter", d["kv_scatter"]["scatter"])
except Exception as e:
    print("no line:", e)
PY
