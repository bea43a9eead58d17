#!/bin/bash
# round 2, session 2: launch floor, fast single-request path, tests, bench with the new default policy
mkdir -p gpurun_out
nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/launch_floor tools/launch_floor.cu && timeout 120 tools/launch_floor > gpurun_out/launch_floor.log 2>&1
cat gpurun_out/launch_floor.log
timeout 300 python tools/probe_latency.py > gpurun_out/probe_latency.log 2>&1; tail -20 gpurun_out/probe_latency.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_gpu.log
tail -15 gpurun_out/tests_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 --ref-replays 1 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_ours.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_ours.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "with_retention", d["e2e"]["with_retention"]["value"])
    print("strict", d["e2e"]["strict"])
    print("retain", d["e2e"]["with_retention"])
    print("roofline", d["roofline"]["frac"], "micro", d["micro_alloc_free"], "cpu", d["cpu_baseline"])
except Exception as e:
    print("no bench line:", e)
PY
