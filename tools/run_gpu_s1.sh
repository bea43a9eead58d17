#!/bin/bash
# round 2, session 1: tests, latency probe, bench N=1 (short reference legs)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_gpu.log
tail -5 gpurun_out/tests_gpu.log
timeout 300 python tools/probe_latency.py > gpurun_out/probe_latency.log 2>&1; tail -20 gpurun_out/probe_latency.log
timeout 900 python bench.py --steps 3 --warmup 3 --ref-replays 1 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_ours.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_ours.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "default", d["e2e"]["with_retention"]["value"])
    print("strict", d["e2e"]["strict"])
    print("roofline", d["roofline"]["frac"], "micro", d["micro_alloc_free"], "cpu", d["cpu_baseline"])
    print("map_unmap", {k: (v.get("map_plus_unmap_vs_raw") if isinstance(v, dict) else v) for k, v in d["map_unmap"].items()})
except Exception as e:
    print("no bench line:", e)
PY
