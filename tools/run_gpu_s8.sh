#!/bin/bash
# round 2, session 8: shared-memory free path + warp-scope alloc run: tests first, then phases, latency, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_gpu.log
tail -25 gpurun_out/tests_gpu.log
timeout 300 python tools/probe_phases.py > gpurun_out/probe_phases.log 2>&1; cat gpurun_out/probe_phases.log
timeout 300 python tools/probe_latency.py > gpurun_out/probe_latency.log 2>&1; tail -18 gpurun_out/probe_latency.log
timeout 1500 python bench.py --steps 3 --warmup 3 --ref-replays 1 --no-two-model > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_ours.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_ours.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms_per_step", d["ms_per_step"], "us/sched step", d["roofline"]["step_kernel"], "e2e", d["e2e"]["value"])
    print("micro", {k: v for k, v in d["micro_alloc_free"].items() if isinstance(v, float)})
    print("per_call", d["e2e_per_call_api"]["value"], "mla", {k: v.get("ours_vs_reference") for k, v in d["mla_pin"].items() if isinstance(v, dict)})
except Exception as e:
    print("no bench line:", e)
PY
