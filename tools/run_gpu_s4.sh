#!/bin/bash
# round 2, session 4: flagged-publish fast path, tests, full bench with the new legs
mkdir -p gpurun_out
timeout 300 python tools/probe_latency.py > gpurun_out/probe_latency.log 2>&1; tail -18 gpurun_out/probe_latency.log
timeout 300 python tools/probe_phases.py > gpurun_out/probe_phases.log 2>&1; head -40 gpurun_out/probe_phases.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_gpu.log
tail -15 gpurun_out/tests_gpu.log
timeout 1200 python bench.py --steps 3 --warmup 3 --ref-replays 1 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_ours.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_ours.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "with_retention", d["e2e"]["with_retention"]["value"])
    print("strict", d["e2e"]["strict"])
    print("roofline", d["roofline"]["frac"], "cpu", d["cpu_baseline"])
    print("micro", json.dumps(d["micro_alloc_free"], indent=1))
    print("per_call", d["e2e_per_call_api"])
    print("mla_pin", json.dumps(d["mla_pin"], indent=1))
    print("shards", d["kv_scatter_tp_shards_on_one_gpu"])
except Exception as e:
    print("no bench line:", e)
PY
