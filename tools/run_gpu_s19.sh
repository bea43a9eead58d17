#!/bin/bash
# round 2, session 19 (2 GPUs): bench N=2 with the driver's K / W after the kernel work of sessions 13-18, reference arm N=2
mkdir -p gpurun_out
nvidia-smi -L | wc -l
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench N=2 rc=$? wall=${SECONDS}s"
tail -c 1500 gpurun_out/bench_n2.err
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29742 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "reference N=2 rc=$? wall=${SECONDS}s"
python - <<'PY'
import json
def last(path):
    try:
        return json.loads([l for l in open(path).read().strip().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        return {"error": str(e)}
d = last("gpurun_out/bench_n2.json")
if "value" in d:
    print("N=2 value", d["value"], "us/step", d["roofline"]["step_kernel"]["us_per_scheduler_step"], "e2e", d["e2e"]["value"], "retention", d["e2e"]["with_retention"]["value"])
    print(" strict", {k: v for k, v in d["e2e"]["strict"].items() if k != "policy"})
    print(" tp", d["e2e"]["tp"])
    print(" scatter", d["kv_scatter"]["scatter"], d["kv_scatter"]["gather"])
    print(" map_unmap", {k: (v.get("map_plus_unmap_vs_raw", v) if isinstance(v, dict) else v) for k, v in d["map_unmap"].items()})
else:
    print(d)
r = last("gpurun_out/bench_ref_n2.json")
print("ref N=2", r.get("value"), r.get("cpu_baseline", {}).get("sample", "")[:200])
PY
