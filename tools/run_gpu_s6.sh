#!/bin/bash
# round 2, session 6 (2 GPUs): all tests incl. multi-GPU TP, bench N=1 (all legs), bench N=2 both transports, reference arm N=2
mkdir -p gpurun_out
nvidia-smi -L
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tests_gpu_2gpu.log
tail -15 gpurun_out/tests_gpu_2gpu.log
timeout 1500 python bench.py --steps 3 --warmup 3 --ref-replays 1 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench N=1 rc=$?"
tail -c 2000 gpurun_out/bench_ours.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench N=2 mailbox rc=$?"
tail -c 2500 gpurun_out/bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 3 --warmup 3 --tp-transport nccl > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err; echo "bench N=2 nccl rc=$?"
tail -c 1500 gpurun_out/bench_n2_nccl.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "reference N=2 rc=$?"
tail -c 1500 gpurun_out/bench_ref_n2.err
python - <<'PY'
import json
def last(path):
    try:
        return json.loads([l for l in open(path).read().strip().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        return {"error": str(e)}
d = last("gpurun_out/bench_ours.json")
if "value" in d:
    print("N=1 value", d["value"], "e2e", d["e2e"]["value"], "retention", d["e2e"]["with_retention"]["value"], "no_zero", d["e2e"]["no_zero_page"])
    print("two_model", json.dumps(d["two_model"], indent=1)[:3500])
    print("init", json.dumps(d["init_and_layout"], indent=1)[:2500])
    print("cpu", d["cpu_baseline"])
for f in ("bench_n2", "bench_n2_nccl"):
    d = last(f"gpurun_out/{f}.json")
    if "value" in d:
        print(f, "value", d["value"], "e2e", d["e2e"]["value"], "retention", d["e2e"]["with_retention"]["value"], "tp", d["e2e"]["tp"], "strict", d["e2e"]["strict"])
        print("  scatter", d["kv_scatter"]["scatter"], "map_unmap", {k: (v.get("map_plus_unmap_vs_raw", v) if isinstance(v, dict) else v) for k, v in d["map_unmap"].items()})
    else:
        print(f, d)
print("ref N=2", last("gpurun_out/bench_ref_n2.json"))
PY
