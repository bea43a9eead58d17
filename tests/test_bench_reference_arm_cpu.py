"""`bench.py --impl reference` prints ONE JSON line with the contract's keys; without a GPU it falls back to the
reference's "cpu" device (bookkeeping only) instead of failing.  Needs oracle/_ref (skipped where it cannot be built)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def test_reference_arm_prints_the_contract_line():
    import build_ref
    if not (build_ref.built() or build_ref.have_reference()):
        pytest.skip("oracle/_ref is not built and /root/reference is not present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "kv_block_alloc_free_ops_per_s" and d["unit"] == "ops/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] == 2
    assert d["e2e"] == {"value": d["value"], "unit": "ops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("T-decode-4096") and d["gpu_launches"] == 0


def test_reference_arm_at_n_gt_1_uses_a_bounded_sample_and_other_ranks_do_nothing():
    import build_ref
    if not (build_ref.built() or build_ref.have_reference()):
        pytest.skip("oracle/_ref is not built and /root/reference is not present")
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""          # rank 1 exits 0 without work
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.strip()][-1])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert "bounded sample" in d["cpu_baseline"]["sample"]
