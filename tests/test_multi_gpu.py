"""Multi-GPU correctness as part of the pytest run (it lived in scripts only the builder ran).  Needs >= 2 visible GPUs;
each test launches one process per GPU with torch.distributed.run on 127.0.0.1 and checks the script's own verdict lines:
  * tests/mgpu_verify_gather.py   gathered draw list on the presenter == the single-GPU list, byte for byte
  * tests/mgpu_instanced.py       gathered per-LOD instance buckets == the oracle's buckets for the whole scene
  * bench.py --gpus 2             the contract line carries gather_ok = true (device-side reduction of the gathered lists)
  * bench.py --workload cfg4      gathered cluster draw lists == sum of the ranks' reductions, all three modes
"""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def torchrun(nproc, script, *args, port=29611, timeout=1500):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, script), *args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    out = []
    for line in r.stdout.splitlines():
        line = line.strip()
        if line.startswith("{"):
            try:
                out.append(json.loads(line))
            except Exception:
                pass
    return out, r.stdout


needs2 = pytest.mark.skipif(gpus() < 2, reason="needs at least 2 GPUs")


@needs2
def test_gathered_draw_list_equals_single_gpu_list(built):
    _, text = torchrun(2, "tests/mgpu_verify_gather.py", port=29611)          # exits non-zero on any mismatch
    assert text.count("IDENTICAL") == 6 and "MISMATCH" not in text, text[-2000:]


@needs2
def test_gathered_instance_buckets_equal_oracle(built):
    lines, _ = torchrun(2, "tests/mgpu_instanced.py", "--objects", "4000000", "--verify-objects", "1500000", "--iters", "3", port=29612)
    ver = [l for l in lines if l.get("what") == "verification run"]
    assert ver and all(l["gathered_buckets_equal_oracle"] for l in ver), lines


@needs2
def test_bench_line_carries_gather_ok(built):
    lines, _ = torchrun(2, "bench.py", "--gpus", "2", "--steps", "10", "--warmup", "3", "--objects", "2097152", "--no-e2e", port=29613)
    assert lines and lines[-1]["n_gpus"] == 2 and lines[-1]["gather_ok"] is True, lines
    assert all(g["equals_sum_of_ranks"] and g["unsorted"] == 0 for g in lines[-1]["detail"]["gathered_lists"])


@needs2
def test_cluster_workload_gather(built):
    lines, _ = torchrun(2, "bench.py", "--gpus", "2", "--workload", "cfg4", "--records", "16777216", "--steps", "3", "--warmup", "3", port=29614)
    modes = lines[-1]["detail"]["modes"]
    assert set(modes) == {"passthrough", "sphere", "sphere_hiz"} and all(m["gathered_equals_sum_of_ranks"] for m in modes.values()), lines


@needs2
def test_instancing_workload_gather(built):
    lines, _ = torchrun(2, "bench.py", "--gpus", "2", "--workload", "cfg3", "--objects", "4194304", "--steps", "3", "--warmup", "3", port=29615)
    assert lines[-1]["detail"]["gathered_buckets_equal_sum_of_ranks"] is True, lines
