"""GPU, >= 2 devices (run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_ddp.py -m gpu`): averaged gradients of
two half-batches == single-rank gradients of the concatenated batch, through NCCL."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_gradients_match_single_rank(cuda_dev):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29700 + (os.getpid() % 200)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "ddp_gpu_worker.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("DDP_RESULT ")]
    assert line, r.stdout[-2000:]
    res = json.loads(line[-1][len("DDP_RESULT "):])
    print(res)
    assert res["segments"] == ["decoder", "fusion_bn", "trunk_deep", "trunk_shallow"]
    # same kernels on both sides; the half-batch losses are means over half as many pixels (gradients scaled by 2, a power of
    # two: identical bf16 roundings), so only summation orders differ
    for k in ("overlap_eager", "overlap_graph", "flat"):
        assert res[k] <= 2e-3, (k, res)
