"""Numerical check of the overlapped data-parallel path on REAL GPUs (the gloo tests cover only the bucket
arithmetic).  Needs two GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp.py -m gpu`; skipped on one."""
import json
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("overlap", ["layer", "none"])
def test_overlapped_allreduce_equals_the_local_sum_and_replicas_stay_in_sync(overlap):
    """Both exchange policies of DataParallelTrainer ("none": one all-reduce after backward, the default; "layer": below).
    Two ranks, different shards, identical weights: the gradients after the per-layer-event all-reduce (NCCL on a side
    stream behind the events the native backward records, head parameters in the last layer's bucket, position embedding
    in the last bucket) equal the sum of the shards' gradients computed on one GPU without any exchange -- every tensor
    within 1e-5 relative (fp32 sums in a different order) -- and after two AdamW steps the replicas' parameter vectors are
    bit-identical."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(here, "dp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "TOME_DP_OVERLAP": overlap})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("DPRESULT ")][-1]
    res = json.loads(line[len("DPRESULT "):])
    print(res)
    assert res["world"] == 2 and res["grad_norm"] > 0 and res["overlap"] == overlap
    assert res["rel_err_all"] <= 1e-5 and res["worst_tensor_rel_err"] <= 1e-5, res
    assert res["params_in_sync"], res
