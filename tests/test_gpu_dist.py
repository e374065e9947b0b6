"""Multi-GPU (N>1) equivalence on real kernels: needs >= 2 GPUs on the box (gpurun --gpus 2); skipped otherwise."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_render_and_train_step_equivalence():
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(here, "dist_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    out_dir = os.path.join(os.path.dirname(here), "gpurun_out")
    if os.path.isdir(out_dir):                      # keep the measured equivalence errors next to the pass/fail line
        with open(os.path.join(out_dir, "dist_worker_stdout.txt"), "w") as f:
            f.write("".join(l + "\n" for l in r.stdout.splitlines() if l.startswith("rank") or "DIST_OK" in l))
