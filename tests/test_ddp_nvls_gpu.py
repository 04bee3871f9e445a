"""Fused reduce-scatter + AdamW + shadow all-gather over NVLink (csrc/ddp_nvls.cu) against NCCL all-reduce -> ub_adamw_dev on
the real 88 M-element arena size: needs >= 2 GPUs on the box (skipped otherwise), launched through torchrun like bench.py."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["push", "p2p", "mc"])
def test_fused_step_matches_nccl_then_adamw(mode):
    n_gpu = torch.cuda.device_count()
    if n_gpu < 2:
        pytest.skip("needs 2 GPUs")
    world = 8 if n_gpu >= 8 else (4 if n_gpu >= 4 else 2)
    env = dict(os.environ, UB_NVLS_MODE=mode)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29561", os.path.join(ROOT, "tools", "nvls_check.py"), "--numel", "8808040", "--iters", "3"],
                       capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0 and "PARITY OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
