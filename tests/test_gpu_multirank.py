"""Multi-rank GPU tests: one torchrun per check, two ranks (needs >= 2 visible GPUs; skipped otherwise).

These run the same scripts used for the multi-GPU profiles under profiles/: fused TP encoder (all-gather -> GEMM,
GEMM -> reduce-scatter), the fused top-k exchange, the VMM/NVLS heap with the multimem kernels, and the tensor-parallel
T5 with the vocab-parallel LM head.  Each script prints ALL OK when every internal comparison against the single-GPU /
NCCL oracle passed."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _run(script: str, *args: str, port: int, nproc: int = 2, timeout: int = 420) -> str:
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "scripts" / script), *args]
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=str(ROOT), env=env)
    text = out.stdout + out.stderr
    assert out.returncode == 0, text[-3000:]
    return text


def test_nvls_heap_and_multimem_kernels_two_ranks():
    text = _run("gpu_check_nvls.py", port=29611)
    assert "ALL OK" in text, text[-3000:]


def test_tp_t5_matches_single_gpu_two_ranks():
    text = _run("gpu_check_tp_t5.py", "t5-tiny", port=29612)
    assert "ALL OK" in text, text[-3000:]


def test_fused_topk_exchange_two_ranks():
    text = _run("gpu_check_p2p.py", port=29613)
    assert "ALL OK" in text and "FAILED" not in text, text[-3000:]


def test_fused_tp_encoder_two_ranks():
    text = _run("gpu_check_tp.py", port=29614)
    assert "ALL OK" in text and "MISMATCH" not in text, text[-3000:]


def test_context_parallel_attention_in_kernel_two_ranks():
    text = _run("gpu_check_cp.py", port=29615)
    assert "ALL OK" in text, text[-3000:]


def test_multigpu_product_serving_matches_single_gpu():
    """[gpu] devices = 2: worker-per-GPU serving returns the single-GPU hit lists (scripts/gpu_check_multigpu_serving.py)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, str(ROOT / "scripts" / "gpu_check_multigpu_serving.py"), "--gpus", "2", "--docs", "6000",
                          "--queries", "128", "--full"], capture_output=True, text=True, timeout=600, cwd=str(ROOT),
                         env=dict(os.environ, PYTHONPATH=str(ROOT)))
    text = out.stdout + out.stderr
    assert out.returncode == 0 and "ALL OK" in text, text[-3000:]


def test_exchange_channels_with_poisoned_slots_two_ranks():
    """Debug build of the channels (INFOMESH_B200_POISON_SLOTS=1): consumers poison every slot they consume and trap if they
    ever read poison; many back-to-back exchanges (graph replays included) must still come out identical to NCCL."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29616", str(ROOT / "scripts" / "gpu_check_p2p.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=str(ROOT),
                         env=dict(os.environ, PYTHONPATH=str(ROOT), INFOMESH_B200_POISON_SLOTS="1"))
    text = out.stdout + out.stderr
    assert out.returncode == 0 and "ALL OK" in text and "POISONED" not in text, text[-3000:]
