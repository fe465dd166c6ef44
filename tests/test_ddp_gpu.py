"""Multi-GPU parity: needs >= 2 GPUs; each check runs under torchrun in a subprocess
(tests/gpu/ddp_check.py) so that a hang is bounded by a timeout and cannot take the test session with it."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, port):
    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip('needs at least 2 GPUs')
    n = 2 if n < 4 else (4 if n < 8 else 8)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}',
           '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tests', 'gpu', 'ddp_check.py')] + args
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=ROOT)
    print(out.stdout[-3000:])
    print(out.stderr[-3000:])
    assert out.returncode == 0


def test_peer_fused_step_equals_single_rank_large_batch_step():
    _run(['step', 'peer'], 29611)


def test_nccl_bucket_allreduce_step_equals_single_rank_large_batch_step():
    _run(['step', 'nccl'], 29612)


def test_peer_fused_exchange_and_sgd_are_exact_on_known_gradients():
    _run(['exchange', 'peer'], 29614)


def test_nccl_exchange_and_sgd_are_exact_on_known_gradients():
    _run(['exchange', 'nccl'], 29615)


def test_sharded_retrieval_equals_single_gpu_search():
    _run(['retrieval'], 29613)
