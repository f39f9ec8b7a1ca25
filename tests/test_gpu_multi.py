"""Multi-GPU parity (skipped on a one-GPU box): the split solve of the sharded SR step against the plain sharded step, under
torchrun on two GPUs (tests/tools_split_solve.py holds the assertions).  The N > 1 host logic is covered on the CPU by the
gloo tests (test_distributed_gloo.py, test_host_logic.py::test_split_solve_shares)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize('engine', ['tc_exact', 'tc'])
def test_split_solve_equals_the_plain_sharded_step(engine):
    if _gpus() < 2:
        pytest.skip('needs two GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29541', os.path.join(ROOT, 'tests', 'tools_split_solve.py'), '--engine', engine]
    out = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    text = out.stdout.decode(errors='replace')
    assert out.returncode == 0 and 'SPLIT SOLVE OK' in text, text[-4000:]
