"""Data-parallel training (BASELINE configs[2]) with two ranks: tools/dp_check.py under torchrun. On a box
with one GPU both ranks share it and the 18 KB gradient travels over gloo; with two or more GPUs the same script runs on
NCCL (gpurun --gpus 2). The sharding / all-reduce host logic is also covered on CPU (tests/test_parallel_cpu.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_training_step_equals_single_process(world):
    port = 29500 + (os.getpid() % 400) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "dp_check.py")]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="4")
    res = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-4000:]
    assert res.stdout.count("DP_CHECK OK") == world, res.stdout[-4000:]
    print("\n".join(l for l in res.stdout.splitlines() if "DP_CHECK" in l))
