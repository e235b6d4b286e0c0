"""Multi-GPU parity inside `pytest -m gpu` (missing #6): when the box shows two or more
GPUs, spawn tests/multi_gpu_parity.py under torchrun (one rank per GPU, NCCL) and require its
verdict -- every case within north_star's tolerance of the CPU oracle run on the SAME partition,
the three stepping modes bit-identical, overlap on / off bit-identical.  On a single-GPU box the
N-rank module path is still covered: tests/test_host_emulation.py runs it with ranks as host
threads against the same partition oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def visible_gpus():
    import ryujin_b200 as rj

    return rj.device_count()


@pytest.mark.gpu
def test_multi_gpu_parity_against_the_partition_oracle():
    n = visible_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (this box shows %d)" % n)
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    env = dict(os.environ)
    env.pop("RYUJIN_B200_LIB", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_parity.py")]
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    sys.stdout.write(out.stdout[-6000:])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "multi-gpu parity ok" in out.stdout
    assert out.stdout.count("vs ORACLE with the same partition") >= 6
