"""The multi-GPU exchange on real GPUs: needs at least two of them on the box (`gpurun --gpus 2 -- pytest
tests/test_gpu_multi.py -m gpu`); skipped on a one-GPU box, where bench.py's own N > 1 check covers the driver's scaling
run. One process per GPU under torch.distributed.run; the check itself is tests/tools/multi_check.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("no_symm", [False, True], ids=["symmetric-memory", "cuda-ipc"])
def test_exchange_kernel_all_gather_and_one_gpu_agree(no_symm):
    n = _gpus()
    if n < 2:
        pytest.skip("needs two GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ)
    env.pop("LUTGEN_NO_SYMM_MEM", None)
    if no_symm:
        env["LUTGEN_NO_SYMM_MEM"] = "1"
    port = 29600 + (os.getpid() % 200) + (50 if no_symm else 0)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "tools", "multi_check.py")],
                       capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0 and "MULTI_CHECK_OK" in p.stdout, (p.stdout[-3000:], p.stderr[-3000:])
