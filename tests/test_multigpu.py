"""Multi-GPU parity (needs >= 2 GPUs on the box: `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _keep_log(kind, world, r):
    """the workers' own report (one line per failed check, the p2p state, the final verdict) is kept next to the other GPU-run
    artefacts so that a multi-GPU session leaves evidence behind (profiles/ holds the committed copies)"""
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, f"mp_worker_{kind}_w{world}.log"), "w") as f:
            f.write(r.stdout[-20000:] + "\n---- stderr ----\n" + r.stderr[-4000:])


@pytest.mark.parametrize("world", [2, 4, 8])
def test_transposes_fft_poisson_multi_gpu(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29510 + world), os.path.join(ROOT, "tests", "mp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    _keep_log("main", world, r)
    assert "MP_WORKER_RESULT PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("world", [2, 4])
def test_late_sections_multi_gpu(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29530 + world), os.path.join(ROOT, "tests", "mp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, PDO_MP_LATE="1"))
    _keep_log("widened", world, r)
    assert "MP_WORKER_LATE PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
