"""The GPU tests written after the round's last GPU session (every gpu test that carries xfail(strict=False)) run HERE, in child
processes, one per test file: device code that has never seen hardware must not be able to take the validated suite down with it
(a faulting kernel poisons the CUDA context of its process for good), nor one unvalidated area another.  The children run them
with --runxfail, i.e. as ordinary tests; these wrappers are themselves non-strict xfail, so the outcome is reported without
turning a green suite red.  tools/late_tests.sh does the same by hand and keeps the full log."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["test_golden.py", "test_igrid_gpu.py", "test_nonperiodic_gpu.py", "test_ops_periodic_gpu.py", "test_spectral_gpu.py",
         "test_stagg_nonperiodic_gpu.py", "test_vecops_gpu.py", "test_multigpu.py"]


@pytest.mark.gpu
@pytest.mark.late_wrapper
@pytest.mark.xfail(strict=False, reason="first hardware run of the tests added after the round's last GPU session")
@pytest.mark.parametrize("fname", FILES)
def test_late_gpu_tests_in_a_child_process(fname):
    env = dict(os.environ, PDO_RUN_LATE="1")
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", fname), "-m", "gpu", "--runxfail", "-q", "-rf", "-p", "no:cacheprovider"]
    # own session: on a hang the whole process group goes (torchrun workers of the multi-GPU file included), nothing keeps a GPU
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, cwd=ROOT, start_new_session=True)
    try:
        out, _ = p.communicate(timeout=300)      # the late tests are small: a hang must not eat the GPU tier
    except subprocess.TimeoutExpired:
        import signal
        os.killpg(p.pid, signal.SIGKILL)
        out, _ = p.communicate()
        raise AssertionError("timed out\n" + out[-4000:])
    assert p.returncode in (0, 5), out[-8000:]   # 5: nothing collected (every late test of the file has been promoted)


def test_every_late_file_is_listed():
    """CPU: a late GPU test in a file the wrapper does not know would never run anywhere."""
    import re
    for fn in sorted(os.listdir(os.path.join(ROOT, "tests"))):
        if not fn.startswith("test_") or not fn.endswith(".py") or fn == "test_late_isolated.py":
            continue
        src = open(os.path.join(ROOT, "tests", fn)).read()
        if re.search(r"pytest\.mark\.xfail\(", src) and "pytest.mark.gpu" in src:
            assert fn in FILES, fn
