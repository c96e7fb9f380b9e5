"""The drop-in boundary exercised from compiled C (tests/c/c_abi_smoke.c): a plain C99 program includes include/padeops_b200.h, links
libpadeops_b200.so and calls it the way the Fortran shim modules do (by-value scalars, void* fields, host and device pointers,
int status codes) — no Python, no ctypes, no CUDA headers in the caller.
CPU: it compiles and links against the built library and every public symbol it uses resolves.  GPU: it runs and passes."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "c_abi_smoke.c")


def _build(tmpdir):
    import padeops_b200
    so = padeops_b200.library_path()
    assert os.path.exists(so), "build the library first (python -c 'import __graft_entry__ as g; g.build()')"
    exe = os.path.join(str(tmpdir), "c_abi_smoke")
    libdir = os.path.dirname(so)
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", libdir, "-lpadeops_b200", "-lm", f"-Wl,-rpath,{libdir}", "-Wl,--no-as-needed"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no C compiler")
def test_c_caller_compiles_and_links_against_the_header(tmp_path):
    exe = _build(tmp_path)
    # the header is valid C99 on its own and the program's undefined pdo_* symbols all come from the library
    r = subprocess.run(["nm", "-u", exe], capture_output=True, text=True)
    used = sorted(l.split()[-1].split("@")[0] for l in r.stdout.splitlines() if " pdo_" in l or l.strip().startswith("U pdo_"))
    assert len(used) >= 20, used
    assert not [s for s in used if s.startswith("pdo_debug")], "the C caller must stay on the public ABI"


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("gcc") is None, reason="no C compiler")
def test_c_caller_runs_on_the_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "C_ABI_SMOKE PASS" in r.stdout, (r.returncode, r.stdout[-2000:], r.stderr[-4000:])
