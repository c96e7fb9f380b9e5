"""CPU: the C-ABI library loads and exports every symbol include/padeops_b200.h declares; the
product refuses to compute without a GPU (no CPU fallback); host-side error codes follow the reference."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "padeops_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pdo_[a-z0-9_A-Z]+)\s*\(", src)))


def test_header_symbols_are_exported(pdo):
    from padeops_b200 import _lib
    L = pdo.lib()
    declared = _declared_symbols()
    assert len(declared) >= 75
    for s in declared:
        assert hasattr(L, s), f"{s} declared in include/padeops_b200.h but not exported"
    assert set(declared) == set(_lib.EXPORTED), set(declared) ^ set(_lib.EXPORTED)


def test_product_library_exports_the_header_and_nothing_else_in_c(pdo):
    """The C symbols of libpadeops_b200.so are exactly the functions of include/padeops_b200.h: no pdo_debug_* test hook ships in
    the product (they live in libpadeops_b200_testhooks.so, wrappers around pdo::hooks::* C++ functions), and the hooks library
    exports nothing but those wrappers."""
    import subprocess
    from padeops_b200 import _lib

    def c_exports(path):
        out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout
        return {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("pdo_")}
    prod = c_exports(pdo.library_path())
    assert not [s for s in prod if s.startswith("pdo_debug")], sorted(s for s in prod if s.startswith("pdo_debug"))
    assert prod == set(_declared_symbols()), sorted(prod ^ set(_declared_symbols()))
    hooks = c_exports(_lib._HOOKS_SO)
    assert hooks and all(s.startswith("pdo_debug_") for s in hooks)
    assert hooks == {k for k in _lib._PROTOS if k.startswith("pdo_debug")}


def test_no_cpu_fallback(pdo):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    op = pdo.cd10()
    assert op.init(64, 0.1) == 1004  # PDO_E_NODEVICE
    assert b"no CPU fallback" in pdo.lib().pdo_last_error()


def test_init_error_codes_match_reference(pdo):
    # these checks run before the device probe, like the Fortran init functions they mirror
    assert pdo.cd10().init(5, 0.1) == 2      # cd10.F90:224
    assert pdo.cd06().init(4, 0.1) == 3      # cd06.F90:158
    assert pdo.cf90().init(9) == 7           # cf90.F90:128
    assert pdo.cd06().init(16, 0.1, periodic_=False, bc1_=1) == 1002  # the reference marks cd06's symmetric closures "Incomplete"
    assert pdo.cd06().init(5, 0.1, periodic_=False) == 3
    assert pdo.cd10().init(16, 0.1, periodic_=False, bc1_=2) == 324   # cd10.F90:2044-2046
    assert pdo.cd10().init(5, 0.1, periodic_=False) == 2
    with pytest.raises(pdo.PadeOpsError) as e:
        pdo.cd06stagg().init(4, 0.1)
    assert e.value.code == 21                # cd06stagg.F90:182-184


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (tier rule)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "padeops_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{fn} mentions the oracle"


def test_fortran_shims_bind_declared_symbols(pdo):
    """Every bind(C, name=...) / macro-instantiated name in fortran/*.F90 must be an exported, declared symbol."""
    declared = set(_declared_symbols())
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "fortran")):
        txt = open(os.path.join(ROOT, "fortran", fn)).read()
        names |= set(re.findall(r'name="(pdo_[A-Za-z0-9_]+)"', txt))
        names |= set(re.findall(r"PDO_[A-Z]+_FN\((pdo_[A-Za-z0-9_]+)\)", txt))
    assert len(names) >= 50
    assert names <= declared, names - declared
