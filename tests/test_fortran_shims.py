"""The Fortran side of the drop-in boundary (fortran/*.F90) — what can be checked WITHOUT a Fortran compiler (the image and the
GPU box have none: profiles/r02_gpu_box_probe.txt):
  * the generated files are what tools/gen_fortran_shims.py produces from the current header (no drift);
  * every bind(C) name is a function of include/padeops_b200.h AND an exported symbol of the built library, and every header
    function has an interface (header symbols == shim bind(C) names);
  * plain free-form Fortran: no line beyond 132 characters, no multi-statement cpp macros, balanced program units;
  * where /root/reference exists: every PUBLIC type-bound procedure / generic of the reference's types on the hot path is
    offered by the shim type of the same module and name, so a caller compiles against either."""
import importlib.util
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FDIR = os.path.join(ROOT, "fortran")
REF = "/root/reference/src"


def _gen():
    spec = importlib.util.spec_from_file_location("gen_fortran_shims", os.path.join(ROOT, "tools", "gen_fortran_shims.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _files():
    return sorted(f for f in os.listdir(FDIR) if f.endswith(".F90"))


def test_generated_files_are_in_step_with_the_header():
    G = _gen()
    for fn, gen in G.FILES.items():
        assert open(os.path.join(FDIR, fn)).read() == gen(), f"fortran/{fn} is stale: run python tools/gen_fortran_shims.py"


def test_bind_c_names_equal_header_symbols_and_library_exports():
    G = _gen()
    header = {name for _, name, _ in G.prototypes()}
    src = re.sub(r"&\n\s*", "", open(os.path.join(FDIR, "padeops_b200_c.F90")).read())     # join continuation lines
    binds = set(re.findall(r'bind\(C, name="(pdo_\w+)"\)', src))
    assert binds == header, (sorted(header - binds), sorted(binds - header))
    import padeops_b200
    out = subprocess.run(["nm", "-D", "--defined-only", padeops_b200.library_path()], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert not (binds - exported), sorted(binds - exported)
    # every call a shim module makes goes through a declared interface
    for fn in _files():
        if fn == "padeops_b200_c.F90":
            continue
        for name in set(re.findall(r"\b(pdo_\w+)\s*\(", open(os.path.join(FDIR, fn)).read())):
            assert name in binds, (fn, name)


def test_plain_free_form_fortran():
    for fn in _files():
        txt = open(os.path.join(FDIR, fn)).read()
        for i, l in enumerate(txt.splitlines(), 1):
            assert len(l) <= 132, (fn, i, len(l))
        assert "#define" not in txt and "\\\n" not in txt, f"{fn}: cpp macros expand to over-long lines with default compiler options"
        code = "\n".join(l.split("!")[0] for l in txt.splitlines())
        for unit in ("module", "subroutine", "function", "type", "interface"):
            if unit == "type":
                opens = len(re.findall(r"^\s*type\s*(,\s*(public|bind\(C\)))?\s*(::)?\s*\w+\s*$", code, flags=re.M | re.I))
            elif unit == "module":
                opens = len(re.findall(r"^\s*module\s+(?!procedure)\w+", code, flags=re.M | re.I))
            elif unit == "interface":
                opens = len(re.findall(r"^\s*interface\b", code, flags=re.M | re.I))
            else:
                opens = len(re.findall(rf"^\s*(pure\s+)?{unit}\s+\w+\s*\(", code, flags=re.M | re.I)) + \
                    len(re.findall(rf"^\s*{unit}\s+\w+\s*$", code, flags=re.M | re.I))
            closes = len(re.findall(rf"^\s*end\s+{unit}\b", code, flags=re.M | re.I))
            assert opens == closes, (fn, unit, opens, closes)


# reference file, type -> shim file; the procedures below are the PUBLIC, in-scope ones (private helpers and the branches
# SURVEY.md marks out of scope — upsampling, test filters, z-base FFTs, ... — are not part of the boundary)
SCOPE = {
    ("derivatives/cd10.F90", "cd10", "cd10.F90"): ["init", "destroy", "GetSize", "dd1", "dd2", "dd3", "d2d1", "d2d2", "d2d3"],
    ("derivatives/cd06.F90", "cd06", "cd06.F90"): ["init", "destroy", "GetSize", "dd1", "dd2", "dd3"],
    ("filters/cf90.F90", "cf90", "cf90.F90"): ["init", "destroy", "filter1", "filter2"],
    ("filters/gaussian.F90", "gaussian", "gaussian.F90"): ["init", "destroy", "filter1", "filter2"],
    ("derivatives/cd06stagg.F90", "cd06stagg", "cd06stagg.F90"): ["destroy", "ddz_E2C", "ddz_C2E", "ddz_E2E", "ddz_C2C", "d2dz2_E2E", "d2dz2_C2C",
                                                                   "InterpZ_E2C", "InterpZ_C2E"],
    ("utilities/fft_3d.F90", "fft_3d", "fft_3d.F90"): ["init", "fft3_x2z", "ifft3_z2x", "ifft2_y2x", "fft2_x2y", "destroy", "get_complex_output_size",
                                                       "get_complex_output_start_end_indices"],
    ("incompressible/spectral.F90", "spectral", "spectral.F90"): ["init", "destroy", "fft", "ifft", "dealias", "dealias_edgeField", "mTimes_ik1_oop",
                                                                  "mTimes_ik1_ip", "mTimes_ik2_oop", "mTimes_ik2_ip", "take_fft1d_z2z_ip", "take_ifft1d_z2z_ip",
                                                                  "shiftz_E2C", "shiftz_C2E", "ddz_C2C_real_inplace", "ddz_C2C_complex_inplace"],
    ("incompressible/PadeDerOps.F90", "Pade6stagg", "PadeDerOps.F90"): ["init", "destroy", "getModifiedWavenumbers", "ddz_C2E", "ddz_E2C", "d2dz2_C2C",
                                                                        "d2dz2_E2E", "interpz_C2E", "interpz_E2C"],
    ("incompressible/PadePoisson.F90", "padepoisson", "PadePoisson.F90"): ["init", "PressureProjection", "destroy", "DivergenceCheck", "getPressure",
                                                                           "getPressureAndUpdateRHS"],
    ("incompressible/igrid_operators_periodic.F90", "Ops_Periodic", "igrid_operators_periodic.F90"): ["init", "destroy", "ddx", "ddy", "ddz", "ddz_cmplx2cmplx",
                                                                                                       "ReadField3D", "WriteField3D", "allocate3Dfield", "SolvePoisson_oop",
                                                                                                       "SolvePoisson_ip", "dealiasField", "link_spect"],
    ("utilities/PoissonPeriodic.F90", "PoissonPeriodic", "PoissonPeriodic.F90"): ["init", "poisson_solve", "destroy"],
}


def _bound_names(path, typename):
    """public type-bound procedure and generic names of `type typename` in a Fortran source"""
    txt = open(path, errors="replace").read()
    m = re.search(rf"^\s*type\s*(,\s*\w+\s*)?(::)?\s*{typename}\s*$(.*?)^\s*end\s+type", txt, flags=re.M | re.S | re.I)
    assert m, (path, typename)
    names = set()
    for l in m.group(3).splitlines():
        l = l.split("!")[0]
        g = re.match(r"\s*(procedure|generic)\s*(,\s*(private|public)\s*)?::\s*(\w+)", l, flags=re.I)
        if g and (g.group(3) or "").lower() != "private":
            names.add(g.group(4).lower())
    return names


@pytest.mark.parametrize("key", sorted(SCOPE))
def test_shim_types_offer_the_reference_procedures(key):
    ref_file, typename, shim = key
    mine = _bound_names(os.path.join(FDIR, shim), typename)
    want = [n.lower() for n in SCOPE[key]]
    assert not [n for n in want if n not in mine], (shim, [n for n in want if n not in mine])
    if os.path.exists(os.path.join(REF, ref_file)):      # the build container: the names really are the reference's public ones
        theirs = _bound_names(os.path.join(REF, ref_file), typename)
        assert not [n for n in want if n not in theirs], (ref_file, [n for n in want if n not in theirs])
        m1 = re.search(r"^\s*module\s+(\w+)", open(os.path.join(REF, ref_file), errors="replace").read(), flags=re.M | re.I).group(1)
        m2 = re.search(r"^\s*module\s+(\w+)", "\n".join(l for l in open(os.path.join(FDIR, shim)).read().splitlines() if not l.lstrip().startswith("!")),
                       flags=re.M | re.I).group(1)
        assert m1.lower() == m2.lower(), (m1, m2)
