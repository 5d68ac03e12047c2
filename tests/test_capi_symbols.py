"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/smilei_b200.h declares; without a GPU the entry points fail loudly (no fallback); and the C++
adapter header type-checks against the reference's own operator base classes."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "smilei_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sb200_[A-Za-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from smilei_b200 import capi
    lib = capi.lib()
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/smilei_b200.h but not exported"
    assert set(capi.SYMBOLS) == set(syms), set(capi.SYMBOLS) ^ set(syms)
    assert lib.sb200_abi_version() == 1


def test_no_cpu_fallback():
    """Without a CUDA device creating a patch fails with an error message; nothing computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import smilei_b200
    with pytest.raises(smilei_b200.SmileiB200Error) as e:
        smilei_b200.Patch((8, 8, 8), (0.1, 0.1, 0.1), 0.05)
    assert "sb200_patch_create" in str(e.value)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under smilei_b200/ or include/ may reference it."""
    bad = []
    for base in ("smilei_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"oracle_lib|oracle_patch|libsmilei_oracle|libsmilei_ref|oracle/", txt):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference headers not present")
def test_adapter_header_compiles_against_reference_headers():
    inc = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle", "ref_build", "stubs")]
    for d in sorted(os.listdir("/root/reference/src")):
        p = os.path.join("/root/reference/src", d)
        if os.path.isdir(p):
            inc.append("-I" + p)
    pyinc = subprocess.run(["python3-config", "--includes"], capture_output=True, text=True).stdout.split()
    src = '#include "smilei_b200_operators.hpp"\n' \
          'void use( Params &p, Patch *pt, Species *s ) {\n' \
          '  smilei_b200::Interpolator3D2OrderB200 i( p, pt ); smilei_b200::PusherB200 pu( p, s );\n' \
          '  smilei_b200::Projector3DB200 pr( p, pt ); smilei_b200::MA_Solver3D_B200 ma( p ); smilei_b200::MF_Solver3D_B200 mf( p, pt );\n' \
          '  smilei_b200::ElectroMagnBC3D_SM_B200 sm( p, pt, 0 ); smilei_b200::forward_particle_bc( pt, s, 0 );\n' \
          '  smilei_b200::Bridge::attach( p, pt, 2, 0 ); }\n'
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-w", "-x", "c++", "-"] + inc + pyinc, input=src,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
