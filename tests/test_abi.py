"""The C-ABI shared library loads and exports every symbol include/sph_b200.h declares; without a GPU the
product fails loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "sph_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sph_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for s in ("sph_create", "sph_upload", "sph_step", "sph_run_until", "sph_download", "sph_download_diag",
              "sph_counters", "sph_destroy", "sph_evaluate", "sph_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol(built_engine):
    lib = C.CDLL(built_engine)
    for s in declared_symbols():
        assert hasattr(lib, s), f"libsph_b200.so does not export {s}"


def test_struct_layout_matches_header(built_engine):
    from summersph_b200 import SphParams, SphCounts, default_params, MODE_VARIABLE_H
    assert C.sizeof(SphParams) == 4 * 4 + 10 * 8 + 2 * 4
    assert C.sizeof(SphCounts) == 8 * 8
    lib = C.CDLL(built_engine)
    p = SphParams()
    assert lib.sph_default_params(MODE_VARIABLE_H, C.byref(p)) == 0
    q = default_params(MODE_VARIABLE_H)
    for k, _ in SphParams._fields_:
        assert getattr(p, k) == getattr(q, k), k


def test_no_cpu_fallback(built_engine):
    """On a box without a CUDA device the engine refuses to start (SPH_ERR_NO_DEVICE), it never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from summersph_b200 import default_params
    from summersph_b200.engine import Engine, SphError
    with pytest.raises(SphError) as ei:
        Engine(default_params())
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    """Nothing under summersph_b200/ or host/ imports, includes, links or dlopens anything under oracle/."""
    pat = re.compile(r"(^\s*(from|import)\s+oracle\b)|(#include\s+\".*oracle)|(libsph_oracle)|(oracle/)|(orc_[a-z_]+\s*\()", re.M)
    for top in ("summersph_b200", "host"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".f90", "Makefile")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    txt = re.sub(r"oracle/oracle\.py\)", "", txt)      # one docstring mention in _abi.py
                    m = pat.search(txt)
                    assert m is None, f"{os.path.join(dirpath, f)} references the oracle: {m.group(0)!r}"
