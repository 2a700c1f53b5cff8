"""CPU-only checks: the C-ABI library loads and exports every symbol include/lancet_b200.h declares
(no compute calls without a GPU), and creating a context without a GPU fails loudly."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "lancet_b200.h")).read()
    return sorted(set(re.findall(r"\b(lb2_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    so = os.path.join(ROOT, "lancet_b200", "_lb2.so")
    assert os.path.exists(so), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(so)
    names = _declared()
    assert {"lb2_create", "lb2_process", "lb2_upload", "lb2_run", "lb2_download", "lb2_destroy"} <= set(names)
    for n in names:
        assert hasattr(lib, n), n


def test_every_exported_symbol_is_declared():
    """the other direction: nothing is exported from the library that the header does not declare"""
    import subprocess
    so = os.path.join(ROOT, "lancet_b200", "_lb2.so")
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    exported = sorted({ln.split()[-1] for ln in out.splitlines() if ln.split()[-1].startswith("lb2_") and " T " in ln})
    declared = set(_declared())
    assert exported, "no lb2_* symbols exported?"
    missing = [n for n in exported if n not in declared]
    assert not missing, f"exported but not declared in include/lancet_b200.h: {missing}"


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    from lancet_b200.api import Context
    try:
        Context(device=0)
    except RuntimeError as e:
        assert "CUDA" in str(e) or "fallback" in str(e)
    else:
        raise AssertionError("Context() must not succeed without a GPU")


def test_struct_sizes_match_numpy_views():
    from lancet_b200.api import VARIANT_DTYPE, WINDOW_DTYPE
    assert VARIANT_DTYPE.itemsize == 40 and WINDOW_DTYPE.itemsize == 16
