"""CPU-only: libdasm.so loads and exports every symbol include/dasm.h declares; host-side (non-compute) calls work."""
import ctypes
import os
import re

import pytest

from __graft_entry__ import ROOT, load_package


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "dasm.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dasm_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_all_declared_symbols():
    pkg = load_package()
    lib = pkg.lib()
    syms = declared_symbols()
    assert len(syms) > 40
    for s in syms:
        assert hasattr(lib, s), "missing symbol " + s


def test_decompose_balanced_matches_reference_table():
    import json
    pkg = load_package()
    rows = json.load(open(os.path.join(ROOT, "tests", "golden", "subdivided_hyper_cube_balanced_01.json")))
    for s, nref, s0, s1, s2, _ in rows:
        assert pkg.decompose_balanced(s) == (nref, [s0, s1, s2])


def test_no_cpu_fallback_without_gpu():
    import torch
    pkg = load_package()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.DasmError):
        pkg.Context(0)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dealii-asm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "dasm_oracle" not in txt and "oracle/" not in txt, f
