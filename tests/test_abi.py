"""CPU tests of the drop-in boundary: libcobsgpu.so loads, exports every symbol that
include/cobsgpu.h declares, and fails loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import cobs_b200
from cobs_b200 import _lib
from conftest import ROOT, golden_path


def header_symbols():
    src = open(os.path.join(ROOT, "include", "cobsgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cobsgpu_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = cobs_b200.lib()
    names = header_symbols()
    assert len(names) == 28
    for n in names:
        assert hasattr(L, n), n
    # and the ctypes table covers the header exactly
    assert sorted(_lib.SYMBOLS) == names
    assert L.cobsgpu_version() == 2


def test_exported_symbols_are_plain_c():
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH]).decode()
    exported = [l.split()[-1] for l in out.splitlines() if " T " in l]
    ours = [s for s in exported if s.startswith("cobsgpu_")]
    assert sorted(ours) == header_symbols()
    # no torch / pybind symbols leak through the boundary
    assert not [s for s in exported if "torch" in s.lower() or "pybind" in s.lower()]


def test_struct_layouts_match_header():
    # sizes implied by include/cobsgpu.h (natural alignment, 64-bit)
    assert C.sizeof(_lib.IndexDesc) == 80
    assert C.sizeof(_lib.IndexInfo) == 104
    assert C.sizeof(_lib.Result) == 24
    assert C.sizeof(_lib.Timers) == 72


@pytest.mark.skipif(cobs_b200.lib().cobsgpu_device_count() > 0, reason="GPU present")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(cobs_b200.CobsGpuError) as e:
        cobs_b200.GpuIndex.procedural(cobs_b200.KIND_CLASSIC, 100, [1000], 3)
    assert e.value.code == _lib.ERR_CUDA
    assert "no CPU fallback" in e.value.msg
    with pytest.raises(cobs_b200.CobsGpuError) as e:
        cobs_b200.GpuIndex.open_file(golden_path("all160.cobs_classic"))
    assert e.value.code == _lib.ERR_CUDA


def test_bad_files_are_rejected_before_touching_the_gpu(tmp_path):
    p = tmp_path / "bad.cobs_classic"
    p.write_bytes(b"COBS:NOT_AN_INDEX" + b"\0" * 64)
    with pytest.raises(cobs_b200.CobsGpuError) as e:
        cobs_b200.GpuIndex.open_file(str(p))
    assert e.value.code == _lib.ERR_BAD_FILE
    with pytest.raises(cobs_b200.CobsGpuError) as e:
        cobs_b200.GpuIndex.open_file(str(tmp_path / "missing.cobs_classic"))
    assert e.value.code == _lib.ERR_IO
    with pytest.raises(cobs_b200.CobsGpuError) as e:
        d = _lib.IndexDesc()
        h = C.c_void_p()
        _lib.check(cobs_b200.lib().cobsgpu_index_open(C.byref(d), C.byref(h)))
    assert e.value.code == _lib.ERR_INVALID_ARG


def test_product_does_not_import_the_oracle():
    """the oracle is test infrastructure: nothing under cobs_b200/ may reference it"""
    bad = []
    walk = list(os.walk(os.path.join(ROOT, "cobs_b200"))) + list(os.walk(os.path.join(ROOT, "cobs_index")))
    for dirpath, _, files in walk:
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle|#include\s+[\"<].*oracle|liboracle|libcobs_ref",
                             txt, flags=re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
