"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/mray_b200.h
declares; without a CUDA device the product path refuses to run (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from mray_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "mray_b200.h")).read()
    return sorted(set(re.findall(r"MRB_API\s+[\w\s\*]+?\b(mrb_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    syms = header_symbols()
    assert len(syms) >= 15
    assert sorted(capi._PROTOTYPES.keys()) == syms


def test_library_exports_every_declared_symbol():
    if not os.path.exists(capi.LIB_PATH):
        capi.build_library()
    lib = ctypes.CDLL(capi.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), f"{s} not exported"
    # the header, the library and the ctypes mirror agree on the ABI version
    header = open(os.path.join(ROOT, "include", "mray_b200.h")).read()
    m = re.search(r"#define MRB_ABI_VERSION \(\((\d+)u << 16\) \| (\d+)u\)", header)
    assert m, "MRB_ABI_VERSION not found in the header"
    assert lib.mrb_abi_version() == (int(m.group(1)) << 16 | int(m.group(2))) == capi.MRB_ABI_VERSION


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.MrbError) as e:
        capi.Context(0)
    assert e.value.status == -1  # MRB_ERR_NO_DEVICE


def test_product_sources_never_touch_the_oracle():
    """The product (mray_b200/, include/) must not import, link or load anything under oracle/."""
    bad = []
    for base in ("mray_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"oracle_lib|liboracle|mray_oracle|oracle/_ref|libref_taps", txt):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
