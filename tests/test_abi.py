"""CPU: the C-ABI library builds, loads and exports every symbol include/vlaser_b200.h declares;
without a GPU every compute entry point fails loudly instead of falling back."""
from __future__ import annotations

import ctypes as C
import os
import re

import pytest
import torch

from vlaser_b200 import _lib, build
from vlaser_b200.config import tiny

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "vlaser_b200.h")).read()
    declared = set(re.findall(r"\b(vlaser_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"vlaser_handle"}
    assert declared, "no declarations parsed"
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in the header but not exported"
    assert declared == set(_lib.EXPORTS)
    assert lib.vlaser_abi_version() == 1


def test_cfg_struct_layout_matches_header():
    hdr = open(os.path.join(ROOT, "include", "vlaser_b200.h")).read()
    body = hdr[hdr.index("typedef struct vlaser_cfg {"): hdr.index("} vlaser_cfg;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in re.findall(r"(?:int32_t|float)\s+([^;]+);", body):
        names += [n.strip() for n in decl.split(",")]
    assert names == [n for n, _ in _lib.VlaserCfg._fields_]
    assert C.sizeof(_lib.VlaserCfg) == 4 * len(names)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    from vlaser_b200.policy import _to_c_cfg, PiZeroInference
    h = C.c_void_p()
    cfg = _to_c_cfg(tiny(), 2)
    rc = lib.vlaser_create(C.byref(cfg), C.byref(h))
    assert rc < 0 and b"no CUDA device" in lib.vlaser_last_error()
    with pytest.raises(RuntimeError):
        PiZeroInference(tiny(), max_batch=2)
