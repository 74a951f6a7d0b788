"""CPU: libfacl_b200.so loads and exports exactly the symbols include/facl_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from facl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "facl_b200.h")).read()
    return sorted(set(re.findall(r"FACL_API\s+[\w\s\*]+?\b(facl_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    declared = _declared_symbols()
    assert declared, "no symbols parsed from the header"
    assert sorted(_lib.SIGNATURES) == declared


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build the library first: make -C facl_b200/csrc"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(handle, name), name
    assert b"sm_100a" in _lib.lib().facl_version()


def test_no_cpu_fallback():
    import pytest
    import torch
    from facl_b200 import ops
    with pytest.raises(_lib.FaclError):
        ops.group_points_raw(torch.zeros(1, 64, 4), 8, 8, 0.1)
