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


def test_phase_constants_match_the_header():
    """facl_b200/dist.py issues facl_train_step in phases: its constants must be the FACL_PHASE_* macros of the header, and the
    finer cuts must be disjoint bits (the C side tests them with `&`)."""
    from facl_b200 import dist
    text = open(os.path.join(ROOT, "include", "facl_b200.h")).read()
    macros = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+FACL_PHASE_(\w+)\s+(\d+)", text)}
    names = ["FORWARD", "LOSS", "BACKWARD", "UPDATE", "BACKWARD_HEAD", "BACKWARD_L1", "FORWARD_X", "FORWARD_G", "BACKWARD_HEAD_G",
             "BACKWARD_HEAD_X"]
    for n in names:
        assert macros[n] == getattr(dist, "PHASE_" + n), n
    bits = [macros[n] for n in names]
    assert all(b & (b - 1) == 0 for b in bits) and len(set(bits)) == len(bits)
    assert macros["ALL"] == macros["FORWARD"] | macros["LOSS"] | macros["BACKWARD"] | macros["UPDATE"]
