"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol include/*.h declares
(no compute calls without a GPU), the product never touches oracle/, and missing-GPU use fails loudly."""
import ctypes
import os
import re

import pytest
import torch

import azg_b200
from azg_b200 import _native as nat

ROOT = os.path.realpath(os.path.join(os.path.dirname(__file__), ".."))


def declared_symbols():
    syms = set()
    inc = os.path.join(ROOT, "include")
    for f in os.listdir(inc):
        if f.endswith(".h"):
            text = open(os.path.join(inc, f)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            syms |= set(re.findall(r"\b((?:spl|mcts)_[a-z0-9_]+)\s*\(", text))
    return syms


def test_library_exports_every_declared_symbol():
    nat.build()
    lib = ctypes.CDLL(nat.LIB)
    decl = declared_symbols()
    assert len(decl) >= 20
    for s in decl:
        assert hasattr(lib, s), f"{s} declared in include/ but not exported"
    assert set(nat.EXPORTS) <= decl
    assert lib.spl_abi_version() == 1


def test_size_helpers_match_reference_shapes():
    lib = nat.lib()
    for n, rows, bytes_ in ((2, 56, 392), (3, 71, 497), (4, 88, 616)):   # observation_size, SplendorLogicNumba.py:25-27
        assert lib.spl_state_rows(n) == rows and lib.spl_state_bytes(n) == bytes_
        assert azg_b200.observation_size(n) == (rows, 7)
    assert lib.spl_lanes_padded(1) == 32 and lib.spl_lanes_padded(33) == 64
    assert lib.spl_planes_bytes(2, 33) == 64 * 392 and lib.spl_mask_planes_bytes(33) == 64 * 52
    assert azg_b200.action_size() == 406


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "alphazero-general-ori_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle/refgen/gen_tables.py", ""), f"{f} mentions the oracle"


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError):
        azg_b200.SplendorEnv(2, 64)
    h = ctypes.c_void_p()
    assert nat.lib().spl_ctx_create(2, 10, 7, 0, ctypes.byref(h)) == -3   # SPL_E_NOGPU
    with pytest.raises(RuntimeError):
        azg_b200.SplendorGame(2)
