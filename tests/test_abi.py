"""No-GPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/dfot_b200.h declares, and rejects bad arguments without touching a device."""
import ctypes
import os
import re

from dfot_b200 import _abi
from helpers import ROOT


def declared_symbols():
    with open(os.path.join(ROOT, "include", "dfot_b200.h")) as f:
        src = f.read()
    return sorted(set(re.findall(r"DFOT_API\s+[\w\s\*]+?\b(dfot_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _abi.lib()
    names = declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dfot_b200.h but not exported"
    assert sorted(_abi.SYMBOLS) == names
    assert lib.dfot_abi_version() == 1


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_abi.FrameUpdate) == 24
    assert ctypes.sizeof(_abi.FramePrepare) == 16
    assert ctypes.sizeof(_abi.GemmEpilogue) == 168


def test_argument_validation_is_host_side():
    lib = _abi.lib()
    # null pointers / bad sizes are rejected before any CUDA call
    assert lib.dfot_sampler_step_hg(None, None, 0, None, 0, None, None, None, None, None, 1, 1, 1, 4, None) == -1
    assert b"sampler_step_hg" in lib.dfot_last_error()
    assert lib.dfot_attention(None, None, 1, 1, 1, 64, None) == -1
    e = _abi.GemmEpilogue()
    assert lib.dfot_gemm_bf16(None, 8, None, 8, None, 8, 1, 8, 8, 0, ctypes.byref(e), None) == -1
    assert lib.dfot_adaln_layernorm(None, None, 0, 0, 0, None, None, 1, 8, 1, 1e-6, None) == -1
