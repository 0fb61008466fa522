"""The C-ABI shared library: it loads, exports every symbol include/agx.h declares, and refuses to
compute without a GPU (no CPU fallback).  No compute calls here."""
import ctypes
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def header_functions():
    text = (ROOT / "include" / "agx.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(?:int|int32_t|int64_t|void|double|const char \*)\s*\*?\s*(\w+)\s*\(", text)
    return sorted(set(n for n in names if n.startswith(("agx_", "sw_", "pairhmm_"))))


def test_header_and_binding_agree(agx):
    assert header_functions() == sorted(agx.capi.SYMBOLS)


def test_library_exports_every_declared_symbol(agx):
    lib = agx.capi.load_library()
    for name in header_functions():
        assert hasattr(lib, name), f"libagx.so does not export {name}"
    assert agx.capi.version().endswith("sm_100a")


def test_library_carries_sm100a_code_only(agx):
    out = subprocess.run(["cuobjdump", "-lelf", str(agx.capi.LIB_PATH)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu(agx):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cap = agx.capi
    with pytest.raises(cap.AgxError) as e:
        cap.sw_score_batch([b"ACGT\n"], [b"ACGT\n"])
    assert e.value.code == -2          # AGX_ENODEVICE
    with pytest.raises(cap.AgxError) as e:
        cap.pairhmm_forward_batch([(b"ACGT", b"IIII", b"IIII", b"IIII", b"++++")], [b"ACGT"])
    assert e.value.code == -2
    assert cap.device_count() == 0


def test_argument_validation(agx):
    cap = agx.capi
    lib = cap.load_library()
    # bad arguments are rejected before any device work
    buf = np.frombuffer(b"ACGT\nACGT\n", dtype=np.uint8)
    off = np.array([0, 50], dtype=np.int64)          # second sequence outside the buffer
    ln = np.array([5, 5], dtype=np.int32)
    with pytest.raises(cap.AgxError) as e:
        cap.sw_score_flat(buf, off, ln)
    assert e.value.code == -1 and b"outside" in lib.agx_last_error()
    assert cap.sw_score_flat(buf, off[:0], ln[:0]).size == 0       # empty batch is a no-op
    assert lib.agx_launch_count() >= 0
