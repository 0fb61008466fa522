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


@pytest.mark.parametrize("bad", [3, 700_001, 1_199_999])
def test_large_batches_are_validated_on_several_host_threads(agx, bad):
    """2^19 and more (offset, length) entries are checked in slices on four host threads; the message still names
    the first offending sequence, for the score and the alignment entry points alike (no device work happens)."""
    cap = agx.capi
    lib = cap.load_library()
    n_seq = 1_200_000
    buf = np.zeros(64, dtype=np.uint8)
    off = np.zeros(n_seq, dtype=np.int64)
    ln = np.full(n_seq, 8, dtype=np.int32)
    off[bad] = 60                                    # 60 + 8 > 64
    off[-1] = 61 if bad != n_seq - 1 else 60         # a later offender must not be the one reported
    for call in (cap.sw_score_flat, cap.sw_ends_flat, cap.sw_align_flat):
        with pytest.raises(cap.AgxError) as e:
            call(buf, off, ln)
        assert e.value.code == -1
        assert f"sequence {bad} lies outside".encode() in lib.agx_last_error()
    ln[bad] = -1
    off[bad] = 0
    with pytest.raises(cap.AgxError) as e:
        cap.sw_score_flat(buf, off, ln)
    assert e.value.code == -1 and f"sequence {bad} ".encode() in lib.agx_last_error()
