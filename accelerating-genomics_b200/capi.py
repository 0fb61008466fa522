"""ctypes binding of libagx.so -- the C ABI declared in include/agx.h.

This is the Python host-side mirror used by the tests and bench.py; the C drivers under drivers/
bind the same symbols directly.  There is no CPU fallback: if libagx.so is missing or no B200 is
visible every compute call raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Iterable, Optional, Sequence

import numpy as np

import os

PKG = Path(__file__).resolve().parent
# AGX_LIB_PATH selects an alternative build of the same library (kernel-tuning experiments only)
LIB_PATH = Path(os.environ["AGX_LIB_PATH"]) if os.environ.get("AGX_LIB_PATH") else PKG / "libagx.so"

# every symbol include/agx.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "agx_init", "agx_init_devices", "agx_device_count", "agx_device_ordinal", "agx_device_name", "agx_shutdown",
    "agx_last_error",
    "agx_version", "agx_launch_count", "agx_reset_launch_count", "agx_set_profiling", "agx_profile_ms",
    "sw_score_batch", "sw_score_batch_flat", "sw_score_batch_device", "sw_score_file_image",
    "pairhmm_forward_batch", "pairhmm_forward_batches_flat", "pairhmm_forward_batches_device",
    "pairhmm_forward_file_image",
    "agx_pairhmm_set_gatk_mode", "agx_pairhmm_set_force_fp64", "agx_pairhmm_rescue_count",
    "sw_score_shards_device", "pairhmm_forward_shards_device",
    "sw_ends_batch_flat", "sw_align_batch_flat", "sw_ends_batch", "sw_align_batch",
    "sw_ends_batch_device", "sw_align_batch_device",
]

# the reference's scoring constants, antidiagonalSmithWaterman.c:40-43
SW_MATCH, SW_MISMATCH, SW_GAP_OPEN, SW_GAP_EXTEND = 1, -1, -3, -1


class AgxError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libagx error {code}: {msg}")
        self.code = code


_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """Load libagx.so (built in-tree by build.py).  Fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    u8p, i32p, i64p, f64p = (C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                             C.POINTER(C.c_double))
    pp = C.POINTER(C.c_void_p)
    lib.agx_init.argtypes = [C.c_int32]
    lib.agx_init_devices.argtypes = [i32p, C.c_int32]
    lib.agx_device_count.restype = C.c_int32
    lib.agx_device_ordinal.argtypes = [C.c_int32]
    lib.agx_device_ordinal.restype = C.c_int32
    lib.agx_device_name.argtypes = [C.c_int32]
    lib.agx_device_name.restype = C.c_char_p
    lib.agx_shutdown.restype = None
    lib.agx_last_error.restype = C.c_char_p
    lib.agx_version.restype = C.c_char_p
    lib.agx_launch_count.restype = C.c_int64
    lib.agx_reset_launch_count.restype = None
    lib.agx_set_profiling.argtypes = [C.c_int32]
    lib.agx_profile_ms.argtypes = [C.c_int32, C.c_int32]
    lib.agx_profile_ms.restype = C.c_double
    lib.sw_score_batch.argtypes = [pp, i32p, pp, i32p, C.c_int64] + [C.c_int32] * 4 + [i32p]
    lib.sw_score_batch_flat.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64] + \
        [C.c_int32] * 4 + [C.c_void_p]
    lib.sw_score_batch_device.argtypes = [C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                          C.c_int64] + [C.c_int32] * 4 + [C.c_void_p, C.c_void_p]
    lib.sw_score_file_image.argtypes = [C.c_void_p, C.c_int64, C.c_int32] + [C.c_int32] * 4 + \
        [C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
         C.POINTER(C.c_int32)]
    lib.sw_score_file_image.restype = C.c_int
    lib.pairhmm_forward_file_image.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                               C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    lib.pairhmm_forward_file_image.restype = C.c_int
    lib.pairhmm_forward_batch.argtypes = [C.c_int32, pp, pp, pp, pp, pp, i32p, C.c_int32, pp, i32p, f64p]
    lib.pairhmm_forward_batches_flat.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                                 C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                                 C.c_void_p, C.c_int64, C.c_void_p]
    lib.pairhmm_forward_batches_device.argtypes = [
        C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
        C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p,
        C.c_void_p]
    lib.agx_pairhmm_rescue_count.argtypes = [C.c_int32]
    lib.agx_pairhmm_rescue_count.restype = C.c_int64
    lib.sw_score_shards_device.argtypes = [C.c_void_p, C.c_int32] + [C.c_int32] * 4
    lib.sw_ends_batch_flat.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_int32] * 4 + \
        [C.c_void_p, C.c_void_p]
    lib.sw_ends_batch_flat.restype = C.c_int
    lib.sw_align_batch_flat.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_int32] * 4 + \
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    lib.sw_align_batch_flat.restype = C.c_int
    lib.sw_ends_batch_device.argtypes = [C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_int32] * 4 + \
        [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sw_ends_batch_device.restype = C.c_int
    lib.sw_align_batch_device.argtypes = [C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_int32] * 4 + \
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.c_void_p]
    lib.sw_align_batch_device.restype = C.c_int
    lib.sw_ends_batch.argtypes = [pp, i32p, pp, i32p, C.c_int64] + [C.c_int32] * 4 + [C.c_void_p, C.c_void_p]
    lib.sw_ends_batch.restype = C.c_int
    lib.sw_align_batch.argtypes = [pp, i32p, pp, i32p, C.c_int64] + [C.c_int32] * 4 + \
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    lib.sw_align_batch.restype = C.c_int
    lib.sw_score_shards_device.restype = C.c_int
    lib.pairhmm_forward_shards_device.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    lib.pairhmm_forward_shards_device.restype = C.c_int
    lib.agx_pairhmm_set_gatk_mode.argtypes = [C.c_int32]
    lib.agx_pairhmm_set_force_fp64.argtypes = [C.c_int32]
    for name in ("agx_init", "agx_init_devices", "sw_score_batch", "sw_score_batch_flat",
                 "sw_score_batch_device", "pairhmm_forward_batch", "pairhmm_forward_batches_flat",
                 "pairhmm_forward_batches_device", "agx_pairhmm_set_gatk_mode",
                 "agx_pairhmm_set_force_fp64"):
        getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def _check(rc: int) -> None:
    if rc != 0:
        raise AgxError(rc, load_library().agx_last_error().decode(errors="replace"))


# ------------------------------------------------------------------------------- runtime
def init(n_gpus: int = 0) -> int:
    _check(load_library().agx_init(int(n_gpus)))
    return device_count()


def init_devices(devices: Sequence[int]) -> int:
    arr = (C.c_int32 * len(devices))(*[int(d) for d in devices])
    _check(load_library().agx_init_devices(arr, len(devices)))
    return device_count()


def device_count() -> int:
    return int(load_library().agx_device_count())


def shutdown() -> None:
    load_library().agx_shutdown()


def version() -> str:
    return load_library().agx_version().decode()


def launch_count() -> int:
    return int(load_library().agx_launch_count())


def reset_launch_count() -> None:
    load_library().agx_reset_launch_count()


PROF_SW_DUO, PROF_SW_WAVE, PROF_HMM_STREAM, PROF_HMM_FP64, PROF_SW_CLASSIFY, PROF_HMM_CLASSIFY, PROF_SW_LONG = range(7)


def set_profiling(on: bool) -> None:
    load_library().agx_set_profiling(1 if on else 0)


def profile_ms(device: int, which: int) -> float:
    return float(load_library().agx_profile_ms(int(device), int(which)))


def set_pairhmm_gatk_mode(on) -> None:
    """0 / False: the reference's priors; 1 / True: mismatch prior Qr/3; 3: that plus GATK's base-quality floor 6"""
    _check(load_library().agx_pairhmm_set_gatk_mode(int(on)))


def pairhmm_rescue_count(device: int) -> int:
    return int(load_library().agx_pairhmm_rescue_count(int(device)))


class SwShard(C.Structure):
    _fields_ = [("device", C.c_int32), ("d_seqs", C.c_void_p), ("seqs_bytes", C.c_int64), ("d_off", C.c_void_p),
                ("d_len", C.c_void_p), ("n_pairs", C.c_int64), ("d_scores_out", C.c_void_p)]


class HmmShard(C.Structure):
    _fields_ = [("device", C.c_int32), ("d_buf", C.c_void_p), ("buf_bytes", C.c_int64), ("d_read_field_off", C.c_void_p),
                ("d_read_len", C.c_void_p), ("d_read_batch", C.c_void_p), ("d_read_out_off", C.c_void_p),
                ("n_reads", C.c_int64), ("d_hap_off", C.c_void_p), ("d_hap_len", C.c_void_p), ("n_haps", C.c_int64),
                ("d_batch_hap_start", C.c_void_p), ("n_batches", C.c_int64), ("n_pairs", C.c_int64),
                ("d_log10_out", C.c_void_p)]


def sw_score_shards_device(shards: Sequence[SwShard],
                           scoring=(1, -1, -3, -1)) -> None:
    """sw_score_shards_device: one device-resident shard per GPU, the library's in-process dispatcher."""
    arr = (SwShard * len(shards))(*shards)
    _check(load_library().sw_score_shards_device(arr, len(shards), *[int(s) for s in scoring]))


def pairhmm_forward_shards_device(shards: Sequence[HmmShard], fp64_rescue: bool = True) -> None:
    arr = (HmmShard * len(shards))(*shards)
    _check(load_library().pairhmm_forward_shards_device(arr, len(shards), 1 if fp64_rescue else 0))


def set_pairhmm_force_fp64(on: bool) -> None:
    _check(load_library().agx_pairhmm_set_force_fp64(1 if on else 0))


# ------------------------------------------------------------------------------- helpers
def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def _as(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


def _byte_ptr_array(items: Sequence[bytes]):
    """(array of void*, keep-alive list) for a sequence of bytes objects."""
    keep = [C.create_string_buffer(bytes(x), len(x)) if len(x) else C.create_string_buffer(1) for x in items]
    arr = (C.c_void_p * len(items))(*[C.addressof(k) for k in keep])
    return arr, keep


# ------------------------------------------------------------------------------- Smith-Waterman
def sw_score_flat(seqs: np.ndarray, off: np.ndarray, length: np.ndarray,
                  scoring=(SW_MATCH, SW_MISMATCH, SW_GAP_OPEN, SW_GAP_EXTEND), out=None) -> np.ndarray:
    """sw_score_batch_flat: seqs uint8 buffer, off/len of 2*n_pairs sequences (a0 b0 a1 b1 ...).
    out: an int32 array to fill (a pinned one receives every chunk's scores by DMA, without a staging copy)."""
    seqs = _as(seqs, np.uint8)
    off = _as(off, np.int64)
    length = _as(length, np.int32)
    assert off.size == length.size and off.size % 2 == 0
    n = off.size // 2
    if out is None:
        out = np.empty(n, dtype=np.int32)
    assert out.dtype == np.int32 and out.size >= n and out.flags["C_CONTIGUOUS"]
    _check(load_library().sw_score_batch_flat(_ptr(seqs), seqs.size, _ptr(off), _ptr(length), n,
                                              *[int(s) for s in scoring], _ptr(out)))
    return out[:n]


def sw_ends_flat(seqs: np.ndarray, off: np.ndarray, length: np.ndarray, scoring=(SW_MATCH, SW_MISMATCH, SW_GAP_OPEN, SW_GAP_EXTEND)):
    """sw_ends_batch_flat: (scores[n], ends[n, 2]) -- ends = index of the last aligned symbol in a and in b"""
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    length = np.ascontiguousarray(length, dtype=np.int32)
    n = off.size // 2
    scores = np.empty(n, dtype=np.int32)
    ends = np.empty((n, 2), dtype=np.int32)
    _check(load_library().sw_ends_batch_flat(_ptr(seqs), seqs.size, _ptr(off), _ptr(length), n,
                                             *[int(s) for s in scoring], _ptr(scores), _ptr(ends)))
    return scores, ends


def sw_align_flat(seqs: np.ndarray, off: np.ndarray, length: np.ndarray, scoring=(SW_MATCH, SW_MISMATCH, SW_GAP_OPEN, SW_GAP_EXTEND),
                  cigar_cap: Optional[int] = None, out=None):
    """sw_align_batch_flat: (scores[n], coords[n, 4] = a_start a_end b_start b_end, cigar_off[n + 1], cigar runs).
    out = (scores, coords, cigar_off, cigar) arrays to fill (e.g. pinned ones: results then come home by DMA while the
    next chunk computes); the returned cigar is a view of the given array."""
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    length = np.ascontiguousarray(length, dtype=np.int32)
    n = off.size // 2
    lib = load_library()
    if out is not None:
        scores, coords, cig_off, cigar = out
        assert scores.size >= n and coords.size >= 4 * n and cig_off.size >= n + 1
        total = C.c_int64(0)
        _check(lib.sw_align_batch_flat(_ptr(seqs), seqs.size, _ptr(off), _ptr(length), n, *[int(s) for s in scoring],
                                       _ptr(scores), _ptr(coords), _ptr(cig_off), _ptr(cigar), cigar.size, C.byref(total)))
        return scores[:n], coords.reshape(-1, 4)[:n], cig_off[:n + 1], cigar[:int(total.value)]
    scores = np.empty(n, dtype=np.int32)
    coords = np.empty((n, 4), dtype=np.int32)
    cig_off = np.empty(n + 1, dtype=np.int64)
    cap = int(cigar_cap) if cigar_cap is not None else max(16, 8 * n)
    while True:
        cigar = np.empty(max(cap, 1), dtype=np.uint32)
        total = C.c_int64(0)
        rc = lib.sw_align_batch_flat(_ptr(seqs), seqs.size, _ptr(off), _ptr(length), n, *[int(s) for s in scoring],
                                     _ptr(scores), _ptr(coords), _ptr(cig_off), _ptr(cigar), cap, C.byref(total))
        if rc == -5 and cigar_cap is None and total.value > cap:      # AGX_ERANGE: the runs did not fit; ask again
            cap = int(total.value)
            continue
        _check(rc)
        return scores, coords, cig_off, cigar[:int(total.value)]


def sw_align_batch(a: Sequence[bytes], b: Sequence[bytes], scoring=(SW_MATCH, SW_MISMATCH, SW_GAP_OPEN, SW_GAP_EXTEND)):
    """sw_align_batch: pointer-array form; (scores, coords[n, 4], cigar_off[n + 1], cigar runs)"""
    n = len(a)
    assert len(b) == n
    pa = (C.c_char_p * n)(*a)
    pb = (C.c_char_p * n)(*b)
    la = (C.c_int32 * n)(*[len(x) for x in a])
    lb = (C.c_int32 * n)(*[len(x) for x in b])
    scores = np.empty(n, dtype=np.int32)
    coords = np.empty((n, 4), dtype=np.int32)
    cig_off = np.empty(n + 1, dtype=np.int64)
    cap = max(16, sum(len(x) + len(y) + 2 for x, y in zip(a, b)))      # no pair has more runs than symbols
    cigar = np.empty(cap, dtype=np.uint32)
    total = C.c_int64(0)
    _check(load_library().sw_align_batch(C.cast(pa, C.POINTER(C.c_void_p)), la, C.cast(pb, C.POINTER(C.c_void_p)), lb, n,
                                         *[int(s) for s in scoring], _ptr(scores), _ptr(coords), _ptr(cig_off),
                                         _ptr(cigar), cap, C.byref(total)))
    return scores, coords, cig_off, cigar[:int(total.value)]


def sw_ends_batch(a: Sequence[bytes], b: Sequence[bytes], scoring=(SW_MATCH, SW_MISMATCH, SW_GAP_OPEN, SW_GAP_EXTEND)):
    """sw_ends_batch: pointer-array form; (scores, ends[n, 2])"""
    n = len(a)
    assert len(b) == n
    pa = (C.c_char_p * n)(*a)
    pb = (C.c_char_p * n)(*b)
    la = (C.c_int32 * n)(*[len(x) for x in a])
    lb = (C.c_int32 * n)(*[len(x) for x in b])
    scores = np.empty(n, dtype=np.int32)
    ends = np.empty((n, 2), dtype=np.int32)
    _check(load_library().sw_ends_batch(C.cast(pa, C.POINTER(C.c_void_p)), la, C.cast(pb, C.POINTER(C.c_void_p)), lb, n,
                                        *[int(s) for s in scoring], _ptr(scores), _ptr(ends)))
    return scores, ends


def sw_score_batch(a: Sequence[bytes], b: Sequence[bytes],
                   scoring=(SW_MATCH, SW_MISMATCH, SW_GAP_OPEN, SW_GAP_EXTEND)) -> np.ndarray:
    """sw_score_batch: pointer-array form, sequences are raw bytes as the reference sees them."""
    assert len(a) == len(b)
    n = len(a)
    pa, ka = _byte_ptr_array(a)
    pb, kb = _byte_ptr_array(b)
    la = (C.c_int32 * n)(*[len(x) for x in a])
    lb = (C.c_int32 * n)(*[len(x) for x in b])
    out = (C.c_int32 * max(n, 1))()
    _check(load_library().sw_score_batch(C.cast(pa, C.POINTER(C.c_void_p)), la,
                                         C.cast(pb, C.POINTER(C.c_void_p)), lb, n,
                                         *[int(s) for s in scoring], out))
    del ka, kb
    return np.array(out[:n], dtype=np.int32)


def sw_score_file_image(image, line_buf: int = 1000,
                        scoring=(SW_MATCH, SW_MISMATCH, SW_GAP_OPEN, SW_GAP_EXTEND), max_pairs=None, out=None,
                        copy: bool = True):
    """sw_score_file_image: the whole file image, chunked into fgets() lines on the GPU.
    Returns (scores, header, dangling_bytes).  `out` (int32, optionally pinned) receives the scores;
    by default a buffer for `max_pairs` (or the worst case image_bytes / 2) is allocated.
    copy=False returns a view of `out` (what the C caller sees) instead of a private copy."""
    img = np.frombuffer(image, dtype=np.uint8) if isinstance(image, (bytes, bytearray)) else _as(image, np.uint8)
    if out is None:
        cap = max(1, img.size // 2 + 2) if max_pairs is None else max(1, int(max_pairs))
        out = np.empty(cap, dtype=np.int32)
    cap = out.size
    n = C.c_int64(0)
    header = C.c_int32(0)
    d_off = C.c_int64(-1)
    d_len = C.c_int32(0)
    _check(load_library().sw_score_file_image(_ptr(img) if img.size else None, img.size, int(line_buf),
                                              *[int(s) for s in scoring], _ptr(out), cap, C.byref(n),
                                              C.byref(header), C.byref(d_off), C.byref(d_len)))
    dangling = img[d_off.value:d_off.value + d_len.value].tobytes() if d_off.value >= 0 else b""
    return (out[:n.value].copy() if copy else out[:n.value]), int(header.value), dangling


def sw_score_device(device: int, d_seqs: int, seqs_bytes: int, d_off: int, d_len: int, n_pairs: int,
                    d_scores: int, stream: int = 0,
                    scoring=(SW_MATCH, SW_MISMATCH, SW_GAP_OPEN, SW_GAP_EXTEND)) -> None:
    """sw_score_batch_device: raw device pointers (e.g. torch tensor .data_ptr())."""
    _check(load_library().sw_score_batch_device(int(device), d_seqs, int(seqs_bytes), d_off, d_len,
                                                int(n_pairs), *[int(s) for s in scoring], d_scores,
                                                stream or None))


def sw_ends_device(device: int, d_seqs: int, seqs_bytes: int, d_off: int, d_len: int, n_pairs: int, d_scores: int, d_ends: int,
                   stream: int = 0, scoring=(SW_MATCH, SW_MISMATCH, SW_GAP_OPEN, SW_GAP_EXTEND)) -> None:
    """sw_ends_batch_device: raw device pointers"""
    _check(load_library().sw_ends_batch_device(int(device), d_seqs, int(seqs_bytes), d_off, d_len, int(n_pairs),
                                               *[int(s) for s in scoring], d_scores, d_ends, stream or None))


def sw_align_device(device: int, d_seqs: int, seqs_bytes: int, d_off: int, d_len: int, n_pairs: int, d_scores: int, d_coords: int,
                    d_cigar_off: int, d_cigar: int, cigar_cap: int, stream: int = 0,
                    scoring=(SW_MATCH, SW_MISMATCH, SW_GAP_OPEN, SW_GAP_EXTEND)) -> int:
    """sw_align_batch_device: raw device pointers; returns the number of CIGAR runs of the batch"""
    total = C.c_int64(0)
    _check(load_library().sw_align_batch_device(int(device), d_seqs, int(seqs_bytes), d_off, d_len, int(n_pairs),
                                                *[int(s) for s in scoring], d_scores, d_coords, d_cigar_off, d_cigar,
                                                int(cigar_cap), C.byref(total), stream or None))
    return int(total.value)


# ------------------------------------------------------------------------------- PairHMM
def pairhmm_forward_batch(reads: Sequence[Sequence[bytes]], haps: Sequence[bytes]) -> np.ndarray:
    """pairhmm_forward_batch: reads = [(bases, q, qi, qd, qg), ...]; returns [n_reads, n_haps]."""
    nr, nh = len(reads), len(haps)
    cols = []
    keeps = []
    for f in range(5):
        arr, keep = _byte_ptr_array([r[f] for r in reads])
        cols.append(C.cast(arr, C.POINTER(C.c_void_p)))
        keeps.append((arr, keep))
    rl = (C.c_int32 * max(nr, 1))(*[len(r[0]) for r in reads])
    ph, kh = _byte_ptr_array(haps)
    hl = (C.c_int32 * max(nh, 1))(*[len(h) for h in haps])
    out = (C.c_double * max(nr * nh, 1))()
    _check(load_library().pairhmm_forward_batch(nr, *cols, rl, nh, C.cast(ph, C.POINTER(C.c_void_p)), hl, out))
    del keeps, kh
    return np.array(out[:nr * nh], dtype=np.float64).reshape(nr, nh)


def pairhmm_forward_flat(buf: np.ndarray, read_field_off: np.ndarray, read_len: np.ndarray,
                         hap_off: np.ndarray, hap_len: np.ndarray, batch_read_start: np.ndarray,
                         batch_hap_start: np.ndarray, out=None) -> np.ndarray:
    """pairhmm_forward_batches_flat on host arrays; returns the flat log10 vector.
    out: a float64 array to fill (a pinned one receives the results by DMA, without a staging copy)."""
    buf = _as(buf, np.uint8)
    rfo = _as(read_field_off, np.int64).reshape(-1)
    rl = _as(read_len, np.int32)
    ho = _as(hap_off, np.int64)
    hl = _as(hap_len, np.int32)
    brs = _as(batch_read_start, np.int64)
    bhs = _as(batch_hap_start, np.int64)
    nb = brs.size - 1
    n_out = int(np.sum((brs[1:] - brs[:-1]) * (bhs[1:] - bhs[:-1])))
    if out is None:
        out = np.empty(max(n_out, 1), dtype=np.float64)
    assert out.dtype == np.float64 and out.size >= n_out and out.flags["C_CONTIGUOUS"]
    _check(load_library().pairhmm_forward_batches_flat(_ptr(buf), buf.size, _ptr(rfo), _ptr(rl), rl.size,
                                                       _ptr(ho), _ptr(hl), hl.size, _ptr(brs), _ptr(bhs),
                                                       nb, _ptr(out)))
    return out[:n_out]


def pairhmm_forward_file_image(image, copy: bool = True):
    """pairhmm_forward_file_image: the raw pairHMM/test_set-format file image, parsed on the GPU.
    Returns (log10 values, pairs per batch, incomplete).  With copy=False the arrays alias the library's
    pinned result buffers (valid until the next PairHMM call)."""
    img = np.frombuffer(image, dtype=np.uint8) if isinstance(image, (bytes, bytearray)) else _as(image, np.uint8)
    p_out, p_bp = C.c_void_p(), C.c_void_p()
    n_out, n_b, inc = C.c_int64(0), C.c_int64(0), C.c_int32(0)
    _check(load_library().pairhmm_forward_file_image(_ptr(img) if img.size else None, img.size, C.byref(p_out),
                                                     C.byref(n_out), C.byref(p_bp), C.byref(n_b), C.byref(inc)))
    vals = np.ctypeslib.as_array(C.cast(p_out, C.POINTER(C.c_double)), shape=(n_out.value,)) if n_out.value else \
        np.empty(0, np.float64)
    bp = np.ctypeslib.as_array(C.cast(p_bp, C.POINTER(C.c_int32)), shape=(n_b.value,)) if n_b.value else \
        np.empty(0, np.int32)
    if copy:
        vals, bp = vals.copy(), bp.copy()
    return vals, bp, int(inc.value)


def pairhmm_forward_device(device: int, d_buf: int, buf_bytes: int, d_read_field_off: int,
                           d_read_len: int, d_read_batch: int, d_read_out_off: int, n_reads: int,
                           d_hap_off: int, d_hap_len: int, n_haps: int, d_batch_hap_start: int,
                           n_batches: int, n_pairs: int, d_out: int, stream: int = 0,
                           fp64_rescue: bool = True) -> None:
    _check(load_library().pairhmm_forward_batches_device(
        int(device), d_buf, int(buf_bytes), d_read_field_off, d_read_len, d_read_batch, d_read_out_off,
        int(n_reads), d_hap_off, d_hap_len, int(n_haps), d_batch_hap_start, int(n_batches), int(n_pairs),
        1 if fp64_rescue else 0, d_out, stream or None))
