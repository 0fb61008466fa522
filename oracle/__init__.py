"""ctypes binding of the CPU oracle (oracle/liboracle.so) and runners for the compiled reference
programs (oracle/_ref/*).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  Nothing under accelerating-genomics_b200/ or
drivers/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle.so"
REF = HERE / "_ref"

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        srcs = [HERE / "oracle.c", HERE / "sw_blocked.c", HERE / "sw_align.c", HERE / "oracle.h"]
        if not LIB.exists() or any(p.exists() and p.stat().st_mtime > LIB.stat().st_mtime for p in srcs):
            subprocess.run(["make", "-C", str(HERE), "liboracle.so"], check=True, capture_output=True)
        l = C.CDLL(str(LIB))
        l.oracle_sw_score.restype = C.c_int32
        l.oracle_sw_score.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32] + [C.c_int32] * 4 + \
            [C.POINTER(C.c_int32)]
        l.oracle_sw_score_blocked.restype = C.c_int32
        l.oracle_sw_score_blocked.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64] + [C.c_int32] * 6
        l.oracle_sw_align.restype = C.c_int32
        l.oracle_sw_align.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32] + [C.c_int32] * 4 + \
            [C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.c_int32, C.POINTER(C.c_int32)]
        l.oracle_sw_ends.restype = C.c_int32
        l.oracle_sw_ends.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32] + [C.c_int32] * 4 + [C.POINTER(C.c_int32)]
        l.oracle_sw_cigar_score.restype = C.c_int32
        l.oracle_sw_cigar_score.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32] + [C.c_int32] * 4 + \
            [C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.c_int32]
        l.oracle_sw_file.restype = C.c_int64
        l.oracle_sw_file.argtypes = [C.c_char_p, C.c_int32, C.c_void_p, C.c_int64, C.POINTER(C.c_int32)]
        l.oracle_pairhmm_prob.restype = C.c_double
        l.oracle_pairhmm_prob.argtypes = [C.c_uint8]
        l.oracle_pairhmm_forward.restype = C.c_double
        l.oracle_pairhmm_forward.argtypes = [C.c_char_p] * 5 + [C.c_int32, C.c_char_p, C.c_int32, C.c_int32]
        l.oracle_pairhmm_file.restype = C.c_int64
        l.oracle_pairhmm_file.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int32)]
        _lib = l
    return _lib


def sw_score(a: bytes, b: bytes, scoring=(1, -1, -3, -1), want_corner: bool = False):
    corner = C.c_int32(0)
    s = lib().oracle_sw_score(a, len(a), b, len(b), *scoring, C.byref(corner))
    return (int(s), int(corner.value)) if want_corner else int(s)


def sw_score_blocked(a: bytes, b: bytes, scoring=(1, -1, -3, -1), tile: int = 4096, threads: int = 0) -> int:
    """oracle_sw_score_blocked: the same recurrence, tile by tile on several host threads (long pairs)."""
    return int(lib().oracle_sw_score_blocked(a, len(a), b, len(b), *scoring, int(tile), int(threads)))


def sw_scores_flat(buf: np.ndarray, off: np.ndarray, ln: np.ndarray, scoring=(1, -1, -3, -1)) -> np.ndarray:
    data = np.ascontiguousarray(buf, dtype=np.uint8).tobytes()
    n = off.size // 2
    out = np.empty(n, dtype=np.int32)
    for p in range(n):
        a = data[off[2 * p]:off[2 * p] + ln[2 * p]]
        b = data[off[2 * p + 1]:off[2 * p + 1] + ln[2 * p + 1]]
        out[p] = sw_score(a, b, scoring)
    return out


def sw_align(a: bytes, b: bytes, scoring=(1, -1, -3, -1)):
    """oracle_sw_align: (score, (a_start, a_end, b_start, b_end), [cigar words]); coordinates -1 when score == 0"""
    cap = len(a) + len(b) + 2
    coords = (C.c_int32 * 4)()
    cig = (C.c_uint32 * cap)()
    n = C.c_int32(0)
    s = lib().oracle_sw_align(a, len(a), b, len(b), *scoring, coords, cig, cap, C.byref(n))
    assert s != -(1 << 31)
    return int(s), tuple(int(x) for x in coords), [int(cig[k]) for k in range(n.value)]


def sw_ends(a: bytes, b: bytes, scoring=(1, -1, -3, -1)):
    """oracle_sw_ends: (score, (a_end, b_end)) with two rolling rows"""
    e = (C.c_int32 * 2)()
    s = lib().oracle_sw_ends(a, len(a), b, len(b), *scoring, e)
    return int(s), (int(e[0]), int(e[1]))


def sw_cigar_score(a: bytes, b: bytes, coords, cigar, scoring=(1, -1, -3, -1)) -> int:
    """score spelled by a CIGAR between the given coordinates (None when the path does not fit them)"""
    c = (C.c_int32 * 4)(*coords)
    g = (C.c_uint32 * max(1, len(cigar)))(*cigar)
    s = int(lib().oracle_sw_cigar_score(a, len(a), b, len(b), *scoring, c, g, len(cigar)))
    return None if s == -(1 << 31) else s


def sw_align_flat(buf: np.ndarray, off: np.ndarray, ln: np.ndarray, scoring=(1, -1, -3, -1)):
    """(scores[n], coords[n, 4], list of cigar lists) over a flat batch"""
    data = np.ascontiguousarray(buf, dtype=np.uint8).tobytes()
    n = off.size // 2
    scores = np.empty(n, dtype=np.int32)
    coords = np.empty((n, 4), dtype=np.int32)
    cigars = []
    for p in range(n):
        a = data[off[2 * p]:off[2 * p] + ln[2 * p]]
        b = data[off[2 * p + 1]:off[2 * p + 1] + ln[2 * p + 1]]
        s, c, g = sw_align(a, b, scoring)
        scores[p] = s
        coords[p] = c
        cigars.append(g)
    return scores, coords, cigars


def cigar_string(cigar) -> str:
    return "".join(f"{w >> 4}{'MID'[w & 15]}" for w in cigar)


def sw_file(path: str, line_buf: int = 1000, cap: int = 1 << 22):
    out = np.empty(cap, dtype=np.int32)
    header = C.c_int32(0)
    n = lib().oracle_sw_file(str(path).encode(), line_buf, out.ctypes.data, cap, C.byref(header))
    if n < 0:
        raise OSError(f"oracle_sw_file failed on {path}")
    return out[:min(n, cap)].copy(), int(header.value)


def pairhmm_forward(read, hap: bytes, gatk=False) -> float:
    """gatk: False / 0 the reference's priors; True / 1 mismatch prior Qr/3; 3 = that plus the base-quality floor 6"""
    bases, q, qi, qd, qg = read
    return float(lib().oracle_pairhmm_forward(bases, q, qi, qd, qg, len(bases), hap, len(hap), int(gatk)))


def pairhmm_file(path: str, cap: int = 1 << 22):
    out = np.empty(cap, dtype=np.float64)
    nb = C.c_int32(0)
    n = lib().oracle_pairhmm_file(str(path).encode(), out.ctypes.data, cap, C.byref(nb))
    if n < 0:
        raise OSError(f"oracle_pairhmm_file failed on {path}")
    return out[:min(n, cap)].copy(), int(nb.value)


def pairhmm_flat(inp, gatk: bool = False, limit: Optional[int] = None) -> np.ndarray:
    """Oracle over an accelerating_genomics_b200.formats.HmmInput (read-major inside each batch)."""
    data = inp.buf.tobytes()
    out: List[float] = []
    for b in range(inp.n_batches):
        r0, r1 = int(inp.batch_read_start[b]), int(inp.batch_read_start[b + 1])
        h0, h1 = int(inp.batch_hap_start[b]), int(inp.batch_hap_start[b + 1])
        haps = [data[inp.hap_off[h]:inp.hap_off[h] + inp.hap_len[h]] for h in range(h0, h1)]
        for r in range(r0, r1):
            L = int(inp.read_len[r])
            fields = tuple(data[int(o):int(o) + L] for o in inp.read_field_off[r])
            for hp in haps:
                out.append(pairhmm_forward(fields, hp, gatk))
                if limit is not None and len(out) >= limit:
                    return np.asarray(out)
    return np.asarray(out)


# ------------------------------------------------------------------ compiled reference programs
def ref_available(name: str = "sw_antidiag") -> bool:
    return (REF / name).exists()


def run_ref_sw(path: str, long_lines: bool = False, timeout: float = 600.0):
    """Run the compiled reference SW program; returns (scores, header, stdout)."""
    exe = REF / ("sw_antidiag_long" if long_lines else "sw_antidiag")
    cmd = f"ulimit -s unlimited 2>/dev/null; exec {exe} {path}" if long_lines else f"exec {exe} {path}"
    r = subprocess.run(["bash", "-c", cmd], capture_output=True, timeout=timeout)
    text = r.stdout.decode(errors="replace")
    scores = [int(l.split()[1]) for l in text.splitlines() if l.startswith("Score:")]
    header = None
    for l in text.splitlines():
        if l.startswith("line_num:"):
            header = int(l.split()[1])
    return np.asarray(scores, dtype=np.int32), header, text


def run_ref_sw_ends(path: str, timeout: float = 600.0):
    """oracle/_ref/sw_antidiag_ends (the reference + position bookkeeping beside its running maximum, see
    oracle/Makefile): rows of (score, end_iy, end_ix, sx_line) -- 1-based DP indices, sx_line 1 or 2."""
    r = subprocess.run([str(REF / "sw_antidiag_ends"), str(path)], capture_output=True, timeout=timeout)
    return parse_ref_sw_ends(r.stdout.decode(errors="replace"))


def parse_ref_sw_ends(text: str):
    rows = []
    for l in text.splitlines():
        if l.startswith("Score:"):
            t = l.split()
            rows.append((int(t[1]), int(t[3]), int(t[4]), int(t[6])))
    return rows


def ref_ends_to_coords(row):
    """(score, iy, ix, sx_line) of the instrumented reference -> (a_end, b_end), 0-based, -1 when score == 0"""
    s, iy, ix, sx = row
    if s == 0:
        return (-1, -1)
    return (ix - 1, iy - 1) if sx == 1 else (iy - 1, ix - 1)


def run_ref_pairhmm(path: str, which: str = "pairhmm_matrix", timeout: float = 600.0):
    """Run a compiled reference PairHMM program; returns (values parsed from its %f output, text)."""
    exe = REF / which
    with tempfile.NamedTemporaryFile(suffix=".out", delete=False) as t:
        outp = t.name
    try:
        subprocess.run([str(exe), str(path), outp], capture_output=True, timeout=timeout, check=True)
        text = Path(outp).read_text()
    finally:
        os.unlink(outp)
    vals = np.asarray([float(x) for x in text.split()], dtype=np.float64)
    return vals, text
