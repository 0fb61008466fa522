"""Host-side mirror of the reference's text formats.

These functions turn a file image into the flat (buffer, offsets, lengths) form the C ABI takes,
reproducing the reference's input-handling rules, quirks included:

  Smith-Waterman  smithWaterman/antidiagonalSmithWaterman.c:205-227
     * line 1 is atoi()'d into the number of LINES to consume (loop `i += 2`, :216), so a
       generator.py file (header = number of alignments) yields half its pairs;
     * each sequence is one fgets() chunk of a 1000-byte buffer (:201-202): at most 999 bytes, a
       longer line is split into several "lines";
     * the trailing '\\n' stays part of the sequence (:229-244);
     * EOF in the middle of a pair stops the run (:219-227).
  PairHMM         pairHMM/antidiagsPairHMM.c:371-491
     * batches of "<num_read> <num_haplotypes>", read lines with five space-separated fields,
       haplotype lines; read length is inferred as (strlen(line) - 4) / 5 (:418).

The C drivers under drivers/ implement the same rules in C; tests check both against the compiled
reference programs.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

SW_LINE_BUF = 1000       # MAX_LINE_LENGTH, antidiagonalSmithWaterman.c:44
HMM_LINE_BUF = 5001      # MAX_READ_LEN*5+1, antidiagsPairHMM.c:8, :353


def _c_atoi(b: bytes) -> int:
    """C atoi(): optional whitespace, optional sign, digits; 0 when there are none."""
    s = b.lstrip(b" \t\n\v\f\r")
    sign = 1
    if s[:1] in (b"+", b"-"):
        sign = -1 if s[:1] == b"-" else 1
        s = s[1:]
    n = 0
    for ch in s:
        if 48 <= ch <= 57:
            n = n * 10 + (ch - 48)
        else:
            break
    return sign * n


def _c_sscanf_d(s: bytes):
    """One "%d" of sscanf on s: (value or None when nothing parses, rest of s)."""
    t = s.lstrip(b" \t\n\v\f\r")
    sign, i = 1, 0
    if t[:1] in (b"+", b"-"):
        sign, i = (-1 if t[:1] == b"-" else 1), 1
    j = i
    while j < len(t) and 48 <= t[j] <= 57:
        j += 1
    if j == i:
        return None, s
    return sign * int(t[i:j]), t[j:]


def sscanf_two_ints(line: bytes, prev):
    """sscanf(line, "%d %d", &num_read, &num_haplotypes) (antidiagsPairHMM.c:378) on variables that keep their
    previous values (prev) where a field does not parse: the reference declares them once, outside the batch
    loop (:345-346)."""
    nr, nh = prev
    v, rest = _c_sscanf_d(line)
    if v is None:
        return nr, nh
    nr = v
    v, _ = _c_sscanf_d(rest)
    if v is not None:
        nh = v
    return nr, nh


def fgets_chunks(data: np.ndarray, start: int, bufsize: int) -> Tuple[np.ndarray, np.ndarray]:
    """(offsets, lengths) of successive fgets(buf, bufsize) results over data[start:]."""
    n = data.size
    nl = np.flatnonzero(data[start:] == 10) + start
    line_start = np.concatenate(([start], nl + 1))
    line_end = np.concatenate((nl + 1, [n]))            # exclusive, '\n' included
    if line_start[-1] >= n:                              # file ends with '\n': no empty last line
        line_start, line_end = line_start[:-1], line_end[:-1]
    length = line_end - line_start
    cap = bufsize - 1
    if length.size == 0 or length.max() <= cap:
        return line_start.astype(np.int64), length.astype(np.int32)
    offs: List[int] = []
    lens: List[int] = []
    for s, l in zip(line_start.tolist(), length.tolist()):
        while l > cap:
            offs.append(s)
            lens.append(cap)
            s += cap
            l -= cap
        if l > 0:
            offs.append(s)
            lens.append(l)
    return np.asarray(offs, dtype=np.int64), np.asarray(lens, dtype=np.int32)


@dataclass
class SwInput:
    buf: np.ndarray        # uint8 file image
    off: np.ndarray        # int64 [2*n_pairs]
    len: np.ndarray        # int32 [2*n_pairs]
    header: int            # atoi(first line)
    dangling: bytes        # first line of an incomplete last pair (the reference echoes it), or b""

    @property
    def n_pairs(self) -> int:
        return self.off.size // 2


def parse_sw(data: bytes | np.ndarray, line_buf: int = SW_LINE_BUF) -> SwInput:
    buf = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else np.asarray(data, np.uint8)
    if buf.size == 0:
        raise ValueError("file is empty")
    off, ln = fgets_chunks(buf, 0, line_buf)
    header = _c_atoi(buf[off[0]:off[0] + ln[0]].tobytes())
    off, ln = off[1:], ln[1:]
    want_pairs = max(0, (header + 1) // 2)               # for (i = 0; i < line_num; i += 2)
    have_pairs = off.size // 2
    n = min(want_pairs, have_pairs)
    dangling = b""
    if want_pairs > have_pairs and off.size % 2 == 1:
        o, l = int(off[-1]), int(ln[-1])
        dangling = buf[o:o + l].tobytes()
    return SwInput(buf, off[:2 * n].copy(), ln[:2 * n].copy(), header, dangling)


def write_sw(pairs: Sequence[Tuple[bytes, bytes]], header: int | None = None,
             trailing_newline: bool = True) -> bytes:
    """A generator.py-style file.  header defaults to the number of LINES (so every pair is scored);
    pass header=len(pairs) to reproduce generator.py's own (halving) header."""
    lines = []
    for a, b in pairs:
        lines.append(a)
        lines.append(b)
    h = 2 * len(pairs) if header is None else header
    body = b"\n".join(lines)
    out = str(h).encode() + b"\n" + body
    if trailing_newline and lines:
        out += b"\n"
    return out


@dataclass
class HmmInput:
    buf: np.ndarray
    read_field_off: np.ndarray   # int64 [n_reads, 5]
    read_len: np.ndarray         # int32 [n_reads]
    hap_off: np.ndarray          # int64 [n_haps]
    hap_len: np.ndarray          # int32 [n_haps]
    batch_read_start: np.ndarray  # int64 [n_batches+1]
    batch_hap_start: np.ndarray   # int64 [n_batches+1]

    @property
    def n_batches(self) -> int:
        return self.batch_read_start.size - 1

    @property
    def n_pairs(self) -> int:
        return int(np.sum(np.diff(self.batch_read_start) * np.diff(self.batch_hap_start)))

    def cells(self) -> int:
        tot = 0
        for b in range(self.n_batches):
            r0, r1 = self.batch_read_start[b], self.batch_read_start[b + 1]
            h0, h1 = self.batch_hap_start[b], self.batch_hap_start[b + 1]
            tot += int(self.read_len[r0:r1].astype(np.int64).sum()) * int(self.hap_len[h0:h1].astype(np.int64).sum())
        return tot


def parse_pairhmm(data: bytes | np.ndarray) -> HmmInput:
    buf = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else np.asarray(data, np.uint8)
    off, ln = fgets_chunks(buf, 0, HMM_LINE_BUF)
    # strip the '\n' (line[strcspn(line, "\n")] = 0, :399, :417)
    has_nl = np.zeros(off.size, dtype=bool)
    if off.size:
        last = off + ln - 1
        has_nl = buf[last] == 10
    ln = ln - has_nl.astype(np.int32)
    rfo: List[List[int]] = []
    rl: List[int] = []
    ho: List[int] = []
    hl: List[int] = []
    brs, bhs = [0], [0]
    i = 0
    n_lines = off.size
    nr = nh = 0
    while i < n_lines:
        nr, nh = sscanf_two_ints(buf[off[i]:off[i] + ln[i]].tobytes(), (nr, nh))
        nr, nh = max(nr, 0), max(nh, 0)
        i += 1
        if i + nr + nh > n_lines:
            break                                            # "Error reading ..." in the reference
        for r in range(nr):
            o, l = int(off[i + r]), int(ln[i + r])
            L = (l - 4) // 5                                 # :418
            line = buf[o:o + l].tobytes()
            fields = line.split()
            if len(fields) >= 5 and all(len(f) >= L for f in fields[:5]):
                # offsets of the five whitespace-separated fields (sscanf "%s %s %s %s %s", :101)
                pos, offs = 0, []
                for f in fields[:5]:
                    pos = line.index(f, pos)
                    offs.append(o + pos)
                    pos += len(f)
            else:
                raise ValueError(f"malformed read line {i + r}")
            rfo.append(offs)
            rl.append(L)
        i += nr
        for h in range(nh):
            ho.append(int(off[i + h]))
            hl.append(int(ln[i + h]))
        i += nh
        brs.append(len(rl))
        bhs.append(len(hl))
    return HmmInput(buf, np.asarray(rfo, dtype=np.int64).reshape(-1, 5), np.asarray(rl, dtype=np.int32),
                    np.asarray(ho, dtype=np.int64), np.asarray(hl, dtype=np.int32),
                    np.asarray(brs, dtype=np.int64), np.asarray(bhs, dtype=np.int64))


def write_pairhmm(batches) -> bytes:
    """batches = [(reads, haps)], reads = [(bases, q, qi, qd, qg)] as bytes, haps = [bytes]."""
    out = []
    for reads, haps in batches:
        out.append(f"{len(reads)} {len(haps)}".encode())
        for r in reads:
            out.append(b" ".join(r))
        out.extend(haps)
    return b"\n".join(out) + b"\n"
