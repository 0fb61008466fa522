"""Seeded synthetic inputs of the BASELINE.json shapes, written in the reference's own file formats.

The reference's generator (smithWaterman/generator.py) is unseeded and ignores argv (SURVEY.md
section 2, row 4); these generators keep its FORMAT (header line, one ACGT line per sequence) and add
a seed and a shape.  Shapes follow SURVEY.md section 8(d).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .formats import HmmInput, SwInput

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _mutate_rows(rng: np.random.Generator, a: np.ndarray, sub: float, indel: float) -> np.ndarray:
    """Row-wise mutated copy of a [n, L] base matrix: substitutions, then deletions/insertions,
    re-padded with random bases to L columns (vectorised with a sort-key trick)."""
    n, L = a.shape
    b = a.copy()
    m = rng.random((n, L)) < sub
    b[m] = ACGT[rng.integers(0, 4, size=int(m.sum()))]
    if indel <= 0:
        return b
    n_ins = max(1, int(np.ceil(L * indel * 4)))
    keys = np.empty((n, 2 * L + n_ins), dtype=np.float32)
    keys[:, :L] = np.arange(L, dtype=np.float32)
    deleted = rng.random((n, L)) < indel
    keys[:, :L][deleted] = np.inf
    ins_on = rng.random((n, n_ins)) < (L * indel / n_ins)
    ins_key = rng.random((n, n_ins)).astype(np.float32) * L
    ins_key[~ins_on] = np.inf
    keys[:, L:L + n_ins] = ins_key
    keys[:, L + n_ins:] = L + 1 + np.arange(L, dtype=np.float32)      # padding, always last
    pool = np.concatenate([b, ACGT[rng.integers(0, 4, size=(n, n_ins + L))]], axis=1)
    idx = np.argsort(keys, axis=1, kind="stable")[:, :L]
    return np.take_along_axis(pool, idx, axis=1)


def sw_uniform_pairs(n_pairs: int, length: int = 150, seed: int = 1, related_frac: float = 0.5,
                     sub: float = 0.05, indel: float = 0.01, chunk: int = 65536) -> SwInput:
    """BASELINE config 3: n_pairs pairs of `length` x `length`, A uniform ACGT, B independent for
    (1 - related_frac) of the pairs and a mutated copy of A for the rest; every line
    newline-terminated; header = number of lines (2 * n_pairs) so every pair is scored."""
    rng = np.random.default_rng(seed)
    header = f"{2 * n_pairs}\n".encode()
    row = length + 1
    buf = np.empty(len(header) + 2 * n_pairs * row, dtype=np.uint8)
    buf[:len(header)] = np.frombuffer(header, dtype=np.uint8)
    body = buf[len(header):].reshape(n_pairs, 2, row)
    body[:, :, length] = 10
    for s in range(0, n_pairs, chunk):
        e = min(n_pairs, s + chunk)
        a = ACGT[rng.integers(0, 4, size=(e - s, length))]
        b = ACGT[rng.integers(0, 4, size=(e - s, length))]
        rel = rng.random(e - s) < related_frac
        if rel.any():
            b[rel] = _mutate_rows(rng, a[rel], sub, indel)
        body[s:e, 0, :length] = a
        body[s:e, 1, :length] = b
    off = len(header) + np.arange(2 * n_pairs, dtype=np.int64) * row
    ln = np.full(2 * n_pairs, row, dtype=np.int32)
    return SwInput(buf, off, ln, 2 * n_pairs, b"")


def sw_random_file(rng: np.random.Generator, n_pairs: int, min_len: int, max_len: int,
                   alphabet: bytes = b"ACGT", related_frac: float = 0.5,
                   trailing_newline: bool = True, header: int | None = None) -> bytes:
    """Small ragged files for parity tests (generator.py format)."""
    alpha = np.frombuffer(alphabet, dtype=np.uint8)
    lines: List[bytes] = []
    for _ in range(n_pairs):
        la = int(rng.integers(min_len, max_len + 1))
        a = alpha[rng.integers(0, alpha.size, size=la)]
        if rng.random() < related_frac and la > 0:
            b = a.copy()
            m = rng.random(la) < 0.08
            b[m] = alpha[rng.integers(0, alpha.size, size=int(m.sum()))]
            keep = rng.random(la) >= 0.03
            b = b[keep]
            lb_target = int(rng.integers(min_len, max_len + 1))
            if b.size > lb_target:
                b = b[:lb_target]
        else:
            lb = int(rng.integers(min_len, max_len + 1))
            b = alpha[rng.integers(0, alpha.size, size=lb)]
        lines.append(a.tobytes())
        lines.append(b.tobytes())
    h = 2 * n_pairs if header is None else header
    out = str(h).encode() + b"\n" + b"\n".join(lines)
    if trailing_newline:
        out += b"\n"
    return out


def sw_long_pair(length: int, seed: int = 1, related: bool = True, sub: float = 0.02,
                 indel: float = 0.005) -> bytes:
    """BASELINE configs 1 / 5: one pair of `length` x `length` in generator.py format ("2\\n" + 2 lines)."""
    rng = np.random.default_rng(seed)
    a = ACGT[rng.integers(0, 4, size=(1, length))]
    b = _mutate_rows(rng, a, sub, indel) if related else ACGT[rng.integers(0, 4, size=(1, length))]
    return b"2\n" + a[0].tobytes() + b"\n" + b[0].tobytes() + b"\n"


# ------------------------------------------------------------------------------------ PairHMM
def _qual_string(rng, n, lo, hi, skew_high=False) -> np.ndarray:
    if skew_high:
        q = hi - np.minimum(rng.geometric(0.18, size=n) - 1, hi - lo)
    else:
        q = rng.integers(lo, hi + 1, size=n)
    return (q + 33).astype(np.uint8)


def pairhmm_batches(n_batches: int, reads_per_batch: int = 200, haps_per_batch: int = 5, seed: int = 1,
                    read_len=(100, 250), hap_len=(200, 500), unrelated_frac: float = 0.0,
                    n_frac: float = 0.002) -> HmmInput:
    """BASELINE config 4 (HaplotypeCaller-shaped): each batch has `haps_per_batch` sibling
    haplotypes (first one uniform ACGT, the others = it + a few SNPs/indels) and
    `reads_per_batch` reads, each a substring of one haplotype with ~1% substitution errors and
    rare 'N'; base quals in [6, 40] skewed high, ins/del quals in [30, 45], gcp = 10 ('+').
    `unrelated_frac` of the reads are uniform random instead (these exercise the FP64 rescue).
    The result is the file image in the pairHMM/test_set text format plus its index arrays."""
    rng = np.random.default_rng(seed)
    parts: List[bytes] = []
    pos = 0
    rfo: List[np.ndarray] = []
    rl: List[int] = []
    ho: List[int] = []
    hl: List[int] = []
    brs, bhs = [0], [0]
    for _ in range(n_batches):
        head = f"{reads_per_batch} {haps_per_batch}\n".encode()
        parts.append(head)
        pos += len(head)
        H0 = int(rng.integers(hap_len[0], hap_len[1] + 1))
        h0 = ACGT[rng.integers(0, 4, size=H0)]
        haps = [h0]
        for _k in range(haps_per_batch - 1):
            h = h0.copy()
            snp = rng.random(H0) < 0.01
            h[snp] = ACGT[rng.integers(0, 4, size=int(snp.sum()))]
            if rng.random() < 0.5 and H0 > hap_len[0] + 8:
                cut = int(rng.integers(1, H0 - 4))
                h = np.delete(h, slice(cut, cut + int(rng.integers(1, 4))))
            elif H0 < hap_len[1] - 8:
                cut = int(rng.integers(1, H0 - 1))
                h = np.insert(h, cut, ACGT[rng.integers(0, 4, size=int(rng.integers(1, 4)))])
            haps.append(h)
        read_lines = []
        for _r in range(reads_per_batch):
            src = haps[int(rng.integers(0, len(haps)))]
            L = int(rng.integers(read_len[0], min(read_len[1], src.size) + 1))
            if rng.random() < unrelated_frac:
                bases = ACGT[rng.integers(0, 4, size=L)]
            else:
                st = int(rng.integers(0, src.size - L + 1))
                bases = src[st:st + L].copy()
                err = rng.random(L) < 0.01
                bases[err] = ACGT[rng.integers(0, 4, size=int(err.sum()))]
            nn = rng.random(L) < n_frac
            bases[nn] = ord("N")
            q = _qual_string(rng, L, 6, 40, skew_high=True)
            qi = _qual_string(rng, L, 30, 45)
            qd = _qual_string(rng, L, 30, 45)
            qg = np.full(L, ord("+"), dtype=np.uint8)
            line = b" ".join(x.tobytes() for x in (bases, q, qi, qd, qg)) + b"\n"
            rfo.append(pos + np.arange(5, dtype=np.int64) * (L + 1))
            rl.append(L)
            read_lines.append(line)
            pos += len(line)
        parts.extend(read_lines)
        for h in haps:
            ho.append(pos)
            hl.append(h.size)
            parts.append(h.tobytes() + b"\n")
            pos += h.size + 1
        brs.append(len(rl))
        bhs.append(len(hl))
    buf = np.frombuffer(b"".join(parts), dtype=np.uint8)
    return HmmInput(buf, np.stack(rfo).astype(np.int64), np.asarray(rl, dtype=np.int32),
                    np.asarray(ho, dtype=np.int64), np.asarray(hl, dtype=np.int32),
                    np.asarray(brs, dtype=np.int64), np.asarray(bhs, dtype=np.int64))
