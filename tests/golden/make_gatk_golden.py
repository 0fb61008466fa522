#!/usr/bin/env python
"""An INDEPENDENT double-precision statement of the GATK-semantics PairHMM forward pass, and the values it gives
on the reference's own test inputs: tests/golden/pairhmm_gatk.json.

Why: `agx_pairhmm_set_gatk_mode` corrects the reference's mismatch prior
(antidiagsPairHMM.c:111-113 uses Qr where the published model uses Qr / 3).  The reference has no such mode, so
its programs cannot produce goldens for it, and no GATK / GKL build exists in this image.  This file is the pin:
it is written from the PUBLISHED algorithm (GATK's LoglessPairHMM as described in its documentation and in the
PairHMM literature: full (R+1) x (H+1) matrices, per-cell prior matrix, per-row transition table, initial condition
2^1020 / H in the deletion row, result log10(sum of match + insertion over the last row) - log10(2^1020)) and shares
nothing with oracle/oracle.c: other loop structure (matrices, not rolling rows), other scale constant, other
association of the match expression, qualities turned into probabilities by a table of exact powers of ten.
tests/test_gatk_pin.py compares oracle.c's gatk branch (CPU) and libagx's GATK mode (GPU) with these values.

Runs anywhere (pure Python + numpy for file parsing); reads only tests/golden/*.in.gz.
"""
from __future__ import annotations

import gzip
import json
import math
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent

INITIAL_CONDITION = 2.0 ** 1020          # GATK: PairHMM.INITIAL_CONDITION
TRISTATE_CORRECTION = 3.0                # a wrong base is one of the three other bases
MIN_USABLE_Q_SCORE = 6                   # GATK reads base qualities below 6 as 6

# qualToErrorProb: 10^(-q/10) for the integer qualities a Phred+33 byte can hold
QUAL_TO_ERROR = [10.0 ** (-q / 10.0) for q in range(-33, 223)]


def err(ch: int) -> float:
    """probability of a Phred+33 quality byte (a signed char in the reference: bytes >= 128 are negative)"""
    q = (ch - 256 if ch >= 128 else ch) - 33
    return QUAL_TO_ERROR[q + 33]


def gatk_forward(bases: bytes, quals: bytes, ins: bytes, dels: bytes, gcp: bytes, hap: bytes, qual_floor: bool) -> float:
    """log10 P(read | haplotype), GATK semantics.  qual_floor: base qualities below 6 count as 6."""
    R, H = len(bases), len(hap)
    # transition table, one row per read position (1-based)
    MM, GM, MX, XX, MY, YY = range(6)
    tr = [[0.0] * 6 for _ in range(R + 1)]
    for i in range(1, R + 1):
        qi, qd, qg = err(ins[i - 1]), err(dels[i - 1]), err(gcp[i - 1])
        tr[i][MM] = 1.0 - (qi + qd)
        tr[i][GM] = 1.0 - qg
        tr[i][MX] = qi
        tr[i][XX] = qg
        tr[i][MY] = qd
        tr[i][YY] = qg
    # prior matrix
    prior = [[0.0] * (H + 1) for _ in range(R + 1)]
    for i in range(1, R + 1):
        qb = quals[i - 1]
        if qual_floor and (qb - 256 if qb >= 128 else qb) - 33 < MIN_USABLE_Q_SCORE:
            qb = 33 + MIN_USABLE_Q_SCORE
        e = err(qb)
        x = bases[i - 1]
        for j in range(1, H + 1):
            y = hap[j - 1]
            prior[i][j] = (1.0 - e) if (x == y or x == ord("N") or y == ord("N")) else e / TRISTATE_CORRECTION
    match = [[0.0] * (H + 1) for _ in range(R + 1)]
    insertion = [[0.0] * (H + 1) for _ in range(R + 1)]
    deletion = [[0.0] * (H + 1) for _ in range(R + 1)]
    for j in range(H + 1):
        deletion[0][j] = INITIAL_CONDITION / H
    for i in range(1, R + 1):
        t = tr[i]
        for j in range(1, H + 1):
            match[i][j] = prior[i][j] * (match[i - 1][j - 1] * t[MM] + insertion[i - 1][j - 1] * t[GM] +
                                         deletion[i - 1][j - 1] * t[GM])
            insertion[i][j] = match[i - 1][j] * t[MX] + insertion[i - 1][j] * t[XX]
            deletion[i][j] = match[i][j - 1] * t[MY] + deletion[i][j - 1] * t[YY]
    final = 0.0
    for j in range(1, H + 1):
        final += match[R][j] + insertion[R][j]
    return math.log10(final) - math.log10(INITIAL_CONDITION) if final > 0 else float("-inf")


def read_batches(path: Path):
    """[(reads, haps)] of a pairHMM/test_set file; reads = 5-tuples of bytes.  Well-formed files only."""
    lines = gzip.open(path, "rb").read().split(b"\n")
    k, out = 0, []
    while k < len(lines) and lines[k].strip():
        nr, nh = (int(x) for x in lines[k].split()[:2])
        reads = [tuple(lines[k + 1 + r].split()[:5]) for r in range(nr)]
        haps = [lines[k + 1 + nr + h].strip() for h in range(nh)]
        out.append((reads, haps))
        k += 1 + nr + nh
    return out


def main() -> None:
    rec = {"source": "tests/golden/make_gatk_golden.py (independent LoglessPairHMM statement, double precision)",
           "modes": {"1": "mismatch prior Qr/3", "3": "Qr/3 + base-quality floor 6"}, "files": {}}
    picks = {"pairhmm_test.in.gz": None, "pairhmm_10s.in.gz": 6}     # None: every batch; k: the first k batches, <= 8 reads each
    for name, lim in picks.items():
        batches = read_batches(HERE / name)
        rows = []
        for b, (reads, haps) in enumerate(batches[:lim] if lim else batches):
            for r, rd in enumerate(reads[:8] if lim else reads):
                for h, hp in enumerate(haps):
                    rows.append({"batch": b, "read": r, "hap": h,
                                 "gatk": gatk_forward(*rd, hp, qual_floor=False),
                                 "gatk_floor": gatk_forward(*rd, hp, qual_floor=True)})
        rec["files"][name] = rows
        print(name, len(rows), "pairs", file=sys.stderr)
    # base qualities below 6 (the test_set files hold none: there the floor changes nothing): seeded synthetic pairs,
    # inputs recorded beside the values
    import random
    rnd = random.Random(20261019)
    low = []
    for _ in range(24):
        H = rnd.randint(12, 60)
        hap = "".join(rnd.choice("ACGT") for _ in range(H))
        R = rnd.randint(4, min(40, H))
        at = rnd.randint(0, H - R)
        bases = "".join(c if rnd.random() > 0.15 else rnd.choice("ACGTN") for c in hap[at:at + R])
        quals = "".join(chr(33 + rnd.randint(0, 14)) for _ in range(R))
        ins = "".join(chr(33 + rnd.randint(25, 45)) for _ in range(R))
        dels = "".join(chr(33 + rnd.randint(25, 45)) for _ in range(R))
        gcp = "".join(chr(33 + rnd.choice((10, 10, 10, 8, 12))) for _ in range(R))
        rd = tuple(x.encode() for x in (bases, quals, ins, dels, gcp))
        low.append({"read": [bases, quals, ins, dels, gcp], "hap": hap,
                    "gatk": gatk_forward(*rd, hap.encode(), qual_floor=False),
                    "gatk_floor": gatk_forward(*rd, hap.encode(), qual_floor=True)})
    rec["synthetic_low_quality"] = low
    (HERE / "pairhmm_gatk.json").write_text(json.dumps(rec, indent=0) + "\n")


if __name__ == "__main__":
    main()
