"""Randomised comparison of sw_ends_batch_flat / sw_align_batch_flat with the oracle (oracle/sw_align.c): length
distributions around every class boundary, long rows on short columns, alphabets with ties and non-ACGT bytes,
missing newlines, several scorings.  usage: python profiles/align_fuzz.py [seconds] [seed]"""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import agxpkg

agx = agxpkg.load()
import oracle

cap = agx.capi
cap.init(1)
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
SCORINGS = [(1, -1, -3, -1), (2, -3, -4, -1), (3, -2, -5, -2), (1, -4, 0, -2), (5, -4, -10, -1), (30, -30, -60, -30), (7, -9, -20, -3)]
EDGES = [1, 2, 31, 32, 33, 63, 64, 65, 95, 96, 97, 127, 128, 129, 151, 152, 153, 191, 192, 193, 255, 256, 257, 383, 384, 385,
         511, 512, 513, 767, 768, 769, 1022, 1023, 1024, 1025, 1100]


def seq(alpha, n):
    return alpha[rng.integers(0, alpha.size, size=n)]


def make_pair(alpha):
    kind = rng.integers(0, 6)
    la = int(rng.choice(EDGES)) + int(rng.integers(-1, 2)) if kind < 3 else int(rng.integers(1, 400))
    la = max(1, la)
    x = seq(alpha, la)
    if kind in (0, 3):                                   # related, similar length
        y = x.copy()
        m = rng.random(la) < 0.07
        y[m] = seq(alpha, int(m.sum()))
        y = y[rng.random(la) > 0.03]
        if rng.random() < 0.5:
            y = np.concatenate([seq(alpha, int(rng.integers(0, 20))), y, seq(alpha, int(rng.integers(0, 20)))])
    elif kind == 1:                                      # long rows on short columns
        y = np.concatenate([seq(alpha, int(rng.integers(0, 1500))), x, seq(alpha, int(rng.integers(0, 1500)))])
    elif kind == 4:                                      # tandem repeats: many co-optimal end cells
        unit = seq(alpha, int(rng.integers(1, 5)))
        x = np.tile(unit, int(rng.integers(2, 60)))
        y = np.tile(unit, int(rng.integers(2, 60)))
    else:
        y = seq(alpha, max(1, int(rng.integers(1, 2 * la + 2))))
    a, b = x.tobytes(), y.tobytes()
    if rng.random() < 0.5:
        a, b = b, a
    r = rng.random()
    if r < 0.7:
        a, b = a + b"\n", b + b"\n"
    elif r < 0.8:
        a = a + b"\n"
    elif r < 0.9:
        b = b + b"\n"
    return a, b


t_end = time.time() + budget
rounds = pairs = bad = 0
while time.time() < t_end:
    alpha = np.frombuffer([b"ACGT", b"ACGT", b"AC", b"ACGTN", b"ACGTNacgt", b"A"][int(rng.integers(0, 6))], np.uint8)
    sc = SCORINGS[int(rng.integers(0, len(SCORINGS)))]
    n = int(rng.integers(20, 160))
    ab = [make_pair(alpha) for _ in range(n)]
    parts, off, ln, at = [], [], [], 0
    for a, b in ab:
        for s in (a, b):
            parts.append(s); off.append(at); ln.append(len(s)); at += len(s)
    buf = np.frombuffer(b"".join(parts) + b"\0", np.uint8)
    off = np.asarray(off, np.int64); ln = np.asarray(ln, np.int32)
    s1, ends = cap.sw_ends_flat(buf, off, ln, sc)
    s2, coords, coff, cig = cap.sw_align_flat(buf, off, ln, sc)
    s0 = cap.sw_score_flat(buf, off, ln, sc)
    for p, (a, b) in enumerate(ab):
        ws, wc, wg = oracle.sw_align(a, b, sc)
        got = (int(s0[p]), int(s1[p]), int(s2[p]), tuple(ends[p].tolist()), tuple(coords[p].tolist()), cig[coff[p]:coff[p + 1]].tolist())
        want = (ws, ws, ws, (wc[1], wc[3]), wc, wg)
        if got != want:
            bad += 1
            if bad <= 5:
                print(json.dumps({"mismatch": {"scoring": sc, "a": a.decode("latin1"), "b": b.decode("latin1"), "got": str(got), "want": str(want)}}), flush=True)
    rounds += 1
    pairs += n
print(json.dumps({"rounds": rounds, "pairs": pairs, "mismatches": bad}))
sys.exit(1 if bad else 0)
