"""One long Smith-Waterman alignment of COLS x ROWS through the host C ABI; prints ms, GCUPS, score.
usage: python profiles/long_probe.py COLS ROWS [GPUS]   (AGX_LONG_K / AGX_LONG_CHAIN / AGX_LONG_SWAP apply)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import agxpkg
agx = agxpkg.load(); cap = agx.capi
cols, rows = int(sys.argv[1]), int(sys.argv[2])
gpus = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cap.init(gpus)
rng = np.random.default_rng(7)
acgt = np.frombuffer(b"ACGT", np.uint8)
a = acgt[rng.integers(0, 4, size=max(cols, rows))]
b = a.copy()
mut = rng.random(b.size) < 0.02
b[mut] = acgt[rng.integers(0, 4, size=int(mut.sum()))]
x, y = a[:cols], b[:rows]
buf = np.concatenate([x, y])
off = np.array([0, cols], dtype=np.int64); ln = np.array([cols, rows], dtype=np.int32)
if cols < rows: os.environ["AGX_LONG_SWAP"] = "1"     # keep `cols` as the striped dimension
score = int(cap.sw_score_flat(buf, off, ln)[0])
best = 1e9
for _ in range(int(os.environ.get("REPS", "2"))):
    t0 = time.perf_counter(); s2 = int(cap.sw_score_flat(buf, off, ln)[0]); dt = time.perf_counter() - t0
    assert s2 == score or os.environ.get("NOCHECK")
    best = min(best, dt)
print("cols=%d rows=%d gpus=%d K=%s chain=%s: %.1f ms  %.0f GCUPS  score %d" % (
    cols, rows, gpus, os.environ.get("AGX_LONG_K", "auto"), os.environ.get("AGX_LONG_CHAIN", "auto"), best * 1e3,
    cols * rows / best / 1e9, score))
