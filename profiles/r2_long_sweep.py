"""Sweep of sw_longr_kernel variants (stripe width K, rows per step R, chain form, dp4a substitution, steps per
hand-off block B) on ONE GPU, in one process (the library reads its AGX_LONG_* knobs at every call).
usage: AGX_LIB_PATH=build/libagx_sweep.so python profiles/r2_long_sweep.py share|full|quick
  share : 125 000 columns x 1 000 000 rows = the per-GPU share of BASELINE configs[4] on 8 GPUs (and x 500 000 rows,
          which separates the stripe-fill term from the row term)
  full  : 1 000 000 x 1 000 000 on one GPU
Every variant's score is compared with the round-1 kernel's (AGX_LONG_OLD=1)."""
import itertools, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import agxpkg
agx = agxpkg.load(); cap = agx.capi
cap.init(1)
mode = sys.argv[1] if len(sys.argv) > 1 else "share"
rng = np.random.default_rng(7)
acgt = np.frombuffer(b"ACGT", np.uint8)
N = 1_000_000
a = acgt[rng.integers(0, 4, size=N)]
b = a.copy()
mut = rng.random(N) < 0.02
b[mut] = acgt[rng.integers(0, 4, size=int(mut.sum()))]

def run(cols, rows, env, reps=2):
    for k in [k for k in os.environ if k.startswith("AGX_LONG_")]:
        del os.environ[k]
    os.environ.update({k: str(v) for k, v in env.items()})
    if cols < rows:
        os.environ["AGX_LONG_SWAP"] = "1"       # keep `cols` as the striped dimension
    buf = np.concatenate([a[:cols], b[:rows]])
    off = np.array([0, cols], dtype=np.int64); ln = np.array([cols, rows], dtype=np.int32)
    best, score = 1e9, None
    for _ in range(reps + 1):
        t0 = time.perf_counter(); s = int(cap.sw_score_flat(buf, off, ln)[0]); dt = time.perf_counter() - t0
        assert score is None or s == score
        score = s; best = min(best, dt)
    return best * 1e3, score

def sweep_pipe(cols, rows, ks, rs, bs, dp=(0, 1)):
    """the software-pipelined kernel (sw_longp_kernel, AGX_LONG_PIPE=1)"""
    ref_ms, ref_score = run(cols, rows, {"AGX_LONG_OLD": 1, "AGX_LONG_K": ks[0]})
    print(json.dumps({"cols": cols, "rows": rows, "variant": "round-1 kernel", "K": ks[0], "ms": round(ref_ms, 2), "score": ref_score}), flush=True)
    for k, r, d, bb in itertools.product(ks, rs, dp, bs):
        if r > k or (r == 6 and not d):
            continue
        try:
            ms, sc = run(cols, rows, {"AGX_LONG_K": k, "AGX_LONG_R": r, "AGX_LONG_PIPE": 1, "AGX_LONG_DP4A": d, "AGX_LONG_B": bb})
        except Exception as e:
            print(json.dumps({"K": k, "R": r, "pipe": 1, "dp4a": d, "B": bb, "error": str(e)[:80]}), flush=True)
            continue
        print(json.dumps({"cols": cols, "rows": rows, "K": k, "R": r, "pipe": 1, "dp4a": d, "B": bb, "ms": round(ms, 2),
                          "gcups": round(cols * rows / ms / 1e6), "ok": sc == ref_score}), flush=True)


def sweep(cols, rows, ks, rs, bs, dp=(0, 1), chains=(0, 1)):
    ref_ms, ref_score = run(cols, rows, {"AGX_LONG_OLD": 1, "AGX_LONG_K": ks[0]})
    print(json.dumps({"cols": cols, "rows": rows, "variant": "round-1 kernel", "K": ks[0], "ms": round(ref_ms, 2), "score": ref_score}), flush=True)
    for k, r, ch, d, bb in itertools.product(ks, rs, chains, dp, bs):
        try:
            ms, sc = run(cols, rows, {"AGX_LONG_K": k, "AGX_LONG_R": r, "AGX_LONG_CHAIN": ch, "AGX_LONG_DP4A": d, "AGX_LONG_B": bb,
                                      "AGX_LONG_PIPE": 0})
        except Exception as e:      # a variant that was not instantiated
            print(json.dumps({"K": k, "R": r, "short": ch, "dp4a": d, "B": bb, "error": str(e)[:80]}), flush=True)
            continue
        print(json.dumps({"cols": cols, "rows": rows, "K": k, "R": r, "short": ch, "dp4a": d, "B": bb, "ms": round(ms, 2),
                          "gcups": round(cols * rows / ms / 1e6), "ok": sc == ref_score}), flush=True)

if mode == "smallk":
    # two warps per scheduler instead of one: narrower stripes on the per-GPU share of an 8-GPU run
    sweep(125_000, 1_000_000, [7, 2, 4, 6], [2, 4], [8, 16, 32], dp=(1,), chains=(0,))
    sweep(125_000, 125_000, [7, 2, 4, 6], [2, 4], [8, 16], dp=(1,), chains=(0,))
elif mode == "tune":
    # 125 kbp x 125 kbp on one GPU has the stripe-fill : row ratio of 1 Mbp x 1 Mbp on eight (x 8 = the 8-GPU time)
    sweep(125_000, 125_000, [7], [2, 4], [2, 4, 8, 16, 32], dp=(1,), chains=(0, 1))
    sweep(125_000, 1_000_000, [7], [2, 4], [4, 8, 32], dp=(1,), chains=(0,))
    sweep(1_000_000, 1_000_000, [27], [2, 4], [8, 32], dp=(1,), chains=(0,))
    sweep_pipe(125_000, 125_000, [7], [2, 4], [4, 8, 16], dp=(1,))
elif mode == "pipe2":
    sweep_pipe(125_000, 1_000_000, [7, 8], [2, 3, 4, 6], [4, 8, 16, 32], dp=(1,))
    sweep_pipe(125_000, 500_000, [7], [2, 3, 4, 6], [8], dp=(1,))
    sweep_pipe(1_000_000, 1_000_000, [27], [2, 4], [8, 32], dp=(1,))
elif mode == "pipe_share":
    sweep_pipe(125_000, 1_000_000, [6, 7, 8], [1, 2, 3, 4, 6], [4, 8, 16, 32])
    sweep_pipe(125_000, 500_000, [7], [1, 2, 3, 4, 6], [8])
elif mode == "pipe_full":
    sweep_pipe(1_000_000, 1_000_000, [14, 27], [1, 2, 3, 4, 6], [8, 32])
elif mode == "quick":
    sweep(125_000, 200_000, [7], [1, 2, 4], [8])
elif mode == "share":
    sweep(125_000, 1_000_000, [6, 7, 8], [1, 2, 4], [4, 8, 16, 32])
    sweep(125_000, 500_000, [7], [1, 2, 4], [8])
else:
    sweep(1_000_000, 1_000_000, [14, 27], [1, 2, 4], [8, 32])
