"""Throughput of sw_ends_batch_flat / sw_align_batch_flat on the 150 x 150 batch (host buffers in, host results out)
and the device-side spans (agx_profile_ms 7 = DP kernels of the last alignment call, 8 = traceback walk)."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import agxpkg

agx = agxpkg.load()
cap = agx.capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 150
cap.init(1)
cap.set_profiling(True)
inp = agx.synth.sw_uniform_pairs(n, L, seed=3)
try:                                     # pinned host buffers, as bench.py's end-to-end legs use
    import torch
    for f in ("buf", "off", "len"):
        setattr(inp, f, torch.from_numpy(np.ascontiguousarray(getattr(inp, f))).pin_memory().numpy())
except Exception as e:                   # pageable buffers still work, only slower
    print(json.dumps({"note": f"not pinned: {e}"}), flush=True)
cells = float(n) * L * L
only = sys.argv[3] if len(sys.argv) > 3 else None
for name, fn in (("score", lambda: cap.sw_score_flat(inp.buf, inp.off, inp.len)),
                 ("ends", lambda: cap.sw_ends_flat(inp.buf, inp.off, inp.len)),
                 ("align", lambda: cap.sw_align_flat(inp.buf, inp.off, inp.len, cigar_cap=16 * n))):
    if only and name not in only.split('+'):
        continue
    fn()
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        r = fn()
        best = min(best, time.perf_counter() - t0)
    rec = {"what": name, "pairs": n, "len": L, "ms": best * 1e3, "gcups_e2e": cells / best / 1e9}
    if name != "score":
        rec["dp_ms"] = cap.profile_ms(0, 7)
        rec["walk_ms"] = cap.profile_ms(0, 8)
    if name == "align":
        rec["cigar_runs"] = int(r[3].size)
    print(json.dumps(rec), flush=True)
