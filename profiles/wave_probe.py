"""Batch of mid-length pairs (both sides > 1024: the one-warp-per-pair s32 kernel); prints GCUPS.
usage: python profiles/wave_probe.py [N_PAIRS] [LEN]   (AGX_WAVE_RAW=1: raw-byte pass only)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import agxpkg
agx = agxpkg.load(); cap = agx.capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
cap.init(1)
inp = agx.synth.sw_uniform_pairs(n, L, seed=3)
cap.sw_score_flat(inp.buf, inp.off, inp.len)
t0 = time.perf_counter(); s = cap.sw_score_flat(inp.buf, inp.off, inp.len); dt = time.perf_counter() - t0
print("%d pairs of %dx%d, raw=%s: %.1f ms  %.0f GCUPS (e2e, flat entry)  checksum %d" % (
    n, L, L, os.environ.get("AGX_WAVE_RAW", "0"), dt * 1e3, n * L * L / dt / 1e9, int(s.sum())))
