import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import agxpkg
agx = agxpkg.load(); cap = agx.capi
cap.init(1)
inp = agx.synth.pairhmm_batches(100, 200, 5, seed=3)
args = (inp.buf, inp.read_field_off, inp.read_len, inp.hap_off, inp.hap_len, inp.batch_read_start, inp.batch_hap_start)
cells = inp.cells()
for mode in (False, True):
    cap.set_pairhmm_force_fp64(mode)
    cap.pairhmm_forward_flat(*args)
    t0 = time.perf_counter(); r = cap.pairhmm_forward_flat(*args); dt = time.perf_counter() - t0
    print("force_fp64=%s: %.1f ms, %.0f GCUPS e2e (100 batches, %.2e cells)" % (mode, dt * 1e3, cells / dt / 1e9, cells))
    if mode: r64 = r
    else: r32 = r
print("max rel diff fp32 path vs fp64 path: %.3g" % np.max(np.abs(r32 - r64) / np.abs(r64)))
