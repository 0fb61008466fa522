"""SW config-3 batch with a fraction of the pairs carrying an 'N' (pairs the s16x2 kernel cannot code).
usage: python profiles/n_probe.py [FRACTION]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import agxpkg
agx = agxpkg.load(); cap = agx.capi
cap.init(1)
n = 1_000_000
for frac in ([float(sys.argv[1])] if len(sys.argv) > 1 else [0.0, 0.001, 0.01, 0.05]):
    inp = agx.synth.sw_uniform_pairs(n, 150, seed=1)
    buf = inp.buf.copy()
    rng = np.random.default_rng(5)
    pick = rng.choice(n, size=int(frac * n), replace=False)
    buf[inp.off[2 * pick] + rng.integers(0, 150, size=pick.size)] = ord("N")
    d_buf, d_off, d_len = (torch.from_numpy(x).cuda() for x in (buf, inp.off, inp.len))
    d_out = torch.empty(n, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    run = lambda: cap.sw_score_device(0, d_buf.data_ptr(), d_buf.numel(), d_off.data_ptr(), d_len.data_ptr(), n, d_out.data_ptr(), st)
    for _ in range(3): run()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): run()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print("%.1f%% of the pairs with an N: %.2f ms per 10^6 pairs = %.0f GCUPS" % (100 * frac, dt * 1e3, n * 22500 / dt / 1e9))
