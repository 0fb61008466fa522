import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import agxpkg
agx = agxpkg.load(); cap = agx.capi
cap.init(1)
n = 1_000_000
inp = agx.synth.sw_uniform_pairs(n, 150, seed=1)
hb = torch.from_numpy(inp.buf).pin_memory(); ho = torch.from_numpy(inp.off).pin_memory(); hl = torch.from_numpy(inp.len).pin_memory()
dev = torch.device('cuda:0')
for _ in range(3):
    d = hb.to(dev, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    d = hb.to(dev, non_blocking=True); d2 = ho.to(dev, non_blocking=True); d3 = hl.to(dev, non_blocking=True); torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print("pure H2D of the batch (326 MB): %.2f ms = %.1f GB/s" % (dt * 1e3, 326e6 / dt / 1e9))
nb, no, nl = hb.numpy(), ho.numpy(), hl.numpy()
for _ in range(3): cap.sw_score_flat(nb, no, nl)
t0 = time.perf_counter()
for _ in range(5): cap.sw_score_flat(nb, no, nl)
print("e2e default: %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
# pageable host memory
pb, po, pl = inp.buf.copy(), inp.off.copy(), inp.len.copy()
for _ in range(2): cap.sw_score_flat(pb, po, pl)
t0 = time.perf_counter()
for _ in range(3): cap.sw_score_flat(pb, po, pl)
print("e2e default, pageable host buffers: %.2f ms" % ((time.perf_counter() - t0) / 3 * 1e3))
