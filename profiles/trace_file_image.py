import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import agxpkg
agx = agxpkg.load(); cap = agx.capi
cap.init(1)
n = 1_000_000
inp = agx.synth.sw_uniform_pairs(n, 150, seed=1)
hb = torch.from_numpy(inp.buf).pin_memory()
out = torch.empty(n, dtype=torch.int32).pin_memory()
for _ in range(3): cap.sw_score_file_image(hb.numpy(), out=out.numpy())
os.environ["AGX_TRACE"] = "1"
cap.sw_score_file_image(hb.numpy(), out=out.numpy())
