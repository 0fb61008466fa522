"""sw_align_batch_device / sw_ends_batch_device with everything resident: host clock around synchronised calls
(AGX_ALIGN_TRACE=1 prints the library's own timeline).  usage: python profiles/align_dev_probe.py [pairs] [len]"""
import json, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import agxpkg
agx = agxpkg.load(); cap = agx.capi
import torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 150
cap.init(1)
inp = agx.synth.sw_uniform_pairs(n, L, seed=3)
dev = torch.device("cuda", 0)
d_buf, d_off, d_len = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (inp.buf, inp.off, inp.len))
d_scores = torch.empty(n, dtype=torch.int32, device=dev); d_ends = torch.empty((n, 2), dtype=torch.int32, device=dev)
d_coords = torch.empty((n, 4), dtype=torch.int32, device=dev); d_coff = torch.empty(n + 1, dtype=torch.int64, device=dev)
d_cig = torch.empty(8 * n, dtype=torch.int32, device=dev)
calls = {"ends": lambda: cap.sw_ends_device(0, d_buf.data_ptr(), d_buf.numel(), d_off.data_ptr(), d_len.data_ptr(), n, d_scores.data_ptr(), d_ends.data_ptr(), 0),
         "align": lambda: cap.sw_align_device(0, d_buf.data_ptr(), d_buf.numel(), d_off.data_ptr(), d_len.data_ptr(), n, d_scores.data_ptr(),
                                              d_coords.data_ptr(), d_coff.data_ptr(), d_cig.data_ptr(), d_cig.numel(), 0)}
for name, fn in calls.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize(dev)
    ms = 1e3 * (time.perf_counter() - t0) / 5
    print(json.dumps({"what": name, "pairs": n, "ms_per_call": ms, "gcups": n * L * L / ms / 1e6}), flush=True)
