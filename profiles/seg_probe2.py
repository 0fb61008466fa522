"""Why bench.py's e2e (7.0 ms) differs from seg_probe.py's (6.3 - 6.4 ms): same call, bench-like timing variants."""
import sys, os, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import agxpkg
agx = agxpkg.load(); cap = agx.capi
cap.init_devices([0])
cap.set_profiling(True)
n = 1_000_000
inp = agx.synth.sw_uniform_pairs(n, 150, seed=1000)
hb = torch.from_numpy(inp.buf).pin_memory()
out = torch.empty(n, dtype=torch.int32).pin_memory()
nb, no = hb.numpy(), out.numpy()
def timed(tag, warm, reps):
    for _ in range(warm): cap.sw_score_file_image(nb, out=no, copy=False)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); cap.sw_score_file_image(nb, out=no, copy=False); ts.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps({"variant": tag, "ms_mean": float(np.mean(ts)), "ms_median": float(np.median(ts)), "all": [round(t, 2) for t in ts]}), flush=True)
timed("cold: 2 warm-up calls, 5 timed", 2, 5)
timed("again: 0 warm-up calls, 7 timed", 0, 7)
# as bench.py does first: the device-resident path on torch's stream, with torch-allocated device buffers
dev = torch.device('cuda:0')
d_buf, d_off, d_len = hb.to(dev), torch.from_numpy(inp.off).to(dev), torch.from_numpy(inp.len).to(dev)
d_out = torch.empty(n, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(8):
    cap.sw_score_device(0, d_buf.data_ptr(), d_buf.numel(), d_off.data_ptr(), d_len.data_ptr(), n, d_out.data_ptr(), st)
torch.cuda.synchronize()
timed("after the resident path: 2 warm-up calls, 5 timed", 2, 5)
cap.set_profiling(False)
timed("profiling spans off", 2, 5)
os.environ["AGX_TRACE"] = "1"
cap.sw_score_file_image(nb, out=no, copy=False)
