"""sw_score_file_image end to end: region timeline (AGX_TRACE) and the upload-segment knob (AGX_SW_IMAGE_SEGMENT)."""
import sys, os, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import agxpkg
agx = agxpkg.load(); cap = agx.capi
cap.init(1)
n = 1_000_000
inp = agx.synth.sw_uniform_pairs(n, 150, seed=1)
hb = torch.from_numpy(inp.buf).pin_memory()
out = torch.empty(n, dtype=torch.int32).pin_memory()
dev = torch.device('cuda:0')
for _ in range(3):
    d = hb.to(dev, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    d = hb.to(dev, non_blocking=True); torch.cuda.synchronize()
print(json.dumps({"h2d_only_ms": (time.perf_counter() - t0) / 5 * 1e3}))
nb, no = hb.numpy(), out.numpy()
def run(tag, reps=7, copy=False):
    for _ in range(3): cap.sw_score_file_image(nb, out=no, copy=copy)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); cap.sw_score_file_image(nb, out=no, copy=copy); ts.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps({"variant": tag, "ms_median": float(np.median(ts)), "ms_min": float(np.min(ts))}), flush=True)
run("default, wrapper copies the scores", copy=True)
run("default")
os.environ["AGX_SW_IMAGE_HALVING"] = "1"
run("halving schedule (the first one)")
del os.environ["AGX_SW_IMAGE_HALVING"]
for seg in (20, 24, 28, 32):
    for tail in (0, 6, 10):
        os.environ["AGX_SW_IMAGE_SEGMENT"] = str(seg << 20)
        os.environ["AGX_SW_IMAGE_TAIL"] = str(tail << 20)
        run("segments of %d MiB, last one %d MiB" % (seg, tail))
del os.environ["AGX_SW_IMAGE_SEGMENT"], os.environ["AGX_SW_IMAGE_TAIL"]
for k in sys.argv[1:]:
    kv = k.split("=")
    os.environ[kv[0]] = kv[1]
    run(k)
    del os.environ[kv[0]]
os.environ["AGX_TRACE"] = "1"
cap.sw_score_file_image(nb, out=no, copy=False)
