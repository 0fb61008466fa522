"""pairhmm_forward_batches_flat end to end on the config-4 batch: parts per shard (AGX_HMM_PARTS) and parity with the
device-resident entry point."""
import sys, os, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import agxpkg
agx = agxpkg.load(); cap = agx.capi
cap.init_devices([0])
inp = agx.synth.pairhmm_batches(1000, 200, 5, seed=2000, unrelated_frac=0.001)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
arrs = [pin(x) for x in (inp.buf, inp.read_field_off, inp.read_len, inp.hap_off, inp.hap_len, inp.batch_read_start, inp.batch_hap_start)]
ref = None
for parts in (1, 2, 4, 6, 8, 0):
    if parts: os.environ["AGX_HMM_PARTS"] = str(parts)
    else: os.environ.pop("AGX_HMM_PARTS", None)
    for _ in range(2): got = cap.pairhmm_forward_flat(*arrs)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); got = cap.pairhmm_forward_flat(*arrs); ts.append((time.perf_counter() - t0) * 1e3)
    if ref is None: ref = got.copy()
    print(json.dumps({"parts": parts or "default", "ms_median": float(np.median(ts)), "equal_to_one_part": bool(np.array_equal(got, ref, equal_nan=True)),
                      "nan": int(np.isnan(got).sum())}), flush=True)
