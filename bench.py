#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric ("SW and PairHMM GCUPS at 1/2/4/8 B200 vs reference C on host cores").

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the reference C programs on the host cores)

A "step" is one pass of the hot path over one batch.  ONE JSON line comes out; its top level is the headline
workload, the other BASELINE configurations are objects inside it:

  top level      configs[2]  SW short-read batch, 10^6 pairs of 150 x 150 per GPU ("scaling": "weak": every rank
                 scores its own batch; pairs are independent, so there is no data-path collective --
                 torch.distributed only carries the barrier and the max-over-ranks of the timings).
  "pairhmm"      configs[3]  10^6 (read, haplotype) pairs per GPU, reads 100-250 x haplotypes 200-500, with the
                 <= 0.1 % tail of unrelated reads that exercises the FP64 rescue (count and kernel ms reported).
  "sw_long"      configs[4]  ONE pair of 1 Mbp x 1 Mbp, columns striped over ALL N GPUs of the box inside ONE
                 process (rank 0; the other ranks park on a CPU barrier): NVLink peer boundary exchange, strong
                 scaling, pipeline efficiency = T(1 GPU) / (N x T(N GPUs)), score checked against the CPU-computed
                 expected score committed in tests/golden/sw_long_expected.json.
  "strong"       configs[2] / [3] again as ONE 10^6-pair batch through the library's own in-process multi-GPU
                 dispatcher (agx_init(N); rank 0): end to end from host buffers and resident (one shard per GPU).
  "pairhmm_gatk" configs[3] with the corrected GATK priors (mismatch prior Qr/3, optional base-quality floor):
                 never used for parity with the reference, reported separately; "pin" = the mode against the values
                 of an independent LoglessPairHMM statement on the reference's test_set inputs (tests/golden).
  "sw_lengths"   the inter-task SW kernel at the published MI210 sweep lengths (64 ... 1024) and at generator.py's
                 own 450-500 bp, one length class each.
  "sw_align"     (rank 0) alignment END CELLS and full alignments (start cell + CIGAR) of the headline batch:
                 sw_ends_batch_flat / sw_align_batch_flat end to end from pinned host buffers, the device spans, the
                 storing kernel against the alu pipe and against HBM; at N = 1 a sample against the oracle.
  "parity"       (N = 1) GPU results against the reference C programs' own output on the cpu_baseline shards.

  value        GCUPS with the batch already resident in HBM; CUDA events on the launching stream, max over ranks.
  e2e          the same metric through the reference-facing C-ABI call on PINNED HOST buffers: H2D of the batch,
               kernels, D2H of the results all inside the timed region; h2d_only_ms = the bare upload of the same
               bytes by every rank at once (the floor that bounds e2e scaling).  These legs run with libagx's own
               kernel spans (agx_set_profiling) off, as a caller's would: their timing events cost 0.65 ms per call.
  roofline     dominant kernel timed alone with CUDA events inside libagx (agx_profile_ms):
               frac = cells/s x ops_per_cell_executed / measured pipe peak (profiles/peaks_r*.json); the SURVEY's
               algorithmic count is given beside it.
  cpu_baseline the reference C programs (oracle/_ref, kind "reference") -- or the oracle port where they are
               absent -- on all host cores, one process per core, each on its own shard of the same generator.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SW_LEN = 150
# lane-ops per cell on the pipe that bounds each kernel: "algorithmic" = SURVEY.md section 8(d)'s count, "executed" =
# what the kernel issues on that pipe
SW_OPS = {"algorithmic": 4.0, "executed": 2.25}        # s16x2: 4.5 alu-pipe (+ 2 fma-pipe) instructions per cell PAIR
HMM_OPS = {"algorithmic": 8.0, "executed": 6.0}        # FP32 lane-instructions per cell (X' / Y' form)
# alu-pipe instructions per cell of the alignment modes (SASS of the unrolled loop, per 32-bit word of two cells:
# 2 VIADDMNMX + 1 VIMNMX3.RELU + 1/2 VIMNMX3 + 1 PRMT as in the score kernel, + per row the widening of the row key;
# MODE 2 adds 1/2 PRMT per word for the byte packing)
ALIGN_OPS = {"ends": 2.45, "align": 2.75}
LONG_OPS = {"algorithmic": 8.0, "executed": 3.75}      # s32 coded cell: 2 VIADDMNMX + VIMNMX3.RELU + 1/2 VIMNMX3 + 1/4 PRMT
NOMINAL_ALU = 148 * 64 * 1.965e9     # INT32/DPX lane-ops/s   
NOMINAL_FP32 = 148 * 128 * 1.965e9   # FP32 lane-instr/s


# --------------------------------------------------------------------------------------- utilities
def env_rank():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def measured_peaks():
    """ALU peaks measured by drivers/bin/agx_peaks on this pool's B200s (committed under profiles/)."""
    best = None
    for p in sorted((ROOT / "profiles").glob("peaks_r*.json")):
        try:
            best = json.loads(p.read_text())
        except Exception:
            pass
    return best


def hbm_peak():
    """(GB/s, source): the driver-written MEASURED_PEAKS.json (torch copy, read + write bytes), else the fallback
    /opt/skills/guides/B200_PROFILING.md states for this pool"""
    try:
        v = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
        return v, "MEASURED_PEAKS.json hbm_gbs (driver-written: torch copy, read + write bytes)"
    except Exception:
        return 6650.0, "fallback of B200_PROFILING.md (6.65 TB/s, an earlier measurement on this pool)"


def pipe_peak(kind: str):
    """(lane-ops/s, where the number comes from) for kind in {"alu", "fp32"}"""
    peaks = measured_peaks()
    key = "alu_tlaneops" if kind == "alu" else "fp32_tlaneops"
    if peaks and peaks.get(key):
        return peaks[key] * 1e12, f"measured ({peaks.get('source', 'profiles/peaks')})"
    nominal = NOMINAL_ALU if kind == "alu" else NOMINAL_FP32
    return nominal, "nominal 148 SM x %d lanes/clk x 1.965 GHz (no measured peak committed)" % (64 if kind == "alu" else 128)


def roofline_obj(kind, kernel, cells, kernel_ms, ops, unit, ncu_key=None, extra=None):
    """frac is computed from the instructions the kernel EXECUTES on its binding pipe, so it cannot exceed 1."""
    peak, src = pipe_peak(kind)
    peaks = measured_peaks() or {}
    executed = cells * ops["executed"] / (kernel_ms * 1e-3)
    out = {"bound": "alu", "pipe": "INT32/DPX alu pipe" if kind == "alu" else "FP32 fma pipe", "kernel": kernel,
           "achieved": executed / 1e12, "peak": peak / 1e12, "unit": unit, "frac": executed / peak, "peak_source": src,
           "ops_per_cell_executed": ops["executed"], "ops_per_cell_algorithmic": ops["algorithmic"],
           "frac_algorithmic": cells * ops["algorithmic"] / (kernel_ms * 1e-3) / peak,
           "kernel_ms": kernel_ms, "kernel_gcups": cells / (kernel_ms * 1e-3) / 1e9}
    if ncu_key and peaks.get("ncu", {}).get(ncu_key) is not None:
        out["pipe_active_ncu"] = peaks["ncu"][ncu_key]
        out["pipe_active_ncu_source"] = peaks["ncu"].get(ncu_key + "_source")
    if extra:
        out.update(extra)
    return out


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._th = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            # the first queries of a process are slow (tens of ms): take them now, not inside a 20 ms timed region
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            try:
                pynvml.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.002)

    def __enter__(self):
        if self.nv is not None:
            self._th = threading.Thread(target=self._run, daemon=True)
            self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._th:
            self._th.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def merge_clocks(a, b):
    if a is None:
        return b
    out = dict(a)
    if b.get("sm_mhz") is not None:
        out["sm_mhz"] = min(x for x in (a.get("sm_mhz"), b.get("sm_mhz")) if x is not None)
    out["reasons"] = sorted(set(a.get("reasons", [])) | set(b.get("reasons", [])))
    return out


# --------------------------------------------------------------------------------------- CPU reference arm
def _run_parallel(cmds, stdouts=None, timeout=1800):
    """One process per command, all at once; stdouts[i] (a path) receives command i's stdout."""
    t0 = time.perf_counter()
    files = [open(s, "wb") if s else subprocess.DEVNULL for s in (stdouts or [None] * len(cmds))]
    procs = [subprocess.Popen(c, stdout=f, stderr=subprocess.DEVNULL) for c, f in zip(cmds, files)]
    for p in procs:
        p.wait(timeout=timeout)
    dt = time.perf_counter() - t0
    for f in files:
        if f is not subprocess.DEVNULL:
            f.close()
    bad = [p.returncode for p in procs if p.returncode != 0]
    if bad:
        raise RuntimeError(f"reference process failed: exit codes {bad[:4]}")
    return dt


def _port_worker(args):
    kind, path = args
    import oracle
    if kind == "sw":
        s, _ = oracle.sw_file(path)
        np.savetxt(path + ".scores", s, fmt="Score: %d")
        return len(s)
    v, _ = oracle.pairhmm_file(path)
    np.savetxt(path + ".out", v, fmt="%f")
    return len(v)


class CpuReference:
    """Times the reference C programs (oracle/_ref) -- or, where they are absent, the oracle port --
    on every host core, one process per core, each on its own shard of the synthetic workload, and KEEPS what
    they print: the GPU is then checked against it on the same shard files (parity())."""

    def __init__(self):
        import agxpkg
        self.agx = agxpkg.load()
        self.cores = os.cpu_count() or 1
        ref = ROOT / "oracle" / "_ref"
        self.sw_exe = ref / "sw_antidiag"
        self.hmm_exe = ref / "pairhmm_matrix"
        self.kind = "reference" if self.sw_exe.exists() and self.hmm_exe.exists() else "port"
        self.tmp = tempfile.TemporaryDirectory(prefix="agx_cpu_")
        self._sw_files = None
        self._hmm_files = None

    def prepare_sw(self, pairs_per_core: int, seed: int = 77):
        files = []
        for c in range(self.cores):
            inp = self.agx.synth.sw_uniform_pairs(pairs_per_core, SW_LEN, seed=seed + c)
            p = Path(self.tmp.name) / f"sw_{c}.in"
            p.write_bytes(inp.buf.tobytes())
            files.append(str(p))
        self._sw_files = files
        self._sw_pairs = pairs_per_core * self.cores
        self._sw_cells = float(self.cores) * pairs_per_core * SW_LEN * SW_LEN
        self._sw_sample = (f"{self.cores} processes x {pairs_per_core} pairs of {SW_LEN}x{SW_LEN} "
                           f"({'oracle/_ref/sw_antidiag = antidiagonalSmithWaterman.c -O3' if self.kind == 'reference' else 'oracle port'})")

    def prepare_hmm(self, batches_per_core: int, seed: int = 99, unrelated_frac: float = 0.001):
        files, cells, pairs = [], 0, 0
        for c in range(self.cores):
            inp = self.agx.synth.pairhmm_batches(batches_per_core, 200, 5, seed=seed + c, unrelated_frac=unrelated_frac)
            p = Path(self.tmp.name) / f"hmm_{c}.in"
            p.write_bytes(inp.buf.tobytes())
            files.append(str(p))
            cells += inp.cells()
            pairs += inp.n_pairs
        self._hmm_files = files
        self._hmm_pairs = pairs
        self._hmm_cells = float(cells)
        self._hmm_sample = (f"{self.cores} processes x {batches_per_core} batches of 200 reads x 5 haplotypes "
                            f"({'oracle/_ref/pairhmm_matrix = pairHMMmatrix.c -O3, output byte-identical to antidiagsPairHMM.c' if self.kind == 'reference' else 'oracle port'})")

    def _time(self, which: str) -> float:
        files = self._sw_files if which == "sw" else self._hmm_files
        if self.kind == "reference":
            if which == "sw":
                return _run_parallel([[str(self.sw_exe), f] for f in files], [f + ".scores" for f in files])
            return _run_parallel([[str(self.hmm_exe), f, f + ".out"] for f in files])
        import multiprocessing as mp
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(self.cores) as pool:
            pool.map(_port_worker, [(which, f) for f in files])
        return time.perf_counter() - t0

    def sw_gcups(self):
        dt = self._time("sw")
        return self._sw_cells / dt / 1e9, dt

    def hmm_gcups(self):
        dt = self._time("hmm")
        return self._hmm_cells / dt / 1e9, dt

    def baseline_obj(self, which: str, value: float):
        return {"value": value, "unit": "GCUPS", "cores": self.cores, "kind": self.kind,
                "sample": self._sw_sample if which == "sw" else self._hmm_sample}

    # ---- the GPU against what the CPU programs printed, on the same shard files (agx arm only) ----
    def parity_sw(self, cap):
        pairs = bad = 0
        for f in self._sw_files:
            want = np.array([int(l.split()[1]) for l in Path(f + ".scores").read_bytes().splitlines() if l.startswith(b"Score:")],
                            dtype=np.int32)
            got, _, _ = cap.sw_score_file_image(np.fromfile(f, dtype=np.uint8), max_pairs=want.size + 1)
            pairs += want.size
            bad += int(want.size != got.size) + int(np.count_nonzero(want[:got.size] != got[:want.size]))
        return {"sw_pairs": pairs, "sw_mismatches": bad}

    def parity_hmm(self, cap):
        pairs = bad = 0
        worst = 0.0
        for f in self._hmm_files:
            want = np.array(Path(f + ".out").read_bytes().split(), dtype=np.float64)     # "%f" lines; "-inf" parses
            got, _, _ = cap.pairhmm_forward_file_image(np.fromfile(f, dtype=np.uint8))
            pairs += want.size
            if want.size != got.size:
                bad += 1
                continue
            fin = np.isfinite(want)
            bad += int(np.count_nonzero(got[~fin] != want[~fin]))
            # the reference prints six decimals: half a unit of the last one is print rounding, not error
            rel = np.maximum(np.abs(got[fin] - want[fin]) - 5e-7, 0.0) / np.maximum(np.abs(want[fin]), 1e-300)
            if rel.size:
                worst = max(worst, float(rel.max()))
                bad += int(np.count_nonzero(rel > 1e-5))
        return {"hmm_pairs": pairs, "hmm_mismatches": bad, "hmm_max_rel": worst, "hmm_tolerance": 1e-5}


def run_reference_arm(args):
    rank, _, world = env_rank()
    if rank != 0:
        return 0
    ref = CpuReference()
    steps, warm = args.steps, args.warmup
    # size the per-step sample so the whole run ends within a few minutes
    budget = max(1.0, min(8.0, 150.0 / (2 * (steps + warm))))
    sw_pairs = max(200, int(budget * 37e6 / (SW_LEN * SW_LEN)))
    hmm_batches = max(1, int(round(budget * 78e6 / 6.1e7)))
    ref.prepare_sw(sw_pairs)
    ref.prepare_hmm(hmm_batches)
    for _ in range(warm):
        ref.sw_gcups()
    sw_dt = [ref.sw_gcups()[1] for _ in range(steps)]
    for _ in range(warm):
        ref.hmm_gcups()
    hmm_dt = [ref.hmm_gcups()[1] for _ in range(steps)]
    sw_val = ref._sw_cells * steps / sum(sw_dt) / 1e9
    hmm_val = ref._hmm_cells * steps / sum(hmm_dt) / 1e9
    line = {
        "impl": "reference", "metric": "SW and PairHMM GCUPS", "value": sw_val, "unit": "GCUPS",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * sum(sw_dt) / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": "sw_short_150x150 (headline) + pairhmm_hc_shaped (\"pairhmm\" object); "
                               "reference C on host cores, bounded sample per step",
                   "sw_pairs_per_step": sw_pairs * ref.cores, "pairhmm_pairs_per_step": hmm_batches * 1000 * ref.cores},
        "cpu_baseline": ref.baseline_obj("sw", sw_val),
        "e2e": {"value": sw_val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "pairhmm": {"value": hmm_val, "unit": "GCUPS", "dtype": "f64", "ms_per_step": 1e3 * sum(hmm_dt) / steps,
                    "cpu_baseline": ref.baseline_obj("hmm", hmm_val),
                    "e2e": {"value": hmm_val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
    }
    emit(line)
    return 0


# --------------------------------------------------------------------------------------- GPU arm
def max_over_ranks(x: float, world: int, device) -> float:
    if world == 1:
        return float(x)
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world: int):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def h2d_only_ms(h_tensor, device, world, reps=3):
    """The bare upload of one step's input bytes from pinned memory, by every rank at the same time."""
    import torch
    d = torch.empty_like(h_tensor, device=device)
    best = 1e30
    for _ in range(reps):
        barrier(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d.copy_(h_tensor, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, max_over_ranks(e0.elapsed_time(e1), world, device))
    del d
    return best


def bench_sw(agx, args, rank, local_rank, world, device):
    import torch
    cap = agx.capi
    n = args.sw_pairs
    inp = agx.synth.sw_uniform_pairs(n, SW_LEN, seed=1000 + rank)
    cells = float(n) * SW_LEN * SW_LEN
    h_buf = torch.from_numpy(inp.buf).pin_memory()
    h_off = torch.from_numpy(inp.off).pin_memory()
    h_len = torch.from_numpy(inp.len).pin_memory()
    d_buf, d_off, d_len = h_buf.to(device), h_off.to(device), h_len.to(device)
    d_out = torch.empty(n, dtype=torch.int32, device=device)
    stream = torch.cuda.current_stream().cuda_stream

    def step_resident():
        cap.sw_score_device(local_rank, d_buf.data_ptr(), d_buf.numel(), d_off.data_ptr(), d_len.data_ptr(), n,
                            d_out.data_ptr(), stream)

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    barrier(world)
    cap.reset_launch_count()
    kern_ms, cls_ms = [], []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sampler as clk:
        ev0.record()
        for _ in range(args.steps):
            step_resident()
            kern_ms.append(cap.profile_ms(local_rank, cap.PROF_SW_DUO))
            cls_ms.append(cap.profile_ms(local_rank, cap.PROF_SW_CLASSIFY))
        ev1.record()
        torch.cuda.synchronize()
    launches = cap.launch_count()
    ms = max_over_ranks(ev0.elapsed_time(ev1) / args.steps, world, device)
    res_scores = d_out.cpu().numpy()

    # end to end through the host C ABI on pinned host buffers.  Headline: sw_score_file_image(), the call
    # the drop-in driver makes -- the raw generator.py-format file image in, scores out; fgets() chunking
    # on the GPU, upload segments overlapped with the DP kernels.  Second: sw_score_batch_flat() with
    # caller-built (offset, length) arrays.
    np_buf, np_off, np_len = h_buf.numpy(), h_off.numpy(), h_len.numpy()
    h_scores = torch.empty(n, dtype=torch.int32).pin_memory()
    np_scores = h_scores.numpy()
    # (libagx's own kernel spans are a measurement aid of the resident section: their timing events sit between the
    # overlapped regions of the end-to-end paths and cost 0.65 ms per call there -- profiles/r2bg_seg_probe2.jsonl)
    cap.set_profiling(False)
    for _ in range(min(args.warmup, 2)):
        img_scores, header, _ = cap.sw_score_file_image(np_buf, out=np_scores, copy=False)
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        img_scores, header, _ = cap.sw_score_file_image(np_buf, out=np_scores, copy=False)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / args.steps, world, device)
    assert header == 2 * n and np.array_equal(img_scores, res_scores), "file-image and device entry points disagree"
    for _ in range(min(args.warmup, 2)):
        e2e_scores = cap.sw_score_flat(np_buf, np_off, np_len, out=np_scores)
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_scores = cap.sw_score_flat(np_buf, np_off, np_len, out=np_scores)
    torch.cuda.synchronize()
    flat_ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / args.steps, world, device)
    assert np.array_equal(e2e_scores, res_scores), "host and device entry points disagree"
    cap.set_profiling(True)
    copy_ms = h2d_only_ms(h_buf, device, world)

    k_ms = float(np.mean(kern_ms))
    peaks = measured_peaks() or {}
    return {
        "value": world * cells / (ms * 1e-3) / 1e9, "ms_per_step": ms,
        "e2e": {"value": world * cells / (e2e_ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(np_buf.nbytes), "d2h_bytes_per_step": int(4 * n),
                "h2d_only_ms": copy_ms, "h2d_only_gbs_per_rank": np_buf.nbytes / (copy_ms * 1e-3) / 1e9,
                "entry_point": "sw_score_file_image (pinned file image in, pinned scores out)"},
        "e2e_flat": {"value": world * cells / (flat_ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": flat_ms,
                     "h2d_bytes_per_step": int(np_buf.nbytes + np_off.nbytes + np_len.nbytes),
                     "d2h_bytes_per_step": int(4 * n), "entry_point": "sw_score_batch_flat"},
        "roofline": roofline_obj(
            "alu", "sw_duo_kernel<8,19> (s16x2 DPX)", cells, k_ms, SW_OPS, "Tlaneop/s (INT32/DPX alu pipe)", "sw_duo_alu_pipe_active_pct",
            {"kernel_share_of_step": k_ms / ms,
             "traffic": (peaks["sw_duo_dram_bytes_per_pair_150x150"] * n if peaks.get("sw_duo_dram_bytes_per_pair_150x150") else None),
             "traffic_note": "DRAM bytes per launch from the committed ncu capture, scaled by pairs; the kernel is ALU-bound, "
                             "HBM time for this is ~0.05 ms",
             "loader": {"kernel": "sw_classify_kernel", "ms": float(np.mean(cls_ms)), "bound": "hbm"}}),
        "gpu_launches": int(launches), "clocks": clk.summary(),
        "pairs_per_gpu": n, "cells_per_gpu": cells,
    }


def hmm_device_arrays(inp, device):
    """Index arrays of an HmmInput as the device entry points take them (torch tensors on `device`)."""
    import torch
    nb = inp.n_batches
    nh_b = np.diff(inp.batch_hap_start)
    read_batch = np.repeat(np.arange(nb, dtype=np.int32), np.diff(inp.batch_read_start))
    out_off = np.concatenate(([0], np.cumsum(nh_b[read_batch])))[:-1].astype(np.int64)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return {"buf": t(inp.buf), "rfo": t(inp.read_field_off.reshape(-1)), "rl": t(inp.read_len), "rb": t(read_batch),
            "roo": t(out_off), "ho": t(inp.hap_off), "hl": t(inp.hap_len), "bhs": t(inp.batch_hap_start),
            "out": torch.empty(inp.n_pairs, dtype=torch.float64, device=device)}


def bench_hmm(agx, args, rank, local_rank, world, device):
    import torch
    cap = agx.capi
    inp = agx.synth.pairhmm_batches(args.hmm_batches, 200, 5, seed=2000 + rank, unrelated_frac=args.hmm_unrelated)
    cells = float(inp.cells())
    n_pairs = inp.n_pairs
    nb = inp.n_batches
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_buf = pin(inp.buf)
    h_rfo, h_rl, h_ho, h_hl = pin(inp.read_field_off.reshape(-1)), pin(inp.read_len), pin(inp.hap_off), pin(inp.hap_len)
    h_brs, h_bhs = pin(inp.batch_read_start), pin(inp.batch_hap_start)
    d = hmm_device_arrays(inp, device)
    stream = torch.cuda.current_stream().cuda_stream

    def step_resident():
        cap.pairhmm_forward_device(local_rank, d["buf"].data_ptr(), d["buf"].numel(), d["rfo"].data_ptr(), d["rl"].data_ptr(),
                                   d["rb"].data_ptr(), d["roo"].data_ptr(), inp.read_len.size, d["ho"].data_ptr(),
                                   d["hl"].data_ptr(), inp.hap_len.size, d["bhs"].data_ptr(), nb, n_pairs,
                                   d["out"].data_ptr(), stream, True)

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    barrier(world)
    cap.reset_launch_count()
    kern_ms, fp64_ms = [], []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sampler as clk:
        ev0.record()
        for _ in range(args.steps):
            step_resident()
            kern_ms.append(cap.profile_ms(local_rank, cap.PROF_HMM_STREAM))
            fp64_ms.append(cap.profile_ms(local_rank, cap.PROF_HMM_FP64))
        ev1.record()
        torch.cuda.synchronize()
    launches = cap.launch_count()
    rescued = cap.pairhmm_rescue_count(local_rank)
    ms = max_over_ranks(ev0.elapsed_time(ev1) / args.steps, world, device)
    res = d["out"].cpu().numpy()

    # end to end.  Headline: pairhmm_forward_file_image(), the call the drop-in driver makes -- the raw
    # pairHMM/test_set-format file image in pinned host memory in, log10 likelihoods in pinned host memory
    # out; batch walk and field splitting on the GPU.  Second: pairhmm_forward_batches_flat() with
    # caller-built index arrays.
    arrs = [x.numpy() for x in (h_buf, h_rfo, h_rl, h_ho, h_hl, h_brs, h_bhs)]
    cap.set_profiling(False)                 # (as in bench_sw: no timing events inside the end-to-end calls)
    for _ in range(min(args.warmup, 2)):
        img_vals, img_bp, img_inc = cap.pairhmm_forward_file_image(arrs[0], copy=False)
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        img_vals, img_bp, img_inc = cap.pairhmm_forward_file_image(arrs[0], copy=False)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / args.steps, world, device)
    assert img_inc == 0 and img_bp.size == nb and np.array_equal(img_vals, res, equal_nan=True), \
        "file-image and device entry points disagree"
    h_flat = torch.empty(n_pairs, dtype=torch.float64).pin_memory().numpy()     # pinned result array, like the inputs
    for _ in range(min(args.warmup, 2)):
        e2e = cap.pairhmm_forward_flat(*arrs, out=h_flat)
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e = cap.pairhmm_forward_flat(*arrs, out=h_flat)
    torch.cuda.synchronize()
    flat_ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / args.steps, world, device)
    assert np.array_equal(e2e, res, equal_nan=True), "host and device entry points disagree"
    assert not np.any(np.isnan(res)), "NaN PairHMM result"
    cap.set_profiling(True)
    copy_ms = h2d_only_ms(h_buf, device, world)

    k_ms = float(np.mean(kern_ms))
    f_ms = [x for x in fp64_ms if x >= 0]
    peaks = measured_peaks() or {}
    return {
        "value": world * cells / (ms * 1e-3) / 1e9, "unit": "GCUPS", "dtype": "f32 (+f64 rescue)", "ms_per_step": ms,
        "e2e": {"value": world * cells / (e2e_ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(arrs[0].nbytes), "d2h_bytes_per_step": int(8 * n_pairs + 4 * nb),
                "h2d_only_ms": copy_ms, "h2d_only_gbs_per_rank": arrs[0].nbytes / (copy_ms * 1e-3) / 1e9,
                "entry_point": "pairhmm_forward_file_image (pinned file image in, pinned results out)"},
        "e2e_flat": {"value": world * cells / (flat_ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": flat_ms,
                     "h2d_bytes_per_step": int(sum(a.nbytes for a in arrs)), "d2h_bytes_per_step": int(8 * n_pairs),
                     "entry_point": "pairhmm_forward_batches_flat"},
        "roofline": roofline_obj(
            "fp32", "hmm_duo_kernel<K=4..8> (FP32, two reads per warp, f32x2) + hmm_stream_kernel for unpaired reads",
            cells, k_ms, HMM_OPS, "Tlaneinstr/s (FP32 pipe)", "hmm_duo_fma_pipe_active_pct",
            {"kernel_share_of_step": k_ms / ms,
             "traffic": (peaks["hmm_dram_bytes_per_step_config4"] * (args.hmm_batches / 1000.0)
                         if peaks.get("hmm_dram_bytes_per_step_config4") else None),
             "traffic_note": "DRAM bytes of the stream-kernel launches of one step from the committed ncu launch list, "
                             "scaled by batches; FP32-bound, HBM time ~0.04 ms",
             # the pipe's measured rate depends on how many distinct registers an instruction reads: the cell is four
             # three-source FFMA2 and two two-source FMUL2
             "operand_limited_peak": (6.0 / (4.0 / peaks["fp32_three_source_tlaneops"] + 2.0 / peaks["fp32_tlaneops"])
                                      if peaks.get("fp32_three_source_tlaneops") else None),
             "operand_limited_peak_note": "measured 21.8 T lane-instr/s for FP32 instructions with three distinct register "
                                          "sources vs 35.09 with two (profiles/r1e_peaks_fp2.jsonl): the instruction mix of "
                                          "the cell cannot exceed this; frac is still quoted against the 35.09 peak"}),
        "fp64_rescue": {"unrelated_read_fraction": args.hmm_unrelated, "pairs_rescued": int(rescued),
                        "fraction_of_pairs": (rescued / n_pairs if rescued >= 0 else None),
                        "fp64_kernel_ms": (float(np.mean(f_ms)) if f_ms else 0.0),
                        "note": "pairs whose FP32 forward sum fell below 2^-100 (or was not finite) re-run by "
                                "hmm_striped_kernel<double,4>; inside the timed region"},
        "gpu_launches": int(launches), "clocks": clk.summary(),
        "pairs_per_gpu": n_pairs, "cells_per_gpu": cells,
    }


# ----------------------------------------------------------------- rank 0, one process, all N GPUs of the box
def expected_long_score(length, seed, related):
    try:
        for rec in json.loads((ROOT / "tests" / "golden" / "sw_long_expected.json").read_text()):
            if (rec["len"], rec["seed"], rec["related"]) == (length, seed, related):
                return rec["score"]
    except Exception:
        pass
    return None


def bench_sw_long(agx, args, n_gpus):
    """BASELINE configs[4]: one pair, columns striped over the GPUs inside this process (sw_long.cu)."""
    cap = agx.capi
    L = args.long_len
    inp = agx.formats.parse_sw(agx.synth.sw_long_pair(L, seed=5, related=True), line_buf=1 << 30)
    cells = float(L) * float(L)
    steps = max(1, min(args.steps, 5))

    def timed(devices):
        cap.shutdown()
        cap.init_devices(devices)
        cap.set_profiling(True)
        for _ in range(2):
            score = int(cap.sw_score_flat(inp.buf, inp.off, inp.len)[0])
        cap.reset_launch_count()
        with ClockSampler(devices[0]) as clk:
            t0 = time.perf_counter()
            for _ in range(steps):
                score = int(cap.sw_score_flat(inp.buf, inp.off, inp.len)[0])
            dt = (time.perf_counter() - t0) / steps
        k_ms = [cap.profile_ms(dv, cap.PROF_SW_LONG) for dv in devices]
        return dt * 1e3, score, k_ms, cap.launch_count() // steps, clk.summary()

    ms1, score1, k1, launches1, clocks1 = timed([0])
    out = {"workload": f"one pair {L} x {L} (synth.sw_long_pair seed 5, 2 % substitutions + 0.5 % indels), through sw_score_batch_flat",
           "unit": "GCUPS", "n_gpus": n_gpus, "scaling": "strong", "dtype": "int32", "steps": steps,
           "ms_1gpu": ms1, "gcups_1gpu": cells / (ms1 * 1e-3) / 1e9, "kernel_ms_1gpu": k1[0]}
    ms, score, k_ms, launches, clocks = ms1, score1, k1, launches1, clocks1
    if n_gpus > 1:
        ms, score, k_ms, launches, clocks = timed(list(range(n_gpus)))
        out["score_matches_1gpu"] = bool(score == score1)
    expected = expected_long_score(L, 5, True)
    out.update({"ms": ms, "value": cells / (ms * 1e-3) / 1e9, "speedup_vs_1gpu": ms1 / ms,
                "pipeline_efficiency": ms1 / (n_gpus * ms), "kernel_ms_per_gpu": k_ms, "gpu_launches": int(launches),
                "score": score, "score_expected": expected,
                "score_expected_source": "tests/golden/sw_long_expected.json: oracle/sw_blocked.c on host cores (tests/golden/make_long_expected.py)",
                "score_ok": (None if expected is None else bool(score == expected and score1 == expected)),
                "h2d_bytes_per_step": int(2 * (L + 1)) if n_gpus == 1 else int((L + 1) * (n_gpus + 1)),
                "d2h_bytes_per_step": 4 * n_gpus, "clocks": clocks,
                "roofline": roofline_obj("alu", "sw_longr_kernel<K,R> (s32 DPX, symbol-coded, dp4a substitution), one GPU",
                                         cells, k1[0] if k1[0] > 0 else ms1, LONG_OPS, "Tlaneop/s (INT32/DPX alu pipe)",
                                         "sw_long_alu_pipe_active_pct", {"traffic": None})})
    return out


def split_hmm(inp, parts):
    """`parts` contiguous groups of whole batches, each as its own HmmInput (offsets rebased)."""
    from types import SimpleNamespace
    nb = inp.n_batches
    cuts = [nb * k // parts for k in range(parts + 1)]
    out = []
    for k in range(parts):
        b0, b1 = cuts[k], cuts[k + 1]
        r0, r1 = int(inp.batch_read_start[b0]), int(inp.batch_read_start[b1])
        h0, h1 = int(inp.batch_hap_start[b0]), int(inp.batch_hap_start[b1])
        lo = int(min(inp.read_field_off[r0:r1].min(), inp.hap_off[h0:h1].min()))
        hi = int(max((inp.read_field_off[r0:r1, 4] + inp.read_len[r0:r1]).max(), (inp.hap_off[h0:h1] + inp.hap_len[h0:h1]).max()))
        nr_b = np.diff(inp.batch_read_start[b0:b1 + 1])
        nh_b = np.diff(inp.batch_hap_start[b0:b1 + 1])
        out.append(SimpleNamespace(
            buf=inp.buf[lo:hi], read_field_off=inp.read_field_off[r0:r1] - lo, read_len=inp.read_len[r0:r1],
            hap_off=inp.hap_off[h0:h1] - lo, hap_len=inp.hap_len[h0:h1],
            batch_read_start=inp.batch_read_start[b0:b1 + 1] - r0, batch_hap_start=inp.batch_hap_start[b0:b1 + 1] - h0,
            n_batches=b1 - b0, n_pairs=int(np.sum(nr_b * nh_b))))
    return out


def bench_strong(agx, args, n_gpus):
    """ONE 10^6-pair batch of each kind through the library's in-process dispatcher bound to all n_gpus GPUs:
    end to end from pinned host buffers, and resident (one device-resident shard per GPU)."""
    import torch
    cap = agx.capi
    steps = max(1, args.steps)
    out = {"n_gpus": n_gpus, "scaling": "strong", "unit": "GCUPS",
           "note": "one process, agx_init over all GPUs, one host thread + stream per GPU, no collective"}

    def wall(fn, warm=2):
        for _ in range(warm):
            r = fn()
        for dv in range(n_gpus):
            torch.cuda.synchronize(dv)
        t0 = time.perf_counter()
        for _ in range(steps):
            r = fn()
        return 1e3 * (time.perf_counter() - t0) / steps, r

    # ---- Smith-Waterman -------------------------------------------------------------------------------
    n = args.sw_pairs
    inp = agx.synth.sw_uniform_pairs(n, SW_LEN, seed=1000)
    cells = float(n) * SW_LEN * SW_LEN
    h_buf = torch.from_numpy(inp.buf).pin_memory()
    np_buf = h_buf.numpy()
    h_scores = torch.empty(n, dtype=torch.int32).pin_memory()
    off_pin = torch.from_numpy(inp.off).pin_memory().numpy()
    len_pin = torch.from_numpy(inp.len).pin_memory().numpy()
    align_keep = {}
    pinned = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
    align_out = (pinned(n, torch.int32), pinned((n, 4), torch.int32), pinned(n + 1, torch.int64),
                 pinned(8 * n, torch.int32).view(np.uint32))
    res = {}
    for label, devs in (("1gpu", [0]), ("ngpu", list(range(n_gpus)))):
        if label == "ngpu" and n_gpus == 1:
            res["ngpu"] = res["1gpu"]
            break
        cap.shutdown()
        cap.init_devices(devs)
        cap.set_profiling(False)                # end-to-end calls without libagx's timing events (see bench_sw)
        img_ms, (img_scores, _, _) = wall(lambda: cap.sw_score_file_image(np_buf, out=h_scores.numpy(), copy=False))
        img_scores = img_scores.copy()       # (a view of h_scores until here)
        flat_ms, flat_scores = wall(lambda: cap.sw_score_flat(np_buf, off_pin, len_pin, out=h_scores.numpy()))
        flat_scores = flat_scores.copy()
        assert np.array_equal(img_scores, flat_scores), "file-image and flat entry points disagree"
        # full alignments (end cell, start cell, CIGAR) of the same batch, pairs sharded over the GPUs
        align_ms, align_res = wall(lambda: cap.sw_align_flat(np_buf, off_pin, len_pin, out=align_out), warm=1)
        assert np.array_equal(align_res[0], flat_scores), "alignment scores and score-only scores disagree"
        align_keep[label] = tuple(a.copy() for a in align_res)      # (views of the pinned result arrays until here)
        cap.set_profiling(True)
        # resident: contiguous shards of equal pair count, uploaded once
        g = len(devs)
        shards, keep = [], []
        for k, dv in enumerate(devs):
            p0, p1 = n * k // g, n * (k + 1) // g
            b0 = int(inp.off[2 * p0]) if p0 < n else int(inp.buf.size)
            b1 = int(inp.off[2 * p1 - 1] + inp.len[2 * p1 - 1])
            dev = torch.device("cuda", dv)
            t_buf = torch.from_numpy(inp.buf[b0:b1]).to(dev)
            t_off = torch.from_numpy(inp.off[2 * p0:2 * p1] - b0).to(dev)
            t_len = torch.from_numpy(inp.len[2 * p0:2 * p1]).to(dev)
            t_out = torch.empty(p1 - p0, dtype=torch.int32, device=dev)
            keep.append((t_buf, t_off, t_len, t_out))
            shards.append(cap.SwShard(dv, t_buf.data_ptr(), t_buf.numel(), t_off.data_ptr(), t_len.data_ptr(), p1 - p0, t_out.data_ptr()))
        res_ms, _ = wall(lambda: cap.sw_score_shards_device(shards), warm=3)
        k_ms = max(cap.profile_ms(dv, cap.PROF_SW_DUO) for dv in devs)
        got = np.concatenate([t[3].cpu().numpy() for t in keep])
        assert np.array_equal(got, flat_scores), "resident shards and host entry points disagree"
        res[label] = {"e2e_file_image_ms": img_ms, "e2e_flat_ms": flat_ms, "resident_ms": res_ms, "kernel_ms_max_over_gpus": k_ms,
                      "align_ms": align_ms}
        del keep, shards
    one, many = res["1gpu"], res["ngpu"]
    out["sw"] = {"workload": f"{n} pairs of {SW_LEN}x{SW_LEN} in ONE batch",
                 "resident": {"value": cells / (many["resident_ms"] * 1e-3) / 1e9, "ms": many["resident_ms"], "ms_1gpu": one["resident_ms"],
                              "speedup_vs_1gpu": one["resident_ms"] / many["resident_ms"],
                              "kernel_ms_max_over_gpus": many["kernel_ms_max_over_gpus"], "entry_point": "sw_score_shards_device"},
                 "e2e": {"value": cells / (many["e2e_file_image_ms"] * 1e-3) / 1e9, "ms": many["e2e_file_image_ms"],
                         "ms_1gpu": one["e2e_file_image_ms"], "speedup_vs_1gpu": one["e2e_file_image_ms"] / many["e2e_file_image_ms"],
                         "h2d_bytes_per_step": int(np_buf.nbytes), "d2h_bytes_per_step": int(4 * n), "entry_point": "sw_score_file_image"},
                 "e2e_flat": {"value": cells / (many["e2e_flat_ms"] * 1e-3) / 1e9, "ms": many["e2e_flat_ms"], "ms_1gpu": one["e2e_flat_ms"],
                              "speedup_vs_1gpu": one["e2e_flat_ms"] / many["e2e_flat_ms"], "entry_point": "sw_score_batch_flat"},
                 "e2e_align": {"value": cells / (many["align_ms"] * 1e-3) / 1e9, "ms": many["align_ms"], "ms_1gpu": one["align_ms"],
                               "speedup_vs_1gpu": one["align_ms"] / many["align_ms"], "entry_point": "sw_align_batch_flat",
                               "note": "scores, start / end cells and CIGAR of every pair; pairs sharded over the GPUs, runs "
                                       "concatenated in pair order",
                               "equal_to_1gpu": bool(len(align_keep) < 2 or all(np.array_equal(x, y) for x, y in
                                                                                 zip(align_keep["1gpu"], align_keep["ngpu"])))}}
    del h_buf, h_scores, align_keep, align_out

    # ---- PairHMM --------------------------------------------------------------------------------------
    hin = agx.synth.pairhmm_batches(args.hmm_batches, 200, 5, seed=2000, unrelated_frac=args.hmm_unrelated)
    hcells = float(hin.cells())
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    arrs = [pin(x).numpy() for x in (hin.buf, hin.read_field_off.reshape(-1), hin.read_len, hin.hap_off, hin.hap_len,
                                     hin.batch_read_start, hin.batch_hap_start)]
    res = {}
    for label, devs in (("1gpu", [0]), ("ngpu", list(range(n_gpus)))):
        if label == "ngpu" and n_gpus == 1:
            res["ngpu"] = res["1gpu"]
            break
        cap.shutdown()
        cap.init_devices(devs)
        cap.set_profiling(False)
        img_ms, (img_vals, _, _) = wall(lambda: cap.pairhmm_forward_file_image(arrs[0], copy=False))
        img_vals = img_vals.copy()           # (a view of the library's pinned result buffer until here)
        flat_ms, flat_vals = wall(lambda: cap.pairhmm_forward_flat(*arrs))
        cap.set_profiling(True)
        assert np.array_equal(img_vals, flat_vals, equal_nan=True), "file-image and flat entry points disagree"
        parts = split_hmm(hin, len(devs))
        shards, keep = [], []
        for part, dv in zip(parts, devs):
            d = hmm_device_arrays(part, torch.device("cuda", dv))
            keep.append(d)
            shards.append(cap.HmmShard(dv, d["buf"].data_ptr(), d["buf"].numel(), d["rfo"].data_ptr(), d["rl"].data_ptr(),
                                       d["rb"].data_ptr(), d["roo"].data_ptr(), part.read_len.size, d["ho"].data_ptr(),
                                       d["hl"].data_ptr(), part.hap_len.size, d["bhs"].data_ptr(), part.n_batches, part.n_pairs,
                                       d["out"].data_ptr()))
        res_ms, _ = wall(lambda: cap.pairhmm_forward_shards_device(shards, True), warm=3)
        k_ms = max(cap.profile_ms(dv, cap.PROF_HMM_STREAM) for dv in devs)
        got = np.concatenate([d["out"].cpu().numpy() for d in keep])
        assert np.array_equal(got, flat_vals, equal_nan=True), "resident shards and host entry points disagree"
        res[label] = {"e2e_file_image_ms": img_ms, "e2e_flat_ms": flat_ms, "resident_ms": res_ms, "kernel_ms_max_over_gpus": k_ms}
        del keep, shards
    one, many = res["1gpu"], res["ngpu"]
    out["pairhmm"] = {"workload": f"{hin.n_pairs} (read, haplotype) pairs in ONE call ({hin.n_batches} batches)",
                      "resident": {"value": hcells / (many["resident_ms"] * 1e-3) / 1e9, "ms": many["resident_ms"], "ms_1gpu": one["resident_ms"],
                                   "speedup_vs_1gpu": one["resident_ms"] / many["resident_ms"],
                                   "kernel_ms_max_over_gpus": many["kernel_ms_max_over_gpus"], "entry_point": "pairhmm_forward_shards_device"},
                      "e2e": {"value": hcells / (many["e2e_file_image_ms"] * 1e-3) / 1e9, "ms": many["e2e_file_image_ms"],
                              "ms_1gpu": one["e2e_file_image_ms"], "speedup_vs_1gpu": one["e2e_file_image_ms"] / many["e2e_file_image_ms"],
                              "h2d_bytes_per_step": int(arrs[0].nbytes), "d2h_bytes_per_step": int(8 * hin.n_pairs),
                              "entry_point": "pairhmm_forward_file_image"},
                      "e2e_flat": {"value": hcells / (many["e2e_flat_ms"] * 1e-3) / 1e9, "ms": many["e2e_flat_ms"], "ms_1gpu": one["e2e_flat_ms"],
                                   "speedup_vs_1gpu": one["e2e_flat_ms"] / many["e2e_flat_ms"], "entry_point": "pairhmm_forward_batches_flat"}}
    return out


def bench_gatk(agx, args, device_index):
    """configs[3] with the corrected GATK priors -- reported separately, never a parity claim against the reference."""
    import torch
    cap = agx.capi
    device = torch.device("cuda", device_index)
    inp = agx.synth.pairhmm_batches(args.hmm_batches, 200, 5, seed=2000, unrelated_frac=args.hmm_unrelated)
    d = hmm_device_arrays(inp, device)
    stream = torch.cuda.current_stream(device).cuda_stream
    cells = float(inp.cells())
    out = {"unit": "GCUPS", "workload": "the \"pairhmm\" workload of rank 0"}

    def step():
        cap.pairhmm_forward_device(device_index, d["buf"].data_ptr(), d["buf"].numel(), d["rfo"].data_ptr(), d["rl"].data_ptr(),
                                   d["rb"].data_ptr(), d["roo"].data_ptr(), inp.read_len.size, d["ho"].data_ptr(),
                                   d["hl"].data_ptr(), inp.hap_len.size, d["bhs"].data_ptr(), inp.n_batches, inp.n_pairs,
                                   d["out"].data_ptr(), stream, True)

    results = {}
    try:
        for mode, name in ((0, "reference"), (1, "gatk"), (3, "gatk_qual_floor")):
            cap.set_pairhmm_gatk_mode(mode)
            for _ in range(2):
                step()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                step()
            e1.record()
            torch.cuda.synchronize(device)
            results[name] = d["out"].cpu().numpy().copy()
            if mode:
                ms = e0.elapsed_time(e1) / args.steps
                out[name] = {"value": cells / (ms * 1e-3) / 1e9, "ms_per_step": ms}
    finally:
        cap.set_pairhmm_gatk_mode(0)
    diff = results["reference"] - results["gatk"]                     # >= 0: the /3 can only lower a likelihood
    fin = np.isfinite(diff)
    out["gatk"].update({"mismatch_prior": "Qr/3", "pairs_that_differ_from_reference_mode": int(np.count_nonzero(np.abs(diff[fin]) > 1e-9)),
                        "mean_log10_shift": float(diff[fin].mean()), "max_log10_shift": float(diff[fin].max()),
                        "log10_3": float(np.log10(3.0))})
    out["gatk_qual_floor"].update({"mismatch_prior": "Qr/3", "base_quality_floor": 6,
                                   "pairs_that_differ_from_gatk_mode": int(np.count_nonzero(results["gatk"] != results["gatk_qual_floor"]))})
    out["pin"] = gatk_pin(agx)
    out["_results"] = results
    out["_input"] = inp
    return out


def gatk_pin(agx):
    """libagx's GATK mode against tests/golden/pairhmm_gatk.json: values of an independent full-matrix LoglessPairHMM
    statement (tests/golden/make_gatk_golden.py) on the reference's own test_set inputs (test.in, 10s.in)."""
    import gzip
    import json as js
    cap = agx.capi
    gold = ROOT / "tests" / "golden"
    if not (gold / "pairhmm_gatk.json").exists():
        return None
    files = js.loads((gold / "pairhmm_gatk.json").read_text())["files"]
    rec = {"source": "tests/golden/pairhmm_gatk.json (independent LoglessPairHMM statement, make_gatk_golden.py)", "pairs": 0,
           "max_rel": 0.0, "tolerance": 1e-5}
    try:
        for name, rows in files.items():
            inp = agx.formats.parse_pairhmm(gzip.decompress((gold / name).read_bytes()))
            nh, nr = np.diff(inp.batch_hap_start), np.diff(inp.batch_read_start)
            base = np.concatenate(([0], np.cumsum(nr * nh)))
            for mode, key in ((1, "gatk"), (3, "gatk_floor")):
                cap.set_pairhmm_gatk_mode(mode)
                got = cap.pairhmm_forward_flat(inp.buf, inp.read_field_off, inp.read_len, inp.hap_off, inp.hap_len,
                                               inp.batch_read_start, inp.batch_hap_start)
                for r in rows:
                    k = int(base[r["batch"]] + r["read"] * nh[r["batch"]] + r["hap"])
                    rec["max_rel"] = max(rec["max_rel"], abs(got[k] - r[key]) / abs(r[key]))
                    rec["pairs"] += 1
    finally:
        cap.set_pairhmm_gatk_mode(0)
    rec["ok"] = bool(rec["max_rel"] <= rec["tolerance"])
    return rec


def bench_sw_lengths(agx, args, device_index):
    """The inter-task kernel one length class at a time: the published MI210 sweep lengths (hiprun.sh:18) and
    generator.py's own 450-500 bp (generator.py:4-5)."""
    import torch
    cap = agx.capi
    device = torch.device("cuda", device_index)
    stream = torch.cuda.current_stream(device).cuda_stream
    peak, _ = pipe_peak("alu")
    rows = []
    rng = np.random.default_rng(4242)
    for spec in args.sw_len.split(","):
        spec = spec.strip()
        if not spec:
            continue
        if "-" in spec:
            lo, hi = (int(x) for x in spec.split("-"))
            n = max(2000, int(args.sw_len_cells / (0.25 * (lo + hi) ** 2)))
            inp = agx.formats.parse_sw(agx.synth.sw_random_file(rng, n, lo, hi), line_buf=1 << 20)
        else:
            lo = hi = int(spec)
            n = max(2000, int(args.sw_len_cells / (lo * lo)))
            inp = agx.synth.sw_uniform_pairs(n, lo, seed=500 + lo)
        plain = inp.len.astype(np.int64) - (inp.buf[inp.off + inp.len - 1] == 10)      # without the newline symbol
        cells = float(np.sum(plain[0::2] * plain[1::2]))
        d_buf, d_off, d_len = (torch.from_numpy(x).to(device) for x in (inp.buf, inp.off, inp.len))
        d_out = torch.empty(inp.off.size // 2, dtype=torch.int32, device=device)
        step = lambda: cap.sw_score_device(device_index, d_buf.data_ptr(), d_buf.numel(), d_off.data_ptr(), d_len.data_ptr(),
                                           inp.off.size // 2, d_out.data_ptr(), stream)
        for _ in range(3):
            step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k_ms = []
        e0.record()
        for _ in range(args.steps):
            step()
            k_ms.append(cap.profile_ms(device_index, cap.PROF_SW_DUO))
        e1.record()
        torch.cuda.synchronize(device)
        ms = e0.elapsed_time(e1) / args.steps
        k = float(np.mean(k_ms))
        rows.append({"len": spec, "pairs": int(inp.off.size // 2), "cells": cells, "ms_per_step": ms, "value": cells / (ms * 1e-3) / 1e9,
                     "kernel_ms": k, "kernel_gcups": cells / (k * 1e-3) / 1e9,
                     "alu_frac_executed": cells * SW_OPS["executed"] / (k * 1e-3) / peak})
    return {"unit": "GCUPS", "ops_per_cell_executed": SW_OPS["executed"],
            "note": "resident batch of one length class; cells = len_a x len_b without the newline symbol; padding columns/rows "
                    "of the class count as lost throughput", "lengths": rows}


def bench_sw_align(agx, args, device_index):
    """The alignment path on top of the headline batch: end cells
    (sw_ends_batch_flat) and full alignments (sw_align_batch_flat: start cells + CIGAR) of the same 150 x 150
    pairs, end to end from pinned host buffers, with the device spans libagx times itself (DP kernels; traceback
    walk) and the two resources the storing kernel leans on."""
    import torch
    cap = agx.capi
    n = min(args.sw_pairs, args.align_pairs)
    inp = agx.synth.sw_uniform_pairs(n, SW_LEN, seed=1234)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    buf, off, ln = pin(inp.buf), pin(inp.off), pin(inp.len)
    cells = float(n) * SW_LEN * SW_LEN
    out = {"unit": "GCUPS", "workload": "%d pairs of %dx%d (sw_uniform_pairs seed 1234)" % (n, SW_LEN, SW_LEN),
           "end_cell": "the cell the reference's running maximum comes from (antidiagonalSmithWaterman.c:335; its visiting "
                       "order); start cell + CIGAR: oracle/sw_align.c's rule"}
    scores = torch.empty(n, dtype=torch.int32).pin_memory().numpy()
    ends = torch.empty((n, 2), dtype=torch.int32).pin_memory().numpy()
    coords = torch.empty((n, 4), dtype=torch.int32).pin_memory().numpy()
    coff = torch.empty(n + 1, dtype=torch.int64).pin_memory().numpy()
    cig = torch.empty(16 * n, dtype=torch.uint32 if hasattr(torch, "uint32") else torch.int32).pin_memory().numpy().view(np.uint32)
    lib = cap.load_library()
    import ctypes as C
    total = C.c_int64(0)
    sc = [cap.SW_MATCH, cap.SW_MISMATCH, cap.SW_GAP_OPEN, cap.SW_GAP_EXTEND]
    P = lambda a: a.ctypes.data

    def run_ends():
        rc = lib.sw_ends_batch_flat(P(buf), buf.size, P(off), P(ln), n, *sc, P(scores), P(ends))
        assert rc == 0, lib.agx_last_error().decode()

    def run_align():
        rc = lib.sw_align_batch_flat(P(buf), buf.size, P(off), P(ln), n, *sc, P(scores), P(coords), P(coff), P(cig), cig.size,
                                     C.byref(total))
        assert rc == 0, lib.agx_last_error().decode()

    # the same two calls with everything resident on the device (sw_ends_batch_device / sw_align_batch_device)
    dev = torch.device("cuda", device_index)
    d_buf, d_off, d_len = (torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (inp.buf, inp.off, inp.len))
    d_scores = torch.empty(n, dtype=torch.int32, device=dev)
    d_ends = torch.empty((n, 2), dtype=torch.int32, device=dev)
    d_coords = torch.empty((n, 4), dtype=torch.int32, device=dev)
    d_coff = torch.empty(n + 1, dtype=torch.int64, device=dev)
    d_cig = torch.empty(8 * n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    resident = {
        "ends": lambda: cap.sw_ends_device(device_index, d_buf.data_ptr(), d_buf.numel(), d_off.data_ptr(), d_len.data_ptr(), n,
                                           d_scores.data_ptr(), d_ends.data_ptr(), stream),
        "align": lambda: cap.sw_align_device(device_index, d_buf.data_ptr(), d_buf.numel(), d_off.data_ptr(), d_len.data_ptr(), n,
                                             d_scores.data_ptr(), d_coords.data_ptr(), d_coff.data_ptr(), d_cig.data_ptr(),
                                             d_cig.numel(), stream)}

    def resident_ms(fn):
        for _ in range(3):
            fn()
        # (torch's current stream is the legacy default stream = "NULL": libagx then works on its own non-blocking
        # stream, which events recorded on torch's stream do not order against -- so: device-wide synchronise and
        # the host clock around the calls)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        torch.cuda.synchronize(dev)
        return 1e3 * (time.perf_counter() - t0) / args.steps

    peak_alu, src = pipe_peak("alu")
    hbm, hbm_src = hbm_peak()
    for name, fn in (("ends", run_ends), ("align", run_align)):
        for _ in range(max(2, args.warmup - 1)):
            fn()
        cap.set_profiling(False)               # the timed calls run without libagx's timing events (see bench_sw) ...
        cap.reset_launch_count()
        t, dp, wk = [], [], []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            fn()
            t.append(time.perf_counter() - t0)
        launches = cap.launch_count() // args.steps
        cap.set_profiling(True)                # ... and two more collect the device spans of the kernels
        for _ in range(2):
            fn()
            dp.append(cap.profile_ms(device_index, 7))
            wk.append(cap.profile_ms(device_index, 8))
        ms, dp_ms = 1e3 * float(np.mean(t)), float(np.mean(dp))
        rec = {"e2e": {"value": cells / (ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": ms,
                       "h2d_bytes_per_step": int(buf.nbytes + off.nbytes + ln.nbytes),
                       "entry_point": "sw_ends_batch_flat" if name == "ends" else "sw_align_batch_flat"},
               "dp_kernel_ms": dp_ms, "gpu_launches": launches}
        ops = ALIGN_OPS[name]
        rec["roofline"] = {"bound": "alu", "pipe": "INT32/DPX alu pipe", "kernel": "sw_duo_kernel<8,19,%d>" % (1 if name == "ends" else 2),
                           "achieved": cells * ops / (dp_ms * 1e-3) / 1e12, "peak": peak_alu / 1e12,
                           "unit": "Tlaneop/s (INT32/DPX alu pipe)", "frac": cells * ops / (dp_ms * 1e-3) / peak_alu,
                           "ops_per_cell_executed": ops, "peak_source": src, "kernel_ms": dp_ms,
                           "kernel_gcups": cells / (dp_ms * 1e-3) / 1e9}
        ncu = (measured_peaks() or {}).get("ncu", {})
        key = "sw_ends_alu_pipe_active_pct" if name == "ends" else "sw_align_alu_pipe_active_pct"
        if ncu.get(key) is not None:
            rec["roofline"]["pipe_active_ncu"] = ncu[key]
            rec["roofline"]["pipe_active_ncu_source"] = ncu.get(key + "_source")
        res_ms = resident_ms(resident[name])
        rec["value"] = cells / (res_ms * 1e-3) / 1e9
        rec["ms_per_step"] = res_ms
        rec["value_note"] = ("the whole device-resident call (%s: length classes, DP kernels%s), CUDA events around %d calls" %
                             ("sw_ends_batch_device" if name == "ends" else "sw_align_batch_device",
                              "" if name == "ends" else ", traceback walk, scan + gather of the runs", args.steps)).replace("CUDA events around", "device synchronised, host clock around")
        if name == "ends":
            rec["e2e"]["d2h_bytes_per_step"] = int(scores.nbytes + ends.nbytes)
        else:
            walk_ms = float(np.mean(wk))
            runs = int(total.value)
            # the storing kernel writes one byte per computed cell: 164 steps x 32 lanes x 40 bytes per warp of 8 pairs
            tb_bytes = float(-(-n // 8)) * (SW_LEN + 1 + 7) * 32 * 40
            rec.update({"walk_kernel_ms": walk_ms, "kernels_gcups": cells / ((dp_ms + walk_ms) * 1e-3) / 1e9,
                        "cigar_runs": runs, "matrix_bytes": tb_bytes})
            rec["e2e"]["d2h_bytes_per_step"] = int(scores.nbytes + coords.nbytes + coff.nbytes + 4 * runs)
            rec["roofline"]["hbm"] = {"bound": "hbm", "achieved": tb_bytes / (dp_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                      "frac": tb_bytes / (dp_ms * 1e-3) / 1e9 / hbm,
                                      "traffic": tb_bytes, "peak_source": hbm_src,
                                      "note": "bulk stores of the H-byte matrices (cp.async.bulk shared -> global); the kernel is "
                                              "shared between this and the alu pipe"}
        out[name] = rec
    # consistency inside the run: alignment scores = end-cell scores = score-only scores; ends agree
    s0 = cap.sw_score_flat(buf, off, ln)
    out["scores_equal_score_only"] = bool(np.array_equal(s0, scores))
    out["ends_equal"] = bool(np.array_equal(coords[:, 1], ends[:, 0]) and np.array_equal(coords[:, 3], ends[:, 1]))
    out["device_entry_equals_host_entry"] = bool(np.array_equal(d_scores.cpu().numpy(), scores) and
                                                 np.array_equal(d_coords.cpu().numpy(), coords) and
                                                 np.array_equal(d_coff.cpu().numpy(), coff))
    out["_check"] = (inp, scores.copy(), coords.copy(), coff.copy(), cig[:int(total.value)].copy())
    return out


def cpu_leg(agx, args, gatk, align=None):
    """cpu_baseline (rank 0, N = 1): the reference C programs on the host cores, the GPU against their output on
    the same shards, and the GATK mode against the oracle's GATK branch.  The one place bench.py's GPU arm runs
    anything under oracle/."""
    cap = agx.capi
    ref = CpuReference()
    out = {}
    # >= 10^6 pairs of each configuration once per run when the box has the cores for it (16 cores: ~20 s each)
    budget_pairs = args.cpu_seconds * ref.cores * 60e6 / (SW_LEN * SW_LEN)
    sw_pairs = -(-int(min(args.sw_pairs, budget_pairs)) // ref.cores)
    ref.prepare_sw(max(200, sw_pairs))
    v, dt = ref.sw_gcups()
    out["cpu_sw"] = ref.baseline_obj("sw", v)
    out["cpu_sw"]["seconds"] = dt
    budget_batches = args.cpu_seconds * ref.cores * 160e6 / 6.1e7
    hmm_batches = -(-int(min(args.hmm_batches, budget_batches)) // ref.cores)
    ref.prepare_hmm(max(1, hmm_batches), unrelated_frac=args.hmm_unrelated)
    v, dt = ref.hmm_gcups()
    out["cpu_hmm"] = ref.baseline_obj("hmm", v)
    out["cpu_hmm"]["seconds"] = dt
    par = {"reference": ("oracle/_ref programs' own output (Score: %d lines / %f file)" if ref.kind == "reference" else "oracle port"),
           "gpu_entry_points": "sw_score_file_image / pairhmm_forward_file_image on the same shard files"}
    par.update(ref.parity_sw(cap))
    par.update(ref.parity_hmm(cap))
    out["parity"] = par
    # BASELINE configs[4]'s CPU side: the reference with a raised line buffer (oracle/_ref/sw_antidiag_long, arithmetic
    # untouched) on a square the core finishes in seconds, the GPU's long-alignment kernel on the same pair, and the
    # time 1 Mbp x 1 Mbp would take -- an EXTRAPOLATION by cell count, labelled as one
    long_exe = ROOT / "oracle" / "_ref" / "sw_antidiag_long"
    if long_exe.exists() and args.long_cpu_len > 0:
        L = args.long_cpu_len
        data = agx.synth.sw_long_pair(L, seed=21, related=True)
        pth = Path(ref.tmp.name) / "long.in"
        pth.write_bytes(data)
        t0 = time.perf_counter()
        r = subprocess.run(["bash", "-c", f"ulimit -s unlimited 2>/dev/null; exec {long_exe} {pth}"], capture_output=True)
        dt = time.perf_counter() - t0
        want = [int(l.split()[1]) for l in r.stdout.decode(errors="replace").splitlines() if l.startswith("Score:")]
        inp = agx.formats.parse_sw(data, line_buf=1 << 30)
        # (4 * 10^8 cells >= 2^28: the pair takes the whole-GPU long-alignment kernel)
        got = cap.sw_score_flat(inp.buf, inp.off, inp.len)
        cells = float(inp.len[0]) * float(inp.len[1])
        out["cpu_sw_long"] = {"value": cells / dt / 1e9, "unit": "GCUPS", "cores": 1, "kind": "reference",
                              "sample": f"one pair {L} x {L} (oracle/_ref/sw_antidiag_long = antidiagonalSmithWaterman.c -O3 with "
                                        f"MAX_LINE_LENGTH raised), {dt:.1f} s",
                              "extrapolated_seconds_1mbp_one_core": dt * (1e12 / cells),
                              "extrapolation": "by cell count from the sample above; not measured",
                              "gpu_score_equals_reference": bool(len(want) == 1 and got.size == 1 and int(got[0]) == want[0]),
                              "score": int(got[0]) if got.size else None}
    if align is not None:
        # alignments of a sample against the oracle (end cell, start cell, CIGAR) and EVERY CIGAR of a larger sample
        # re-scored to its Smith-Waterman score
        import oracle
        inp, scores, coords, coff, cig = align["_check"]
        data = inp.buf.tobytes()
        rng = np.random.default_rng(9)
        n = scores.size
        exact = rescored = bad = 0
        for k, p in enumerate(rng.choice(n, size=min(n, 6000), replace=False).tolist()):
            a = data[inp.off[2 * p]:inp.off[2 * p] + inp.len[2 * p]]
            b = data[inp.off[2 * p + 1]:inp.off[2 * p + 1] + inp.len[2 * p + 1]]
            runs = cig[coff[p]:coff[p + 1]].tolist()
            if k < 1500:
                ws, wc, wg = oracle.sw_align(a, b)
                bad += int((ws, list(wc), wg) != (int(scores[p]), coords[p].tolist(), runs))
                exact += 1
            elif scores[p] > 0:
                bad += int(oracle.sw_cigar_score(a, b, coords[p].tolist(), runs) != int(scores[p]))
                rescored += 1
        align["parity"] = {"pairs_vs_oracle": exact, "cigars_rescored": rescored, "mismatches": bad}
    if gatk is not None:
        import oracle
        inp, results = gatk["_input"], gatk["_results"]
        n_check = 400
        for name, mode in (("gatk", 1), ("gatk_qual_floor", 3)):
            want = oracle.pairhmm_flat(inp, gatk=mode, limit=n_check)
            got = results[name][:want.size]
            fin = np.isfinite(want)
            assert np.array_equal(got[~fin], want[~fin]), "GATK mode: non-finite results differ from the oracle's"
            gatk[name]["max_rel_vs_oracle_gatk_branch"] = float(np.max(np.abs(got[fin] - want[fin]) / np.abs(want[fin])))
            gatk[name]["pairs_checked"] = int(want.size)
    return out


def run_sw_long(args):
    """--workload sw_long: BASELINE configs[4] alone, as the headline of the line."""
    import torch
    import agxpkg
    agx = agxpkg.load()
    n_gpus = min(args.gpus, torch.cuda.device_count())
    obj = bench_sw_long(agx, args, n_gpus)
    L = args.long_len
    line = {"metric": "SW and PairHMM GCUPS", "value": obj["value"], "unit": "GCUPS", "n_gpus": n_gpus,
            "steps": obj["steps"], "warmup": 2, "ms_per_step": obj["ms"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": f"sw_long: one pair {L}x{L} (BASELINE configs[4]), column stripes over {n_gpus} GPU(s) in one process"},
            "e2e": {"value": obj["value"], "unit": "GCUPS", "h2d_bytes_per_step": obj["h2d_bytes_per_step"],
                    "d2h_bytes_per_step": obj["d2h_bytes_per_step"]},
            "roofline": obj["roofline"], "gpu_launches": obj["gpu_launches"], "clocks": obj["clocks"], "sw_long": obj}
    emit(line)
    agx.capi.shutdown()
    return 0


def bind_to_gpu_numa_node(index: int):
    """Pin this rank to the CPUs next to its GPU (NVML's ideal affinity) before any host buffer is
    touched, so pinned staging memory is allocated on the GPU's NUMA node."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        allowed = os.sched_getaffinity(0)
        if cpus & allowed:
            os.sched_setaffinity(0, cpus & allowed)
    except Exception:
        pass


def run_gpu_arm(args):
    import torch
    import agxpkg
    agx = agxpkg.load()
    rank, local_rank, world = env_rank()
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (libagx has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    bind_to_gpu_numa_node(local_rank)
    park = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
        park = dist.new_group(backend="gloo")          # CPU-side barrier: parked ranks must not hold an SM
    cap = agx.capi
    cap.init_devices([local_rank])
    cap.set_profiling(True)

    sw = bench_sw(agx, args, rank, local_rank, world, device) if args.workload in ("both", "sw") else None
    hmm = bench_hmm(agx, args, rank, local_rank, world, device) if args.workload in ("both", "pairhmm") else None

    # ---- sections that one process runs over ALL GPUs of the box: the other ranks release their GPUs and wait ----
    extras = {}
    if world > 1:
        barrier(world)
        if rank != 0:
            cap.shutdown()
            torch.cuda.empty_cache()
    if rank == 0 and args.workload == "both":
        n_box = world if world > 1 else 1
        try:
            if not args.no_sw_long:
                extras["sw_long"] = bench_sw_long(agx, args, n_box)
            if not args.no_strong:
                extras["strong"] = bench_strong(agx, args, n_box)
        finally:
            cap.shutdown()
            cap.init_devices([local_rank])
            cap.set_profiling(True)
        gatk = bench_gatk(agx, args, local_rank) if not args.no_gatk else None
        if world == 1 and args.sw_len:
            extras["sw_lengths"] = bench_sw_lengths(agx, args, local_rank)
        align = bench_sw_align(agx, args, local_rank) if not args.no_align else None
        if world == 1 and not args.no_cpu_baseline:
            extras.update(cpu_leg(agx, args, gatk, align))
        if align is not None:
            extras["sw_align"] = {k: v for k, v in align.items() if not k.startswith("_")}
        if gatk is not None:
            extras["pairhmm_gatk"] = {k: v for k, v in gatk.items() if not k.startswith("_")}
    if world > 1:
        import torch.distributed as dist
        dist.barrier(group=park)

    if rank == 0:
        head = sw if sw is not None else hmm
        clocks = None
        for r in (sw, hmm):
            if r is not None:
                clocks = merge_clocks(clocks, r["clocks"])
        line = {
            "metric": "SW and PairHMM GCUPS", "value": head["value"], "unit": "GCUPS", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "s16x2 (int16 pairs in int32 lanes)" if sw is not None else "f32", "data": "synthetic",
            "config": {
                "workload": ("sw_short: %d pairs of %dx%d per GPU (BASELINE configs[2], headline value)" %
                             (args.sw_pairs, SW_LEN, SW_LEN) if sw is not None else "") +
                            (" + pairhmm_hc: %d batches x 200 reads x 5 haplotypes per GPU (BASELINE configs[3], "
                             "\"pairhmm\" object)" % args.hmm_batches if hmm is not None else "") +
                            (" + sw_long 1 Mbp x 1 Mbp over all GPUs in one process (BASELINE configs[4], \"sw_long\" object)"
                             if "sw_long" in extras else ""),
                "parallelism": f"pair-sharded x{world}, no collective",
                "l2": "inputs larger than L2 (302 MB SW batch, 176 MB PairHMM batch; 126 MB L2); no explicit flush",
                "cells_counted": "len_a*len_b per pair, newline row/column excluded",
            },
            "e2e": head["e2e"], "roofline": head["roofline"], "gpu_launches": head["gpu_launches"],
            "clocks": clocks,
        }
        if "e2e_flat" in head:
            line["e2e_flat"] = head["e2e_flat"]
        if "cpu_sw" in extras:
            line["cpu_baseline"] = extras["cpu_sw"]
        elif "cpu_hmm" in extras and sw is None:
            line["cpu_baseline"] = extras["cpu_hmm"]
        if hmm is not None and sw is not None:
            sub = {k: hmm[k] for k in ("value", "unit", "dtype", "ms_per_step", "e2e", "e2e_flat", "roofline", "fp64_rescue",
                                       "gpu_launches", "clocks", "pairs_per_gpu", "cells_per_gpu")}
            if "cpu_hmm" in extras:
                sub["cpu_baseline"] = extras["cpu_hmm"]
            line["pairhmm"] = sub
        if "cpu_sw_long" in extras and "sw_long" in extras:
            extras["sw_long"]["cpu_baseline"] = extras["cpu_sw_long"]
        for key in ("sw_long", "strong", "pairhmm_gatk", "sw_lengths", "sw_align", "parity"):
            if key in extras:
                line[key] = extras[key]
        emit(line)
    cap.shutdown()
    if world > 1:
        import torch.distributed as dist
        dist.barrier(group=park)
        dist.destroy_process_group()
    return 0


class OneLineStdout:
    """The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, for one) write to
    fd 1 too, so fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved fd."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line: str):
        sys.stdout.flush()
        os.write(self.saved, (line + "\n").encode())

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


OUT = None


def emit(obj):
    line = json.dumps(obj)
    if OUT is not None:
        OUT.emit(line)
    else:
        print(line, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["agx", "reference"], default="agx")
    ap.add_argument("--workload", choices=["both", "sw", "pairhmm", "sw_long"], default="both")
    ap.add_argument("--long-len", type=int, default=1_000_000, help="sw_long: side of the single pair")
    ap.add_argument("--sw-pairs", type=int, default=1_000_000, help="SW pairs per GPU per step")
    ap.add_argument("--hmm-batches", type=int, default=1000, help="PairHMM batches (200 reads x 5 haps) per GPU per step")
    ap.add_argument("--hmm-unrelated", type=float, default=0.001, help="fraction of unrelated reads (the FP64-rescue tail)")
    ap.add_argument("--sw-len", default="64,128,256,512,1024,450-500",
                    help="length classes of the \"sw_lengths\" sweep (L or LO-HI, comma separated; empty = skip)")
    ap.add_argument("--sw-len-cells", type=float, default=5e9, help="cells per length class of that sweep")
    ap.add_argument("--cpu-seconds", type=float, default=25.0,
                    help="CPU-baseline sample: seconds of reference work per core (capped at one full 10^6-pair batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sw-long", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-gatk", action="store_true")
    ap.add_argument("--no-align", action="store_true")
    ap.add_argument("--long-cpu-len", type=int, default=20000,
                    help="side of the square the reference C is timed on for the sw_long CPU baseline (0 = skip)")
    ap.add_argument("--align-pairs", type=int, default=1_000_000, help="pairs of the \"sw_align\" object (at most --sw-pairs)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "agx":
        args.warmup = max(args.warmup, 3)
    global OUT
    with OneLineStdout() as OUT:
        if args.impl == "reference":
            return run_reference_arm(args)
        if args.workload == "sw_long":
            return run_sw_long(args)
        return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
