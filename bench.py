#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric ("SW and PairHMM GCUPS at 1/2/4/8 B200 vs reference C on host
cores") measured on BASELINE.json's two batch configurations:

    configs[2]  SW short-read batch: 10^6 pairs 150 bp x 150 bp (synthetic)          <- headline
    configs[3]  PairHMM HaplotypeCaller-shaped batch: 10^6 (read, haplotype) pairs,
                reads 100-250 bp x haplotypes 200-500 bp (synthetic)                  <- "pairhmm" object

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the reference C programs on the host cores)

A "step" is one pass of the hot path over one batch.  Per-GPU work is fixed as N grows (every rank
scores its own 10^6-pair shard: "scaling": "weak"); pairs are independent, so there is no data-path
collective -- torch.distributed is used only for the barrier and the max-over-ranks of the timings.

  value        GCUPS with the batch already resident in HBM (raw file bytes + offsets): classify kernel,
               grid-sizing read-back, DP kernels; CUDA events on the launching stream, max over ranks.
  e2e          the same metric through the host C-ABI call (sw_score_batch_flat /
               pairhmm_forward_batches_flat) on PINNED HOST buffers: H2D of the batch, kernels, D2H of
               the results all inside the timed region.
  roofline     dominant kernel (SW: sw_duo_kernel<8,19>; PairHMM: hmm_stream_kernel<K>) timed alone with
               CUDA events inside libagx (agx_profile_ms): algorithmic ALU lane-ops / s over the measured
               ALU peak of profiles/peaks_*.json (nominal fallback stated).
  cpu_baseline the reference C programs (oracle/_ref, kind "reference") or the oracle port, on all host
               cores, on a bounded sample of the same workload.
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
SW_OPS_PER_CELL = 4.0       # s16x2: 8 INT32-pipe lane-ops per cell pair
HMM_OPS_PER_CELL = 8.0      # 8 FP32-pipe lane-instructions = 11 FLOP per cell
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
def _run_parallel(cmds, timeout=900):
    t0 = time.perf_counter()
    procs = [subprocess.Popen(c, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for c in cmds]
    for p in procs:
        p.wait(timeout=timeout)
    dt = time.perf_counter() - t0
    bad = [p.returncode for p in procs if p.returncode != 0]
    if bad:
        raise RuntimeError(f"reference process failed: exit codes {bad[:4]}")
    return dt


def _port_worker(args):
    kind, path = args
    import oracle
    if kind == "sw":
        s, _ = oracle.sw_file(path)
        return len(s)
    v, _ = oracle.pairhmm_file(path)
    return len(v)


class CpuReference:
    """Times the reference C programs (oracle/_ref) -- or, where they are absent, the oracle port --
    on every host core, one process per core, each on its own shard of the synthetic workload."""

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
        self._sw_cells = float(self.cores) * pairs_per_core * SW_LEN * SW_LEN
        self._sw_sample = (f"{self.cores} processes x {pairs_per_core} pairs of {SW_LEN}x{SW_LEN} "
                           f"({'oracle/_ref/sw_antidiag = antidiagonalSmithWaterman.c -O3' if self.kind == 'reference' else 'oracle port'})")

    def prepare_hmm(self, batches_per_core: int, seed: int = 99):
        files, cells = [], 0
        for c in range(self.cores):
            inp = self.agx.synth.pairhmm_batches(batches_per_core, 200, 5, seed=seed + c)
            p = Path(self.tmp.name) / f"hmm_{c}.in"
            p.write_bytes(inp.buf.tobytes())
            files.append(str(p))
            cells += inp.cells()
        self._hmm_files = files
        self._hmm_cells = float(cells)
        self._hmm_sample = (f"{self.cores} processes x {batches_per_core} batches of 200 reads x 5 haplotypes "
                            f"({'oracle/_ref/pairhmm_matrix = pairHMMmatrix.c -O3, output byte-identical to antidiagsPairHMM.c' if self.kind == 'reference' else 'oracle port'})")

    def _time(self, which: str) -> float:
        files = self._sw_files if which == "sw" else self._hmm_files
        if self.kind == "reference":
            if which == "sw":
                cmds = [[str(self.sw_exe), f] for f in files]
            else:
                cmds = [[str(self.hmm_exe), f, f + ".out"] for f in files]
            return _run_parallel(cmds)
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
    for _ in range(min(args.warmup, 2)):
        img_scores, header, _ = cap.sw_score_file_image(np_buf, out=np_scores)
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        img_scores, header, _ = cap.sw_score_file_image(np_buf, out=np_scores)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / args.steps, world, device)
    assert header == 2 * n and np.array_equal(img_scores, res_scores), "file-image and device entry points disagree"
    for _ in range(min(args.warmup, 2)):
        e2e_scores = cap.sw_score_flat(np_buf, np_off, np_len)
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_scores = cap.sw_score_flat(np_buf, np_off, np_len)
    torch.cuda.synchronize()
    flat_ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / args.steps, world, device)
    assert np.array_equal(e2e_scores, res_scores), "host and device entry points disagree"

    k_ms = float(np.mean(kern_ms))
    peaks = measured_peaks()
    if peaks and peaks.get("alu_tlaneops"):
        peak, peak_src = peaks["alu_tlaneops"] * 1e12, f"measured ({peaks.get('source', 'profiles/peaks')})"
    else:
        peak, peak_src = NOMINAL_ALU, "nominal 148 SM x 64 lanes/clk x 1.965 GHz (no measured ALU peak committed yet)"
    achieved = cells * SW_OPS_PER_CELL / (k_ms * 1e-3)
    return {
        "value": world * cells / (ms * 1e-3) / 1e9, "ms_per_step": ms,
        "e2e": {"value": world * cells / (e2e_ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(np_buf.nbytes), "d2h_bytes_per_step": int(4 * n),
                "entry_point": "sw_score_file_image (pinned file image in, pinned scores out)"},
        "e2e_flat": {"value": world * cells / (flat_ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": flat_ms,
                     "h2d_bytes_per_step": int(np_buf.nbytes + np_off.nbytes + np_len.nbytes),
                     "d2h_bytes_per_step": int(4 * n), "entry_point": "sw_score_batch_flat"},
        "roofline": {"bound": "alu", "kernel": "sw_duo_kernel<8,19> (s16x2 DPX)", "achieved": achieved / 1e12,
                     "peak": peak / 1e12, "unit": "Tlaneop/s (INT32/DPX pipe)", "frac": achieved / peak,
                     "peak_source": peak_src, "ops_per_cell": SW_OPS_PER_CELL, "kernel_ms": k_ms,
                     "kernel_gcups": cells / (k_ms * 1e-3) / 1e9, "kernel_share_of_step": k_ms / ms,
                     "traffic": (peaks["sw_duo_dram_bytes_per_pair_150x150"] * n
                                 if peaks and peaks.get("sw_duo_dram_bytes_per_pair_150x150") else None),
                     "traffic_note": "DRAM bytes per launch from the committed ncu capture (profiles/r1f_sw_duo_ncu.txt), "
                                     "scaled by pairs; the kernel is ALU-bound, HBM time for this is ~0.05 ms",
                     "loader": {"kernel": "sw_classify_kernel", "ms": float(np.mean(cls_ms)), "bound": "hbm"}},
        "gpu_launches": int(launches), "clocks": clk.summary(),
        "pairs_per_gpu": n, "cells_per_gpu": cells,
    }


def bench_hmm(agx, args, rank, local_rank, world, device):
    import torch
    cap = agx.capi
    inp = agx.synth.pairhmm_batches(args.hmm_batches, 200, 5, seed=2000 + rank)
    cells = float(inp.cells())
    n_pairs = inp.n_pairs
    nb = inp.n_batches
    nh_b = np.diff(inp.batch_hap_start)
    read_batch = np.repeat(np.arange(nb, dtype=np.int32), np.diff(inp.batch_read_start))
    out_off = np.concatenate(([0], np.cumsum(nh_b[read_batch])))[:-1].astype(np.int64)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_buf = pin(inp.buf)
    h_rfo, h_rl, h_ho, h_hl = pin(inp.read_field_off.reshape(-1)), pin(inp.read_len), pin(inp.hap_off), pin(inp.hap_len)
    h_brs, h_bhs = pin(inp.batch_read_start), pin(inp.batch_hap_start)
    d_buf, d_rfo, d_rl, d_ho, d_hl, d_bhs = (x.to(device) for x in (h_buf, h_rfo, h_rl, h_ho, h_hl, h_bhs))
    d_rb, d_roo = torch.from_numpy(read_batch).to(device), torch.from_numpy(out_off).to(device)
    d_out = torch.empty(n_pairs, dtype=torch.float64, device=device)
    stream = torch.cuda.current_stream().cuda_stream

    def step_resident():
        cap.pairhmm_forward_device(local_rank, d_buf.data_ptr(), d_buf.numel(), d_rfo.data_ptr(), d_rl.data_ptr(),
                                   d_rb.data_ptr(), d_roo.data_ptr(), inp.read_len.size, d_ho.data_ptr(),
                                   d_hl.data_ptr(), inp.hap_len.size, d_bhs.data_ptr(), nb, n_pairs,
                                   d_out.data_ptr(), stream, True)

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    barrier(world)
    cap.reset_launch_count()
    kern_ms = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sampler as clk:
        ev0.record()
        for _ in range(args.steps):
            step_resident()
            kern_ms.append(cap.profile_ms(local_rank, cap.PROF_HMM_STREAM))
        ev1.record()
        torch.cuda.synchronize()
    launches = cap.launch_count()
    ms = max_over_ranks(ev0.elapsed_time(ev1) / args.steps, world, device)
    res = d_out.cpu().numpy()

    # end to end.  Headline: pairhmm_forward_file_image(), the call the drop-in driver makes -- the raw
    # pairHMM/test_set-format file image in pinned host memory in, log10 likelihoods in pinned host memory
    # out; batch walk and field splitting on the GPU.  Second: pairhmm_forward_batches_flat() with
    # caller-built index arrays.
    arrs = [x.numpy() for x in (h_buf, h_rfo, h_rl, h_ho, h_hl, h_brs, h_bhs)]
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
    for _ in range(min(args.warmup, 2)):
        e2e = cap.pairhmm_forward_flat(*arrs)
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e = cap.pairhmm_forward_flat(*arrs)
    torch.cuda.synchronize()
    flat_ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / args.steps, world, device)
    assert np.array_equal(e2e, res, equal_nan=True), "host and device entry points disagree"
    assert np.all(np.isfinite(res)), "non-finite PairHMM result"

    k_ms = float(np.mean(kern_ms))
    peaks = measured_peaks()
    if peaks and peaks.get("fp32_tlaneops"):
        peak, peak_src = peaks["fp32_tlaneops"] * 1e12, f"measured ({peaks.get('source', 'profiles/peaks')})"
    else:
        peak, peak_src = NOMINAL_FP32, "nominal 148 SM x 128 lanes/clk x 1.965 GHz (no measured FP32 peak committed yet)"
    achieved = cells * HMM_OPS_PER_CELL / (k_ms * 1e-3)
    return {
        "value": world * cells / (ms * 1e-3) / 1e9, "unit": "GCUPS", "dtype": "f32 (+f64 rescue)", "ms_per_step": ms,
        "e2e": {"value": world * cells / (e2e_ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(arrs[0].nbytes), "d2h_bytes_per_step": int(8 * n_pairs + 4 * nb),
                "entry_point": "pairhmm_forward_file_image (pinned file image in, pinned results out)"},
        "e2e_flat": {"value": world * cells / (flat_ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": flat_ms,
                     "h2d_bytes_per_step": int(sum(a.nbytes for a in arrs)), "d2h_bytes_per_step": int(8 * n_pairs),
                     "entry_point": "pairhmm_forward_batches_flat"},
        "roofline": {"bound": "alu", "kernel": "hmm_duo_kernel<K=4..8> (FP32, two reads per warp, f32x2) + hmm_stream_kernel for unpaired reads", "achieved": achieved / 1e12,
                     "peak": peak / 1e12, "unit": "Tlaneinstr/s (FP32 pipe)", "frac": achieved / peak,
                     "peak_source": peak_src, "ops_per_cell": HMM_OPS_PER_CELL, "kernel_ms": k_ms,
                     "kernel_gcups": cells / (k_ms * 1e-3) / 1e9, "kernel_share_of_step": k_ms / ms,
                     "traffic": (peaks["hmm_dram_bytes_per_step_config4"] * (args.hmm_batches / 1000.0)
                                 if peaks and peaks.get("hmm_dram_bytes_per_step_config4") else None),
                     "traffic_note": "DRAM bytes of the stream-kernel launches of one step from the committed ncu launch "
                                     "list (profiles/r1f_launch_shares.txt), scaled by batches; FP32-bound, HBM time ~0.04 ms"},
        "gpu_launches": int(launches), "clocks": clk.summary(),
        "pairs_per_gpu": n_pairs, "cells_per_gpu": cells,
    }


def run_sw_long(args):
    """BASELINE configs[4]: ONE pair of --long-len x --long-len, columns striped over --gpus B200s of this
    box inside one process (NVLink peer boundary exchange), through the host C ABI."""
    import torch
    import agxpkg
    import oracle
    agx = agxpkg.load()
    cap = agx.capi
    n_gpus = min(args.gpus, torch.cuda.device_count())
    cap.init(n_gpus)
    cap.set_profiling(True)
    L = args.long_len
    inp = agx.formats.parse_sw(agx.synth.sw_long_pair(L, seed=5, related=True), line_buf=1 << 30)
    cells = float(L) * float(L)
    for _ in range(max(1, min(args.warmup, 2))):
        score = int(cap.sw_score_flat(inp.buf, inp.off, inp.len)[0])
    cap.reset_launch_count()
    with ClockSampler(0) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            score = int(cap.sw_score_flat(inp.buf, inp.off, inp.len)[0])
        dt = (time.perf_counter() - t0) / args.steps
    launches = cap.launch_count()
    # CPU: the MAX_LINE_LENGTH-raised reference build on a smaller square, extrapolation labelled
    cpu = None
    ref_exe = ROOT / "oracle" / "_ref" / "sw_antidiag_long"
    if not args.no_cpu_baseline and ref_exe.exists():
        n_small = 20000
        with tempfile.TemporaryDirectory() as td:
            pth = Path(td) / "small.in"
            pth.write_bytes(agx.synth.sw_long_pair(n_small, seed=6, related=True))
            t1 = time.perf_counter()
            oracle.run_ref_sw(str(pth), long_lines=True)
            cdt = time.perf_counter() - t1
        g = n_small * n_small / cdt / 1e9
        cpu = {"value": g, "unit": "GCUPS", "cores": 1, "kind": "reference",
               "sample": f"oracle/_ref/sw_antidiag_long (MAX_LINE_LENGTH raised, arithmetic untouched) on one {n_small}x{n_small} pair; "
                         f"the reference is single-threaded per pair; {L}x{L} would take ~{cells / (g * 1e9) / 3600:.1f} core-hours (extrapolated)"}
    peaks = measured_peaks()
    peak = peaks["alu_tlaneops"] * 1e12 if peaks and peaks.get("alu_tlaneops") else NOMINAL_ALU
    line = {"metric": "SW and PairHMM GCUPS", "value": cells / dt / 1e9, "unit": "GCUPS", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": f"sw_long: one pair {L}x{L} (BASELINE configs[4]), column stripes over {n_gpus} GPU(s) in one process",
                       "score": score},
            "e2e": {"value": cells / dt / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": int(2 * (L + 1)) * n_gpus,
                    "d2h_bytes_per_step": 4 * n_gpus},
            "roofline": {"bound": "alu", "kernel": "sw_long_kernel<K> (s32 DPX)", "achieved": cells * 8 / dt / 1e12 / n_gpus,
                         "peak": peak / 1e12, "unit": "Tlaneop/s per GPU (INT32/DPX pipe)", "frac": cells * 8 / dt / n_gpus / peak,
                         "ops_per_cell": 8.0, "traffic": None},
            "gpu_launches": int(launches), "clocks": clk.summary()}
    if cpu:
        line["cpu_baseline"] = cpu
    emit(line)
    cap.shutdown()
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
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    cap = agx.capi
    cap.init_devices([local_rank])
    cap.set_profiling(True)

    sw = bench_sw(agx, args, rank, local_rank, world, device) if args.workload in ("both", "sw") else None
    hmm = bench_hmm(agx, args, rank, local_rank, world, device) if args.workload in ("both", "pairhmm") else None

    cpu_sw = cpu_hmm = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref = CpuReference()
        if sw is not None:
            ref.prepare_sw(max(200, int(args.cpu_seconds * 37e6 / (SW_LEN * SW_LEN))))
            v, _ = ref.sw_gcups()
            cpu_sw = ref.baseline_obj("sw", v)
        if hmm is not None:
            ref.prepare_hmm(max(1, int(round(args.cpu_seconds * 78e6 / 6.1e7))))
            v, _ = ref.hmm_gcups()
            cpu_hmm = ref.baseline_obj("hmm", v)

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
                             "\"pairhmm\" object)" % args.hmm_batches if hmm is not None else ""),
                "parallelism": f"pair-sharded x{world}, no collective",
                "l2": "inputs larger than L2 (302 MB SW batch, 176 MB PairHMM batch; 126 MB L2); no explicit flush",
                "cells_counted": "len_a*len_b per pair, newline row/column excluded",
            },
            "e2e": head["e2e"], "roofline": head["roofline"], "gpu_launches": head["gpu_launches"],
            "clocks": clocks,
        }
        if "e2e_flat" in head:
            line["e2e_flat"] = head["e2e_flat"]
        if cpu_sw is not None:
            line["cpu_baseline"] = cpu_sw
        elif cpu_hmm is not None and sw is None:
            line["cpu_baseline"] = cpu_hmm
        if hmm is not None and sw is not None:
            sub = {k: hmm[k] for k in ("value", "unit", "dtype", "ms_per_step", "e2e", "e2e_flat", "roofline", "gpu_launches",
                                       "clocks", "pairs_per_gpu", "cells_per_gpu")}
            if cpu_hmm is not None:
                sub["cpu_baseline"] = cpu_hmm
            line["pairhmm"] = sub
        emit(line)
    cap.shutdown()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
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
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="CPU-baseline sample size, seconds of work per core")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
