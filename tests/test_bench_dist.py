"""The N > 1 plumbing of bench.py on CPU: world_size-2 gloo processes exercise the rank -> shard mapping,
the barrier and the max-over-ranks reduction that the NCCL run uses (no GPU, no kernels)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    import agxpkg
    import bench
    import oracle
    agx = agxpkg.load()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert bench.env_rank() == (rank, rank, world)
    # every rank scores its OWN shard (weak scaling): same seeds as bench_sw / bench_hmm
    sw = agx.synth.sw_uniform_pairs(64, 150, seed=1000 + rank)
    scores = oracle.sw_scores_flat(sw.buf, sw.off, sw.len)
    digest = torch.tensor([int(np.frombuffer(sw.buf.tobytes()[-4000:], dtype=np.uint8).sum()), int(scores.sum())],
                          dtype=torch.int64)
    gathered = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(gathered, digest)
    assert len({tuple(g.tolist()) for g in gathered}) == world, "ranks must not score the same shard"
    # timings: max over ranks, aggregate = units of ALL ranks / that time
    my_ms = 10.0 + 5.0 * rank
    dist.barrier()
    worst = bench.max_over_ranks(my_ms, world, torch.device("cpu"))
    assert worst == 10.0 + 5.0 * (world - 1)
    cells = 64 * 150 * 150
    value = world * cells / (worst * 1e-3) / 1e9
    if rank == 0:
        Path(out_dir, "value.txt").write_text(repr(value))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_plumbing(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    value = float((tmp_path / "value.txt").read_text())
    assert value == pytest.approx(2 * 64 * 150 * 150 / 15e-3 / 1e9)


def test_host_dispatcher_cuts_are_balanced():
    """The library's in-process sharding rule (api.cu: balanced_cuts) restated: contiguous ranges of
    equal cell weight."""
    rng = np.random.default_rng(0)
    w = rng.integers(1, 1000, size=5000).astype(np.float64)
    prefix = np.concatenate(([0.0], np.cumsum(w)))
    parts = 8
    cuts = [0] + [int(np.searchsorted(prefix, prefix[-1] * k / parts, side="left")) for k in range(1, parts)] + [w.size]
    loads = [prefix[cuts[k + 1]] - prefix[cuts[k]] for k in range(parts)]
    assert cuts == sorted(cuts) and cuts[-1] == w.size
    assert max(loads) / (prefix[-1] / parts) < 1.01
