"""CPU: host-side multi-GPU logic (utterance sharding, bucketing, weight broadcast) with the gloo backend,
world_size 2 -- the N>1 path has no data-path collective, only the start-up broadcast."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))

from viettts_b200 import parallel, synthetic, weights  # noqa: E402


def test_lpt_shard_balances_and_covers():
    rng = np.random.default_rng(0)
    nf = rng.integers(156, 938, size=256)
    shards = parallel.lpt_shard(nf, 8)
    flat = sorted(i for s in shards for i in s)
    assert flat == list(range(256))
    loads = [int(nf[s].sum()) for s in shards]
    assert max(loads) - min(loads) <= nf.max()
    assert parallel.lpt_shard([5], 4) == [[0], [], [], []]


def test_bucketing_respects_padding_budget():
    rng = np.random.default_rng(1)
    nf = rng.integers(156, 938, size=256)
    buckets = parallel.bucket_by_length(nf, max_pad_frac=0.08, max_rows=32)
    assert sorted(i for b in buckets for i in b) == list(range(256))
    for b in buckets:
        assert len(b) <= 32
        longest = nf[b].max()
        assert (longest * len(b) - nf[b].sum()) <= 0.08 * longest * len(b) + 1e-9


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hp = synthetic.hifigan_params(1234) if rank == 0 else None
        blob = weights.pack_hifigan(hp) if rank == 0 else 13_926_017
        t = parallel.broadcast_blob(blob, torch.device("cpu"), src=0)
        # the duration model's blob: rank 1 only knows its size, from the library
        from viettts_b200 import _lib
        dblob = weights.pack_duration(synthetic.duration_ckpt(1234)) if rank == 0 else int(_lib.load().vtts_duration_blob_floats())
        td = parallel.broadcast_blob(dblob, torch.device("cpu"), src=0)
        # every rank shards the same global work list identically and takes its own slice
        nf = np.random.default_rng(5).integers(156, 938, size=64)
        mine = parallel.lpt_shard(nf, world)[rank]
        q.put((rank, float(t.double().sum()) + float(td.double().sum()) * 1e-3, int(t.numel()), sorted(mine)))
    finally:
        dist.destroy_process_group()


def test_weight_broadcast_and_sharding_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, n0, m0), (r1, s1, n1, m1) = res
    assert n0 == n1 == 13_926_017
    assert s0 == s1                      # rank 1 received rank 0's blob bit for bit
    assert not set(m0) & set(m1) and sorted(m0 + m1) == list(range(64))
