"""Multi-GPU plumbing: utterances are independent, so the path shards by utterance
with NO collective on the hot path.  The only communication is
one broadcast of the packed weight blobs from rank 0 at start-up
(torch.distributed, NCCL over NVLink on GPUs / gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def lpt_shard(costs, world_size: int):
    """Longest-processing-time-first assignment of utterances to ranks.
    costs: per-utterance cost (n_frames).  Returns list (len world_size) of index lists,
    each sorted by decreasing cost so that length buckets stay contiguous."""
    costs = np.asarray(costs, dtype=np.int64)
    order = np.argsort(-costs, kind="stable")
    loads = np.zeros(world_size, dtype=np.int64)
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(loads))
        shards[r].append(int(i))
        loads[r] += int(costs[i])
    return shards


def bucket_by_length(n_frames, max_pad_frac: float = 0.08, max_rows: int = 128):
    """Group utterance indices (any order) into batches whose padding overhead
    (rows padded to the longest member) stays below max_pad_frac.  Returns list of index lists."""
    n_frames = np.asarray(n_frames, dtype=np.int64)
    order = list(np.argsort(-n_frames, kind="stable"))
    buckets, cur = [], []
    for i in order:
        if not cur:
            cur = [int(i)]
            continue
        longest = int(n_frames[cur[0]])
        total = int(n_frames[cur].sum() + n_frames[i])
        padded = longest * (len(cur) + 1)
        if len(cur) < max_rows and (padded - total) <= max_pad_frac * padded:
            cur.append(int(i))
        else:
            buckets.append(cur)
            cur = [int(i)]
    if cur:
        buckets.append(cur)
    return buckets


def broadcast_blob(blob_np, device, src: int = 0, group=None):
    """Broadcast a float32 blob from rank `src`; returns a torch tensor on `device`
    on every rank.  `blob_np` is only read on rank `src` (other ranks pass the size)."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if rank == src:
        t = torch.from_numpy(np.ascontiguousarray(blob_np, dtype=np.float32)).to(device)
    else:
        n = int(blob_np) if np.isscalar(blob_np) else int(np.asarray(blob_np).size)
        t = torch.empty(n, dtype=torch.float32, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(t, src=src, group=group)
    return t


def load_weights_distributed(engine, hifigan_params=None, acoustic_ckpt=None, device=None, src: int = 0, group=None,
                             duration_ckpt=None, with_duration: bool = False):
    """Rank `src` packs the Haiku-layout checkpoints; every rank receives the packed
    blobs by one broadcast each and loads them from DEVICE memory (no host round trip).
    `with_duration` (same value on every rank) adds the duration model's blob (55.7 + 50.1 + 7.4 MB)."""
    import torch
    import torch.distributed as dist

    from . import _lib, weights

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lib = _lib.load()
    device = device if device is not None else torch.device("cuda", engine.device)
    hb = weights.pack_hifigan(hifigan_params) if rank == src else int(lib.vtts_hifigan_blob_floats())
    ab = weights.pack_acoustic(acoustic_ckpt) if rank == src else int(lib.vtts_acoustic_blob_floats())
    ht = broadcast_blob(hb, device, src, group)
    at = broadcast_blob(ab, device, src, group)
    torch.cuda.synchronize(device)
    engine.load_hifigan(ht)
    engine.load_acoustic(at)
    total = ht.numel() * 4 + at.numel() * 4
    if with_duration or duration_ckpt is not None:
        db = weights.pack_duration(duration_ckpt) if rank == src else int(lib.vtts_duration_blob_floats())
        dt = broadcast_blob(db, device, src, group)
        torch.cuda.synchronize(device)
        engine.load_duration(dt)
        total += dt.numel() * 4
    return total
