"""Multi-GPU sharding of the hot path: utterances (offline) and streams (realtime) are
independent, so ranks share nothing on the data path.  One process per GPU;
``torch.distributed`` is used only to gather variable-length per-frame decisions when a caller
wants them on one rank (off the timed path), and for the bench's timing reduction."""
import numpy as np
import torch


def shard_contiguous(n_items, rank, world):
    """Contiguous index range [lo, hi) of rank (equal-length utterances, cfg3)."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_balanced(lengths, world):
    """Greedy longest-first assignment of ragged utterances to ranks, balanced by output frames
    (cfg5).  Returns a list of index arrays, one per rank, each sorted ascending so every rank's
    packed buffer keeps utterance order."""
    lengths = np.asarray(lengths, dtype=np.int64)
    cost = np.maximum((lengths - 401) // 160 + 1, 0)
    order = np.argsort(-cost, kind="stable")
    load = np.zeros(world, dtype=np.int64)
    owner = np.empty(len(lengths), dtype=np.int64)
    for i in order:
        r = int(np.argmin(load))
        owner[i] = r
        load[r] += cost[i]
    return [np.flatnonzero(owner == r) for r in range(world)]


def stream_owner(stream_ids, world):
    """Realtime: stream s lives on rank s % world."""
    return np.asarray(stream_ids, dtype=np.int64) % world


def gather_labels(local_labels, local_ids, dst=0, group=None):
    """Collect per-utterance uint8 label tensors on ``dst``.

    ``local_labels``: list of 1-D uint8 tensors (this rank's utterances), ``local_ids``: their
    global utterance ids.  Counts are exchanged first (variable lengths), then one padded
    all_gather moves the bytes (NCCL over NVLink on GPUs, gloo on CPU).  Returns
    {utt_id: tensor} on ``dst`` and None elsewhere."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = local_labels[0].device if local_labels else torch.device("cpu")
    ids = torch.as_tensor(np.asarray(local_ids, dtype=np.int64), device=dev)
    lens = torch.as_tensor(np.asarray([int(t.numel()) for t in local_labels], dtype=np.int64), device=dev)
    meta = torch.tensor([ids.numel(), int(lens.sum().item()) if lens.numel() else 0], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    max_n = max(int(m[0].item()) for m in metas)
    max_b = max(int(m[1].item()) for m in metas)

    def pad(t, n, dtype):
        out = torch.zeros(n, dtype=dtype, device=dev)
        out[: t.numel()] = t
        return out

    flat = torch.cat(local_labels) if local_labels else torch.zeros(0, dtype=torch.uint8, device=dev)
    g_ids = [torch.zeros(max_n, dtype=torch.int64, device=dev) for _ in range(world)]
    g_lens = [torch.zeros(max_n, dtype=torch.int64, device=dev) for _ in range(world)]
    g_flat = [torch.zeros(max_b, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(g_ids, pad(ids, max_n, torch.int64), group=group)
    dist.all_gather(g_lens, pad(lens, max_n, torch.int64), group=group)
    dist.all_gather(g_flat, pad(flat, max_b, torch.uint8), group=group)
    if rank != dst:
        return None
    out = {}
    for r in range(world):
        n = int(metas[r][0].item())
        pos = 0
        for i in range(n):
            ln = int(g_lens[r][i].item())
            out[int(g_ids[r][i].item())] = g_flat[r][pos:pos + ln].clone()
            pos += ln
    return out
