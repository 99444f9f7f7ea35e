"""CPU, world_size 2 over gloo: the N>1 host logic (sharding + label gather).  The data path
has no collective; each rank produces its own utterances' decisions (here with the oracle as
the stand-in producer) and the gathered stream must equal the single-process result byte for
byte."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_math as rm
from vad_b200 import shard
from vad_b200.synth import synth_utterance

LENS = [16000, 4001, 401, 9000, 24000, 7, 1041, 12345]


def _labels(utt_id):
    w = rm.glorot_ffn(0)
    return torch.from_numpy(rm.vad_utterance(synth_utterance(5, utt_id, LENS[utt_id]), w)[3].copy())


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.shard_balanced(LENS, world)[rank]
    got = shard.gather_labels([_labels(int(u)) for u in mine], mine, dst=0)
    if rank == 0:
        q.put({k: v.numpy().tobytes() for k, v in got.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_shard_helpers():
    assert [shard.shard_contiguous(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    parts = shard.shard_balanced(LENS, 2)
    assert sorted(np.concatenate(parts).tolist()) == list(range(len(LENS)))
    cost = np.maximum((np.array(LENS) - 401) // 160 + 1, 0)
    loads = [cost[p].sum() for p in parts]
    assert abs(loads[0] - loads[1]) <= cost.max()
    assert shard.stream_owner([0, 1, 9], 8).tolist() == [0, 1, 1]


def test_two_rank_gather_equals_single_process():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sorted(got) == list(range(len(LENS)))
    for u in range(len(LENS)):
        assert got[u] == _labels(u).numpy().tobytes()
