#!/usr/bin/env python
"""configs[3]: realtime streaming -- 4096 concurrent 16 kHz streams fed in 10 ms (160-sample)
chunks through StreamBank (the many-stream SKLearnAnalyzer.feed_frame), p50/p99/max chunk latency.

Latency = host wall clock from "the tick's chunk batch is in pinned host memory" to "the labels
are readable on the host" (H2D copy + stream_feed_kernel + D2H copy + stream sync), per tick.
Prints one JSON line.  `python tools/bench_stream.py [--streams 4096] [--ticks 10000]`."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vad_b200 import runtime  # noqa: E402
from vad_b200.analyser import StreamBank  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=4096)
    ap.add_argument("--ticks", type=int, default=10000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--graph", type=int, default=1, help="replay the tick as a CUDA graph")
    a = ap.parse_args()
    h = runtime.Handle(0, ffn_weights=runtime.glorot_ffn(0))
    bank = StreamBank(a.streams, handle=h)
    rng = np.random.default_rng(0)
    pool = torch.from_numpy((rng.standard_normal((64, a.streams, 160)) * 3000).astype(np.int16))
    lat = np.zeros(a.ticks)
    graph = None
    if a.graph:
        bank.h_chunks.copy_(pool[0])
        for _ in range(3):
            bank.feed_pinned()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=bank.stream):
            bank._enqueue(False)
        graph = g
        bank.reset()
    speech = 0
    for t in range(a.warmup + a.ticks):
        bank.h_chunks.copy_(pool[t % 64])           # the tick's audio arrives in pinned memory
        t0 = time.perf_counter()
        if graph is not None:
            graph.replay()
            bank.stream.synchronize()
            labels = bank.h_labels
        else:
            labels = bank.feed_pinned()
        dt = time.perf_counter() - t0
        if t >= a.warmup:
            lat[t - a.warmup] = dt
            speech += int((labels == 1).sum())
    lat_ms = np.sort(lat) * 1e3
    out = {
        "metric": "stream_chunk_latency_ms", "streams": a.streams, "ticks": a.ticks, "chunk_ms": 10.0,
        "cuda_graph": bool(a.graph),
        "p50_ms": float(lat_ms[len(lat_ms) // 2]), "p99_ms": float(lat_ms[int(len(lat_ms) * 0.99)]),
        "max_ms": float(lat_ms[-1]), "mean_ms": float(lat_ms.mean()),
        "keeps_up_with_realtime": bool(lat_ms[int(len(lat_ms) * 0.99)] < 10.0),
        "realtime_factor_at_p99": float(10.0 / lat_ms[int(len(lat_ms) * 0.99)]),
        "audio_s_per_s_at_mean": float(a.streams * 0.010 / (lat_ms.mean() * 1e-3)),
        "speech_decisions": speech, "gpu": torch.cuda.get_device_name(0),
        "algorithmic_lag_frames": 3, "h2d_bytes_per_tick": a.streams * 320, "d2h_bytes_per_tick": a.streams,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
