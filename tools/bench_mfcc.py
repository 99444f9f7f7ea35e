#!/usr/bin/env python
"""configs[1]: MFCC-only batch of 1024 x 10 s synthetic utterances on one B200 -- feature-kernel
throughput and its FFT/mel/DCT roofline (13,878 algorithmic flop per frame).
Input 327.68 MB (> L2), output f32[1024, 998, 13] = 53.1 MB.  Prints one JSON line."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vad_b200 import batch, runtime  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    a = ap.parse_args()
    h = runtime.Handle(0)
    L = 160000
    off, ln, stride = batch.uniform_layout(a.utts, L)
    pcm = h.synth_pcm(a.utts, L, seed=0, first_utt=0, utt_stride=stride)
    plan = runtime.Plan(h, off, ln, runtime.MODE_MFCC)
    out = torch.empty((plan.total_rows, 13), dtype=torch.float32, device=h.device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=h.device)
    for _ in range(a.warmup):
        plan.mfcc(pcm, out=out)
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(a.steps):
        flush.fill_(1)                                   # evict the 328 MB input's tail from the 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.mfcc(pcm, out=out)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / a.steps
    frames = plan.total_rows
    peak = max(h.fp32_peak(0, 2048), h.fp32_peak(1, 2048))
    ach = frames * 13878 / (ms * 1e-3) / 1e12
    print(json.dumps({
        "metric": "audio_seconds_per_second", "workload": "cfg2: MFCC-only, %d x 10 s utterances" % a.utts,
        "value": a.utts * 10.0 / (ms * 1e-3), "ms_per_step": ms, "frames": int(frames),
        "roofline": {"bound": "fp32", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                     "algorithmic_flop_per_frame": 13878,
                     "hbm_GBps": frames * (320 + 52) / (ms * 1e-3) / 1e9},
        "l2": "flush buffer of 256 MB written between iterations", "gpu": torch.cuda.get_device_name(0)}))


if __name__ == "__main__":
    main()
