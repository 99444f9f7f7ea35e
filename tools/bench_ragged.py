#!/usr/bin/env python
"""configs[4]: ragged TED-LIUM-shaped synthetic utterances (2-30 s, uniform, seed 7) with deltas +
delta-deltas (MODE_DATASET: [sum(T_u - 5), 39] rows) and packed batching; also the fused VAD over the
same ragged batch.  Checks three sampled utterances against the oracle, prints one JSON line."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vad_b200 import runtime  # noqa: E402
from vad_b200.synth import synth_utterance  # noqa: E402


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--check", type=int, default=1)
    a = ap.parse_args()
    h = runtime.Handle(0, ffn_weights=runtime.glorot_ffn(0))
    h.set_ffn_impl("tc")
    rng = np.random.default_rng(7)
    lengths = rng.integers(2 * 16000, 30 * 16000 + 1, size=a.utts).astype(np.int64)
    padded = (lengths + 7) // 8 * 8
    offsets = np.zeros(a.utts, dtype=np.int64)
    offsets[1:] = np.cumsum(padded[:-1])
    total = int(padded.sum()) + 8
    # device-side synthesis, one launch per utterance length class is overkill: generate at max length
    # into a scratch [utts, Lmax] is too big, so synthesise per utterance in chunks of equal stride.
    pcm = torch.zeros(total, dtype=torch.int16, device=h.device)
    order = np.argsort(lengths, kind="stable")
    for u in order:                                   # variable lengths: one synth launch each (setup only)
        h.synth_pcm(1, int(lengths[u]), seed=7, first_utt=int(u), utt_stride=int(lengths[u]),
                    out=pcm[int(offsets[u]):int(offsets[u]) + int(lengths[u])])
    torch.cuda.synchronize()
    dplan = runtime.Plan(h, offsets, lengths, runtime.MODE_DATASET)
    vplan = runtime.Plan(h, offsets, lengths, runtime.MODE_VAD)
    rows = torch.empty((dplan.total_rows, 39), dtype=torch.float32, device=h.device)
    labels = torch.empty((vplan.total_rows,), dtype=torch.uint8, device=h.device)
    ms_rows = timed(lambda: dplan.mfcc(pcm, out=rows), a.steps, a.warmup)
    ms_vad = timed(lambda: vplan.vad(pcm, labels=labels), a.steps, a.warmup)
    audio_s = float(lengths.sum()) / 16000.0
    ok = None
    if a.check:
        from oracle import ref_math as rm
        ok = True
        for u in (int(order[0]), int(order[len(order) // 2]), int(order[-1])):
            ref = synth_utterance(7, u, int(lengths[u]))
            got = pcm[int(offsets[u]):int(offsets[u]) + int(lengths[u])].cpu().numpy()
            ok &= bool(np.array_equal(ref, got))
            r0, r1 = int(dplan.row_offsets[u]), int(dplan.row_offsets[u + 1])
            want = rm.dataset_features(rm.mfcc_utterance(ref))
            ok &= bool(np.all(np.abs(rows[r0:r1].cpu().numpy() - want) <= 3e-4 + 1e-4 * np.abs(want)))
    print(json.dumps({
        "metric": "audio_seconds_per_second", "workload": "cfg5: %d ragged utterances, 2-30 s uniform (seed 7), "
        "packed with 8-sample-aligned offsets" % a.utts, "audio_hours": audio_s / 3600.0,
        "rows": int(dplan.total_rows), "dataset39_ms": ms_rows, "dataset39_audio_s_per_s": audio_s / (ms_rows * 1e-3),
        "dataset39_out_GBps": dplan.total_rows * 156 / (ms_rows * 1e-3) / 1e9,
        "vad_ms": ms_vad, "vad_audio_s_per_s": audio_s / (ms_vad * 1e-3), "oracle_check_passed": ok,
        "gpu": torch.cuda.get_device_name(0)}))


if __name__ == "__main__":
    main()
