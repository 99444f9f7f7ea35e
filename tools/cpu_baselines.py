#!/usr/bin/env python
"""CPU baselines of SURVEY.md 8(d) on the current host (reported, not optimised).  Uses the
reference-shaped port oracle/ref_loop.py (the reference itself cannot travel to the GPU box):
 (i)   single core, cfg1: one 60 s waveform through the per-frame MFCC loop + window features + FFN;
 (ii)  the reference's own parallel shape, multiprocessing.Pool(4) (config.py:30), and Pool(all cores),
       over 10 s utterances in 30-file steps (config.py:32);
 (iii) LoopAnalyser.feed_frame (SKLearnAnalyzer recipe without its dead logging/spectral subtraction)
       p50 / p99 per call for one stream with the FFN classifier.
Prints one JSON line."""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loop, ref_math as rm  # noqa: E402
from vad_b200.synth import synth_utterance  # noqa: E402


def pool_run(procs, utts):
    out = subprocess.run([sys.executable, "-m", "oracle.cpu_bench", "--utts", str(utts), "--procs", str(procs),
                          "--steps", "1", "--warmup", "1"], cwd=ROOT, stdout=subprocess.PIPE, text=True,
                         env=dict(os.environ, OMP_NUM_THREADS="1"))
    return json.loads(out.stdout.strip().splitlines()[-1])["audio_s_per_s"]


def main():
    w = rm.glorot_ffn(0)
    fb = rm.get_mel_filterbanks()
    pcm = synth_utterance(1234, 0, 960000)
    t0 = time.perf_counter()
    labels = ref_loop.vad_pcm_loop(pcm, w, fb)
    t_single = time.perf_counter() - t0
    an = ref_loop.LoopAnalyser(ref_loop.FFNClassifier(w), fb)
    frames = rm.split_into_frames(pcm[:16000 * 12])
    lat = []
    for fr in frames:
        f32 = fr.astype(np.float32)
        t0 = time.perf_counter()
        an.feed_frame(f32)
        lat.append(time.perf_counter() - t0)
    lat = np.sort(np.array(lat[10:])) * 1e3
    cores = os.cpu_count() or 1
    print(json.dumps({
        "host_cpus": cores,
        "single_core_cfg1_audio_s_per_s": 60.0 / t_single, "cfg1_decisions": int(labels.shape[0]),
        "pool4_audio_s_per_s": pool_run(4, 30), "pool_all_audio_s_per_s": pool_run(cores, 30 * max(1, cores // 4)),
        "feed_frame_p50_ms": float(lat[len(lat) // 2]), "feed_frame_p99_ms": float(lat[int(len(lat) * 0.99)]),
        "streams_sustainable_per_core_at_10ms_hop": float(10.0 / lat[int(len(lat) * 0.99)]),
        "note": "port of the reference loops (oracle/ref_loop.py); the reference's own SKLearnAnalyzer adds eager "
                "str(ndarray) logging and a dead spectral subtraction (5.3 ms p50 measured in the survey)"}))


if __name__ == "__main__":
    main()
