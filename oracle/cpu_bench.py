"""Time the reference-shaped CPU path (oracle/ref_loop.py) on host cores.  TEST / BENCH
INFRASTRUCTURE: used only by bench.py's ``cpu_baseline`` leg and ``--impl reference`` arm.

    python -m oracle.cpu_bench --utts 64 --procs 8 --steps 3 --warmup 1

Each step maps ``vad_pcm_loop`` (per-frame numpy MFCC loop as in dataset/file_processing.py:47-70
+ analyser window features + FFN + argmax==1) over ``--utts`` synthetic 10 s utterances with a
``multiprocessing.Pool`` -- the reference's own parallel shape (dataset_creator.py:63-65,84,
config.py:30 uses Pool(4); here ``--procs``).  Prints one JSON line.
No torch / CUDA is imported in this process or its workers."""
import argparse
import json
import multiprocessing
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_loop, ref_math as rm  # noqa: E402
from vad_b200.synth import synth_utterance  # noqa: E402

_state = {}


def _init():
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    _state["fb"] = rm.get_mel_filterbanks()
    _state["w"] = rm.glorot_ffn(0)


def _work(args):
    seed, utt, n = args
    pcm = synth_utterance(seed, utt, n)
    labels = ref_loop.vad_pcm_loop(pcm, _state["w"], _state["fb"])
    return int(labels.sum()), int(labels.shape[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=64)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--utt-seconds", type=float, default=10.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--seed", type=int, default=1234)
    a = ap.parse_args()
    n = int(round(a.utt_seconds * 16000))
    jobs = [(a.seed, u, n) for u in range(a.utts)]
    times = []
    speech = frames = 0
    if a.procs > 1:
        pool = multiprocessing.Pool(a.procs, initializer=_init)
        run = lambda: pool.map(_work, jobs, chunksize=max(1, a.utts // (4 * a.procs)))  # noqa: E731
    else:
        _init()
        pool = None
        run = lambda: list(map(_work, jobs))  # noqa: E731
    for it in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        res = run()
        dt = time.perf_counter() - t0
        if it >= a.warmup:
            times.append(dt)
            speech = sum(r[0] for r in res)
            frames = sum(r[1] for r in res)
    if pool is not None:
        pool.close()
        pool.join()
    audio_s = a.utts * a.utt_seconds
    total = sum(times)
    print(json.dumps({
        "audio_s_per_s": audio_s * len(times) / total, "ms_per_step": 1e3 * total / len(times),
        "cores": a.procs, "utts_per_step": a.utts, "utt_seconds": a.utt_seconds, "steps": len(times),
        "decisions": frames, "speech": speech, "host_cpus": os.cpu_count()}))


if __name__ == "__main__":
    main()
