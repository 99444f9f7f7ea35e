"""Time the reference's CPU implementation of the path on host cores.  TEST / BENCH
INFRASTRUCTURE: used only by bench.py's ``cpu_baseline`` leg and ``--impl reference`` arm.

    python -m oracle.cpu_bench --utts 64 --procs 8 --steps 3 --warmup 1 [--kind reference|port]

Each step maps one worker function over ``--utts`` synthetic 10 s utterances with a
``multiprocessing.Pool`` -- the reference's own parallel shape (dataset_creator.py:63-65,84,
config.py:30 uses Pool(4); here ``--procs``).

kind "reference" (default when the reference is mounted or staged under oracle/_ref): the UNMODIFIED
reference code through its own entry point ``process_file`` (dataset/file_processing.py:14-77: wav read,
split_into_frames, per-frame mfcc.get_mfcc, 5-slot ring, deltas), imported through the python-3 shim.
The reference never runs its FFN, so the classifier stage -- 5-frame analyser features,
Dense 39-64-32-16-3, argmax == 1 -- is appended from the oracle restatement, vectorised per file (< 2 % of
the time).  The 10 s wav files are written to a temporary directory before the timed region.
kind "port": oracle/ref_loop.py, the per-frame port of the same loops (for boxes without oracle/_ref).
Prints one JSON line.  No torch / CUDA is imported in this process or its workers."""
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


class _Counter(object):          # the counter_queue argument of process_file (file_processing.py:72-75)
    def __init__(self):
        self.v = 1               # never a multiple of 5: no progress prints inside the timed region

    def get(self):
        return self.v

    def put(self, v):
        self.v = 1


def _init_reference():
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import reference_shim
    ref = reference_shim.load()
    _state["ref"] = ref
    _state["fb"] = ref.mfcc.get_mel_filterbanks(300, 8000, ref.FFT_N, 26, 16000)
    _state["w"] = rm.glorot_ffn(0)
    _state["q"] = _Counter()


def _work_reference(path):
    ref = _state["ref"]
    feats = ref.file_processing.process_file([path, 400, 160, ref.FFT_N, _state["fb"], 13, _state["q"], None])
    # classifier stage (not in the reference): analyser features need the MFCC rows t-2 .. t+2; process_file's
    # triples carry c[t] for t = 2 .. T-4, which is every row the windows of t = 4 .. T-6 use
    c = np.array([f[0] for f in feats])
    x = rm.analyser_features(c)
    labels = rm.decide(rm.ffn_forward(x, _state["w"])[0])
    return int(labels.sum()), len(feats)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=64)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--utt-seconds", type=float, default=10.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--kind", default="auto", choices=["auto", "reference", "port"])
    a = ap.parse_args()
    n = int(round(a.utt_seconds * 16000))
    kind = a.kind
    if kind == "auto":
        from oracle import reference_shim
        kind = "reference" if reference_shim.available() else "port"
    tmpdir = None
    if kind == "reference":
        import tempfile
        from scipy.io import wavfile
        tmpdir = tempfile.mkdtemp(prefix="vadb200_cpu_")
        jobs = []
        for u in range(a.utts):
            path = os.path.join(tmpdir, "u%05d.wav" % u)
            wavfile.write(path, 16000, synth_utterance(a.seed, u, n))
            jobs.append(path)
        init, work = _init_reference, _work_reference
    else:
        jobs = [(a.seed, u, n) for u in range(a.utts)]
        init, work = _init, _work
    times = []
    speech = frames = 0
    if a.procs > 1:
        pool = multiprocessing.Pool(a.procs, initializer=init)
        run = lambda: pool.map(work, jobs, chunksize=max(1, a.utts // (4 * a.procs)))  # noqa: E731
    else:
        init()
        pool = None
        run = lambda: list(map(work, jobs))  # noqa: E731
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
    if tmpdir is not None:
        import shutil
        shutil.rmtree(tmpdir, ignore_errors=True)
    audio_s = a.utts * a.utt_seconds
    total = sum(times)
    print(json.dumps({
        "audio_s_per_s": audio_s * len(times) / total, "ms_per_step": 1e3 * total / len(times),
        "cores": a.procs, "utts_per_step": a.utts, "utt_seconds": a.utt_seconds, "steps": len(times),
        "decisions": frames, "speech": speech, "host_cpus": os.cpu_count(), "kind": kind}))


if __name__ == "__main__":
    main()
