"""Freeze golden vectors by executing the UNMODIFIED reference (build container only).

    python -m oracle.make_golden        # writes tests/golden/*.npz

The reference has no tests / fixtures of its own; these files pin the
oracle restatement (and through it the CUDA path) to outputs of the reference's own code:
mfcc.py, dataset/file_processing.py:process_file, realtime_analysis/sklearn_analyser.py.
The FFN rows are produced by the oracle restatement only (parity unpinned, no reference run).
"""
import os
import pickle
import sys
import tempfile

import numpy as np
from scipy.io import wavfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import reference_shim, ref_math as rm  # noqa: E402
from vad_b200.synth import synth_utterance  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


class _Queue(object):
    def __init__(self):
        self.v = 0

    def get(self):
        return self.v

    def put(self, v):
        self.v = v


class RecordingClassifier(object):
    """Stub with the ``.predict`` duck type; records the (1,39) float64 rows it is given."""

    def __init__(self):
        self.rows = []

    def predict(self, x):
        self.rows.append(np.array(x[0], dtype=np.float64))
        return np.array([len(self.rows) % 2])


def ref_process_file(ref, pcm, fb):
    d = tempfile.mkdtemp(prefix="vadgold_")
    path = os.path.join(d, "u.wav")
    wavfile.write(path, 16000, pcm)
    feats = ref.file_processing.process_file([path, 400, 160, ref.FFT_N, fb, 13, _Queue(), None])
    if not feats:
        return np.zeros((0, 39))
    return np.array([np.concatenate(f) for f in feats])


def ref_analyser_rows(ref, pcm, fb):
    d = tempfile.mkdtemp(prefix="vadgold_")
    path = os.path.join(d, "cls.pkl")
    with open(path, "wb") as f:
        pickle.dump(RecordingClassifier(), f)
    an = ref.sklearn_analyser.SKLearnAnalyzer(path, fft_n=ref.FFT_N)
    frames = ref.file_processing.split_into_frames(pcm, 400, 160)
    an.load_init_inactive_frames([fr.astype(np.float32) for fr in frames[:5]])
    returned = []
    with np.errstate(all="ignore"):
        for fr in frames:
            r = an.feed_frame(fr.astype(np.float32))
            returned.append(r is not None)
    rows = an.classifier.rows
    return (np.array(rows) if rows else np.zeros((0, 39))), np.array(returned)


def main():
    ref = reference_shim.load()
    os.makedirs(OUT, exist_ok=True)
    fb = ref.mfcc.get_mel_filterbanks(300, 8000, ref.FFT_N, 26, 16000)

    # 1. single-frame known-answer test
    rng = np.random.default_rng(0)
    frame = (rng.standard_normal(400) * 3000).astype(np.int16)
    spec = ref.mfcc.get_spec_mag(frame, ref.FFT_N)
    np.savez_compressed(
        os.path.join(OUT, "kat_frame.npz"),
        frame=frame, filterbank=fb, spec=np.asarray(spec, dtype=np.float64),
        mfcc=ref.mfcc.get_mfcc(frame, ref.FFT_N, fb, 13),
        mfcc_from_spec=ref.mfcc.get_mfcc_from_spec(spec, fb, 13),
        mfcc_zero=ref.mfcc.get_mfcc(np.zeros(400, np.int16), ref.FFT_N, fb, 13),
        lifter13=ref.mfcc.lifter(np.ones(13)),
        mel_points=np.array(ref.mfcc.mel_from_hz(300, 8000, 26)),
        bins=np.array(ref.mfcc.convert_to_fft_bins(
            16000, ref.mfcc.hz_from_mel(ref.mfcc.mel_from_hz(300, 8000, 26)), ref.FFT_N)))

    # 2. utterances through process_file (dataset mode) and SKLearnAnalyzer (analyser mode)
    cases = {
        "synth_1p5s": synth_utterance(1234, 0, 24000),
        "synth_ragged": synth_utterance(1234, 7, 7777),
        "exact_fit": synth_utterance(1234, 3, 400 + 160 * 9),      # last frame fits exactly -> dropped
        "too_short": synth_utterance(1234, 4, 400 + 160 * 4 + 1),  # T = 5 -> no rows
        "silence_dc": np.concatenate([synth_utterance(1234, 5, 3200),
                                      np.zeros(2400, np.int16),
                                      np.full(2400, 1000, np.int16),
                                      synth_utterance(1234, 6, 3200)]),
    }
    tone_t = np.arange(16000)
    cases["tone_noise"] = np.clip(
        6000 * np.sin(2 * np.pi * 220 * tone_t / 16000.0) * (np.sin(2 * np.pi * 1.5 * tone_t / 16000.0) > 0)
        + np.random.default_rng(5).standard_normal(16000) * 200, -32768, 32767).astype(np.int16)
    blob = {}
    for name, pcm in cases.items():
        frames = ref.file_processing.split_into_frames(pcm, 400, 160)
        with np.errstate(all="ignore"):
            c = np.array([ref.mfcc.get_mfcc(fr, ref.FFT_N, fb, 13) for fr in frames]) \
                if len(frames) else np.zeros((0, 13))
            ds = ref_process_file(ref, pcm, fb)
            an, returned = ref_analyser_rows(ref, pcm, fb)
        blob[name + "/pcm"] = pcm
        blob[name + "/n_frames"] = np.array(len(frames))
        blob[name + "/mfcc"] = c
        blob[name + "/dataset_rows"] = ds
        blob[name + "/analyser_rows"] = an
        print(name, "L", len(pcm), "T", len(frames), "rows", ds.shape, an.shape)
    np.savez_compressed(os.path.join(OUT, "utterances.npz"), **blob)

    # 3. FFN rows -- oracle only (parity unpinned)
    w = rm.glorot_ffn(0)
    feats = blob["synth_1p5s/analyser_rows"]
    logits, probs = rm.ffn_forward(feats, w)
    np.savez_compressed(os.path.join(OUT, "ffn_oracle.npz"), logits=logits, probs=probs,
                        labels=rm.decide(logits), **w)
    print("golden written to", OUT)


if __name__ == "__main__":
    main()
