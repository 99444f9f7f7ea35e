"""Per-frame port of the reference's CPU loops (TEST INFRASTRUCTURE / CPU BASELINE).

Where ``ref_math`` is vectorised for checking at scale, this module keeps the
reference's *shape* -- one numpy call chain per 25 ms frame, a python 5-slot ring --
so that timing it on the GPU box's host cores reproduces what the reference's own
CPU implementation costs (``bench.py --impl reference`` and ``cpu_baseline``).
/root/reference itself cannot travel to the GPU box, so this port stands in for it
(cpu_baseline.kind = "port").  Line citations are relative to /root/reference.
"""
import numpy as np
from scipy.fftpack import dct

from . import ref_math as rm


def get_spec_mag(frame, fft_n):
    """mfcc.py:59-61 (verbatim semantics: float32 cast, complex FFT, first half, /n, |.|^2)."""
    frame = frame.astype(np.float32)
    return np.square(np.absolute(np.fft.fft(frame, fft_n)[0:fft_n // 2] / np.float32(fft_n)))


def lifter(cepstra, L=22):
    """mfcc.py:85-93."""
    ncoeff = np.shape(cepstra)[0]
    n = np.arange(ncoeff)
    lift = 1 + (L / 2.) * np.sin(np.pi * n / L)
    return lift * cepstra


def get_mfcc_from_spec(spec, filterbank, mfcc_n):
    """mfcc.py:72-78."""
    e = np.dot(spec, filterbank.T)
    e = np.where(e == 0, np.finfo(float).eps, e)
    return lifter(dct(np.log10(e), type=2, norm='ortho')[:mfcc_n])


def get_mfcc(frame, fft_n, filterbank, mfcc_n):
    """mfcc.py:67-69."""
    return get_mfcc_from_spec(get_spec_mag(frame, fft_n), filterbank, mfcc_n)


def process_pcm(pcm, filterbank, frame_size=400, frame_step=160, fft_n=512, mfcc_num=13):
    """dataset/file_processing.py:38-70: framing loop + 5-slot ring + difference deltas.
    Returns a list of (mfcc, d1, d2) tuples, T-5 long."""
    frames = []
    offset = 0
    while len(pcm) - offset > frame_size:          # :99
        frames.append(pcm[offset:offset + frame_size])
        offset += frame_step
    buf = []
    feats = []
    for frame in frames:                           # :47
        if len(buf) < 5:
            buf.append(get_mfcc(frame, fft_n, filterbank, mfcc_num))
        else:
            c = buf[2]
            prev_d = np.subtract(c, buf[0])
            next_d = np.subtract(buf[4], c)
            feats.append((c, np.subtract(buf[3], buf[1]), np.subtract(next_d, prev_d)))
            buf.pop(0)
            buf.append(get_mfcc(frame, fft_n, filterbank, mfcc_num))
    return feats


class LoopAnalyser(object):
    """realtime_analysis/sklearn_analyser.py:15-130 without the dead spectral-subtraction /
    logging (results do not depend on them).  ``classifier`` is any object
    with ``.predict(X[1,39]) -> [cls]``."""

    def __init__(self, classifier, filterbank=None, fft_n=512, mfcc_num=13):
        self.classifier = classifier
        self.fft_n = fft_n
        self.mfcc_num = mfcc_num
        self.filterbank = rm.get_mel_filterbanks() if filterbank is None else filterbank
        self.frames_buffer = []
        self.frames_mfcc_buffer = []
        self.noise_buffer = []

    def load_init_inactive_frames(self, frames):   # :37-44
        if len(frames) != 5:
            raise ValueError("Number of inactive frame must be the same as BUFFER SIZE")
        self.noise_buffer = [get_spec_mag(f, self.fft_n) for f in frames]

    def _update(self, frame):                      # :120-130
        m = get_mfcc_from_spec(get_spec_mag(frame, self.fft_n), self.filterbank, self.mfcc_num)
        if len(self.frames_buffer) == 5:
            self.frames_buffer.pop(0)
            self.frames_mfcc_buffer.pop(0)
        self.frames_buffer.append(frame)
        self.frames_mfcc_buffer.append(m)

    def features(self):                            # :52-69,103-107
        b = self.frames_mfcc_buffer
        with np.errstate(divide="ignore", invalid="ignore"):
            z = np.divide(np.subtract(b[2], np.mean(b, axis=0)), np.std(b, axis=0))
        prev_d = np.subtract(z, b[0])
        next_d = np.subtract(b[4], z)
        f = np.array([], dtype=np.float32)
        return np.append(f, (z, np.subtract(b[3], b[1]), np.subtract(next_d, prev_d)))

    def feed_frame(self, frame):                   # :46-82
        if len(self.frames_buffer) < 5:
            self._update(frame)
            return None
        processing_frame = self.frames_buffer[2]
        cls = self.classifier.predict(self.features().reshape(1, -1))
        self._update(frame)
        if cls == 1:
            return processing_frame
        elif cls == 0:
            return None
        raise AssertionError('Wrong classifier class')


class FFNClassifier(object):
    """``.predict`` duck type (sklearn_analyser.py:71) around the FFN of
    learning/ffn_trainer.py:104-116; class 2 (music) maps to 0 (non-speech)."""

    def __init__(self, weights):
        self.w = weights

    def predict(self, x):
        logits, _ = rm.ffn_forward(x, self.w)
        return rm.decide(logits).astype(np.int64)


def vad_pcm_loop(pcm, weights, filterbank):
    """Whole reference-shaped CPU path for one utterance: per-frame MFCC loop, analyser
    window features per frame, batched FFN at the end (the FFN has no reference loop).
    Returns uint8 labels [T-5]."""
    frames = []
    offset = 0
    while len(pcm) - offset > 400:
        frames.append(pcm[offset:offset + 400])
        offset += 160
    c = [get_mfcc(f, 512, filterbank, 13) for f in frames]
    if len(c) < 6:
        return np.zeros((0,), dtype=np.uint8)
    feats = rm.analyser_features(np.asarray(c))
    logits, _ = rm.ffn_forward(feats, weights)
    return rm.decide(logits)
