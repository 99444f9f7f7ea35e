"""Drop-ins for realtime_analysis/: the ``Analyser`` plugin interface (analyser.py:4-15), a
fused-GPU ``FusedAnalyser`` with SKLearnAnalyzer's exact feed_frame contract
(sklearn_analyser.py:15-130), a ``.predict`` duck-typed ``FFNClassifier``
(sklearn_analyser.py:71) and a many-stream ``StreamBank`` for 10 ms hop chunks."""
import abc
import pickle

import numpy as np
import torch

from . import runtime
from .runtime import FEAT_ANALYSER


def _handle_for(weights, handle):
    """One analyser == one classifier (sklearn_analyser.py:21-35): weights given without an explicit
    handle get a PRIVATE handle, so constructing a second object can never replace the classifier of
    an earlier one.  The process-wide default handle serves only the weight-free mfcc.* functions and
    objects that ask for it explicitly."""
    if weights is None:
        return handle or runtime.default_handle()
    w = runtime.load_ffn_npz(weights) if isinstance(weights, str) else weights
    if handle is None:
        return runtime.Handle(ffn_weights=w)
    handle.set_ffn_weights(w)      # explicit handle: the caller owns it and asked for this
    return handle


class Analyser:
    """realtime_analysis/analyser.py:4-15 (the reference's plugin boundary)."""

    def __init__(self):
        pass

    @abc.abstractmethod
    def load_init_inactive_frames(self, frames):
        return

    @abc.abstractmethod
    def feed_frame(self, frame):
        return


class FFNClassifier(object):
    """Object with ``.predict(X[n,39]) -> ndarray[n]`` of classes in {0, 1}
    (1 = VOICED, config.py:46; the FFN's music class 2 maps to 0), evaluated on the GPU."""

    def __init__(self, weights=None, handle=None):
        self.handle = _handle_for(weights, handle)
        if not self.handle.has_ffn:
            raise RuntimeError("FFNClassifier needs FFN weights")

    def predict(self, x):
        labels, _ = self.handle.ffn_predict(np.asarray(x, dtype=np.float32))
        return labels.cpu().numpy().astype(np.int64)

    def predict_logits(self, x):
        return self.handle.ffn_predict(np.asarray(x, dtype=np.float32))[1].cpu().numpy()


class FusedAnalyser(Analyser):
    """SKLearnAnalyzer (sklearn_analyser.py:15-130) on the GPU.

    ``feed_frame(frame)`` takes one whole frame (float / int array of <= 512 samples, vad.py
    feeds 400), returns None for the first five calls, then the raw frame fed three calls
    earlier when it is classified speech, else None.  ``classifier`` may be: None (use the fused
    FFN with ``ffn_weights``), a path to a pickled object with ``.predict`` (as the reference's
    ``fname``), or such an object; an external classifier receives the reference's (1, 39)
    float64 feature row and anything other than class 0 / 1 raises AssertionError
    (sklearn_analyser.py:76-82).  The reference's spectral subtraction is computed and discarded
    there (sklearn_analyser.py:121-123), so it is not reproduced."""

    FRAMES_BUFFER_SIZE = 5
    NOISE_BUFFER_SIZE = 5
    PROCESSING_FRAME_INDEX = 2

    def __init__(self, classifier=None, sample_rate=16000, fft_n=512, mfcc_num=13, low_hz=300, high_hz=8000,
                 fbank_num=26, ffn_weights=None, handle=None):
        Analyser.__init__(self)
        if (sample_rate, int(fft_n), mfcc_num, low_hz, high_hz, fbank_num) != (16000, 512, 13, 300, 8000, 26):
            raise NotImplementedError("vad_b200 kernels are compiled for the reference configuration only")
        self.sample_rate, self.fft_n, self.mfcc_num = sample_rate, int(fft_n), mfcc_num
        self.low_hz, self.high_hz, self.fbank_num = low_hz, high_hz, fbank_num
        if isinstance(classifier, str):
            with open(classifier, "rb") as f:
                classifier = pickle.load(f)
        self.classifier = classifier
        self.handle = _handle_for(ffn_weights if classifier is None else None, handle)
        self.filterbank = self.handle.filterbank()
        if classifier is None:
            if not self.handle.has_ffn:
                raise RuntimeError("FusedAnalyser needs FFN weights or an external classifier")
        self.frames_buffer = []
        self.frames_mfcc_buffer = []  # device tensors [13]
        self.noise_buffer = []
        self.last_logits = None

    def load_init_inactive_frames(self, frames):
        if len(frames) != FusedAnalyser.NOISE_BUFFER_SIZE:
            raise ValueError("Number of inactive frame must be the same as BUFFER SIZE")
        self.noise_buffer = list(self.handle.spec_frames(np.asarray(frames, dtype=np.float32)))

    def _update_frames_buffers(self, frame):
        m = self.handle.mfcc_frames(np.asarray(frame, dtype=np.float32))[0]
        if len(self.frames_buffer) == FusedAnalyser.FRAMES_BUFFER_SIZE:
            self.frames_buffer.pop(0)
            self.frames_mfcc_buffer.pop(0)
        self.frames_buffer.append(frame)
        self.frames_mfcc_buffer.append(m)

    def feed_frame(self, frame):
        if len(self.frames_buffer) < FusedAnalyser.FRAMES_BUFFER_SIZE:
            self._update_frames_buffers(frame)
            return None
        processing_frame = self.frames_buffer[self.PROCESSING_FRAME_INDEX]
        window = torch.stack(self.frames_mfcc_buffer)[None]
        labels, logits, feats = self.handle.vad_windows(window, FEAT_ANALYSER, want_feats=self.classifier is not None)
        if self.classifier is None:
            cls = int(labels[0].item())
            self.last_logits = logits[0]
        else:
            cls = self.classifier.predict(feats.cpu().numpy().astype(np.float64).reshape(1, -1))
        self._update_frames_buffers(frame)
        if cls == 1:
            return processing_frame
        elif cls == 0:
            return None
        else:
            raise AssertionError('Wrong classifier class')


class StreamBank(object):
    """n concurrent 16 kHz streams fed in 160-sample (10 ms) chunks.

    ``feed(chunks[n,160] int16)`` returns uint8[n]: the decision for the frame completed three
    frames earlier (the reference's feed_frame timing), or 255 while a stream's ring is filling.
    Chunk j completes frame j-2 (= 400 samples ending 80 samples into chunk j).  Per-stream state
    (320-sample history, 5-row MFCC ring) lives on the device; chunk input and label output go
    through pinned host buffers.  A tick is captured once into a CUDA graph and replayed: one graph
    launch per tick.  With ``zero_copy`` (default) the graph is the feed kernel alone: it reads the
    chunks from, and writes the labels to, the pinned host buffers directly over PCIe (each chunk
    sample is read once), which removes the two copy nodes and their scheduling gaps from the tick."""

    NOT_READY = 255

    def __init__(self, n_streams, ffn_weights=None, handle=None, use_graph=True, zero_copy=True):
        self.handle = _handle_for(ffn_weights, handle)
        if not self.handle.has_ffn:
            raise RuntimeError("StreamBank needs FFN weights")
        self.n = int(n_streams)
        self.bank = runtime.StreamBankHandle(self.handle, self.n)
        dev = self.handle.device
        self.h_chunks = torch.zeros((self.n, 160), dtype=torch.int16).pin_memory()
        self.h_labels = torch.zeros((self.n,), dtype=torch.uint8).pin_memory()
        self.h_logits = torch.zeros((self.n, 3), dtype=torch.float32).pin_memory()
        self.d_chunks = torch.zeros((self.n, 160), dtype=torch.int16, device=dev)
        self.d_labels = torch.zeros((self.n,), dtype=torch.uint8, device=dev)
        self.d_logits = torch.zeros((self.n, 3), dtype=torch.float32, device=dev)
        self.stream = torch.cuda.Stream(device=dev)
        self.use_graph = bool(use_graph)
        self.zero_copy = bool(zero_copy)
        self._graphs = {}

    def reset(self):
        self.bank.reset()
        torch.cuda.synchronize(self.handle.device)

    def _enqueue(self, want_logits):
        if self.zero_copy:   # pinned host memory is device-visible under UVA: no copy nodes
            self.bank.feed_ptr(self.h_chunks.data_ptr(), self.h_labels.data_ptr(),
                               self.h_logits.data_ptr() if want_logits else 0)
            return
        self.d_chunks.copy_(self.h_chunks, non_blocking=True)
        self.bank.feed_ptr(self.d_chunks.data_ptr(), self.d_labels.data_ptr(),
                           self.d_logits.data_ptr() if want_logits else 0)
        self.h_labels.copy_(self.d_labels, non_blocking=True)
        if want_logits:
            self.h_logits.copy_(self.d_logits, non_blocking=True)

    def _tick(self, want_logits):
        """One H2D + kernel + D2H; returns after the labels are on the host."""
        if self.use_graph:
            g = self._graphs.get(want_logits)
            if g is None:  # capture once (capturing does not execute: no chunk is consumed)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.stream):
                    self._enqueue(want_logits)
                self._graphs[want_logits] = g
            g.replay()                                   # one cudaGraphLaunch on the caller's current stream
            torch.cuda.current_stream(self.handle.device).synchronize()
        else:
            with torch.cuda.stream(self.stream):
                self._enqueue(want_logits)
            self.stream.synchronize()

    def feed(self, chunks, want_logits=False):
        """chunks: array-like int16 [n, 160] on the host.  Blocks until labels are on the host."""
        src = torch.as_tensor(np.asarray(chunks, dtype=np.int16) if not torch.is_tensor(chunks) else chunks)
        if tuple(src.shape) != (self.n, 160):
            raise ValueError("chunks must have shape (%d, 160)" % self.n)
        self.h_chunks.copy_(src)
        self._tick(want_logits)
        if want_logits:
            return self.h_labels.numpy().copy(), self.h_logits.numpy().copy()
        return self.h_labels.numpy().copy()

    def feed_pinned(self):
        """Low-latency tick: the caller has already written ``self.h_chunks`` (pinned); one graph
        launch (H2D + kernel + D2H) is issued and awaited.  Returns a view of ``self.h_labels``."""
        self._tick(False)
        return self.h_labels
