"""Thin object layer over the C ABI: Handle (device tables + FFN weights), Plan (ragged packed
batch -> segment table), StreamBankHandle.  PyTorch is used only to own device / pinned memory
and to name the CUDA stream; every numeric result comes from libvadb200.so kernels."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check

MODE_MFCC, MODE_DATASET, MODE_VAD = 0, 1, 2
FEAT_ANALYSER, FEAT_DATASET = 0, 1
FFN_KEYS = ("W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4")
FFN_SHAPES = {"W1": (39, 64), "b1": (64,), "W2": (64, 32), "b2": (32,), "W3": (32, 16), "b3": (16,),
              "W4": (16, 3), "b4": (3,)}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def frames_for_length(n_samples):
    """dataset/file_processing.py:99 framing rule (strict '>')."""
    return int(_lib.load().vadb200_frames_for_length(int(n_samples)))


def outputs_for_length(n_samples):
    return int(_lib.load().vadb200_outputs_for_length(int(n_samples)))


def glorot_ffn(seed=0):
    """Keras-1 Dense default init (glorot_uniform, zero bias) for the 39-64-32-16-3 FFN of
    learning/ffn_trainer.py:104-116 -- stands in for the trained weights the reference never
    ships.  Same generator as oracle.ref_math.glorot_ffn (kept separate: the product may not
    import the oracle)."""
    rng = np.random.default_rng(seed)
    w = {}
    dims = (39, 64, 32, 16, 3)
    for i in range(4):
        lim = np.sqrt(6.0 / (dims[i] + dims[i + 1]))
        w["W%d" % (i + 1)] = rng.uniform(-lim, lim, size=(dims[i], dims[i + 1])).astype(np.float32)
        w["b%d" % (i + 1)] = np.zeros(dims[i + 1], dtype=np.float32)
    return w


def load_ffn_npz(path):
    """Weight container: .npz with W1[39,64] b1[64] W2[64,32] b2 W3[32,16] b3 W4[16,3] b4
    (Keras layout y = x.W + b); replaces model.load_weights (ffn_trainer.py:159-175)."""
    z = np.load(path)
    return {k: np.asarray(z[k], dtype=np.float32) for k in FFN_KEYS}


class Handle(object):
    """One per (process, device): owns the mel / DCT / twiddle tables and the FFN weights."""

    def __init__(self, device=None, ffn_weights=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("vad_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else int(
            device.index if isinstance(device, torch.device) else device))
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)  # make sure the primary context exists
        self._h = C.c_void_p()
        cfg = _lib.Config()
        self.lib.vadb200_default_config(C.byref(cfg))
        check(self.lib.vadb200_create(C.byref(cfg), self.device.index, C.byref(self._h)))
        self.has_ffn = False
        if ffn_weights is not None:
            self.set_ffn_weights(ffn_weights)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.vadb200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return _stream(self.device)

    def filterbank(self):
        out = np.empty((26, 256), dtype=np.float64)
        check(self.lib.vadb200_get_filterbank(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def set_ffn_weights(self, w):
        arrs = []
        for k in FFN_KEYS:
            a = np.ascontiguousarray(np.asarray(w[k], dtype=np.float32))
            if a.shape != FFN_SHAPES[k]:
                raise ValueError("FFN weight %s must have shape %r, got %r" % (k, FFN_SHAPES[k], a.shape))
            arrs.append(a)
        check(self.lib.vadb200_set_ffn_weights(self._h, *[a.ctypes.data_as(C.c_void_p) for a in arrs]))
        self.has_ffn = True

    def set_ffn_impl(self, impl):
        """0 / "fp32": CUDA-core FFMA; 1 / "tc": tcgen05 kind::tf32 on hi/lo operands; 2 / "tc16" (default):
        tcgen05 kind::f16 on statically scaled fp16 hi/lo operands (half the MMAs; falls back to "tc" for
        weights whose scales would be out of range)."""
        impl = {"fp32": 0, "tc": 1, "tc16": 2}.get(impl, impl)
        check(self.lib.vadb200_set_ffn_impl(self._h, int(impl)))

    @property
    def ffn_impl(self):
        return int(self.lib.vadb200_get_ffn_impl(self._h))

    # ---- per-frame API --------------------------------------------------------------------
    def _frames_tensor(self, frames):
        t = torch.as_tensor(np.asarray(frames, dtype=np.float32) if not torch.is_tensor(frames) else frames)
        t = t.to(device=self.device, dtype=torch.float32)
        if t.dim() == 1:
            t = t[None]
        return t.contiguous()

    def spec_frames(self, frames):
        t = self._frames_tensor(frames)
        out = torch.empty((t.shape[0], 256), dtype=torch.float32, device=self.device)
        check(self.lib.vadb200_spec_frames(self._h, _ptr(t), t.shape[0], t.shape[1], _ptr(out), self.stream))
        return out

    def mfcc_frames(self, frames):
        t = self._frames_tensor(frames)
        out = torch.empty((t.shape[0], 13), dtype=torch.float32, device=self.device)
        check(self.lib.vadb200_mfcc_frames(self._h, _ptr(t), t.shape[0], t.shape[1], _ptr(out), self.stream))
        return out

    def mfcc_from_spec(self, spec):
        t = self._frames_tensor(spec)
        if t.shape[1] != 256:
            raise ValueError("spectrum must have fft_n/2 = 256 bins")
        out = torch.empty((t.shape[0], 13), dtype=torch.float32, device=self.device)
        check(self.lib.vadb200_mfcc_from_spec(self._h, _ptr(t), t.shape[0], _ptr(out), self.stream))
        return out

    def vad_windows(self, windows, feat_mode=FEAT_ANALYSER, want_feats=False):
        t = torch.as_tensor(windows).to(device=self.device, dtype=torch.float32).contiguous().reshape(-1, 5, 13)
        n = t.shape[0]
        labels = torch.empty((n,), dtype=torch.uint8, device=self.device)
        logits = torch.empty((n, 3), dtype=torch.float32, device=self.device)
        feats = torch.empty((n, 39), dtype=torch.float32, device=self.device) if want_feats else None
        check(self.lib.vadb200_vad_windows(self._h, _ptr(t), n, feat_mode, _ptr(labels), _ptr(logits), _ptr(feats),
                                           self.stream))
        return labels, logits, feats

    def ffn_predict(self, x):
        t = torch.as_tensor(x).to(device=self.device, dtype=torch.float32).contiguous().reshape(-1, 39)
        n = t.shape[0]
        labels = torch.empty((n,), dtype=torch.uint8, device=self.device)
        logits = torch.empty((n, 3), dtype=torch.float32, device=self.device)
        check(self.lib.vadb200_ffn_predict(self._h, _ptr(t), n, _ptr(labels), _ptr(logits), self.stream))
        return labels, logits

    # ---- bench support ----------------------------------------------------------------------
    def synth_pcm(self, n_utt, utt_samples, seed=1234, first_utt=0, utt_stride=None, out=None):
        utt_stride = utt_samples if utt_stride is None else utt_stride
        if out is None:
            out = torch.zeros((n_utt * utt_stride + 8,), dtype=torch.int16, device=self.device)
        check(self.lib.vadb200_synth_pcm(self._h, _ptr(out), n_utt, utt_samples, utt_stride, seed & 0xFFFFFFFF,
                                         first_utt, self.stream))
        return out

    def fp32_peak(self, variant=0, iters=4096):
        v = C.c_double()
        check(self.lib.vadb200_fp32_peak(self._h, variant, iters, C.byref(v)))
        return v.value

    def set_host_chunk_samples(self, samples):
        check(self.lib.vadb200_set_host_chunk_samples(self._h, int(samples)))

    def set_plan_segment_frames(self, frames):
        """Override the segment length of plans created afterwards (0 = automatic)."""
        check(self.lib.vadb200_set_plan_segment_frames(self._h, int(frames)))

    # ---- feature sink / ingest ------------------------------------------------------------------
    def scale_rows(self, rows, want_stats=True):
        """dataset/utils.py:5-32 on a packed [n, 39] float32 CUDA tensor, in place.  Returns the six
        statistics (mean x3, population std x3) as a float64 array when want_stats."""
        if not (torch.is_tensor(rows) and rows.is_cuda and rows.dtype == torch.float32 and rows.is_contiguous()
                and rows.dim() == 2 and rows.shape[1] == 39):
            raise TypeError("rows must be a contiguous [n, 39] float32 CUDA tensor")
        stats = np.zeros(6, dtype=np.float64) if want_stats else None
        check(self.lib.vadb200_scale_rows(self._h, _ptr(rows), rows.shape[0],
                                          stats.ctypes.data_as(C.c_void_p) if want_stats else C.c_void_p(0),
                                          self.stream))
        return stats

    def ingest_pcm(self, raw, src_start, dst_start, lengths, out, big_endian=False):
        """Device decode + gather: ``raw`` uint8 CUDA tensor holding 16-bit samples (2-byte aligned),
        segment i = ``lengths[i]`` samples from sample index ``src_start[i]`` to ``out[dst_start[i]:]``."""
        src = torch.as_tensor(np.asarray(src_start, dtype=np.int64)).to(self.device)
        dst = torch.as_tensor(np.asarray(dst_start, dtype=np.int64)).to(self.device)
        ln_host = np.asarray(lengths, dtype=np.int64)
        ln = torch.as_tensor(ln_host).to(self.device)
        if ln_host.size == 0:
            return out
        check(self.lib.vadb200_ingest_pcm(self._h, _ptr(raw), 1 if big_endian else 0, _ptr(src), _ptr(dst), _ptr(ln),
                                          int(ln_host.size), int(ln_host.max()), _ptr(out), self.stream))
        return out


class Plan(object):
    """Segment table for one ragged packed batch (offsets / lengths in samples, host side)."""

    def __init__(self, handle, offsets, lengths, mode):
        self.handle = handle
        self.lib = handle.lib
        self.mode = mode
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        if self.offsets.shape != self.lengths.shape or self.offsets.ndim != 1:
            raise ValueError("offsets and lengths must be 1-D arrays of equal length")
        self.n_utt = int(self.offsets.shape[0])
        self._p = C.c_void_p()
        check(self.lib.vadb200_plan_create(handle._h, self.offsets.ctypes.data_as(C.c_void_p),
                                           self.lengths.ctypes.data_as(C.c_void_p), self.n_utt, mode,
                                           C.byref(self._p)))
        self.total_rows = int(self.lib.vadb200_plan_total_rows(self._p))
        self.segment_frames = int(self.lib.vadb200_plan_segment_frames(self._p))
        self.segment_count = int(self.lib.vadb200_plan_segment_count(self._p))
        self.row_offsets = np.empty(self.n_utt + 1, dtype=np.int64)
        check(self.lib.vadb200_plan_row_offsets(self._p, self.row_offsets.ctypes.data_as(C.c_void_p)))

    def close(self):
        if getattr(self, "_p", None) is not None and self._p:
            self.lib.vadb200_plan_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check_pcm(self, pcm):
        if not (torch.is_tensor(pcm) and pcm.dtype == torch.int16 and pcm.is_contiguous() and pcm.dim() == 1):
            raise TypeError("pcm must be a contiguous 1-D torch.int16 tensor")

    def mfcc(self, pcm, out=None):
        """MODE_MFCC -> [rows, 13]; MODE_DATASET -> [rows, 39] (float32, on pcm's device)."""
        self._check_pcm(pcm)
        width = 13 if self.mode == MODE_MFCC else 39
        if out is None:
            out = torch.empty((self.total_rows, width), dtype=torch.float32, device=pcm.device)
        check(self.lib.vadb200_mfcc_packed(self._p, _ptr(pcm), pcm.numel(), _ptr(out), self.handle.stream))
        return out

    def vad(self, pcm, labels=None, want_logits=False, want_feats=False, feat_mode=FEAT_ANALYSER, logits=None):
        self._check_pcm(pcm)
        dev = pcm.device
        if labels is None:
            labels = torch.empty((self.total_rows,), dtype=torch.uint8, device=dev)
        if logits is None and want_logits:
            logits = torch.empty((self.total_rows, 3), dtype=torch.float32, device=dev)
        feats = torch.empty((self.total_rows, 39), dtype=torch.float32, device=dev) if want_feats else None
        check(self.lib.vadb200_vad_packed(self._p, _ptr(pcm), pcm.numel(), _ptr(labels), _ptr(logits), _ptr(feats),
                                          feat_mode, self.handle.stream))
        return labels, logits, feats

    def mfcc_host(self, pcm_host, out_host=None):
        """MODE_MFCC / MODE_DATASET end to end with HOST tensors: chunked H2D, fused kernel, D2H of the rows."""
        if not (torch.is_tensor(pcm_host) and pcm_host.dtype == torch.int16 and not pcm_host.is_cuda):
            raise TypeError("pcm_host must be a CPU torch.int16 tensor")
        width = 13 if self.mode == MODE_MFCC else 39
        if out_host is None:
            out_host = torch.empty((self.total_rows, width), dtype=torch.float32).pin_memory()
        check(self.lib.vadb200_mfcc_host(self._p, _ptr(pcm_host), pcm_host.numel(), _ptr(out_host)))
        return out_host

    def vad_host(self, pcm_host, labels_host=None, logits_host=None, feat_mode=FEAT_ANALYSER):
        """End to end with HOST tensors (pinned for full PCIe speed): H2D, fused kernel, D2H."""
        if not (torch.is_tensor(pcm_host) and pcm_host.dtype == torch.int16 and not pcm_host.is_cuda):
            raise TypeError("pcm_host must be a CPU torch.int16 tensor")
        if labels_host is None:
            labels_host = torch.empty((self.total_rows,), dtype=torch.uint8).pin_memory()
        check(self.lib.vadb200_vad_host(self._p, _ptr(pcm_host), pcm_host.numel(), _ptr(labels_host),
                                        _ptr(logits_host), feat_mode))
        return labels_host, logits_host


class StreamBankHandle(object):
    def __init__(self, handle, n_streams):
        self.handle = handle
        self.lib = handle.lib
        self.n = int(n_streams)
        self._b = C.c_void_p()
        check(self.lib.vadb200_stream_bank_create(handle._h, self.n, C.byref(self._b)))

    def close(self):
        if getattr(self, "_b", None) is not None and self._b:
            self.lib.vadb200_stream_bank_destroy(self._b)
            self._b = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        check(self.lib.vadb200_stream_bank_reset(self._b, self.handle.stream))

    def feed_ptr(self, chunks_ptr, labels_ptr, logits_ptr=0, stream=None):
        check(self.lib.vadb200_stream_feed(self._b, C.c_void_p(chunks_ptr), C.c_void_p(labels_ptr),
                                           C.c_void_p(logits_ptr), stream if stream is not None else self.handle.stream))


_default = {}


def default_handle(device=None):
    """Process-wide handle per device for the function-style drop-in API (mfcc.py functions)."""
    idx = torch.cuda.current_device() if device is None else int(device)
    if idx not in _default:
        _default[idx] = Handle(idx)
    return _default[idx]


def _handle_get_deltas(self, a, b):
    ta = torch.as_tensor(np.asarray(a, dtype=np.float32) if not torch.is_tensor(a) else a).to(
        device=self.device, dtype=torch.float32).contiguous()
    tb = torch.as_tensor(np.asarray(b, dtype=np.float32) if not torch.is_tensor(b) else b).to(
        device=self.device, dtype=torch.float32).contiguous()
    if ta.shape != tb.shape:
        raise ValueError("operands must have equal shapes")
    out = torch.empty_like(ta)
    check(self.lib.vadb200_get_deltas(self._h, _ptr(ta), _ptr(tb), ta.numel(), _ptr(out), self.stream))
    return out


def _handle_lifter(self, cepstra, L=22):
    t = torch.as_tensor(np.asarray(cepstra, dtype=np.float32) if not torch.is_tensor(cepstra) else cepstra).to(
        device=self.device, dtype=torch.float32).contiguous()
    ncoef = int(t.shape[0]) if t.dim() == 1 else int(t.shape[-1])
    out = torch.empty_like(t)
    check(self.lib.vadb200_lifter(self._h, _ptr(t), t.numel() // max(ncoef, 1), ncoef, int(L), _ptr(out), self.stream))
    return out


Handle.get_deltas = _handle_get_deltas
Handle.lifter = _handle_lifter
