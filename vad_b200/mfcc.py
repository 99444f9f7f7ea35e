"""Drop-in for the reference's ``mfcc.py`` -- same eleven function names, argument meaning and
return shapes (mfcc.py:5-93) -- with the per-frame numerics running on the B200 through
libvadb200.so.  The table constructors (mel points, bin edges, filterbank) are init-time host
math exactly as in the reference; everything that touches samples launches a CUDA kernel.

The kernels are specialised for the reference configuration (config.py:20-27); any other
fft_n / filterbank / mfcc_n raises NotImplementedError rather than falling back to the CPU.
"""
import numpy as np

from . import runtime


# ---- init-time table constructors (host) -----------------------------------------------------
def mel_from_hz(first_hz, upper_hz, n_bins):
    """mfcc.py:5-18: n_bins + 2 mel points from first_hz to upper_hz."""
    mels = []
    first_mel = 1125.0 * np.log(1.0 + first_hz / 700.0)
    last_mel = 1125.0 * np.log(1.0 + upper_hz / 700.0)
    delta = (last_mel - first_mel) / (n_bins + 1)
    for i in range(n_bins + 1):
        mels.append(first_mel + i * delta)
    mels.append(last_mel)
    mels.sort()
    return mels


def one_hz_from_mel(mel):
    """mfcc.py:21-22."""
    return 700 * (np.exp(mel / 1125) - 1)


def hz_from_mel(mels):
    """mfcc.py:25-27."""
    return list(map(one_hz_from_mel, mels))


def convert_to_fft_bins(sample_rate, hzs, fft_n):
    """mfcc.py:30-36: floor((fft_n + 1) * hz / sample_rate), floats holding integers."""
    return [np.floor((fft_n + 1) * hz / sample_rate) for hz in hzs]


def get_mel_filterbanks(low_hz, up_hz, fft_n, n_filters, sample_rate):
    """mfcc.py:39-56: un-normalised triangles, ndarray [n_filters, fft_n // 2] float64."""
    mels_bin = convert_to_fft_bins(sample_rate, hz_from_mel(mel_from_hz(low_hz, up_hz, n_filters)), fft_n)
    half = int(fft_n) // 2
    filterbank = np.zeros((n_filters, half))
    for m in range(1, n_filters + 1):
        for k in range(half):
            if (k >= mels_bin[m - 1]) and (k <= mels_bin[m]):
                filterbank[m - 1, k] = (k - mels_bin[m - 1] + 0.0) / (mels_bin[m] - mels_bin[m - 1] + 0.0)
            elif (k >= mels_bin[m]) and (k <= mels_bin[m + 1]):
                filterbank[m - 1, k] = (mels_bin[m + 1] - k + 0.0) / (mels_bin[m + 1] - mels_bin[m] + 0.0)
    return filterbank


# ---- per-frame numerics (device) ---------------------------------------------------------------
def _require(fft_n=None, filterbank=None, mfcc_n=None):
    if fft_n is not None and int(fft_n) != 512:
        raise NotImplementedError("vad_b200 kernels are compiled for fft_n = 512 (config.py:27)")
    if mfcc_n is not None and int(mfcc_n) != 13:
        raise NotImplementedError("vad_b200 kernels are compiled for mfcc_n = 13 (config.py:26)")
    if filterbank is not None:
        fb = np.asarray(filterbank)
        ref = runtime.default_handle().filterbank()
        if fb.shape != ref.shape or not np.array_equal(fb, ref):
            raise NotImplementedError(
                "vad_b200 kernels are compiled for get_mel_filterbanks(300, 8000, 512, 26, 16000)")


def get_spec_mag(frame, fft_n):
    """mfcc.py:59-61: |FFT_n(frame)[0:n/2] / n|^2 -> float32[256] (numpy>=2 keeps float32)."""
    _require(fft_n=fft_n)
    frame = np.asarray(frame)
    if frame.ndim != 1 or frame.shape[0] > 512:
        raise ValueError("frame must be 1-D with at most fft_n samples")
    return runtime.default_handle().spec_frames(frame.astype(np.float32))[0].cpu().numpy()


def get_mfcc(frame, fft_n, filterbank, mfcc_n):
    """mfcc.py:67-69 -> float64[13]."""
    _require(fft_n, filterbank, mfcc_n)
    frame = np.asarray(frame)
    if frame.ndim != 1 or frame.shape[0] > 512:
        raise ValueError("frame must be 1-D with at most fft_n samples")
    return runtime.default_handle().mfcc_frames(frame.astype(np.float32))[0].cpu().numpy().astype(np.float64)


def get_mfcc_from_spec(spec, filterbank, mfcc_n):
    """mfcc.py:72-78 -> float64[13]."""
    _require(None, filterbank, mfcc_n)
    spec = np.asarray(spec, dtype=np.float32)
    return runtime.default_handle().mfcc_from_spec(spec)[0].cpu().numpy().astype(np.float64)


def get_deltas(mfcc2, mfcc1):
    """mfcc.py:81-82: mfcc2 - mfcc1."""
    out = runtime.default_handle().get_deltas(mfcc2, mfcc1)
    return out.cpu().numpy().astype(np.result_type(np.asarray(mfcc2).dtype, np.asarray(mfcc1).dtype, np.float32))


def lifter(cepstra, L=22):
    """mfcc.py:85-93: cepstra * (1 + (L/2) sin(pi n / L)); L <= 0 returns cepstra unchanged."""
    if L <= 0:
        return cepstra
    return runtime.default_handle().lifter(cepstra, L).cpu().numpy().astype(np.float64)
