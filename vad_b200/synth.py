"""Counter-based synthetic 16 kHz int16 PCM, bit-identical on host (numpy) and device.

Integer arithmetic only, so the CUDA generator (csrc/synth.cu, ``vadb200_synth_pcm``)
and this numpy mirror agree bit-for-bit: parity tests regenerate on the host exactly what
the bench generated in HBM.  Signal: sum of four 16-bit uniforms (Irwin-Hall, ~Gaussian)
times a per-128 ms-block gain that alternates loud (std 870..3200 LSB) and quiet
(std 23..96 LSB) blocks -- a speech-like on/off envelope so both VAD classes occur.
"""
import numpy as np

_M = np.uint64(0xFFFFFFFF)


def _mix32(x):
    """lowbias32 integer hash on uint64 arrays holding 32-bit values."""
    x = x & _M
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & _M
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & _M
    x ^= x >> np.uint64(16)
    return x


def utt_key(seed, utt_id):
    s = _mix32(np.uint64(seed & 0xFFFFFFFF) ^ np.uint64(0x9E3779B9))
    return _mix32((s + np.uint64(utt_id & 0xFFFFFFFF)) & _M)


def synth_utterance(seed, utt_id, n_samples, start=0):
    """int16[n_samples] for samples start .. start+n_samples-1 of utterance ``utt_id``."""
    key = utt_key(seed, utt_id)
    i = (np.arange(n_samples, dtype=np.uint64) + np.uint64(start)) & _M
    a = _mix32((key + i) & _M)
    b = _mix32(a ^ np.uint64(0x85EBCA6B))
    s = ((a & np.uint64(0xFFFF)) + (a >> np.uint64(16)) + (b & np.uint64(0xFFFF))
         + (b >> np.uint64(16))).astype(np.int64) - 131070
    g = _mix32(((key ^ np.uint64(0x5BD1E995)) + (i >> np.uint64(11))) & _M)
    loud = (g & np.uint64(0x1000)) != 0
    gain = np.where(loud, np.uint64(1500) + (g & np.uint64(0xFFF)),
                    np.uint64(40) + (g & np.uint64(0x7F))).astype(np.int64)
    return ((s * gain) >> 16).astype(np.int16)


def synth_batch(seed, first_utt, n_utt, n_samples):
    """[n_utt, n_samples] int16, utterance ids first_utt .. first_utt+n_utt-1."""
    return np.stack([synth_utterance(seed, first_utt + u, n_samples) for u in range(n_utt)])
