"""Float64 vectorised restatement of the reference's MFCC + FFN VAD path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the
reference lines it restates (paths relative to /root/reference).  All math is
float64 numpy; this is the canonical oracle the CUDA path is compared with.
"""
import numpy as np

# config.py:20-27
SAMPLERATE = 16000
FRAME_SIZE = 400
FRAME_STEP = 160
LOW_HZ = 300
HIGH_HZ = 8000
FILTERBANKS_NUM = 26
MFCC_NUM = 13
FFT_N = 512
# config.py:45-47
NONE_VOICED, VOICED, MUSIC = 0, 1, 2
EPS = float(np.finfo(float).eps)  # mfcc.py:74


# ----------------------------------------------------------------------------- framing
def n_frames(n_samples, frame_size=FRAME_SIZE, step=FRAME_STEP):
    """dataset/file_processing.py:99-101 -- ``while len(data) - offset > frame_size``
    (strict '>': a final frame that fits exactly is dropped)."""
    if n_samples <= frame_size:
        return 0
    return (n_samples - frame_size - 1) // step + 1


def n_outputs(n_samples, frame_size=FRAME_SIZE, step=FRAME_STEP):
    """dataset/file_processing.py:40-70 -- the 5-slot ring emits one row per frame
    fed after the ring is full and is never flushed: T - 5 rows (t = 2 .. T-4)."""
    return max(n_frames(n_samples, frame_size, step) - 5, 0)


def split_into_frames(data, frame_size=FRAME_SIZE, step=FRAME_STEP):
    """dataset/file_processing.py:80-103 (no transcription): [T, frame_size] view."""
    data = np.asarray(data)
    t = n_frames(len(data), frame_size, step)
    if t == 0:
        return np.zeros((0, frame_size), dtype=data.dtype)
    return np.lib.stride_tricks.sliding_window_view(data, frame_size)[::step][:t]


# ----------------------------------------------------------------------------- filterbank
def mel_from_hz(first_hz, upper_hz, n_bins):
    """mfcc.py:5-18."""
    first_mel = 1125.0 * np.log(1.0 + first_hz / 700.0)
    last_mel = 1125.0 * np.log(1.0 + upper_hz / 700.0)
    delta = (last_mel - first_mel) / (n_bins + 1)
    mels = [first_mel + i * delta for i in range(n_bins + 1)]
    mels.append(last_mel)
    mels.sort()
    return mels


def hz_from_mel(mels):
    """mfcc.py:21-27."""
    return [700 * (np.exp(m / 1125) - 1) for m in mels]


def convert_to_fft_bins(sample_rate, hzs, fft_n):
    """mfcc.py:30-36 -- note the (fft_n + 1) factor."""
    return [np.floor((fft_n + 1) * hz / sample_rate) for hz in hzs]


def mel_bin_edges(low_hz=LOW_HZ, up_hz=HIGH_HZ, fft_n=FFT_N, n_filters=FILTERBANKS_NUM,
                  sample_rate=SAMPLERATE):
    return [int(b) for b in convert_to_fft_bins(
        sample_rate, hz_from_mel(mel_from_hz(low_hz, up_hz, n_filters)), fft_n)]


def get_mel_filterbanks(low_hz=LOW_HZ, up_hz=HIGH_HZ, fft_n=FFT_N, n_filters=FILTERBANKS_NUM,
                        sample_rate=SAMPLERATE):
    """mfcc.py:39-56 -- un-normalised triangles, shape (n_filters, fft_n//2); the
    rising branch (``if``) wins at the peak; bins limited to 0..fft_n//2-1."""
    b = convert_to_fft_bins(sample_rate, hz_from_mel(mel_from_hz(low_hz, up_hz, n_filters)), fft_n)
    half = fft_n // 2
    fb = np.zeros((n_filters, half))
    k = np.arange(half, dtype=np.float64)
    for m in range(1, n_filters + 1):
        up = (k >= b[m - 1]) & (k <= b[m])
        down = (~up) & (k >= b[m]) & (k <= b[m + 1])
        with np.errstate(divide="ignore", invalid="ignore"):
            fb[m - 1, up] = ((k - b[m - 1]) / (b[m] - b[m - 1] + 0.0))[up]
            fb[m - 1, down] = ((b[m + 1] - k) / (b[m + 1] - b[m] + 0.0))[down]
    return fb


# ----------------------------------------------------------------------------- spectrum / cepstrum
def get_spec_mag(frames, fft_n=FFT_N):
    """mfcc.py:59-61 -- |FFT_n(zero-padded frame)[0:n/2] / n|^2 (rectangular window,
    Nyquist bin dropped).  ``frames``: [..., frame_len] any real dtype; the reference
    casts to float32 first (exact for int16 PCM), then numpy>=2 keeps single precision
    and older numpy used double; the oracle is float64 throughout."""
    x = np.asarray(frames).astype(np.float32).astype(np.float64)
    spec = np.fft.fft(x, fft_n, axis=-1)[..., : fft_n // 2] / float(fft_n)
    return spec.real ** 2 + spec.imag ** 2


def lifter_coefs(n_coef, L=22):
    """mfcc.py:85-90."""
    n = np.arange(n_coef)
    return 1 + (L / 2.0) * np.sin(np.pi * n / L)


def dct2_ortho_matrix(n_out, n_in):
    """scipy.fftpack.dct(type=2, norm='ortho')[:n_out] as a matrix (mfcc.py:76)."""
    n = np.arange(n_in)
    k = np.arange(n_out)[:, None]
    m = np.cos(np.pi * k * (2 * n + 1) / (2.0 * n_in)) * np.sqrt(2.0 / n_in)
    m[0] *= np.sqrt(0.5)
    return m


def _rowwise_dot(x, m):
    """x[..., n] . m[k, n] -> [..., k] with a reduction whose rounding depends only on the
    row's values (never on BLAS blocking / row position), so bit-identical frames give
    bit-identical outputs exactly as the reference's one-frame-at-a-time calls do."""
    x = np.asarray(x, dtype=np.float64)
    out = np.empty(x.shape[:-1] + (m.shape[0],))
    for k in range(m.shape[0]):
        nz = np.flatnonzero(m[k])
        if nz.size == 0:
            out[..., k] = 0.0
            continue
        lo, hi = nz[0], nz[-1] + 1
        out[..., k] = np.sum(x[..., lo:hi] * m[k, lo:hi], axis=-1)
    return out


def get_mfcc_from_spec(spec, filterbank, mfcc_n=MFCC_NUM):
    """mfcc.py:72-78 -- fbank dot, exact-zero -> eps, log10, DCT-II ortho [:n], lifter."""
    filterbank = np.asarray(filterbank, dtype=np.float64)
    e = _rowwise_dot(spec, filterbank)
    e = np.where(e == 0, EPS, e)
    coefs = np.log10(e)
    c = _rowwise_dot(coefs, dct2_ortho_matrix(mfcc_n, filterbank.shape[0]))
    return c * lifter_coefs(mfcc_n)


def get_mfcc(frames, filterbank, fft_n=FFT_N, mfcc_n=MFCC_NUM):
    """mfcc.py:67-69."""
    return get_mfcc_from_spec(get_spec_mag(frames, fft_n), filterbank, mfcc_n)


def mfcc_utterance(pcm, filterbank=None):
    """All frame MFCCs of one utterance: split_into_frames + get_mfcc -> [T, 13]."""
    if filterbank is None:
        filterbank = get_mel_filterbanks()
    fr = split_into_frames(pcm)
    if fr.shape[0] == 0:
        return np.zeros((0, MFCC_NUM))
    return get_mfcc(fr, filterbank)


# ----------------------------------------------------------------------------- 5-frame window features
def dataset_features(c):
    """dataset/file_processing.py:47-70 (+ mfcc.py:81-82): rows t = 2..T-4 of
    [c[t], c[t+1]-c[t-1], (c[t+2]-c[t]) - (c[t]-c[t-2])] -> [T-5, 39]."""
    c = np.asarray(c, dtype=np.float64)
    t = c.shape[0]
    if t < 6:
        return np.zeros((0, 3 * c.shape[1]))
    ctr = c[2:t - 3]
    d1 = c[3:t - 2] - c[1:t - 4]
    d2 = (c[4:t - 1] - ctr) - (ctr - c[0:t - 5])
    return np.concatenate([ctr, d1, d2], axis=1)


def analyser_features(c):
    """realtime_analysis/sklearn_analyser.py:52-69,103-107: z = (c[t]-mean5)/std5
    (population std over frames t-2..t+2, per coefficient), d1 = c[t+1]-c[t-1],
    d2 = (c[t+2]-z) - (z-c[t-2]); emitted for t = 2..T-4 (feed_frame call i = t+3)."""
    c = np.asarray(c, dtype=np.float64)
    t = c.shape[0]
    if t < 6:
        return np.zeros((0, 3 * c.shape[1]))
    win = np.lib.stride_tricks.sliding_window_view(c, 5, axis=0)[: t - 5]  # [T-5, 13, 5]
    mean = win.mean(axis=2)
    std = win.std(axis=2)
    with np.errstate(divide="ignore", invalid="ignore"):
        z = (c[2:t - 3] - mean) / std
    d1 = c[3:t - 2] - c[1:t - 4]
    d2 = (c[4:t - 1] - z) - (z - c[0:t - 5])
    return np.concatenate([z, d1, d2], axis=1)


def scale_features(groups):
    """dataset/utils.py:5-32: z-score with ONE scalar mean / population std per group
    (mfcc, d1, d2) over all frames x coefficients of the step.  ``groups``: list of
    [n_i, 39] arrays (one per file); returns new arrays."""
    allf = np.concatenate([np.asarray(g, dtype=np.float64) for g in groups], axis=0)
    nc = allf.shape[1] // 3
    out = []
    stats = []
    for g in range(3):
        blk = allf[:, g * nc:(g + 1) * nc]
        stats.append((blk.mean(), blk.std()))
    for f in groups:
        f = np.array(f, dtype=np.float64)
        for g in range(3):
            f[:, g * nc:(g + 1) * nc] = (f[:, g * nc:(g + 1) * nc] - stats[g][0]) / stats[g][1]
        out.append(f)
    return out, stats


# ----------------------------------------------------------------------------- FFN (parity unpinned)
FFN_DIMS = (39, 64, 32, 16, 3)  # learning/ffn_trainer.py:104-116


def glorot_ffn(seed=0, dtype=np.float32):
    """Keras-1 Dense defaults: init='glorot_uniform' (limit sqrt(6/(fan_in+fan_out))),
    zero bias; weights stored (in, out) so y = x.W + b.  Seeded random init stands in
    for the trained weights the reference never ships (.MISSING_LARGE_BLOBS)."""
    rng = np.random.default_rng(seed)
    w = {}
    for i in range(4):
        fi, fo = FFN_DIMS[i], FFN_DIMS[i + 1]
        lim = np.sqrt(6.0 / (fi + fo))
        w["W%d" % (i + 1)] = rng.uniform(-lim, lim, size=(fi, fo)).astype(dtype)
        w["b%d" % (i + 1)] = np.zeros(fo, dtype=dtype)
    return w


def ffn_forward(x, w):
    """learning/ffn_trainer.py:106-116: Dense64-relu(-relu)-Dense32-relu-Dense16-relu-
    Dense3-softmax.  Returns (logits, probs) in float64.  NaN inputs propagate."""
    h = np.asarray(x, dtype=np.float64)
    with np.errstate(invalid="ignore", over="ignore"):
        for i in (1, 2, 3):
            h = h @ w["W%d" % i].astype(np.float64) + w["b%d" % i].astype(np.float64)
            h = np.where(np.isnan(h), h, np.maximum(h, 0.0))
        logits = h @ w["W4"].astype(np.float64) + w["b4"].astype(np.float64)
        mx = np.max(np.where(np.isnan(logits), -np.inf, logits), axis=-1, keepdims=True) \
            if logits.size else logits
        z = logits - np.where(np.isfinite(mx), mx, 0.0) if logits.size else logits
        e = np.exp(z)
        probs = e / e.sum(axis=-1, keepdims=True)
    return logits, probs


def decide(logits):
    """realtime_analysis/sklearn_analyser.py:76 with config.py:45-47: speech <=>
    class == VOICED(1).  For the 3-class FFN: argmax(logits) == 1; numpy argmax of an
    all-NaN row is 0 -> non-speech."""
    logits = np.asarray(logits)
    if logits.shape[0] == 0:
        return np.zeros((0,), dtype=np.uint8)
    return (np.argmax(logits, axis=-1) == VOICED).astype(np.uint8)


def vad_utterance(pcm, w, filterbank=None, mode="analyser"):
    """Call stack D of SURVEY.md: PCM -> MFCC -> 5-frame features -> FFN -> label."""
    c = mfcc_utterance(pcm, filterbank)
    feats = analyser_features(c) if mode == "analyser" else dataset_features(c)
    logits, _ = ffn_forward(feats, w)
    return c, feats, logits, decide(logits)
