"""Batched / ragged entry points around the fused kernels, plus drop-ins for the reference's
offline driver: ``split_into_frames`` / ``process_file`` (dataset/file_processing.py:14-103) and
``scale_features`` (dataset/utils.py:5-32).  Host code only packs buffers and shapes results."""
import numpy as np
import torch

from . import runtime
from .runtime import MODE_MFCC, MODE_DATASET, MODE_VAD, FEAT_ANALYSER, FEAT_DATASET

ALIGN = 8  # utterance starts are multiples of 8 samples (16 bytes) for the TMA bulk copies


def pack_utterances(utterances, pin=False):
    """List of 1-D int16 arrays -> (flat int16 CPU tensor, offsets int64[n], lengths int64[n]).
    Each utterance starts at a multiple of 8 samples; the tail of the buffer is padded too.
    Replaces the per-file pickling of dataset_creator.process_files (dataset_creator.py:53-65)."""
    lengths = np.array([len(u) for u in utterances], dtype=np.int64)
    padded = (lengths + ALIGN - 1) // ALIGN * ALIGN
    offsets = np.zeros(len(utterances), dtype=np.int64)
    if len(utterances) > 1:
        offsets[1:] = np.cumsum(padded[:-1])
    total = int(padded.sum()) + ALIGN
    flat = torch.zeros(total, dtype=torch.int16)
    if pin:
        flat = flat.pin_memory()
    buf = flat.numpy()
    for u, o, n in zip(utterances, offsets, lengths):
        a = np.asarray(u)
        if a.dtype != np.int16:
            raise TypeError("utterances must be int16 PCM")
        buf[o:o + n] = a
    return flat, offsets, lengths


def uniform_layout(n_utt, utt_samples):
    stride = (utt_samples + ALIGN - 1) // ALIGN * ALIGN
    offsets = np.arange(n_utt, dtype=np.int64) * stride
    lengths = np.full(n_utt, utt_samples, dtype=np.int64)
    return offsets, lengths, stride


def _split_rows(t, row_offsets):
    return [t[int(row_offsets[i]):int(row_offsets[i + 1])] for i in range(len(row_offsets) - 1)]


def mfcc_batch(utterances, deltas=False, handle=None):
    """MFCC of every frame of every utterance.  deltas=False -> list of [T_u, 13] tensors;
    deltas=True -> list of [T_u - 5, 39] rows [c, d1, d2] (process_file's output)."""
    handle = handle or runtime.default_handle()
    flat, offsets, lengths = pack_utterances(utterances)
    plan = runtime.Plan(handle, offsets, lengths, MODE_DATASET if deltas else MODE_MFCC)
    out = plan.mfcc(flat.to(handle.device))
    return _split_rows(out, plan.row_offsets)


def vad_batch(utterances, handle=None, want_logits=False, feat_mode=FEAT_ANALYSER):
    """Per-frame speech / non-speech labels (uint8) for every utterance; optionally logits."""
    handle = handle or runtime.default_handle()
    if not handle.has_ffn:
        raise RuntimeError("set FFN weights first (Handle.set_ffn_weights)")
    flat, offsets, lengths = pack_utterances(utterances)
    plan = runtime.Plan(handle, offsets, lengths, MODE_VAD)
    labels, logits, _ = plan.vad(flat.to(handle.device), want_logits=want_logits, feat_mode=feat_mode)
    lab = _split_rows(labels, plan.row_offsets)
    if want_logits:
        return lab, _split_rows(logits, plan.row_offsets)
    return lab


# ---- drop-ins for dataset/file_processing.py ---------------------------------------------------
def parse_transcription(path, frame_rate):
    """dataset/stm_parser.py:5-26 (via file_processing.py:106-107): TED-LIUM .stm segment bounds
    in samples.  Lines with fewer than 7 fields or labelled ignore_time_segment_in_scoring are
    skipped; seconds are float32 and the product is truncated to int32, as in the reference."""
    starts, ends = [], []
    with open(path, "r") as f:
        for line in f:
            items = line.split(" ")
            if len(items) < 7 or items[6].strip() == "ignore_time_segment_in_scoring":
                continue
            starts.append(np.float32(items[3]))
            ends.append(np.float32(items[4]))
    s = (np.array(starts, dtype=np.float32) * frame_rate).astype(np.int32)
    e = (np.array(ends, dtype=np.float32) * frame_rate).astype(np.int32)
    return s, e


def _gather_segments(data, transcription_path, frame_rate):
    if transcription_path and frame_rate:
        starts, ends = parse_transcription(transcription_path, frame_rate)
        parts = [data[s:e] for s, e in zip(starts, ends)]
        return np.concatenate(parts).astype(np.int16) if parts else np.array([], dtype=np.int16)
    if transcription_path and frame_rate is None:
        raise Exception('You must specify frame_rate')
    return data


def split_into_frames(data, frame_size, step, transcription_path=None, frame_rate=None):
    """dataset/file_processing.py:80-103: list of overlapping views, strict '>' end rule."""
    data = _gather_segments(np.asarray(data), transcription_path, frame_rate)
    frames = []
    offset = 0
    while len(data) - offset > frame_size:
        frames.append(data[offset:offset + frame_size])
        offset += step
    return frames


def _read_audio(fname):
    if fname.endswith(".wav"):
        from scipy.io import wavfile
        return wavfile.read(fname)
    if fname.endswith(".sph"):
        from .io import read_sph
        return read_sph(fname)
    raise ValueError("Wrong file format: " + str(fname))


def process_file(args):
    """dataset/file_processing.py:14-77.  args = [fname, frame_size, frame_step, fft_n,
    mel_filterbank, mfcc_num, counter_queue, transcription_path]; returns the list of
    (mfcc, d1, d2) float64 triples, T-5 long.  The frame loop, ring and deltas run fused on
    the GPU (MODE_DATASET)."""
    fname, frame_size, frame_step, fft_n, mel_filterbank, mfcc_num, counter_queue, transcription_path = args[:8]
    sample_rate, f_raw = _read_audio(fname)
    if int(frame_size) != 400 or int(frame_step) != 160:
        raise NotImplementedError("vad_b200 kernels are compiled for frame 400 / step 160 (config.py:21-22)")
    from . import mfcc as _mfcc
    _mfcc._require(fft_n, mel_filterbank, mfcc_num)
    data = _gather_segments(np.asarray(f_raw), transcription_path, sample_rate)
    rows = mfcc_batch([np.ascontiguousarray(data, dtype=np.int16)], deltas=True)[0].cpu().numpy().astype(np.float64)
    features = [(r[:13].copy(), r[13:26].copy(), r[26:].copy()) for r in rows]
    if counter_queue is not None:
        processed_files = counter_queue.get() + 1
        if processed_files % 5 == 0:
            print("Processed " + str(processed_files) + ' files')
        counter_queue.put(processed_files)
    return features


def scale_features(features):
    """dataset/utils.py:5-32: in-place z-score with one scalar mean / population std per group
    (mfcc, d1, d2) over every frame x coefficient of the step.  Statistics are reduced on the
    device in float64; the nested-list container is updated in place like the reference."""
    dev = runtime.default_handle().device
    rows = [np.concatenate(fr) for ff in features for fr in ff]
    if not rows:
        return features
    t = torch.as_tensor(np.asarray(rows), dtype=torch.float64, device=dev).reshape(len(rows), 3, -1)
    mean = t.mean(dim=(0, 2), keepdim=True)
    std = t.std(dim=(0, 2), unbiased=False, keepdim=True)
    scaled = ((t - mean) / std).cpu().numpy()
    i = 0
    for ff in features:
        for fr in ff:
            for g in range(3):
                fr[g][:] = scaled[i, g]
            i += 1
    return features


# ---- feature sink: the CSV row layout of dataset/file_processing.py:110-148 -------
def create_table_header(mfcc_len):
    """dataset/file_processing.py:110-127: 13 'MFCC Coef', 13 'First delta', 13 'Second delta', 'voiced'."""
    header = ['MFCC Coef' + str(i + 1) for i in range(mfcc_len)]
    header += ['First delta' + str(i + 1) for i in range(mfcc_len)]
    header += ['Second delta' + str(i + 1) for i in range(mfcc_len)]
    header.append('voiced')
    return header


def write_features(writer, features, label):
    """dataset/file_processing.py:130-148: one CSV row per frame, 39 features then the class label."""
    for file_features in features:
        writer.writerows([np.concatenate((ff[0], ff[1], ff[2], [label])) for ff in file_features])


def write_feature_rows(writer, rows, label):
    """Same sink for the packed [n, 39] row tensors of ``mfcc_batch(deltas=True)``."""
    arr = rows.detach().cpu().numpy() if torch.is_tensor(rows) else np.asarray(rows)
    lab = np.full((arr.shape[0], 1), label, dtype=arr.dtype)
    writer.writerows(np.concatenate([arr, lab], axis=1))
