"""Ingest side of the hot path: container parsing on the host (a few header bytes per file), sample decode
and segment gathering on the device.

* ``read_sph`` / ``sph_layout`` restate dataset/sph.py:33-63 (NIST SPHERE).  The reference decodes
  ``sample_count`` big-endian samples starting RIGHT AFTER THE NINTH HEADER LINE -- not at the declared header
  size -- and always most-significant-byte first; ``reference_compat=True`` (default) reproduces exactly that
  byte stream so files give the reference's samples, ``False`` parses the header properly.
* ``wav_layout`` finds the 16-bit mono PCM payload of a RIFF/WAVE file (scipy.io.wavfile.read in
  dataset/file_processing.py:26-27).
* ``parse_stm`` restates dataset/stm_parser.py:5-26.
* ``DeviceIngest`` stages the raw bytes of many files in ONE pinned buffer, uploads them with one asynchronous
  copy and runs ``vadb200_ingest_pcm`` (byte-swap + STM range gather, dataset/file_processing.py:87-94) into the
  packed, 8-sample-aligned int16 layout the fused kernels read.
"""
import os

import numpy as np
import torch

ALIGN = 8


# ------------------------------------------------------------------------------------------ containers
def sph_layout(fname, reference_compat=True):
    """-> (sample_rate, data_byte_offset, n_samples, big_endian) of a 16-bit NIST SPHERE file."""
    size = os.path.getsize(fname)
    with open(fname, "rb") as f:
        if reference_compat:
            header = [f.readline(1024) for _ in range(9)]                 # dataset/sph.py:36-38
            try:
                samples_num = int(header[2].split(b" ")[2])               # :40
                sample_width = int(header[3].split(b" ")[2])              # :41
                framerate = int(header[6].split(b" ")[2])                 # :43
            except (IndexError, ValueError):
                raise ValueError("not a NIST SPHERE file the reference reader accepts: " + fname)
            if sample_width != 2:
                raise NotImplementedError("only 16-bit SPHERE PCM is supported")
            offset = f.tell()                                             # decoding starts here (:47)
            # samples_num samples, zeros where the file ends early (np.zeros at :45); always MSB first (:55-57)
            return framerate, offset, samples_num, True
        head = f.read(1024)
        if not head.startswith(b"NIST_1A"):
            raise ValueError("not a NIST SPHERE file: " + fname)
        hsize = int(head.split(b"\n")[1].strip())
        if hsize > 1024:
            head += f.read(hsize - 1024)
    fields = {}
    for line in head.split(b"\n")[2:]:
        parts = line.strip().split()
        if not parts or parts[0] == b"end_head":
            break
        if len(parts) >= 3:
            fields[parts[0].decode()] = parts[2].decode()
    if int(fields.get("sample_n_bytes", 2)) != 2:
        raise NotImplementedError("only 16-bit SPHERE PCM is supported")
    n = min(int(fields.get("sample_count", (size - hsize) // 2)), (size - hsize) // 2)
    return int(fields.get("sample_rate", 16000)), hsize, n, fields.get("sample_byte_format", "10") == "10"


def read_sph(fname, reference_compat=True):
    """Host decode (small files, tests): -> (sample_rate, int16[n])."""
    rate, off, n, be = sph_layout(fname, reference_compat)
    avail = max(0, min(2 * n, (os.path.getsize(fname) - off) // 2 * 2))
    raw = np.zeros(2 * n, dtype=np.uint8)
    raw[:avail] = np.fromfile(fname, dtype=np.uint8, count=avail, offset=off)
    return rate, raw.view(">i2" if be else "<i2").astype(np.int16)


def wav_layout(fname):
    """-> (sample_rate, data_byte_offset, n_samples) of a 16-bit mono PCM RIFF/WAVE file."""
    with open(fname, "rb") as f:
        head = f.read(12)
        if len(head) < 12 or head[:4] != b"RIFF" or head[8:12] != b"WAVE":
            raise ValueError("not a RIFF/WAVE file: " + fname)
        fmt = None
        while True:
            ck = f.read(8)
            if len(ck) < 8:
                raise ValueError("no data chunk in " + fname)
            cid, csz = ck[:4], int.from_bytes(ck[4:8], "little")
            if cid == b"fmt ":
                body = f.read(csz + (csz & 1))
                fmt = (int.from_bytes(body[0:2], "little"), int.from_bytes(body[2:4], "little"),
                       int.from_bytes(body[4:8], "little"), int.from_bytes(body[14:16], "little"))
            elif cid == b"data":
                if fmt is None:
                    raise ValueError("data chunk before fmt chunk in " + fname)
                tag, channels, rate, bits = fmt
                if tag not in (1, 0xFFFE) or channels != 1 or bits != 16:
                    raise NotImplementedError("vad_b200 ingests 16-bit mono PCM wav files only: " + fname)
                off = f.tell()
                avail = os.path.getsize(fname) - off
                return rate, off, min(csz, avail) // 2
            else:
                f.seek(csz + (csz & 1), 1)


def parse_stm(path, frame_rate):
    """dataset/stm_parser.py:5-26: TED-LIUM .stm segment bounds in samples.  Lines with fewer than 7 fields or
    labelled ignore_time_segment_in_scoring are skipped; seconds are float32 and the product is truncated to int32."""
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


def clip_ranges(starts, ends, n):
    """numpy slice semantics of ``data[start:end]`` (file_processing.py:91-92) for non-negative bounds."""
    s = np.clip(np.asarray(starts, dtype=np.int64), 0, n)
    e = np.clip(np.asarray(ends, dtype=np.int64), 0, n)
    return s, np.maximum(e, s)


# ------------------------------------------------------------------------------------------ device ingest
class DeviceIngest(object):
    """Raw files -> packed int16 PCM on the device.  ``add(path, transcription_path)`` registers a file;
    ``run()`` returns (d_pcm int16 tensor, offsets, lengths, sample_rates)."""

    def __init__(self, handle):
        self.handle = handle
        self.files = []          # (path, byte_offset, n_samples, big_endian, seg_starts, seg_ends, rate)

    def add(self, fname, transcription_path=None):
        if fname.endswith(".wav"):
            rate, off, n = wav_layout(fname)
            be = False
        elif fname.endswith(".sph"):
            rate, off, n, be = sph_layout(fname)
        else:
            raise ValueError("Wrong file format: " + str(fname))
        if transcription_path:
            s, e = parse_stm(transcription_path, rate)
            if np.any(s < 0) or np.any(e < 0):
                raise NotImplementedError("negative .stm bounds (python negative slicing) are not supported")
            s, e = clip_ranges(s, e, n)
        else:
            s, e = np.array([0], dtype=np.int64), np.array([n], dtype=np.int64)
        self.files.append((fname, off, n, be, s, e, rate))

    def run(self):
        h = self.handle
        raw_off, pos = [], 0
        for (_, _, n, _, _, _, _) in self.files:          # 2-byte aligned slot per file in the raw stream
            raw_off.append(pos)
            pos += 2 * n + (2 * n) % 4
        raw = torch.zeros(max(pos, 4), dtype=torch.uint8).pin_memory()   # zeros: a short file leaves a silent tail
        rawv = raw.numpy()
        lengths = np.zeros(len(self.files), dtype=np.int64)
        seg_src, seg_dst, seg_len, seg_be = [], [], [], []
        for i, (fname, off, n, be, s, e, _) in enumerate(self.files):
            with open(fname, "rb") as f:
                f.seek(off)
                f.readinto(memoryview(rawv[raw_off[i]:raw_off[i] + 2 * n]))   # reads what the file has
            lengths[i] = int((e - s).sum())
        padded = (lengths + ALIGN - 1) // ALIGN * ALIGN
        offsets = np.zeros(len(self.files), dtype=np.int64)
        if len(self.files) > 1:
            offsets[1:] = np.cumsum(padded[:-1])
        for i, (_, _, _, be, s, e, _) in enumerate(self.files):
            d = offsets[i]
            for a, b in zip(s, e):
                if b > a:
                    seg_src.append(raw_off[i] // 2 + int(a))
                    seg_dst.append(int(d))
                    seg_len.append(int(b - a))
                    seg_be.append(be)
                    d += int(b - a)
        total = int(padded.sum()) + ALIGN
        d_raw = raw.to(h.device, non_blocking=True)
        d_pcm = torch.zeros(total, dtype=torch.int16, device=h.device)
        seg_src, seg_dst, seg_len = (np.asarray(x, dtype=np.int64) for x in (seg_src, seg_dst, seg_len))
        seg_be = np.asarray(seg_be, dtype=bool)
        for be in (False, True):
            m = seg_be == be
            for lo in range(0, int(m.sum()), 65535):
                sl = slice(lo, lo + 65535)
                h.ingest_pcm(d_raw, seg_src[m][sl], seg_dst[m][sl], seg_len[m][sl], d_pcm, big_endian=be)
        rates = [f[6] for f in self.files]
        return d_pcm, offsets, lengths, rates
