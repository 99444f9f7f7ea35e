"""CPU: host-side ingest logic (container layouts, .stm parsing, feature sinks) against the oracle restatement of
dataset/sph.py, dataset/stm_parser.py and file_processing.py's writers.  No CUDA call is made here."""
import csv
import os

import numpy as np
import pytest

from oracle import ref_io
from vad_b200.synth import synth_utterance

SPH_HEAD = ("NIST_1A\n   1024\nsample_count -i %d\nsample_n_bytes -i 2\nchannel_count -i 1\n"
            "sample_byte_format -s2 10\nsample_rate -i 16000\nsample_coding -s3 pcm\nend_head\n")


def write_sph(path, data, n_declared=None):
    head = (SPH_HEAD % (len(data) if n_declared is None else n_declared)).encode()
    with open(path, "wb") as f:
        f.write(head + b" " * (1024 - len(head)) + np.asarray(data, dtype=">i2").tobytes())


def test_sph_layout_reproduces_the_reference_reader(tmp_path):
    from vad_b200.io import read_sph, sph_layout
    data = synth_utterance(3, 1, 5000)
    p = str(tmp_path / "a.sph")
    write_sph(p, data)
    _, rate, width, want = ref_io.read_sph(p)
    got_rate, got = read_sph(p)                         # default: the reference's byte stream
    assert (rate, width) == (16000, 2) and got_rate == rate
    assert np.array_equal(got, want)
    # the reference starts decoding right after the ninth header line, i.e. inside the 1024-byte header
    r, off, n, be = sph_layout(p)
    assert off < 1024 and n == len(data) and be
    assert np.array_equal(got[(1024 - off) // 2:], data[: len(data) - (1024 - off) // 2])
    # conforming parse
    r2, got2 = read_sph(p, reference_compat=False)
    assert r2 == 16000 and np.array_equal(got2, data)
    # a file shorter than its declared sample_count: zeros at the end, as np.zeros in the reference
    q = str(tmp_path / "short.sph")
    write_sph(q, data[:1000], n_declared=3000)
    _, _, _, want_q = ref_io.read_sph(q)
    assert np.array_equal(read_sph(q)[1], want_q) and len(want_q) == 3000 and not want_q[-100:].any()
    (tmp_path / "bad.sph").write_bytes(b"RIFFxxxx")
    with pytest.raises(ValueError):
        read_sph(str(tmp_path / "bad.sph"))


def test_wav_layout_matches_scipy(tmp_path):
    from scipy.io import wavfile
    from vad_b200.io import wav_layout
    data = synth_utterance(3, 2, 7001)
    p = str(tmp_path / "a.wav")
    wavfile.write(p, 16000, data)
    rate, off, n = wav_layout(p)
    assert (rate, n) == (16000, len(data))
    assert np.array_equal(np.fromfile(p, dtype="<i2", count=n, offset=off), data)
    wavfile.write(str(tmp_path / "st.wav"), 16000, np.stack([data, data], axis=1))
    with pytest.raises(NotImplementedError):
        wav_layout(str(tmp_path / "st.wav"))
    (tmp_path / "bad.wav").write_bytes(b"NIST_1A\n")
    with pytest.raises(ValueError):
        wav_layout(str(tmp_path / "bad.wav"))


def test_stm_bounds_match_reference_parser(tmp_path):
    from vad_b200.io import parse_stm, clip_ranges
    stm = tmp_path / "t.stm"
    stm.write_text("a 1 spk 0.50 1.75 <o,f0,male> hello world\n"
                   "a 1 spk 1.75 2.00 <o> ignore_time_segment_in_scoring\n"
                   "short line\n"
                   "a 1 spk 3.10 5.20 <o,f0,male> more words\n"
                   "a 1 spk 5.333 99.0 <o,f0,male> beyond the end\n")
    s, e = parse_stm(str(stm), 16000)
    ws, we = ref_io.stm_sample_indices(str(stm), 16000)
    assert np.array_equal(s, ws) and np.array_equal(e, we) and s.dtype == np.int32
    assert list(s[:2]) == [8000, 49600] and list(e[:2]) == [28000, 83200]
    cs, ce = clip_ranges(s, e, 96000)
    data = np.arange(96000)
    for a, b, c, d in zip(s, e, cs, ce):
        assert np.array_equal(data[a:b], data[c:d])      # numpy slice clipping


def test_feature_store_and_csv_sink(tmp_path):
    from vad_b200 import batch
    rows = np.random.default_rng(0).standard_normal((50, 39))
    feats = batch.FileFeatures(rows)
    assert len(feats) == 50 and feats[3][1].shape == (13,) and feats[3][1].base is not None
    feats.rows[3, 13] = 42.0                              # the triples are views of .rows
    assert feats[3][1][0] == 42.0
    with batch.FeatureStore(str(tmp_path / "fs")) as st:
        batch.write_features(st, [feats, [(np.zeros(13), np.ones(13), np.full(13, 2.0))]], 1)
        batch.write_feature_rows(st, rows[:4], 0)
    x, y = batch.load_feature_store(str(tmp_path / "fs"))
    assert x.shape == (55, 39) and y.shape == (55,) and list(y[-4:]) == [0, 0, 0, 0] and y[0] == 1
    np.testing.assert_allclose(x[:50], feats.rows.astype(np.float32))
    assert x[50, 13] == 1.0 and x[50, 26] == 2.0
    with open(tmp_path / "f.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(batch.create_table_header(13))
        batch.write_features(w, [feats], 2)
    lines = list(csv.reader(open(tmp_path / "f.csv")))
    assert len(lines) == 51 and len(lines[1]) == 40 and float(lines[4][13]) == 42.0 and float(lines[1][-1]) == 2.0


def test_list_audio_files_order_and_cap(tmp_path):
    from vad_b200 import batch
    d = tmp_path / "d"
    d.mkdir()
    for n in ("b.wav", "a.sph", "c.txt", "d.wav"):
        (d / n).write_bytes(b"")
    got = batch.list_audio_files([str(d)], 2)
    want = [str(d) + "/" + f for f in os.listdir(str(d)) if f.endswith(".wav") or f.endswith(".sph")][:2]
    assert got == want and len(got) == 2
