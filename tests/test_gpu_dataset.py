"""GPU parity of the offline flow around the hot path (f1 file-list driver, f2 device ingest, f3 feature
sink, and row a8 scale_features), through the C ABI, against the oracle restatement (oracle/ref_io.py, ref_math.py)."""
import csv
import os

import numpy as np
import pytest

from oracle import ref_io, ref_math as rm
from vad_b200.synth import synth_utterance
from _parity import rows_close
from test_io_cpu import write_sph

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def h():
    import torch
    assert torch.cuda.is_available()
    from vad_b200 import runtime
    return runtime.default_handle()


def test_scale_rows_kernel_matches_numpy_float64(h):
    """dataset/utils.py:5-32 on a 30-file step: device statistics and rows vs the float64 oracle on the SAME rows."""
    import torch
    from vad_b200 import batch
    utts = [synth_utterance(61, i, 16000 * (2 + i % 5) + 13 * i) for i in range(30)]
    rows = torch.cat(batch.mfcc_batch(utts, deltas=True, handle=h))
    host = rows.cpu().numpy().astype(np.float64)
    want, stats = rm.scale_features([host])
    got_stats = h.scale_rows(rows)
    np.testing.assert_allclose(got_stats[:3], [s[0] for s in stats], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(got_stats[3:], [s[1] for s in stats], rtol=1e-12)
    got = rows.cpu().numpy()
    assert np.all(np.abs(got - want[0]) <= 1e-6 + 1e-6 * np.abs(want[0]))
    # a constant group has std 0: numpy gives nan (0/0) rows, so does the kernel
    z = torch.ones((7, 39), dtype=torch.float32, device=h.device)
    h.scale_rows(z, want_stats=False)
    assert bool(torch.isnan(z).all())
    assert h.scale_rows(torch.zeros((0, 39), dtype=torch.float32, device=h.device)) is not None


def test_ingest_kernel_decodes_and_gathers_bit_exact(h):
    import torch
    rng = np.random.default_rng(4)
    pcm = rng.integers(-32768, 32768, size=50000).astype(np.int16)
    for be in (False, True):
        raw = pcm.astype(">i2" if be else "<i2").tobytes()
        d_raw = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(h.device)
        src = np.array([0, 777, 40000, 123, 49999], dtype=np.int64)
        ln = np.array([100, 9001, 10000, 0, 1], dtype=np.int64)
        dst = np.concatenate([[8], 8 + np.cumsum(ln)[:-1]]).astype(np.int64)
        out = torch.full((int(ln.sum()) + 16,), 77, dtype=torch.int16, device=h.device)
        h.ingest_pcm(d_raw, src, dst, ln, out, big_endian=be)
        got = out.cpu().numpy()
        want = np.concatenate([pcm[s:s + n] for s, n in zip(src, ln)])
        assert np.array_equal(got[8:8 + len(want)], want)
        assert np.all(got[:8] == 77) and np.all(got[8 + len(want):] == 77)


def _make_corpus(root, n_files):
    from scipy.io import wavfile
    audio = os.path.join(root, "audio")
    stm = os.path.join(root, "stm")
    os.makedirs(audio)
    os.makedirs(stm)
    for i in range(n_files):
        data = synth_utterance(71, i, 16000 * (3 + i % 4) + 101 * i)
        name = "f%03d" % i
        if i % 3 == 0:
            write_sph(os.path.join(audio, name + ".sph"), data)
        else:
            wavfile.write(os.path.join(audio, name + ".wav"), 16000, data)
        dur = len(data) / 16000.0
        with open(os.path.join(stm, name + ".stm"), "w") as f:
            f.write("%s 1 spk 0.25 %.2f <o,f0,male> first segment\n" % (name, dur * 0.4))
            f.write("%s 1 spk %.2f %.2f <o> ignore_time_segment_in_scoring\n" % (name, dur * 0.4, dur * 0.5))
            f.write("%s 1 spk %.2f %.2f <o,f0,male> second segment\n" % (name, dur * 0.55, dur + 1.0))
    open(os.path.join(audio, "notes.txt"), "w").write("not audio")
    return audio, stm


def test_process_files_over_35_wav_sph_stm_files(h, tmp_path):
    """dataset_creator.process_files: 35 files (wav + NIST SPHERE, each with an .stm), max_files 33, steps of 30 ->
    CSV rows equal the oracle's scaled rows."""
    from vad_b200 import batch
    audio, stm = _make_corpus(str(tmp_path), 35)
    want, files = ref_io.process_files_rows([audio], 33, stm)
    assert len(files) == 33
    out = tmp_path / "rows.csv"
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(batch.create_table_header(13))
        n = batch.process_files([audio], 1, 33, w, transcription_dir=stm, handle=h, verbose=False)
    got = np.loadtxt(out, delimiter=",", skiprows=1)
    assert n == want.shape[0] == got.shape[0] and got.shape[1] == 40
    assert np.all(got[:, 39] == 1)
    assert rows_close(got[:, :39], want)
    # binary store sink, no transcriptions, odd step size
    want2, _ = ref_io.process_files_rows([audio], 1000, None, files_per_step=7)
    with batch.FeatureStore(str(tmp_path / "store")) as st:
        batch.process_files([audio], 0, 1000, st, handle=h, files_per_step=7, verbose=False)
    x, y = batch.load_feature_store(str(tmp_path / "store"))
    assert x.shape == want2.shape and not y.any() and rows_close(np.asarray(x, dtype=np.float64), want2)


def test_process_file_sph_with_stm_and_scale_features_in_place(h, tmp_path):
    from vad_b200 import batch, mfcc as vm
    audio, stm = _make_corpus(str(tmp_path), 4)
    fb = vm.get_mel_filterbanks(300, 8000, 512, 26, 16000)
    files = []
    for name in ("f000.sph", "f001.wav"):
        path = os.path.join(audio, name)
        tp = os.path.join(stm, name.split(".")[0] + ".stm")
        feats = batch.process_file([path, 400, 160, 512, fb, 13, None, tp])
        want = rm.dataset_features(rm.mfcc_utterance(ref_io.file_samples(path, tp)))
        got = np.array([np.concatenate(f) for f in feats])
        assert got.shape == want.shape and rows_close(got, want)
        files.append(feats)
    foreign = [(r[:13].copy(), r[13:26].copy(), r[26:].copy()) for r in files[1].rows[:50]]   # a plain python container
    mixed = [files[0], foreign]
    groups = [files[0].rows.copy(), files[1].rows[:50].copy()]
    want, _ = rm.scale_features(groups)
    keep = files[0][5][2]                                   # a view handed out before scaling
    out = batch.scale_features(mixed, handle=h)
    assert out is mixed
    assert np.all(np.abs(files[0].rows - want[0]) <= 1e-6 + 1e-6 * np.abs(want[0]))
    assert np.array_equal(keep, files[0].rows[5, 26:])      # in place: the old view sees the scaled values
    got_f = np.array([np.concatenate(f) for f in foreign])
    assert np.all(np.abs(got_f - want[1]) <= 1e-6 + 1e-6 * np.abs(want[1]))
