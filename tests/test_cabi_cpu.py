"""CPU: the C-ABI library loads and exports every symbol include/vadb200.h declares; host-only
helpers work; numeric entry points fail loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import ref_math as rm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from vad_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_all_exported_and_bound(lib):
    from vad_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "vadb200.h")).read()
    declared = set(re.findall(r"\b(vadb200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(lib, name), "library does not export " + name
        assert name in _lib.SIGNATURES, "python binding misses " + name
    assert set(_lib.SIGNATURES) == declared


def test_framing_helpers_match_oracle(lib):
    for n in [0, 1, 399, 400, 401, 560, 561, 1040, 1041, 1840, 160000, 960000, 7777]:
        assert lib.vadb200_frames_for_length(n) == rm.n_frames(n)
        assert lib.vadb200_outputs_for_length(n) == rm.n_outputs(n)


def test_default_config_is_reference_config(lib):
    from vad_b200 import _lib, config
    cfg = _lib.Config()
    lib.vadb200_default_config(C.byref(cfg))
    assert (cfg.sample_rate, cfg.frame_size, cfg.frame_step, cfg.fft_n, cfg.n_filters, cfg.n_mfcc) == \
        (config.SAMPLERATE, config.FRAME_SIZE, config.FRAME_STEP, config.FFT_N, config.FILTERBANKS_NUM, config.MFCC_NUM)
    assert (cfg.low_hz, cfg.high_hz, cfg.lifter_l) == (300.0, 8000.0, 22)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vad_b200 import _lib
    h = C.c_void_p()
    cfg = _lib.Config()
    lib.vadb200_default_config(C.byref(cfg))
    rc = lib.vadb200_create(C.byref(cfg), 0, C.byref(h))
    assert rc == -3 and not h.value
    assert b"CUDA" in lib.vadb200_last_error() or b"cuda" in lib.vadb200_last_error()
    from vad_b200 import runtime
    with pytest.raises(RuntimeError):
        runtime.Handle()


def test_unsupported_config_rejected(lib):
    from vad_b200 import _lib
    h = C.c_void_p()
    cfg = _lib.Config()
    lib.vadb200_default_config(C.byref(cfg))
    cfg.n_filters = 40
    assert lib.vadb200_create(C.byref(cfg), 0, C.byref(h)) == -2
    assert b"reference configuration" in lib.vadb200_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "vad_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_mfcc_table_constructors_match_reference_golden(golden_dir):
    # host-side init tables of the drop-in module (no GPU needed)
    from vad_b200 import mfcc as vm
    kat = np.load(os.path.join(golden_dir, "kat_frame.npz"))
    np.testing.assert_array_equal(np.array(vm.mel_from_hz(300, 8000, 26)), kat["mel_points"])
    np.testing.assert_array_equal(
        np.array(vm.convert_to_fft_bins(16000, vm.hz_from_mel(vm.mel_from_hz(300, 8000, 26)), 512)), kat["bins"])
    np.testing.assert_array_equal(vm.get_mel_filterbanks(300, 8000, 512, 26, 16000), kat["filterbank"])


def test_pack_utterances_alignment():
    from vad_b200.batch import pack_utterances, split_into_frames
    utts = [np.arange(n, dtype=np.int16) for n in (401, 7, 1000, 0, 4001)]
    flat, off, ln = pack_utterances(utts)
    assert np.all(off % 8 == 0) and list(ln) == [401, 7, 1000, 0, 4001]
    for u, o, n in zip(utts, off, ln):
        np.testing.assert_array_equal(flat.numpy()[o:o + n], u)
    assert len(split_into_frames(utts[2], 400, 160)) == rm.n_frames(1000)
    with pytest.raises(Exception):
        split_into_frames(utts[2], 400, 160, transcription_path="x.stm", frame_rate=None)


def test_synth_matches_itself_and_is_int16():
    from vad_b200.synth import synth_utterance
    a = synth_utterance(1234, 5, 4096)
    b = synth_utterance(1234, 5, 2048, start=2048)
    assert a.dtype == np.int16 and np.array_equal(a[2048:], b)
    big = synth_utterance(1234, 5, 160000).astype(np.float64)
    assert 500 < big.std() < 5000 and np.abs(big).max() <= 32767


def test_shard_balanced_property():
    from hypothesis import given, settings, strategies as st
    from vad_b200 import shard

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.integers(min_value=0, max_value=500000), min_size=1, max_size=60),
           st.integers(min_value=1, max_value=8))
    def prop(lengths, world):
        parts = shard.shard_balanced(lengths, world)
        assert len(parts) == world
        allidx = np.concatenate(parts) if parts else np.array([])
        assert sorted(allidx.tolist()) == list(range(len(lengths)))
        for p in parts:
            assert list(p) == sorted(p)
        cost = np.maximum((np.array(lengths) - 401) // 160 + 1, 0)
        loads = np.array([cost[p].sum() for p in parts])
        assert loads.max() - loads.min() <= max(int(cost.max()), 1)
        lo_hi = [shard.shard_contiguous(len(lengths), r, world) for r in range(world)]
        assert lo_hi[0][0] == 0 and lo_hi[-1][1] == len(lengths)
        assert all(a[1] == b[0] for a, b in zip(lo_hi, lo_hi[1:]))

    prop()


def test_framing_rule_property():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.integers(min_value=0, max_value=2000000))
    def prop(n):
        t = rm.n_frames(n)
        # reference loop (dataset/file_processing.py:99-101)
        if n <= 5000:
            k, off = 0, 0
            while n - off > 400:
                k += 1
                off += 160
            assert t == k
        assert (t == 0) == (n <= 400)
        if t:
            assert 160 * (t - 1) + 400 < n <= 160 * t + 400
        assert rm.n_outputs(n) == max(t - 5, 0)

    prop()


def test_csv_sink_layout(tmp_path):
    import csv
    from vad_b200 import batch
    hdr = batch.create_table_header(13)
    assert len(hdr) == 40 and hdr[0] == "MFCC Coef1" and hdr[13] == "First delta1" and hdr[26] == "Second delta1" \
        and hdr[-1] == "voiced"
    feats = [[(np.arange(13.0), np.arange(13.0) + 100, np.arange(13.0) + 200)] * 2]
    with open(tmp_path / "f.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(hdr)
        batch.write_features(w, feats, 1)
    rows = list(csv.reader(open(tmp_path / "f.csv")))
    assert len(rows) == 3 and len(rows[1]) == 40 and float(rows[1][13]) == 100.0 and float(rows[1][-1]) == 1.0
