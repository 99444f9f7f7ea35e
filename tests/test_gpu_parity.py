"""GPU parity: the CUDA path, called through the C ABI (vad_b200.runtime -> libvadb200.so),
against the float64 oracle and the golden vectors frozen from the unmodified reference.

Tolerances (north_star / SURVEY.md 8d): MFCC |d| <= 1e-4 + 1e-4 |ref|; logits
|d| <= 1e-3 + 1e-3 |ref|; labels identical except frames whose oracle top-2 logit margin is
below twice the logit tolerance; rows with sigma5 == 0 are NaN / label 0 on both sides."""
import os

import numpy as np
import pytest

from oracle import ref_math as rm, ref_loop
from vad_b200.synth import synth_utterance

pytestmark = pytest.mark.gpu

from _parity import (MFCC_ATOL, MFCC_RTOL, LOGIT_ATOL, LOGIT_RTOL, ROW_ATOL, ROW_RTOL, mfcc_close, rows_close,
                     check_vad)

CASES = ["synth_1p5s", "synth_ragged", "exact_fit", "too_short", "silence_dc", "tone_noise"]


@pytest.fixture(scope="module", params=["tc16", "tc", "fp32"])
def env(request):
    """Every test runs under all three FFN implementations: tcgen05 fp16 hi/lo (default), tcgen05 tf32 hi/lo and
    FP32 CUDA cores."""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from vad_b200 import runtime
    w = rm.glorot_ffn(0)
    h = runtime.default_handle()
    h.set_ffn_weights(w)
    h.set_ffn_impl(request.param)
    assert h.ffn_impl == {"fp32": 0, "tc": 1, "tc16": 2}[request.param]
    yield h, w
    h.set_ffn_impl("tc16")


@pytest.fixture(scope="module")
def utts(golden_dir):
    return np.load(os.path.join(golden_dir, "utterances.npz"))


# ---- per-frame drop-in API (mfcc.py) ---------------------------------------------------------
def test_kat_frame_dropin_api(env, golden_dir):
    from vad_b200 import mfcc as vm
    kat = np.load(os.path.join(golden_dir, "kat_frame.npz"))
    fb = vm.get_mel_filterbanks(300, 8000, 512, 26, 16000)
    assert np.array_equal(fb, kat["filterbank"]) and np.array_equal(env[0].filterbank(), fb)
    spec = vm.get_spec_mag(kat["frame"], 512)
    assert spec.shape == (256,) and np.max(np.abs(spec - kat["spec"])) <= 2e-6 * kat["spec"].max()
    out = vm.get_mfcc(kat["frame"], 512, fb, 13)
    assert out.shape == (13,) and out.dtype == np.float64 and mfcc_close(out, kat["mfcc"])
    assert mfcc_close(vm.get_mfcc_from_spec(kat["spec"], fb, 13), kat["mfcc_from_spec"])
    assert mfcc_close(vm.get_mfcc(np.zeros(400, np.int16), 512, fb, 13), kat["mfcc_zero"])
    np.testing.assert_allclose(vm.lifter(np.ones(13)), kat["lifter13"], rtol=2e-6)
    np.testing.assert_allclose(vm.get_deltas(np.arange(13.0), np.ones(13)), np.arange(13.0) - 1)
    with pytest.raises(NotImplementedError):
        vm.get_mfcc(kat["frame"], 1024, fb, 13)
    with pytest.raises(NotImplementedError):
        vm.get_mfcc(kat["frame"], 512, fb * 2, 13)


def test_float_frames_short_and_noninteger(env):
    h, _ = env
    rng = np.random.default_rng(2)
    for n in (400, 320, 512, 1):
        fr = (rng.standard_normal((5, n)) * 1234.5).astype(np.float32)
        got = h.spec_frames(fr).cpu().numpy()
        ref = rm.get_spec_mag(fr)
        assert np.max(np.abs(got - ref)) <= 3e-6 * ref.max()
        assert mfcc_close(h.mfcc_frames(fr).cpu().numpy(), rm.get_mfcc(fr, rm.get_mel_filterbanks()))


# ---- packed batches ---------------------------------------------------------------------------
@pytest.mark.parametrize("name", CASES)
def test_golden_utterances_all_modes(env, utts, name):
    import torch
    from vad_b200 import batch
    h, w = env
    pcm = utts[name + "/pcm"]
    mf = batch.mfcc_batch([pcm], handle=h)[0].cpu().numpy()
    assert mf.shape == utts[name + "/mfcc"].shape
    assert mfcc_close(mf, utts[name + "/mfcc"]) and mfcc_close(mf, rm.mfcc_utterance(pcm))
    ds = batch.mfcc_batch([pcm], deltas=True, handle=h)[0].cpu().numpy()
    ref_ds = utts[name + "/dataset_rows"]
    assert ds.shape == ref_ds.shape
    assert rows_close(ds, ref_ds)
    labels, logits = batch.vad_batch([pcm], handle=h, want_logits=True)
    check_vad(labels[0].cpu().numpy(), logits[0].cpu().numpy(), pcm, w)


def test_analyser_features_vs_reference_golden(env, utts):
    import torch
    from vad_b200 import batch, runtime
    h, w = env
    for name in ("synth_1p5s", "silence_dc", "tone_noise"):
        pcm = utts[name + "/pcm"]
        flat, off, ln = batch.pack_utterances([pcm])
        plan = runtime.Plan(h, off, ln, runtime.MODE_VAD)
        _, _, feats = plan.vad(flat.to(h.device), want_logits=True, want_feats=True)
        feats = feats.cpu().numpy()
        ref = utts[name + "/analyser_rows"]
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(feats), fin)
        c = utts[name + "/mfcc"]
        sig = np.lib.stride_tricks.sliding_window_view(c, 5, axis=0)[: c.shape[0] - 5].std(axis=2)
        tol = 3e-4 + 2e-4 / np.maximum(np.tile(sig, 3), 1e-12)   # z = (c - mu) / sigma5 amplifies MFCC error
        assert np.all(np.abs(feats - ref)[fin] <= tol[fin])


def test_ragged_batch_matches_per_utterance_oracle(env):
    from vad_b200 import batch
    h, w = env
    rng = np.random.default_rng(7)
    lens = [0, 5, 400, 401, 1040, 1041, 1201, 16000, 32000 + 3, 48017, 4000, 561, 7, 25000]
    utts = [synth_utterance(7, i, n) for i, n in enumerate(lens)]
    mf = batch.mfcc_batch(utts, handle=h)
    labels, logits = batch.vad_batch(utts, handle=h, want_logits=True)
    for u, m, la, lo in zip(utts, mf, labels, logits):
        assert m.shape[0] == rm.n_frames(len(u)) and la.shape[0] == rm.n_outputs(len(u))
        if m.shape[0]:
            assert mfcc_close(m.cpu().numpy(), rm.mfcc_utterance(u))
        check_vad(la.cpu().numpy(), lo.cpu().numpy(), u, w)


def test_cfg1_single_60s_waveform(env):
    """configs[0]: one 60 s waveform (5,998 frames, 5,993 decisions) split over many segments."""
    from vad_b200 import batch
    h, w = env
    pcm = synth_utterance(1234, 0, 960000)
    pcm[200000:216000] = 0                      # 1 s digital silence: eps branch, sigma5 == 0
    pcm[300000:316000] = 1000                   # DC segment
    labels, logits = batch.vad_batch([pcm], handle=h, want_logits=True)
    assert labels[0].shape[0] == 5993
    check_vad(labels[0].cpu().numpy(), logits[0].cpu().numpy(), pcm, w)
    mf = batch.mfcc_batch([pcm], handle=h)[0].cpu().numpy()
    assert mf.shape == (5998, 13) and mfcc_close(mf, rm.mfcc_utterance(pcm))


def test_cfg2_shape_batch_properties(env):
    """configs[1] shape (10 s utterances) on a 256-utterance slice: sampled oracle parity plus
    size-independent properties -- batch position must not matter, duplicates give equal rows."""
    import torch
    from vad_b200 import batch, runtime
    h, w = env
    n_utt, L = 256, 160000
    off, ln, stride = batch.uniform_layout(n_utt, L)
    pcm = h.synth_pcm(n_utt, L, seed=42, first_utt=0, utt_stride=stride)
    body = pcm[: n_utt * stride].view(n_utt, stride)
    body[200] = body[3]                                                  # duplicate utterance
    plan = runtime.Plan(h, off, ln, runtime.MODE_MFCC)
    out = plan.mfcc(pcm).view(n_utt, 998, 13)
    assert torch.equal(out[200], out[3])
    host = body.cpu().numpy()
    for u in (0, 3, 77, 255):
        assert np.array_equal(host[u, :L], synth_utterance(42, 3 if u == 200 else u, L))   # synth parity
        assert mfcc_close(out[u].cpu().numpy(), rm.mfcc_utterance(host[u, :L]))
    # same utterances alone, in another order -> bit-identical rows
    sub = [host[77, :L], host[0, :L]]
    alone = batch.mfcc_batch(sub, handle=h)
    assert torch.equal(alone[0], out[77]) and torch.equal(alone[1], out[0])
    vplan = runtime.Plan(h, off, ln, runtime.MODE_VAD)
    labels, logits, _ = vplan.vad(pcm, want_logits=True)
    labels = labels.view(n_utt, 993)
    assert torch.equal(labels[200], labels[3])
    check_vad(labels[77].cpu().numpy(), logits.view(n_utt, 993, 3)[77].cpu().numpy(), host[77, :L], w)
    # labels-only launch (the bench configuration) is bit-identical to the logits-returning one
    labels2, _, _ = vplan.vad(pcm)
    assert torch.equal(labels2.view(n_utt, 993), labels)


def test_unaligned_buffer_tail(env):
    """pcm_len not a multiple of 8 samples: the last partial 16-byte chunk is fetched by plain
    loads instead of the TMA bulk copy."""
    import torch
    from vad_b200 import runtime
    h, w = env
    for L in (4003, 4001, 5367, 10727):
        u = synth_utterance(9, L, L)
        pcm = torch.from_numpy(u.copy()).to(h.device)
        plan = runtime.Plan(h, np.array([0]), np.array([L]), runtime.MODE_MFCC)
        out = plan.mfcc(pcm).cpu().numpy()
        assert mfcc_close(out, rm.mfcc_utterance(u))


def test_host_pipeline_equals_device_path(env):
    import torch
    from vad_b200 import batch, runtime
    h, w = env
    utts = [synth_utterance(11, i, 16000 * (1 + i % 4) + 37 * i) for i in range(40)]
    flat, off, ln = batch.pack_utterances(utts, pin=True)
    plan = runtime.Plan(h, off, ln, runtime.MODE_VAD)
    dev_labels, dev_logits, _ = plan.vad(flat.to(h.device), want_logits=True)
    h.set_host_chunk_samples(100000)            # force several chunks through the 3-deep pipeline
    try:
        logits_host = torch.empty((plan.total_rows, 3), dtype=torch.float32).pin_memory()
        host_labels, _ = plan.vad_host(flat, logits_host=logits_host)
    finally:
        h.set_host_chunk_samples(32 << 20)
    assert torch.equal(host_labels, dev_labels.cpu())
    assert torch.equal(torch.nan_to_num(logits_host), torch.nan_to_num(dev_logits.cpu()))


# ---- tensor-core FFN (tcgen05, tf32 x3) ----------------------------------------------------------
@pytest.fixture(params=["tc16", "tc"])
def tc(env, request):
    h, w = env
    prev = h.ffn_impl
    h.set_ffn_impl(request.param)
    yield h, w
    h.set_ffn_impl(prev)


def test_tc_ffn_rows_vs_oracle(tc):
    h, w = tc
    assert h.ffn_impl in (1, 2)
    rng = np.random.default_rng(5)
    for n in (1, 127, 128, 129, 1000):
        x = (rng.standard_normal((n, 39)) * np.array([1.0] * 13 + [3.0] * 13 + [30.0] * 13)).astype(np.float32)
        x[0, :] = 0.0
        if n > 5:
            x[5, 7] = np.nan
        labels, logits = h.ffn_predict(x)
        labels, logits = labels.cpu().numpy(), logits.cpu().numpy()
        ref_logits, _ = rm.ffn_forward(x, w)
        fin = np.isfinite(x).all(axis=1)
        assert np.array_equal(np.isfinite(logits).all(axis=1), fin) and np.all(labels[~fin] == 0)
        err = np.abs(logits[fin] - ref_logits[fin])
        assert np.all(err <= LOGIT_ATOL + LOGIT_RTOL * np.abs(ref_logits[fin])), err.max()
        srt = np.sort(ref_logits[fin], axis=1)
        dec = (srt[:, -1] - srt[:, -2]) > 2 * (LOGIT_ATOL + LOGIT_RTOL * np.abs(srt[:, -1]))
        assert np.array_equal(labels[fin][dec], rm.decide(ref_logits)[fin][dec])


@pytest.mark.parametrize("name", CASES)
def test_tc_fused_golden_utterances(tc, utts, name):
    from vad_b200 import batch
    h, w = tc
    pcm = utts[name + "/pcm"]
    labels, logits = batch.vad_batch([pcm], handle=h, want_logits=True)
    check_vad(labels[0].cpu().numpy(), logits[0].cpu().numpy(), pcm, w)


def test_tc_fused_ragged_and_60s(tc):
    from vad_b200 import batch
    h, w = tc
    lens = [0, 401, 1041, 1201, 16000, 32003, 48017, 561, 25000, 160000]
    utts = [synth_utterance(7, i, n) for i, n in enumerate(lens)]
    labels, logits = batch.vad_batch(utts, handle=h, want_logits=True)
    for u, la, lo in zip(utts, labels, logits):
        check_vad(la.cpu().numpy(), lo.cpu().numpy(), u, w)
    pcm = synth_utterance(1234, 0, 960000)
    pcm[200000:216000] = 0
    labels, logits = batch.vad_batch([pcm], handle=h, want_logits=True)
    check_vad(labels[0].cpu().numpy(), logits[0].cpu().numpy(), pcm, w)


def test_tc_and_fp32_paths_agree_on_batch(env):
    import torch
    from vad_b200 import batch, runtime
    h, w = env
    n_utt, L = 128, 160000
    off, ln, stride = batch.uniform_layout(n_utt, L)
    pcm = h.synth_pcm(n_utt, L, seed=99, first_utt=0, utt_stride=stride)
    plan = runtime.Plan(h, off, ln, runtime.MODE_VAD)
    prev = h.ffn_impl
    h.set_ffn_impl("fp32")
    la0, lo0, _ = plan.vad(pcm, want_logits=True)
    try:
        for impl in ("tc", "tc16"):
            h.set_ffn_impl(impl)
            la1, lo1, _ = plan.vad(pcm, want_logits=True)
            la2, _, _ = plan.vad(pcm)
            assert torch.equal(la1, la2)
            d = (lo0 - lo1).abs()
            fin = torch.isfinite(lo0).all(dim=1)
            assert torch.equal(torch.isfinite(lo1).all(dim=1), fin)
            assert float(d[fin].max()) < 5e-4
            assert float((la0 != la1).float().mean()) < 1e-3
    finally:
        h.set_ffn_impl(prev)


# ---- analyser / streaming ------------------------------------------------------------------------
def test_fused_analyser_feed_frame_contract(env):
    from vad_b200.analyser import FusedAnalyser, FFNClassifier
    h, w = env
    pcm = synth_utterance(21, 0, 16000)
    frames = rm.split_into_frames(pcm)
    an = FusedAnalyser(handle=h)
    ref = ref_loop.LoopAnalyser(ref_loop.FFNClassifier(w))
    with pytest.raises(ValueError):
        an.load_init_inactive_frames(list(frames[:4]))
    an.load_init_inactive_frames([f.astype(np.float32) for f in frames[:5]])
    _, feats, ref_logits, _ = rm.vad_utterance(pcm, w)
    srt = np.sort(ref_logits, axis=1)
    decisive = (srt[:, -1] - srt[:, -2]) > 2 * (LOGIT_ATOL + LOGIT_RTOL * np.abs(srt[:, -1]))
    for i, fr in enumerate(frames):
        f32 = fr.astype(np.float32)
        got, exp = an.feed_frame(f32), ref.feed_frame(f32)
        if i < 5:
            assert got is None and exp is None
        elif decisive[i - 5]:
            assert (got is None) == (exp is None)
            if got is not None:
                assert np.array_equal(got, frames[i - 3].astype(np.float32))
    # external classifier plug point: receives the reference's (1, 39) float64 row
    seen = []

    class Stub(object):
        def predict(self, x):
            seen.append(x.copy())
            return np.array([2])

    an2 = FusedAnalyser(classifier=Stub(), handle=h)
    for fr in frames[:5]:
        assert an2.feed_frame(fr.astype(np.float32)) is None
    with pytest.raises(AssertionError):
        an2.feed_frame(frames[5].astype(np.float32))
    assert seen[0].shape == (1, 39) and seen[0].dtype == np.float64
    assert np.all(np.abs(seen[0][0] - feats[0]) <= 1e-3 + 1e-3 * np.abs(feats[0]))
    clf = FFNClassifier(handle=h)
    _, lg = rm.ffn_forward(feats, w)
    assert np.array_equal(clf.predict(feats)[decisive], rm.decide(rm.ffn_forward(feats, w)[0])[decisive])


def test_stream_bank_equals_offline_batch(env):
    from vad_b200 import batch
    from vad_b200.analyser import StreamBank
    h, w = env
    n_streams, n_chunks = 70, 60                    # not a multiple of 32 streams
    L = 160 * n_chunks
    utts = [synth_utterance(31, s, L) for s in range(n_streams)]
    bank = StreamBank(n_streams, handle=h)
    got = np.full((n_streams, n_chunks), 255, np.uint8)
    lg = np.zeros((n_streams, n_chunks, 3), np.float32)
    for j in range(n_chunks):
        chunk = np.stack([u[160 * j:160 * (j + 1)] for u in utts])
        got[:, j], lg[:, j] = bank.feed(chunk, want_logits=True)
    # chunk j completes frame j-2 and (reference timing) emits the decision for frame j-5
    assert np.all(got[:, :7] == 255)
    for s in (0, 1, 33, 69):
        c, feats, ref_logits, ref_labels = rm.vad_utterance(utts[s], w)
        n_avail = n_chunks - 7                        # decisions for frames 2 .. n_chunks-6
        rows = np.arange(n_avail)
        fin = np.isfinite(feats[rows]).all(axis=1)
        err = np.abs(lg[s, 7:7 + n_avail][fin] - ref_logits[rows][fin])
        assert np.all(err <= LOGIT_ATOL + LOGIT_RTOL * np.abs(ref_logits[rows][fin]))
    # and bit-identical to the offline fused kernel on the same samples
    off_labels = batch.vad_batch(utts, handle=h)
    for s in range(n_streams):
        k = min(n_chunks - 7, off_labels[s].shape[0])
        assert np.array_equal(got[s, 7:7 + k], off_labels[s].cpu().numpy()[:k])


def test_error_codes(env):
    import torch
    from vad_b200 import runtime
    from vad_b200._lib import VadB200Error
    h, _ = env
    with pytest.raises(VadB200Error):
        runtime.Plan(h, np.array([0, 4]), np.array([100, 100]), runtime.MODE_VAD)      # offset not multiple of 8
    with pytest.raises(VadB200Error):
        runtime.Plan(h, np.array([0, 8]), np.array([100, 100]), runtime.MODE_VAD)      # overlapping
    plan = runtime.Plan(h, np.array([0]), np.array([1000]), runtime.MODE_VAD)
    with pytest.raises(VadB200Error):
        plan.vad(torch.zeros(500, dtype=torch.int16, device=h.device))                  # buffer too short
    with pytest.raises(VadB200Error):
        plan.mfcc(torch.zeros(1000, dtype=torch.int16, device=h.device))                # wrong mode
    empty = runtime.Plan(h, np.array([], dtype=np.int64), np.array([], dtype=np.int64), runtime.MODE_VAD)
    assert empty.total_rows == 0


# ---- drop-ins around the hot path -------------------------------------------------------------------
def test_process_file_and_scale_features_dropins(env, utts, tmp_path):
    from scipy.io import wavfile
    from vad_b200 import batch, mfcc as vm

    class Q(object):
        def __init__(self):
            self.v = 4

        def get(self):
            return self.v

        def put(self, v):
            self.v = v

    fb = vm.get_mel_filterbanks(300, 8000, 512, 26, 16000)
    files = []
    for i, name in enumerate(("synth_ragged", "tone_noise")):
        path = str(tmp_path / ("u%d.wav" % i))
        wavfile.write(path, 16000, utts[name + "/pcm"])
        q = Q()
        feats = batch.process_file([path, 400, 160, 512, fb, 13, q, None])
        assert q.v == 5
        ref = utts[name + "/dataset_rows"]
        got = np.array([np.concatenate(f) for f in feats])
        assert got.shape == ref.shape and rows_close(got, ref)
        files.append(feats)
    with pytest.raises(ValueError):
        batch.process_file(["x.mp3", 400, 160, 512, fb, 13, Q(), None])
    with pytest.raises(NotImplementedError):
        batch.process_file([str(tmp_path / "u0.wav"), 256, 160, 512, fb, 13, Q(), None])
    groups = [np.array([np.concatenate(f) for f in ff]) for ff in files]
    want, _ = rm.scale_features(groups)
    out = batch.scale_features(files)
    assert out is files
    for ff, w in zip(files, want):
        got = np.array([np.concatenate(f) for f in ff])
        assert np.all(np.abs(got - w) <= 1e-6 + 1e-6 * np.abs(w))   # float32 rows, float64 statistics (vadb200_scale_rows)


def test_stm_segments_gather(env, tmp_path):
    from vad_b200 import batch
    h, _ = env
    pcm = synth_utterance(77, 2, 16000 * 6)
    stm = tmp_path / "t.stm"
    stm.write_text("a 1 spk 0.50 1.75 <o,f0,male> hello world\n"
                   "a 1 spk 1.75 2.00 <o> ignore_time_segment_in_scoring\n"
                   "short line\n"
                   "a 1 spk 3.10 5.20 <o,f0,male> more words\n")
    starts, ends = batch.parse_transcription(str(stm), 16000)
    assert list(starts) == [8000, 49600] and list(ends) == [28000, 83200]
    frames = batch.split_into_frames(pcm, 400, 160, str(stm), 16000)
    glued = np.concatenate([pcm[8000:28000], pcm[49600:83200]])
    assert len(frames) == rm.n_frames(len(glued)) and np.array_equal(frames[3], glued[480:880])
    rows = batch.mfcc_batch([glued], deltas=True, handle=h)[0].cpu().numpy()
    want = rm.dataset_features(rm.mfcc_utterance(glued))
    assert rows_close(rows, want)


def test_every_row_of_a_ragged_batch_mfcc_and_dataset_modes(env):
    """All rows (not a sample) of a ragged batch in both row modes against the oracle: utterances from below one step to
    several block phases long, so the flushes start at every alignment of the output pointer within 16 bytes (rows are
    13 / 39 floats) and the staged dataset flush runs with 1 .. 256 valid centres."""
    from vad_b200 import batch
    h, _ = env
    lens = [400 + 160 * k for k in (5, 6, 7, 8, 31, 32, 33, 34, 127, 130, 255, 258, 259, 300, 611)] + [401, 40123, 99999]
    utts_ = [synth_utterance(29, i, n) for i, n in enumerate(lens)]
    rows39 = batch.mfcc_batch(utts_, deltas=True, handle=h)
    rows13 = batch.mfcc_batch(utts_, deltas=False, handle=h)
    starts = set()
    acc = 0
    for u, r39, r13 in zip(utts_, rows39, rows13):
        c = rm.mfcc_utterance(u)
        want = rm.dataset_features(c)
        got = r39.cpu().numpy()
        assert got.shape == want.shape, (len(u), got.shape, want.shape)
        if want.shape[0]:
            assert rows_close(got, want), (len(u), float(np.abs(got - want).max()))
        g13 = r13.cpu().numpy()
        assert g13.shape == c.shape and mfcc_close(g13, c), len(u)
        starts.add((acc * 39 * 4) % 16)
        acc += want.shape[0]
    assert starts == {0, 4, 8, 12}   # the batch really covers every start alignment of the dataset rows


def test_two_handles_with_different_weights(env):
    """Constant-bank ownership switches between handles (and FFN implementations) without mixing weights."""
    import torch
    from vad_b200 import batch, runtime
    h, w = env
    w2 = rm.glorot_ffn(123)
    h2 = runtime.Handle(h.device.index, ffn_weights=w2)
    h2.set_ffn_impl(h.ffn_impl)
    utts_ = [synth_utterance(3, i, 16000 + 500 * i) for i in range(6)]
    for _ in range(2):                                   # alternate owners twice
        for hh, ww in ((h, w), (h2, w2)):
            labels, logits = batch.vad_batch(utts_, handle=hh, want_logits=True)
            for u, la, lo in zip(utts_, labels, logits):
                check_vad(la.cpu().numpy(), lo.cpu().numpy(), u, ww)
            x = np.random.default_rng(1).standard_normal((50, 39)).astype(np.float32)
            ref, _ = rm.ffn_forward(x, ww)
            got = hh.ffn_predict(x)[1].cpu().numpy()
            assert np.all(np.abs(got - ref) <= LOGIT_ATOL + LOGIT_RTOL * np.abs(ref))
    h2.close()


def test_concurrent_streams_share_one_handle(env):
    import torch
    from vad_b200 import batch, runtime
    h, w = env
    n_utt, L = 64, 80000
    off, ln, stride = batch.uniform_layout(n_utt, L)
    pcm = h.synth_pcm(n_utt, L, seed=5, utt_stride=stride)
    plans = [runtime.Plan(h, off, ln, runtime.MODE_VAD) for _ in range(3)]
    base, _, _ = plans[0].vad(pcm)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=h.device) for _ in plans]
    outs = []
    for p, st in zip(plans, streams):
        with torch.cuda.stream(st):
            outs.append(p.vad(pcm)[0])
    torch.cuda.synchronize()
    for o in outs:
        assert torch.equal(o, base)


def test_pure_c_consumer_of_the_abi(tmp_path):
    """gcc-compiled C program linking libvadb200.so: the boundary works without Python or CUDA headers."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "c_abi_smoke")
    subprocess.check_call(["gcc", "-O1", "-o", exe, os.path.join(root, "tests", "c_abi_smoke.c"),
                           "-L" + os.path.join(root, "vad_b200"), "-lvadb200",
                           "-Wl,-rpath," + os.path.join(root, "vad_b200")])
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert out.returncode == 0, (out.returncode, out.stderr)
    rows, speech, checksum, diff, launches = out.stdout.split()
    assert int(rows) == rm.n_outputs(48000) + rm.n_outputs(20001) and 0 <= int(speech) <= int(rows)
    assert int(launches) >= 2


def test_outputs_stay_inside_their_buffers(env):
    """compute-sanitizer is closed on this pool, so check bounds ourselves: every output mode on a
    ragged batch writes only its own rows (guard bands before / after each buffer stay intact) and
    every row is written (no sentinel survives)."""
    import torch
    from vad_b200 import batch, runtime
    h, w = env
    lens = [401, 16003, 0, 1041, 48017, 7777, 400, 32000, 561, 9999]
    utts_ = [synth_utterance(13, i, n) for i, n in enumerate(lens)]
    flat, off, ln = batch.pack_utterances(utts_)
    pcm = flat.to(h.device)
    G = 1024

    def guarded(n, dtype, fill):
        buf = torch.full((n + 2 * G,), fill, dtype=dtype, device=h.device)
        return buf, buf[G:G + n]

    def bands_intact(buf, n, fill):
        return bool((buf[:G] == fill).all()) and bool((buf[G + n:] == fill).all())

    for mode, width in ((runtime.MODE_MFCC, 13), (runtime.MODE_DATASET, 39)):
        plan = runtime.Plan(h, off, ln, mode)
        buf, view = guarded(plan.total_rows * width, torch.float32, -777.0)
        plan.mfcc(pcm, out=view.view(plan.total_rows, width))
        torch.cuda.synchronize()
        assert bands_intact(buf, plan.total_rows * width, -777.0)
        assert not bool((view == -777.0).any())
    plan = runtime.Plan(h, off, ln, runtime.MODE_VAD)
    n = plan.total_rows
    assert n == sum(rm.n_outputs(x) for x in lens)
    lbuf, lview = guarded(n, torch.uint8, 77)
    gbuf, gview = guarded(n * 3, torch.float32, -777.0)
    fbuf, fview = guarded(n * 39, torch.float32, -777.0)
    from vad_b200._lib import check
    import ctypes as C
    check(h.lib.vadb200_vad_packed(plan._p, C.c_void_p(pcm.data_ptr()), pcm.numel(), C.c_void_p(lview.data_ptr()),
                                   C.c_void_p(gview.data_ptr()), C.c_void_p(fview.data_ptr()), 0, h.stream))
    torch.cuda.synchronize()
    assert bands_intact(lbuf, n, 77) and bands_intact(gbuf, n * 3, -777.0) and bands_intact(fbuf, n * 39, -777.0)
    assert bool((lview <= 1).all())
    assert not bool((gview == -777.0).any()) and not bool((fview == -777.0).any())
