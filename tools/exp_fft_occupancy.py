"""Occupancy experiment (hook build): FFT phase alone at 1..4 CTAs per SM.
python -m vad_b200.build --debug-hooks && VADB200_LIB=$PWD/vad_b200/libvadb200_dbg.so python tools/exp_fft_occupancy.py"""
import ctypes as C, os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vad_b200 import batch, runtime
from vad_b200._lib import check
h = runtime.Handle(0, ffn_weights=runtime.glorot_ffn(0))
fn = h.lib.vadb200_exp_fft
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]; fn.restype = C.c_int
n_utt, L = 72000, 160000          # 200 h
off, ln, stride = batch.uniform_layout(n_utt, L)
pcm = h.synth_pcm(n_utt, L, utt_stride=stride)
plan = runtime.Plan(h, off, ln, runtime.MODE_MFCC)
out = {}
for minb in (1, 2, 3, 4):
    for _ in range(2): check(fn(plan._p, pcm.data_ptr(), pcm.numel(), minb, h.stream))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): check(fn(plan._p, pcm.data_ptr(), pcm.numel(), minb, h.stream))
    e1.record(); torch.cuda.synchronize()
    out["ctas_per_sm_%d" % minb] = round(e0.elapsed_time(e1) / 3 * 5, 2)   # ms per 1000 h equivalent
print(json.dumps(out))
