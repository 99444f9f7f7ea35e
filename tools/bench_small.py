"""Small-batch timings (configs[0]: one 60 s waveform; 64 x 10 s): CUDA-event ms per fused VAD launch."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vad_b200 import batch, runtime
h = runtime.Handle(0, ffn_weights=runtime.glorot_ffn(0))
out = {}
for name, n_utt, L in (("cfg1_1x60s", 1, 960000), ("64x10s", 64, 160000), ("1024x10s", 1024, 160000)):
    off, ln, stride = batch.uniform_layout(n_utt, L)
    pcm = h.synth_pcm(n_utt, L, utt_stride=stride)
    plan = runtime.Plan(h, off, ln, runtime.MODE_VAD)
    lab = torch.empty(plan.total_rows, dtype=torch.uint8, device=h.device)
    for _ in range(10): plan.vad(pcm, labels=lab)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): plan.vad(pcm, labels=lab)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    out[name] = {"ms": round(ms, 4), "audio_s_per_s": round(n_utt * L / 16000 / (ms * 1e-3))}
print(json.dumps(out))
