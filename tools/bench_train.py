#!/usr/bin/env python
"""FFN training step on the device (vadb200_train_on_batch: forward, backward, Adadelta): CUDA-event ms per step and
rows per second for a few batch sizes.  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from vad_b200 import runtime, trainer  # noqa: E402

h = runtime.Handle(0)
out = {}
for n in (4096, 65536, 1048576):
    t = trainer.FFNTrainer(handle=h, seed=0, max_batch=n)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(n, 39, device="cuda", generator=g)
    y = torch.randint(0, 3, (n,), device="cuda", generator=g, dtype=torch.int64).to(torch.uint8)
    for _ in range(5):
        t.train_on_batch(x, y, want_loss=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        t.train_on_batch(x, y, want_loss=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    # forward + backward = 3 x 2 x 5104 MAC-flops per row (activations, input grads, weight grads)
    out[str(n)] = {"ms_per_step": round(ms, 4), "rows_per_s": round(n / (ms * 1e-3)),
                   "tflops": round(n * 3 * 2 * 5104 / (ms * 1e-3) / 1e12, 2)}
    t.close()
print(json.dumps(out))
