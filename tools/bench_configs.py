#!/usr/bin/env python
"""One of BASELINE.json's secondary configs in isolation (the same code bench.py runs for its sub-records), for
ncu captures and A/B work:   python tools/bench_configs.py {cfg2|cfg5|stream|analyser} [--stream-ticks N]
Prints one JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from vad_b200 import runtime  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["cfg2", "cfg5", "stream", "analyser"])
    ap.add_argument("--stream-ticks", type=int, default=10000)
    ap.add_argument("--streams", type=int, default=4096)
    a = ap.parse_args()
    h = runtime.Handle(0, ffn_weights=runtime.glorot_ffn(0))
    peak = max(h.fp32_peak(0, 2048), h.fp32_peak(1, 2048), h.fp32_peak(2, 2048))
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm = float(json.load(f)["hbm_gbs"])
    except Exception:
        hbm = 6650.0
    if a.which == "cfg2":
        out = bench.cfg2_record(h, peak, hbm, torch)
    elif a.which == "cfg5":
        out = bench.cfg5_record(h, peak, hbm, torch)
    elif a.which == "stream":
        out = bench.stream_record(h, a, torch)
    else:
        out = bench.analyser_record(h, torch)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
