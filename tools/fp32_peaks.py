#!/usr/bin/env python
"""FP32 pipe microbenchmarks behind the roofline denominator (vadb200_fp32_peak variants): scalar FFMA
(register / constant operand), packed FFMA2 (register / immediate operand) and FFMA2 + FFMA mixes.
Prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vad_b200 import runtime  # noqa: E402

NAMES = ["ffma_reg", "ffma_const", "ffma2_reg", "ffma2_imm", "mix_8ffma2_8ffma", "mix_8ffma2_4ffma", "only_8ffma2"]


def main():
    h = runtime.Handle(0)
    print(json.dumps({n: round(h.fp32_peak(i, 4096), 2) for i, n in enumerate(NAMES)}))


if __name__ == "__main__":
    main()
