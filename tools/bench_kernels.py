#!/usr/bin/env python
"""Kernel micro-benchmarks on the GPU box (CUDA events inside libsdtf, operands rotated through > L2 worth of buffers).
Prints one JSON line per case: attention shapes of the SD1.5 UNet at batch 16 and the main conv / linear shapes."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minsdtf_b200.engine import Engine  # noqa: E402

PEAK = 1714.9


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    e = Engine(0)
    B = 16
    if what in ("all", "attn"):
        for (nq, nk, d, tag) in [(4096, 4096, 40, "self 64x64"), (4096, 77, 40, "cross 64x64"), (1024, 1024, 80, "self 32x32"),
                                 (1024, 77, 80, "cross 32x32"), (256, 256, 160, "self 16x16"), (9216, 9216, 40, "self 96x96 (B=4)")]:
            b = 4 if nq == 9216 else B
            for legacy in ((False, True) if d in (40, 80) else (False,)):
                ms = e.bench_attention(b, 8, nq, nk, d, reps=10, legacy=legacy)
                fl = 4.0 * nq * nk * 8 * d * b
                print(json.dumps({"kernel": "attention", "case": tag, "legacy": legacy, "B": b, "Nq": nq, "Nk": nk, "d": d, "ms": round(ms, 4),
                                  "tflops": round(fl / ms / 1e9, 1), "exp_per_ns": round(nq * nk * 8 * b / ms / 1e6, 1)}), flush=True)
    if what in ("all", "conv"):
        for (hw, cin, cout, k) in [(64, 320, 320, 3), (64, 640, 320, 3), (64, 960, 320, 3), (32, 640, 640, 3), (32, 1280, 640, 3),
                                   (16, 1280, 1280, 3), (16, 2560, 1280, 3), (8, 1280, 1280, 3), (8, 2560, 1280, 3),
                                   (64, 320, 320, 1), (64, 320, 2560, 1), (64, 1280, 320, 1), (32, 640, 5120, 1), (32, 2560, 640, 1),
                                   (16, 1280, 10240, 1), (16, 5120, 1280, 1), (64, 320, 1536, 1), (32, 640, 1920, 1)]:
            ms = e.bench_conv(B, hw, cin, cout, k, reps=20)
            fl = 2.0 * B * hw * hw * cin * cout * k * k
            print(json.dumps({"kernel": "conv_gemm", "B": B, "hw": hw, "cin": cin, "cout": cout, "k": k, "ms": round(ms, 4),
                              "tflops": round(fl / ms / 1e9, 1), "frac_of_burst_peak": round(fl / ms / 1e9 / PEAK, 3)}), flush=True)
    e.close()


if __name__ == "__main__":
    main()
