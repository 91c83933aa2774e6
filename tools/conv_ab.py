"""A/B timing of the VAE's hot 3x3x3 convolutions: CTA-pair kernel (aether_conv3d_bf16) vs the one-CTA kernel
(aether_conv3d_bf16_1cta) at the decoder's real tile shapes.  usage: python tools/conv_ab.py [--json out]"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from aether_b200 import _lib  # noqa: E402
from aether_b200._lib import check, current_stream, ptr  # noqa: E402

BF16 = torch.bfloat16
SHAPES = [  # (name, T_out, H, W, Cin, Cout)   decoder levels of a 30 x 45 latent tile; encoder levels are the same shapes
    ("dec 512->512 @ 3x30x45", 3, 30, 45, 512, 512),
    ("dec 512->512 @ 5x60x90", 5, 60, 90, 512, 512),
    ("dec 256->256 @ 9x120x180", 9, 120, 180, 256, 256),
    ("dec 256->256 @ 8x120x180", 8, 120, 180, 256, 256),
    ("dec 128->128 @ 9x240x360", 9, 240, 360, 128, 128),
    ("dec 128->128 @ 8x240x360", 8, 240, 360, 128, 128),
    ("dec 256->128 @ 8x240x360", 8, 240, 360, 256, 128),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", type=int, default=None, help="index into SHAPES (ncu captures)")
    args = ap.parse_args()
    lib = _lib.require_device()
    g = torch.Generator(device="cuda").manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    for name, T, H, W, cin, cout in (SHAPES if args.only is None else [SHAPES[args.only]]):
        x = torch.randn(T + 2, H, W, cin, device="cuda", generator=g).to(BF16)
        w = (torch.randn(cout, 27 * cin, device="cuda", generator=g) / (27 * cin) ** 0.5).to(BF16)
        b = torch.zeros(cout, device="cuda")
        y = torch.empty(T, H, W, cout, dtype=BF16, device="cuda")
        res = {}
        for label, fn in (("pair", lib.aether_conv3d_bf16), ("one_cta", lib.aether_conv3d_bf16_1cta)):
            def run():
                check(fn(ptr(x), T + 2, H, W, cin, ptr(w), ptr(b), None, ptr(y), T, H, W, cout, 3, 3, 3, 1, 1, 1,
                         current_stream()), "conv")
            for _ in range(3):
                run()
            ts = []
            for _ in range(args.reps):
                flush.zero_()
                a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                run()
                e.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(e))
            ts.sort()
            res[label] = ts[len(ts) // 2]
        flop = 2.0 * T * H * W * cout * 27 * cin
        row = {"shape": name, "pair_ms": res["pair"], "one_cta_ms": res["one_cta"],
               "pair_tflops": flop / res["pair"] / 1e9, "one_cta_tflops": flop / res["one_cta"] / 1e9,
               "speedup": res["one_cta"] / res["pair"]}
        rows.append(row)
        print(json.dumps(row), flush=True)
    if args.json:
        Path(args.json).write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
