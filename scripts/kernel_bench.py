#!/usr/bin/env python
"""Per-kernel throughput of the C-ABI entry points around K1 at BASELINE-sized inputs (CUDA events,
rotating buffers where they fit).  Prints one JSON line per kernel; algorithmic bytes per pixel as in
SURVEY §8(d) / DESIGN.md §3."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvcs_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, iters=50, warm=5):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(name, px, bpp, ms):
    gbs = px * bpp / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "pixels": px, "bytes_per_pixel": bpp, "ms": round(ms, 4),
                      "gpixel_s": round(px / (ms * 1e-3) / 1e9, 2), "gb_s": round(gbs, 1), "frac_of_measured_peak": round(gbs / PEAK, 3)}),
          flush=True)


def main():
    B, C, H, W = 16, 7, 1024, 1024
    px = B * H * W
    g = torch.Generator(device=dev).manual_seed(0)
    xs = [torch.randn(B, C, H, W, generator=g, device=dev) * 3 for _ in range(3)]
    t = torch.randint(0, C, (B, H // 32, W // 32), generator=g, device=dev, dtype=torch.uint8)
    t = t.repeat_interleave(32, 1).repeat_interleave(32, 2).contiguous()
    # K2 argmax
    report("K2 argmax f32 C=7 -> u8", px, C * 4 + 1, timed(lambda i: ops.argmax(xs[i % 3], torch.uint8)))
    xb = [x.to(torch.bfloat16) for x in xs]
    report("K2 argmax bf16 C=7 -> u8", px, C * 2 + 1, timed(lambda i: ops.argmax(xb[i % 3], torch.uint8)))
    # K3 confusion matrix from index maps
    pred = ops.argmax(xs[0], torch.uint8)
    cm = torch.zeros((C, C), dtype=torch.int64, device=dev)
    report("K3 confmat u8/u8 C=7", px, 2, timed(lambda i: ops.confmat_update(cm, pred, t, C, None)))
    cm16 = torch.zeros((16, 16), dtype=torch.int64, device=dev)
    report("K3 confmat u8/u8 C=16", px, 2, timed(lambda i: ops.confmat_update(cm16, pred, t, 16, 0)))
    # K4
    hist = torch.zeros(C + 2, dtype=torch.int64, device=dev)
    report("K4 label histogram u8 C=7", px, 1, timed(lambda i: ops.label_hist(t, C, 255, hist=hist)))
    tw = torch.zeros(2, dtype=torch.float64, device=dev)
    w = torch.rand(C, device=dev) + 0.5
    report("K4 lean total weight u8 C=7", px, 1, timed(lambda i: ops.label_hist(t, C, 255, weight=w, total_weight_out=tw)))
    t64 = t.long()
    report("K4 lean total weight i64 C=7", px, 8, timed(lambda i: ops.label_hist(t64, C, 255, weight=w, total_weight_out=tw)))
    # N3 vote over 5 maps, N4 colourise, N2 stitch
    maps = torch.randint(0, 16, (5, px), generator=g, device=dev, dtype=torch.uint8)
    report("N3 vote 5 x u8", px, 6, timed(lambda i: ops.vote(maps)))
    lut = torch.rand(16, 3, device=dev)
    idx = maps[0].reshape(B * H, W)
    report("N4 colorize u8 -> f32 RGB", px, 13, timed(lambda i: ops.colorize(idx, lut)))
    tiles = maps[1].reshape(B, H, W)
    yx = torch.tensor([((i // 4) * H, (i % 4) * W) for i in range(B)], dtype=torch.int32, device=dev)
    out = torch.zeros((4 * H, 4 * W), dtype=torch.uint8, device=dev)
    report("N2 stitch u8 tiles -> scene", px, 2, timed(lambda i: ops.stitch(tiles, yx, (4 * H, 4 * W), out=out)))


if __name__ == "__main__":
    main()
