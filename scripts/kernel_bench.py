#!/usr/bin/env python
"""Device time of the small C-ABI kernels around K1 (K2 argmax, K3 confusion, K4 histogram / total weight, N2 stitch,
N3 vote, N4 colourise / context) at BASELINE-sized inputs.

Each kernel runs over ROTATING input sets whose total exceeds the 126 MB L2 (so every launch reads HBM, not L2), and a
round of launches is captured into ONE CUDA graph and replayed: the time per launch is then the device's, not the
Python call rate (round 1 timed host calls on one L2-resident buffer).  `--plain` skips the graph (for `ncu
--metrics gpu__time_duration.sum`, scripts/gpu_small_kernels.sh).  One JSON line per kernel; algorithmic bytes per
pixel as in SURVEY §8(d) / DESIGN.md §3."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvcs_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
_pk = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")
PEAK = json.load(open(_pk))["hbm_gbs"] if os.path.exists(_pk) else 6650.0
PLAIN = "--plain" in sys.argv
ONCE = "--once" in sys.argv          # two launches per kernel, untimed: for `ncu --set full` captures


def timed(fns, replays=10):
    """fns: one callable per rotating buffer set.  Returns ms per launch."""
    if ONCE:
        fns[0]()
        fns[-1]()
        torch.cuda.synchronize()
        return 1.0
    for f in fns:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if PLAIN:
        e0.record()
        for _ in range(2):
            for f in fns:
                f()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (2 * len(fns))
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for f in fns:
                f()
    torch.cuda.current_stream(dev).wait_stream(side)
    graph.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(replays):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (replays * len(fns))


def report(name, px, bpp, ms, sets):
    gbs = px * bpp / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "pixels": px, "bytes_per_pixel": bpp, "us": round(ms * 1e3, 2),
                      "gpixel_s": round(px / (ms * 1e-3) / 1e9, 2), "gb_s": round(gbs, 1), "frac_of_measured_peak": round(gbs / PEAK, 3),
                      "rotating_sets": sets, "rotating_mb": round(sets * px * bpp / 1e6), "timing": "plain launches" if PLAIN else "CUDA graph replay"}),
          flush=True)


def main():
    B, C, H, W = 16, 7, 1024, 1024
    px = B * H * W
    g = torch.Generator(device=dev).manual_seed(0)

    def labels(n, classes):
        out = []
        for _ in range(n):
            t = torch.randint(0, classes, (B, H // 32, W // 32), generator=g, device=dev, dtype=torch.uint8)
            out.append(t.repeat_interleave(32, 1).repeat_interleave(32, 2).contiguous())
        return out

    # ---- K2 argmax (470 MB per fp32 set: 2 sets already exceed L2)
    xs = [torch.randn(B, C, H, W, generator=g, device=dev) * 3 for _ in range(2)]
    am = [torch.empty((B, H, W), dtype=torch.uint8, device=dev) for _ in xs]
    report("K2 argmax f32 C=7 -> u8", px, C * 4 + 1, timed([lambda i=i: ops.argmax(xs[i], out=am[i]) for i in range(2)]), 2)
    xb = [x.to(torch.bfloat16) for x in xs]
    report("K2 argmax bf16 C=7 -> u8", px, C * 2 + 1, timed([lambda i=i: ops.argmax(xb[i], out=am[i]) for i in range(2)]), 2)
    del xs, xb
    # ---- index-map kernels: 16 label sets (+ 16 prediction sets) of 16.8 MB each
    n = 16
    ts, ps = labels(n, C), labels(n, C)
    for p_ in ps:                                               # predictions: i.i.d. within the blocks, like an argmax map
        p_.copy_(torch.randint(0, C, p_.shape, generator=g, device=dev, dtype=torch.uint8))
    cm = torch.zeros((C, C), dtype=torch.int64, device=dev)
    report("K3 confmat u8/u8 C=7", px, 2, timed([lambda i=i: ops.confmat_update(cm, ps[i], ts[i], C, None) for i in range(n)]), n)
    cm16 = torch.zeros((16, 16), dtype=torch.int64, device=dev)
    report("K3 confmat u8/u8 C=16 (shared bins)", px, 2,
           timed([lambda i=i: ops.confmat_update(cm16, ps[i], ts[i], 16, 0) for i in range(n)]), n)
    hist = torch.zeros(C + 2, dtype=torch.int64, device=dev)
    report("K4 label histogram u8 C=7", px, 1, timed([lambda i=i: ops.label_hist(ts[i], C, 255, hist=hist) for i in range(n)]), n)
    tw = torch.zeros(2, dtype=torch.float64, device=dev)
    w = torch.rand(C, device=dev) + 0.5
    report("K4 lean total weight u8 C=7", px, 1,
           timed([lambda i=i: ops.label_hist(ts[i], C, 255, weight=w, total_weight_out=tw) for i in range(n)]), n)
    t64 = [t.long() for t in ts[:4]]
    t8 = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    report("K4 labels_prepare i64 -> total weight + u8 labels", px, 9,
           timed([lambda i=i: ops.labels_prepare(t64[i], C, 255, w, tw, t8) for i in range(4)]), 4)
    del t64
    # ---- N3 vote over 5 maps (84 MB per set), N4 colourise, N2 stitch, N4 context
    maps = [torch.randint(0, 16, (5, px), generator=g, device=dev, dtype=torch.uint8) for _ in range(3)]
    vout = torch.empty(px, dtype=torch.uint8, device=dev)
    report("N3 vote 5 x u8", px, 6, timed([lambda i=i: ops.vote(maps[i], out=vout) for i in range(3)]), 3)
    lut = torch.rand(16, 3, device=dev)
    cout = torch.empty((B * H, W, 3), dtype=torch.float32, device=dev)
    report("N4 colorize u8 -> f32 RGB", px, 13, timed([lambda i=i: ops.colorize(ts[i].reshape(B * H, W), lut, out=cout) for i in range(n)]), n)
    yx = torch.tensor([((i // 4) * H, (i % 4) * W) for i in range(B)], dtype=torch.int32, device=dev)
    outs = [torch.zeros((4 * H, 4 * W), dtype=torch.uint8, device=dev) for _ in range(n)]
    report("N2 stitch u8 tiles -> scene", px, 2, timed([lambda i=i: ops.stitch(ts[i], yx, (4 * H, 4 * W), out=outs[i]) for i in range(n)]), n)
    del maps, outs
    # context: 64 patches of 224 x 224 from 4-band 6800 x 7200 scenes (the reference's GID tiles), 9 B read + 1 B written per px·band
    scenes = [torch.randint(0, 256, (4, 6800, 7200), generator=g, device=dev, dtype=torch.uint8) for _ in range(2)]
    p = 224
    cyx = torch.tensor([((i // 8) * 3 * p, (i % 8) * 3 * p) for i in range(64)], dtype=torch.int32, device=dev)
    couts = torch.empty((64, 4, p, p), dtype=torch.uint8, device=dev)
    cpx = 64 * p * p
    report("N4 context 3p x 3p -> p (p=224, 4 bands, 64 patches)", cpx, 4 * 10, timed([lambda i=i: ops.tile_context(scenes[i], cyx, p, out=couts) for i in range(2)]), 2)


if __name__ == "__main__":
    main()
