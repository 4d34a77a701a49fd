#!/usr/bin/env python
"""Experiment (needs the TIMING build: python -m cvcs_b200.build --tag TIMING -D CVCS_X_TIMING and
CVCS_B200_LIB pointing at it): per-CTA globaltimer stamps of K1 launches — where the fixed per-launch cost is.

    CVCS_B200_LIB=cvcs_b200/libcvcs_b200_TIMING.so python scripts/cta_timing.py [cfg2|cfg3] [extra option=value ...]
"""
import os, sys, struct
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvcs_b200 import ops, _lib
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
for a in sys.argv[2:]:
    k, v = a.split("=")
    _lib.set_option(getattr(_lib, "OPT_" + k.upper()), int(v))
dev = torch.device("cuda", 0)
B, C, H, W = 16, 7, 1024, 1024
dt = torch.float32 if wl == "cfg2" else torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
xs = [(torch.randn(B, C, H, W, generator=g, device=dev) * 3).to(dt) for _ in range(3)]
t = torch.randint(0, C, (B, H // 32, W // 32), generator=g, device=dev, dtype=torch.uint8).repeat_interleave(32, 1).repeat_interleave(32, 2).contiguous()
ii = -100
weight = None
if wl == "cfg3":
    t[torch.rand(B, H, W, generator=g, device=dev) < 0.1] = 255
    ii = 255
    weight = torch.rand(C, generator=g, device=dev) + 0.5
dl = [torch.empty_like(x) for x in xs]
am = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
cm = torch.zeros((C, C), dtype=torch.int64, device=dev)
ws = ops.workspace(dev)
HIST_OFF = 64 + 2 * 4096 * 8          # Workspace: 64 B header + partial[2*kMaxGrid] doubles
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for it in range(8):
    torch.cuda.synchronize()
    ev[0].record()
    ops.ce_fused(xs[it % 3], t, weight, ii, want_grad=True, inv_total_weight=1.0 / (B * H * W), dlogits=dl[it % 3], argmax=am, confmat=cm)
    ev[1].record()
    torch.cuda.synchronize()
    raw = ws.cpu().numpy().tobytes()
    h = struct.unpack_from("<1032Q", raw, HIST_OFF)
    ws[HIST_OFF:HIST_OFF + 1032 * 8] = 0
    torch.cuda.synchronize()
    if it < 3:
        continue
    n = sum(1 for v in h[:256] if v)
    st, fi, lo, en, fin = h[0:n], h[256:256 + n], h[512:512 + n], h[768:768 + n], h[1024]
    t0 = min(st)
    q = lambda xs_, f: sorted(xs_)[min(int(f * len(xs_)), len(xs_) - 1)]
    us = lambda v: (v - t0) / 1e3
    print(f"{wl} launch {it}: events {ev[0].elapsed_time(ev[1]) * 1e3:.1f} us | CTAs sampled {n} | start max {us(max(st)):.1f} | "
          f"first stage ready min/med/max {us(min(fi)):.1f}/{us(q(fi, .5)):.1f}/{us(max(fi)):.1f} | "
          f"last chunk done min/med/max {us(min(lo)):.1f}/{us(q(lo, .5)):.1f}/{us(max(lo)):.1f} | "
          f"before epilogue max {us(max(en)):.1f} | grid end {us(fin):.1f} us")
