#!/usr/bin/env python
"""Experiment (needs the TIMING build: python -m cvcs_b200.build --tag TIMING -D CVCS_X_TIMING and
CVCS_B200_LIB pointing at it): per-CTA start/end globaltimer stamps of one K1 launch of cfg2."""
import os, sys, struct
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvcs_b200 import ops, _lib
dev = torch.device("cuda", 0)
B, C, H, W = 16, 7, 1024, 1024
g = torch.Generator(device=dev).manual_seed(0)
xs = [torch.randn(B, C, H, W, generator=g, device=dev) * 3 for _ in range(2)]
t = torch.randint(0, C, (B, H, W), generator=g, device=dev, dtype=torch.uint8)
dl = [torch.empty_like(x) for x in xs]
am = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
cm = torch.zeros((C, C), dtype=torch.int64, device=dev)
ws = ops.workspace(dev)
HIST_OFF = 64 + 2 * 4096 * 8          # Workspace: 64 B header + partial[2*kMaxGrid] doubles
for it in range(6):
    ops.ce_fused(xs[it % 2], t, None, -100, want_grad=True, inv_total_weight=1.0 / (B * H * W), dlogits=dl[it % 2], argmax=am, confmat=cm)
    torch.cuda.synchronize()
    raw = ws.cpu().numpy().tobytes()
    h = struct.unpack_from("<1032Q", raw, HIST_OFF)
    ends = [v for v in h[:512] if v]
    starts = [v for v in h[512:1024] if v]
    ws[HIST_OFF:HIST_OFF + 1032 * 8] = 0
    torch.cuda.synchronize()
    if it >= 2 and ends:
        t0 = min(starts)
        e = sorted(v - t0 for v in ends)
        s = sorted(v - t0 for v in starts)
        n = len(e)
        print(f"launch {it}: CTAs {n}; start spread {s[-1]/1e3:.1f} us; end min {e[0]/1e3:.1f} p10 {e[n//10]/1e3:.1f} median {e[n//2]/1e3:.1f} p90 {e[9*n//10]/1e3:.1f} max {e[-1]/1e3:.1f} us")
