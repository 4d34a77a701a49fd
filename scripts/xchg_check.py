#!/usr/bin/env python
"""N-rank check of the in-kernel Σw exchange (cvcs_ce_fused_tw + cvcs_b200.shard.WeightExchange).  Run under torchrun.

Every rank holds a different batch (different label mix -> different Σ v·w[y]).  For several steps each rank runs
  (a) K4 + NCCL all-reduce of Σw + K1                  (the round-1 path, collective (1) through NCCL)
  (b) K1 with the label pre-pass and the exchange inside the kernel   (one launch, no NCCL)
  (c) K4 locally, exchange inside K1                     (tw_mode 2)
and (b), (c) must give the SAME global Σw on every rank — bit-identical across ranks — and gradients equal to (a)'s to
float32 rounding.  Also times the three forms (CUDA events, max over ranks)."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvcs_b200 import ops, shard  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, C, H, W = 16, 7, 1024, 1024
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    x = (torch.randn(B, C, H, W, generator=g, device=dev) * 3).to(torch.bfloat16)
    t = torch.randint(0, C, (B, H // 32, W // 32), generator=g, device=dev, dtype=torch.uint8)
    t = t.repeat_interleave(32, 1).repeat_interleave(32, 2).contiguous()
    t[torch.rand(B, H, W, generator=g, device=dev) < 0.05 * (rank + 1)] = 255         # each rank ignores a different share
    w = (torch.arange(C, device=dev, dtype=torch.float32) + 1) / C
    xc = shard.WeightExchange(device=dev)
    ok = True
    tw_a = torch.zeros(2, dtype=torch.float64, device=dev)
    tw_b = torch.zeros(2, dtype=torch.float64, device=dev)
    tw_c = torch.zeros(2, dtype=torch.float64, device=dev)
    loc = torch.zeros(2, dtype=torch.float64, device=dev)
    for step in range(3):
        x.mul_(1.0 + 0.01 * step)
        # (a)
        ops.label_hist(t, C, 255, weight=w, total_weight_out=tw_a)
        dist.all_reduce(tw_a[0:1])
        torch.reciprocal(tw_a[0:1], out=tw_a[1:2])
        _, _, da = ops.ce_fused(x, t, w, 255, inv_total_weight_dev=tw_a[1:])
        # (b)
        _, _, db = ops.ce_fused(x, t, w, 255, total_weight="kernel", xchg=xc.handle_for(dev), total_weight_out=tw_b)
        # (c)
        ops.label_hist(t, C, 255, weight=w, total_weight_out=loc)
        _, _, dc = ops.ce_fused(x, t, w, 255, total_weight="kernel", xchg=xc.handle_for(dev), local_total_weight=loc[0:1],
                                total_weight_out=tw_c)
        torch.cuda.synchronize()
        # the totals: identical bits on every rank
        both = torch.stack((tw_b[0], tw_c[0]))
        gathered = [torch.zeros_like(both) for _ in range(world)]
        dist.all_gather(gathered, both)
        same_across_ranks = all(torch.equal(gathered[0], gq) for gq in gathered)
        rel_b = abs(float(tw_b[0]) - float(tw_a[0])) / float(tw_a[0])
        rel_c = abs(float(tw_c[0]) - float(tw_a[0])) / float(tw_a[0])
        gmax = float(da.float().abs().max())
        eb = float((db.float() - da.float()).abs().max()) / gmax
        ec = float((dc.float() - da.float()).abs().max()) / gmax
        good = same_across_ranks and rel_b < 1e-6 and rel_c < 1e-12 and eb < 8e-3 and ec < 8e-3
        ok &= good
        if rank == 0:
            print(f"step {step}: global Σw {float(tw_a[0]):.6f} | kernel pre-pass rel {rel_b:.1e}, K4+exchange rel {rel_c:.1e} | "
                  f"identical on all ranks: {same_across_ranks} | grad diff vs NCCL path {eb:.1e} / {ec:.1e} (bf16 grads)")
    seq, err = xc.state()
    ok &= (seq == 6 and err == 0)
    # timing of the three forms, back to back launches
    def timed(fn, n=50):
        for _ in range(5):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        host = (time.perf_counter() - h0) / n * 1e3
        torch.cuda.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1) / n, host], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt[0]), float(tt[1])

    d_buf = torch.empty_like(x)

    def form_a():
        ops.label_hist(t, C, 255, weight=w, total_weight_out=tw_a)
        dist.all_reduce(tw_a[0:1])
        torch.reciprocal(tw_a[0:1], out=tw_a[1:2])
        ops.ce_fused(x, t, w, 255, inv_total_weight_dev=tw_a[1:], dlogits=d_buf)

    def form_b():
        ops.ce_fused(x, t, w, 255, total_weight="kernel", xchg=xc.handle_for(dev), total_weight_out=tw_b, dlogits=d_buf)

    def form_c():
        ops.label_hist(t, C, 255, weight=w, total_weight_out=loc)
        ops.ce_fused(x, t, w, 255, total_weight="kernel", xchg=xc.handle_for(dev), local_total_weight=loc[0:1], total_weight_out=tw_c,
                     dlogits=d_buf)

    # (d) pipelined: launch i also sums the weights over the next batch's labels (staged with its chunks) and publishes
    # that sum for exchange i+1 as it ends; launch i+1 starts from it.  One launch per step, nobody waits for a peer.
    nxt = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in range(2)]
    ops.label_hist(t, C, 255, weight=w, total_weight_out=nxt[0])
    cnt = {"n": 0}

    def form_d():
        g_ = cnt["n"]
        cnt["n"] += 1
        ops.ce_fused(x, t, w, 255, total_weight="kernel", xchg=xc.handle_for(dev), local_total_weight=nxt[g_ % 2][0:1],
                     total_weight_out=tw_c, next_target=t, next_total_weight_out=nxt[(g_ + 1) % 2], dlogits=d_buf)

    # pass-end sums through the one-shot exchange vs NCCL: same result on every rank, bit for bit
    gen = torch.Generator(device=dev).manual_seed(7 + rank)
    v = torch.rand(649, generator=gen, device=dev, dtype=torch.float64) * 1e6
    cmx = torch.randint(0, 1 << 40, (7, 7), generator=gen, device=dev, dtype=torch.int64)
    v_nccl, cm_nccl = v.clone(), cmx.clone()
    dist.all_reduce(v_nccl)
    dist.all_reduce(cm_nccl)
    v_x, cm_x = v.clone(), cmx.clone()
    xc.all_reduce_(v_x)
    xc.all_reduce_(cm_x)
    torch.cuda.synchronize()
    gath = [torch.zeros_like(v_x) for _ in range(world)]
    dist.all_gather(gath, v_x)
    same_vec = all(torch.equal(gath[0], gq) for gq in gath)
    ok_ar = same_vec and torch.equal(cm_x, cm_nccl) and float((v_x - v_nccl).abs().max()) <= 1e-9 * float(v_nccl.abs().max())
    ok &= ok_ar
    t_ar_x = timed(lambda: xc.all_reduce_(v_x))
    t_ar_n = timed(lambda: dist.all_reduce(v_nccl))
    if rank == 0:
        print(f"pass-end all-reduce of 649 f64: exchange == NCCL (counts exact, floats to 1e-9, identical on all ranks): {ok_ar}; "
              f"per call device ms exchange {t_ar_x[0]:.4f} vs NCCL {t_ar_n[0]:.4f}")
    ta, tb, tc, td = timed(form_a), timed(form_b), timed(form_c), timed(form_d)
    torch.cuda.synchronize()
    rel_d = abs(float(tw_c[0]) - float(tw_a[0])) / float(tw_a[0])
    ok &= rel_d < 1e-6
    if rank == 0:
        print(f"pipelined (next-batch sum staged + pre-published): {td[0]:.4f} / {td[1]:.4f} ms, global Σw rel {rel_d:.1e}")
    if rank == 0:
        print(f"per step, same stream, max over ranks (device ms / host enqueue ms): K4 + NCCL all-reduce + K1 {ta[0]:.4f} / {ta[1]:.4f} | "
              f"one K1 launch with pre-pass + exchange {tb[0]:.4f} / {tb[1]:.4f} | K4 + K1 with exchange {tc[0]:.4f} / {tc[1]:.4f}")
        print("XCHG CHECK", "OK" if ok else "FAILED", f"world={world} exchanges={xc.state()}")
    xc.close()
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
