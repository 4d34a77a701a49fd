#!/usr/bin/env python
"""N-rank NCCL check of cvcs_b200.shard.ShardedScenePass (cfg4-style): tile-sharded scenes, global
loss and confusion matrix must equal the single-rank result on all tiles.  Run under torchrun."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cvcs_b200 import shard  # noqa: E402


def full_size(rank, world, dev):
    """cfg4 at full size: 2 x world scenes of 10000 x 10000 (81 tiles of 1024 x 1024 each), dealt round-robin by global
    tile id; the all-reduced confusion matrix must be bit-identical to the one a single process computes over ALL tiles."""
    import time
    C, p, HW = 7, 1024, 10000
    n_scenes = 2 * world
    scenes = []
    for s in range(n_scenes):
        g = torch.Generator(device=dev).manual_seed(3 + s)          # every rank generates the same scenes
        img = torch.randint(0, 256, (3, HW, HW), generator=g, device=dev, dtype=torch.uint8)
        lab = torch.randint(0, C, (HW // 40, HW // 40), generator=g, device=dev, dtype=torch.uint8)
        lab = lab.repeat_interleave(40, 0).repeat_interleave(40, 1).contiguous()
        scenes.append((img, lab))
    g = torch.Generator(device=dev).manual_seed(1)
    proj = torch.randn(C, 3, generator=g, device=dev) * 0.02

    def logits_fn(x, y):                                              # stub segmenter (not the hot path)
        return torch.einsum("kc,bchw->bkhw", proj, x).contiguous()

    t0 = time.perf_counter()
    sp = shard.ShardedScenePass(scenes, p, C, logits_fn, ignore_index=-100, batch_size=16, device=dev, want_grad=False).run()
    loss, cm = sp.finish()
    torch.cuda.synchronize()
    t_sharded = time.perf_counter() - t0
    sp1 = shard.ShardedScenePass(scenes, p, C, logits_fn, ignore_index=-100, batch_size=16, device=dev, want_grad=False,
                                 single_process=True).run()
    loss1, cm1 = sp1.finish()
    same = bool(torch.equal(cm, cm1))
    rel = abs(float(loss) - float(loss1)) / abs(float(loss1))
    if rank == 0:
        print(f"FULL SIZE world={world}: {n_scenes} scenes of {HW}x{HW}, {sp.n_tiles_done} of {sp1.n_tiles_done} tiles on rank 0, "
              f"{int(cm.sum())} pixels; confusion bit-identical to the single-process pass: {same}; loss rel diff {rel:.1e}; "
              f"sharded pass {t_sharded:.2f} s wall (with the stub segmenter)")
    return same and rel < 1e-6


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    if "--full" in sys.argv:
        ok = full_size(rank, world, dev)
        if rank == 0:
            print("SHARD CHECK FULL", "OK" if ok else "FAILED", f"world={world}")
        dist.barrier()
        dist.destroy_process_group()
        sys.exit(0 if ok else 1)
    C, p, n_scenes, HW = 7, 256, 3, 1100                     # 4 x 4 = 16 whole tiles per scene
    g = torch.Generator().manual_seed(0)
    scenes = []
    for s in range(n_scenes):
        img = torch.randint(0, 256, (3, HW, HW), generator=g, dtype=torch.uint8)
        lab = torch.randint(0, C, (HW // 20, HW // 20), generator=g, dtype=torch.uint8).repeat_interleave(20, 0).repeat_interleave(20, 1)
        lab[torch.rand(HW, HW, generator=g) < 0.1] = 255
        scenes.append((img, lab.contiguous()))
    proj = (torch.randn(C, 3, generator=g) * 0.02).to(dev)

    def logits_fn(x, y):                                      # a stub "segmenter": per-pixel linear map of the bands
        return torch.einsum("kc,bchw->bkhw", proj, x).contiguous()

    weight = (torch.arange(C, dtype=torch.float32) + 1).to(dev) / C
    results = {}
    for policy in ("round_robin", "scene"):
        sp = shard.ShardedScenePass(scenes, p, C, logits_fn, weight=weight, ignore_index=255, batch_size=8, device=dev,
                                    policy=policy, want_grad=True).run()
        loss, cm = sp.finish()
        results[policy] = (float(loss), cm)
    # single-process reference: every rank redoes ALL tiles with collectives switched off
    sp1 = shard.ShardedScenePass(scenes, p, C, logits_fn, weight=weight, ignore_index=255, batch_size=8, device=dev,
                                 want_grad=True, single_process=True).run()
    l1_t, cm1 = sp1.finish()
    l1 = float(l1_t)
    ok = True
    for policy, (l, cm) in results.items():
        same_cm = bool(torch.equal(cm, cm1))
        rel = abs(l - l1) / abs(l1)
        if rank == 0:
            print(f"{policy}: loss {l:.7f} vs single-process {l1:.7f} (rel {rel:.2e}); confusion equal: {same_cm}; pixels {int(cm.sum())}")
        ok &= same_cm and rel < 1e-6
    if rank == 0:
        print("SHARD CHECK", "OK" if ok else "FAILED", f"world={world}")
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
