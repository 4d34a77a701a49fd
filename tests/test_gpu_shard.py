"""cfg4-style pass (cvcs_b200/shard.py::ShardedScenePass) on one GPU: K5 tiling -> stub segmenter -> K4 -> K1,
against the oracle computed on the host over ALL tiles.  The N-rank NCCL run of the same pass is
scripts/shard_check.py (needs N GPUs); the N-rank host logic is tests/test_shard_gloo.py."""
import numpy as np
import pytest
import torch

from oracle import c_oracle

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


@pytest.mark.parametrize("policy", ["round_robin", "scene"])
def test_sharded_scene_pass_matches_oracle(policy):
    from cvcs_b200 import shard
    C, p, HW, n_scenes = 7, 64, 300, 3                       # 4 x 4 whole tiles per scene, 44-pixel remainder dropped
    g = torch.Generator().manual_seed(5)
    scenes = []
    for _ in range(n_scenes):
        img = torch.randint(0, 256, (3, HW, HW), generator=g, dtype=torch.uint8)
        lab = torch.randint(0, C, (HW // 10, HW // 10), generator=g, dtype=torch.uint8).repeat_interleave(10, 0).repeat_interleave(10, 1)
        lab[torch.rand(HW, HW, generator=g) < 0.1] = 255
        scenes.append((img, lab.contiguous()))
    proj = torch.randn(C, 3, generator=g) * 0.02
    weight = (torch.arange(C, dtype=torch.float32) + 1) / C
    proj_d = proj.to(DEV)

    def logits_fn(x, y):
        return torch.einsum("kc,bchw->bkhw", proj_d, x).contiguous()

    sp = shard.ShardedScenePass(scenes, p, C, logits_fn, weight=weight.to(DEV), ignore_index=255, batch_size=5,
                                device=DEV, policy=policy, want_grad=True, single_process=True).run()
    loss, cm = sp.finish()
    assert sp.n_tiles_done == n_scenes * 16

    # oracle over all tiles, in global tile order
    tiles = shard.local_tiles(n_scenes, [HW, HW], p, 0, 1)
    xs, ts = [], []
    for gid, s, tly, tlx in tiles:
        yx = np.array([[tly, tlx]], dtype=np.int32)
        x, t = c_oracle.tile(scenes[s][0].numpy(), yx, p, p, labels=scenes[s][1].numpy())
        xs.append(torch.einsum("kc,bchw->bkhw", proj, torch.from_numpy(x)).numpy())
        ts.append(t.astype(np.int64))
    xs, ts = np.concatenate(xs), np.concatenate(ts)
    l_ref, sums_ref, _ = c_oracle.cross_entropy(xs, ts, weight.numpy(), 255)
    cm_ref, _ = c_oracle.confmat(c_oracle.argmax(xs), ts, C, 255)
    assert np.array_equal(cm.numpy(), cm_ref)
    assert abs(float(loss) - l_ref) <= 1e-5 * abs(l_ref)
    hist_ref = c_oracle.label_hist(ts.astype(np.uint8), C, 255)
    assert np.array_equal(sp.hist.cpu().numpy(), hist_ref)
