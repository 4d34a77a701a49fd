"""N > 1 host logic (cvcs_b200/shard.py) with world_size-2 `gloo` on the CPU: tile ownership and the
three collectives.  Each rank computes its shard's partial results with the oracle (the checker),
reduces them through shard.py, and the result must equal the single-process oracle on all tiles —
which is what 'same result as the single-process reference on identical inputs' means."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import c_oracle

C, P, H, W, N_SCENES = 7, 32, 100, 140, 3     # 3 x 4 whole tiles per scene; remainder dropped


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _scene(s):
    g = np.random.RandomState(100 + s)
    lab = g.randint(0, C, ((H + 7) // 8, (W + 7) // 8)).repeat(8, 0).repeat(8, 1)[:H, :W].astype(np.uint8)
    lab[g.rand(H, W) < 0.1] = 255
    return lab


def _tile_data(g_id, s, tly, tlx):
    """Label tile cut from the scene + a deterministic logits tile any rank can regenerate."""
    lab = _scene(s)[tly:tly + P, tlx:tlx + P]
    x = (np.random.RandomState(1000 + g_id).randn(1, C, P, P) * 3).astype(np.float32)
    return x, lab[None].astype(np.int64)


def _partials(tiles, weight, inv_total=None):
    hist = np.zeros(C + 2, dtype=np.int64)
    cm = np.zeros((C, C), dtype=np.int64)
    sums = np.zeros(3, dtype=np.float64)
    grads = {}
    for g_id, s, tly, tlx in tiles:
        x, t = _tile_data(g_id, s, tly, tlx)
        hist += c_oracle.label_hist(t.astype(np.uint8), C, 255)
        _, sm, d = c_oracle.cross_entropy(x, t, weight, 255, want_grad=True)
        sums += sm
        c_oracle.confmat(c_oracle.argmax(x), t, C, 255, into=cm)
        if inv_total is not None:
            # oracle gradients are normalised by the LOCAL Σw; rescale to the global one
            grads[g_id] = d * (sm[1] * inv_total)
    return hist, cm, sums, grads


def _worker(rank, world, port, policy, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cvcs_b200 import shard
        weight = (np.arange(C, dtype=np.float32) + 1) / C
        tiles = shard.local_tiles(N_SCENES, [H, W], P, rank, world, policy)
        hist, cm, sums, _ = _partials(tiles, weight)
        hist_t = torch.from_numpy(hist.copy())
        tw = shard.global_total_weight(hist_t, torch.from_numpy(weight), C, 255)          # collective (1)
        loss = shard.global_loss(torch.from_numpy(sums))                                    # collective (2)
        cm_t = shard.global_confmat(torch.from_numpy(cm.copy()))                            # collective (3)
        _, _, _, grads = _partials(tiles[:2], weight, inv_total=float(tw[1]))
        q.put((rank, [t[0] for t in tiles], hist_t.numpy(), tw.numpy(), float(loss), cm_t.numpy(),
               {k: v for k, v in grads.items()}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("policy", ["round_robin", "scene"])
def test_two_rank_gloo_equals_single_process(policy):
    from cvcs_b200 import shard
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, policy, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0

    # single-process reference over ALL tiles
    weight = (np.arange(C, dtype=np.float32) + 1) / C
    all_tiles = shard.local_tiles(N_SCENES, [H, W], P, 0, 1)
    assert len(all_tiles) == N_SCENES * 12 and [t[0] for t in all_tiles] == list(range(N_SCENES * 12))
    hist, cm, sums, _ = _partials(all_tiles, weight)
    sw = float((hist[:C].astype(np.float64) * weight.astype(np.float64)).sum())
    xs = np.concatenate([_tile_data(*t)[0] for t in all_tiles])
    ts = np.concatenate([_tile_data(*t)[1] for t in all_tiles])
    loss_1p, sums_1p, d_1p = c_oracle.cross_entropy(xs, ts, weight, 255, want_grad=True)

    owned = sorted(g for r in results for g in r[1])
    assert owned == list(range(N_SCENES * 12))                       # a partition: every tile exactly once
    for rank, ids, hist_r, tw_r, loss_r, cm_r, grads_r in results:
        assert np.array_equal(hist_r, hist)                          # integer sums: bit-exact
        assert np.array_equal(cm_r, cm)
        assert abs(tw_r[0] - sw) <= 1e-12 * sw and abs(tw_r[1] * sw - 1.0) <= 1e-12
        assert abs(loss_r - loss_1p) <= 1e-6 * abs(loss_1p)          # f32 result of fp64 sums
        for g_id, d in grads_r.items():                              # gradients use the GLOBAL Σw
            assert np.abs(d - d_1p[g_id:g_id + 1]).max() <= 1e-5 * np.abs(d_1p).max()


def test_ownership_policies_balance():
    from cvcs_b200 import shard
    for world in (1, 2, 4, 8):
        counts = [len(shard.local_tiles(4, [10000, 10000], 1024, r, world)) for r in range(world)]
        assert sum(counts) == 4 * 81 and max(counts) - min(counts) <= 1      # cfg4: 81 tiles / scene, round-robin
        by_scene = [shard.local_tiles(8, [10000, 10000], 1024, r, world, "scene") for r in range(world)]
        assert sum(len(b) for b in by_scene) == 8 * 81
        for r, b in enumerate(by_scene):
            assert {t[1] % world for t in b} <= {r}
    assert shard.world_info() == (0, 1)
    t = torch.ones(3, dtype=torch.float32)
    with pytest.raises(TypeError):
        shard.all_reduce_sum_(t)


def _metric_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pickle
        from cvcs_b200 import metrics, shard
        # each rank's share of the tiles -> its own confusion matrix (oracle), restored into a metric object the way a
        # checkpoint restores it (host state), then ONE sync() per rank
        tiles = shard.local_tiles(N_SCENES, [H, W], P, rank, world, "round_robin")
        _, cm, _, _ = _partials(tiles, None)
        m = metrics.MulticlassConfusionMatrix(num_classes=C, ignore_index=None)
        m = pickle.loads(pickle.dumps(m))
        m._s["host"] = torch.from_numpy(cm.copy())
        before = m.compute().clone()
        m.sync()                                   # collective: every rank, exactly once
        q.put((rank, before.numpy(), m.compute().numpy(), m.view("true").compute().numpy()))
    finally:
        dist.destroy_process_group()


def test_metric_sync_is_a_collective_over_disjoint_shards():
    """MulticlassConfusionMatrix.sync(): the sum of the ranks' matrices (each rank evaluated its own tiles) equals the
    single-process matrix; eval_model does NOT call it unless asked (sync_ranks=False by default: an unsharded loader
    evaluated on every rank would otherwise be counted world_size times, and a rank-0-only validate() would hang)."""
    import inspect
    from cvcs_b200 import metrics, shard
    assert inspect.signature(metrics.eval_model).parameters["sync_ranks"].default is False
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_metric_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    all_tiles = shard.local_tiles(N_SCENES, [H, W], P, 0, 1, "round_robin")
    _, cm_single, _, _ = _partials(all_tiles, None)
    assert np.array_equal(out[0][1] + out[1][1], cm_single)          # the shards were disjoint and complete
    for _, _, synced, normed in out:
        assert np.array_equal(synced, cm_single)
        ref = cm_single / np.maximum(cm_single.sum(1, keepdims=True), 1)
        assert np.allclose(normed, ref.astype(np.float32))
