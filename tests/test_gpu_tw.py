"""K1 with the total weight computed inside the kernel (cvcs_ce_fused_tw): the label pre-pass + grid barrier must give
the gradients of the K4 -> K1 chain, and the cross-GPU exchange must divide by the sum of every rank's Σ v·w[y]
(SURVEY §8e collective (1); reference: the 'mean' of nn.CrossEntropyLoss over the whole batch, utils.py:230,238).

Several ranks on ONE GPU cannot wait for each other (nothing guarantees that their kernels run at the same time), so
the peers' parts of an exchange are played from the host (Exchange.poke) before the kernel is launched; the real
N-process path runs in scripts/shard_check.py under torchrun."""
import math

import numpy as np
import pytest
import torch

from oracle import c_oracle

pytestmark = pytest.mark.gpu

DEV = torch.device("cuda", 0)


def make_case(B, C, H, W, seed, dtype=torch.float32, ignore=255, frac_ignored=0.1, label_dtype=torch.uint8):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(B, C, H, W, generator=g) * 3).to(dtype)
    t = torch.randint(0, C, (B, H, W), generator=g, dtype=torch.uint8)
    t[torch.rand(B, H, W, generator=g) < frac_ignored] = ignore
    w = torch.rand(C, generator=g) + 0.1
    return x.to(DEV), t.to(label_dtype).to(DEV), w.to(DEV)


def chain(x, t, w, ii):
    """K4 -> K1 (the two-launch form)."""
    from cvcs_b200 import ops
    tw = torch.empty(2, dtype=torch.float64, device=DEV)
    ops.label_hist(t, x.shape[1], ii, weight=w, total_weight_out=tw)
    cm = torch.zeros((x.shape[1],) * 2, dtype=torch.int64, device=DEV)
    loss, sums, d = ops.ce_fused(x, t, w, ii, inv_total_weight_dev=tw[1:], confmat=cm)
    torch.cuda.synchronize()
    return float(loss), d.float().cpu().numpy(), tw.cpu().numpy(), cm.cpu().numpy()


def fused(x, t, w, ii, xchg=None):
    from cvcs_b200 import ops
    tw = torch.full((2,), -1.0, dtype=torch.float64, device=DEV)
    cm = torch.zeros((x.shape[1],) * 2, dtype=torch.int64, device=DEV)
    am = torch.empty(t.shape, dtype=torch.uint8, device=DEV)
    loss, sums, d = ops.ce_fused(x, t, w, ii, total_weight="kernel", xchg=xchg, total_weight_out=tw, confmat=cm, argmax=am)
    torch.cuda.synchronize()
    return float(loss), d.float().cpu().numpy(), tw.cpu().numpy(), cm.cpu().numpy()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("layout", ["NCHW", "NHWC"])
@pytest.mark.parametrize("shape", [(2, 7, 64, 64), (3, 16, 48, 80), (1, 20, 128, 96), (16, 7, 256, 256)])
def test_in_kernel_total_weight_equals_the_two_launch_chain(dtype, layout, shape):
    B, C, H, W = shape
    x, t, w = make_case(B, C, H, W, seed=C * 7 + B, dtype=dtype)
    if layout == "NHWC":
        x = x.contiguous(memory_format=torch.channels_last)
    l0, g0, tw0, cm0 = chain(x, t, w, 255)
    l1, g1, tw1, cm1 = fused(x, t, w, 255)
    assert abs(tw1[0] - tw0[0]) <= 1e-6 * abs(tw0[0]) and abs(tw1[1] * tw1[0] - 1.0) < 1e-12
    assert l1 == l0 or abs(l1 - l0) <= 1e-7 * abs(l0)
    tol = 2e-6 if dtype == torch.float32 else 8e-3          # bf16: one rounding step of the stored gradient
    assert np.abs(g1 - g0).max() <= tol * np.abs(g0).max()
    assert np.array_equal(cm0, cm1)
    # and against the oracle (fp64 restatement)
    xf = x.float().cpu().contiguous().numpy()
    l_ref, _, g_ref = c_oracle.cross_entropy(xf, t.cpu().numpy().astype(np.int64), w.cpu().numpy(), 255)
    assert abs(l1 - l_ref) <= 1e-5 * abs(l_ref)
    assert np.abs(g1 - g_ref).max() <= (1e-5 if dtype == torch.float32 else 1e-2) * np.abs(g_ref).max()


def test_back_to_back_launches_leave_the_workspace_clean():
    """The grid barrier counter and the per-CTA partials are reused by every launch on the stream."""
    x, t, w = make_case(4, 7, 128, 128, seed=3)
    ref = fused(x, t, w, 255)
    for _ in range(5):
        out = fused(x, t, w, 255)
        assert out[0] == ref[0] and np.array_equal(out[1], ref[1]) and np.array_equal(out[2], ref[2])
    # the K4 -> K1 chain still works on the same workspace afterwards
    l0, g0, _, _ = chain(x, t, w, 255)
    assert abs(l0 - ref[0]) <= 1e-7 * abs(ref[0])


@pytest.mark.parametrize("case", ["odd_size", "int64_labels", "wide_C", "all_ignored"])
def test_shapes_outside_the_staged_kernel_fall_back_to_two_launches(case):
    if case == "odd_size":
        x, t, w = make_case(2, 7, 9, 7, seed=5)
    elif case == "int64_labels":
        x, t, w = make_case(2, 7, 64, 64, seed=6, label_dtype=torch.int64)
    elif case == "wide_C":
        x, t, w = make_case(1, 33, 32, 32, seed=7)
    else:
        x, t, w = make_case(2, 7, 64, 64, seed=8, frac_ignored=1.1)
    l0, g0, tw0, cm0 = chain(x, t, w, 255)
    l1, g1, tw1, cm1 = fused(x, t, w, 255)
    if case == "all_ignored":
        assert math.isnan(l0) and math.isnan(l1) and tw1[0] == 0.0 and np.all(g1 == 0)   # exact zeros at ignored pixels
    else:
        assert abs(l1 - l0) <= 1e-7 * abs(l0) and np.abs(g1 - g0).max() <= 2e-6 * np.abs(g0).max()
    assert np.array_equal(cm0, cm1)


class Ranks:
    """This process as rank `rank` of `world`; the other ranks' exchange blocks exist (the kernel stores into them) but
    their kernels never run — their values are played into the local block with poke()."""

    def __init__(self, world, rank):
        from cvcs_b200 import ops
        self.me = ops.Exchange(world, rank, DEV)
        self.others = {q: ops.Exchange(world, q, DEV) for q in range(world) if q != rank}
        for q, o in self.others.items():
            self.me.set_peer(q, o)

    def close(self):
        self.me.close()
        for o in self.others.values():
            o.close()


def test_exchange_adds_the_peers_total_in_rank_order():
    """Two ranks, this process is rank 0; rank 1's Σw arrives in the local block before the launch (played from the host)."""
    from cvcs_b200 import ops
    x, t, w = make_case(4, 7, 128, 128, seed=11)
    l0, g0, _, _ = chain(x, t, w, 255)
    _, _, tw0, _ = fused(x, t, w, 255)                      # this rank's Σw as the kernel's own pre-pass sums it
    rk = Ranks(2, 0)
    xc = rk.me
    try:
        for seq, peer_value in ((1, 12345.678), (2, 0.25 * tw0[0]), (3, 7.0)):
            xc.poke(1, seq, peer_value)
            l1, g1, tw1, _ = fused(x, t, w, 255, xchg=xc)
            total = tw0[0] + peer_value                      # rank order: own value first, then rank 1's
            assert tw1[0] == total and tw1[1] == 1.0 / total
            assert abs(l1 - l0) <= 1e-7 * abs(l0)            # the loss of THIS rank's batch is unaffected
            assert np.abs(g1 - g0 * (tw0[0] / total)).max() <= 2e-6 * np.abs(g0).max() * (tw0[0] / total)
            assert xc.state() == (seq, 0)
    finally:
        rk.close()


def test_exchange_as_rank_one_and_three_ranks():
    from cvcs_b200 import ops
    x, t, w = make_case(2, 7, 64, 64, seed=12)
    _, g0, _, _ = chain(x, t, w, 255)
    _, _, tw0, _ = fused(x, t, w, 255)
    rk = Ranks(3, 1)
    xc = rk.me
    try:
        xc.poke(0, 1, 100.0)
        xc.poke(2, 1, 0.5)
        _, g1, tw1, _ = fused(x, t, w, 255, xchg=xc)
        total = (100.0 + tw0[0]) + 0.5                       # ranks 0, 1, 2 in that order
        assert tw1[0] == total
        assert np.abs(g1 - g0 * (tw0[0] / total)).max() <= 2e-6 * np.abs(g0).max() * (tw0[0] / total)
    finally:
        rk.close()


def test_a_missing_peer_times_out_with_nan_instead_of_hanging():
    from cvcs_b200 import ops
    x, t, w = make_case(1, 7, 64, 64, seed=13)
    rk = Ranks(2, 0)
    xc = rk.me
    try:
        l1, g1, tw1, _ = fused(x, t, w, 255, xchg=xc)        # rank 1 never shows up: ~4 s, then NaN
        assert math.isnan(tw1[0]) and np.isnan(g1[np.broadcast_to((t.cpu().numpy() != 255)[:, None], g1.shape)]).all()
        seq, errors = xc.state()
        assert seq == 1 and errors >= 1
    finally:
        rk.close()


def test_two_handles_of_one_process_wired_together():
    """set_peer: both blocks live in this process (the single-process multi-GPU form); rank 1 runs after rank 0 was
    played, rank 0's own store into rank 1's block is checked through rank 1's result."""
    from cvcs_b200 import ops
    x, t, w = make_case(2, 7, 64, 64, seed=14)
    _, _, tw0, _ = fused(x, t, w, 255)
    a, b = ops.Exchange(2, 0, DEV), ops.Exchange(2, 1, DEV)
    try:
        a.set_peer(1, b)
        b.set_peer(0, a)
        a.poke(1, 1, 3.0)                                    # rank 1's value, ahead of time, for rank 0's wait
        _, _, twa, _ = fused(x, t, w, 255, xchg=a)           # rank 0 also stores ITS value into rank 1's block
        assert twa[0] == tw0[0] + 3.0
        _, _, twb, _ = fused(x, t, w, 255, xchg=b)           # rank 1 finds rank 0's value there
        assert twb[0] == tw0[0] + tw0[0]
    finally:
        a.close()
        b.close()


def test_local_total_weight_given_only_the_exchange_runs_in_the_kernel():
    """tw_mode 2: this rank's Σw comes from a K4 launch (any label dtype); the kernel skips its pre-pass and exchanges."""
    from cvcs_b200 import ops
    for label_dtype in (torch.uint8, torch.int64):
        x, t, w = make_case(3, 7, 96, 64, seed=21, label_dtype=label_dtype)
        l0, g0, tw0, cm0 = chain(x, t, w, 255)
        local = torch.tensor([tw0[0]], dtype=torch.float64, device=DEV)
        # single GPU: identical to the chain
        tw = torch.zeros(2, dtype=torch.float64, device=DEV)
        cm = torch.zeros((7, 7), dtype=torch.int64, device=DEV)
        loss, _, d = ops.ce_fused(x, t, w, 255, total_weight="kernel", local_total_weight=local, total_weight_out=tw, confmat=cm)
        torch.cuda.synchronize()
        assert tw.cpu().numpy()[0] == tw0[0] and float(loss) == l0 and np.array_equal(d.float().cpu().numpy(), g0)
        assert np.array_equal(cm.cpu().numpy(), cm0)
        # two ranks: rank 1's value played from the host
        rk = Ranks(2, 0)
        try:
            rk.me.poke(1, 1, 1000.0)
            loss, _, d = ops.ce_fused(x, t, w, 255, total_weight="kernel", xchg=rk.me, local_total_weight=local, total_weight_out=tw)
            torch.cuda.synchronize()
            total = tw0[0] + 1000.0
            assert tw.cpu().numpy()[0] == total and rk.me.state() == (1, 0)
            g1 = d.float().cpu().numpy()
            assert np.abs(g1 - g0 * (tw0[0] / total)).max() <= 2e-6 * np.abs(g0).max() * (tw0[0] / total)
        finally:
            rk.close()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_next_batch_scan_pipelines_the_pre_pass_across_launches(dtype):
    """Launch i also sums the weights over batch i+1's labels; launch i+1 takes that as its local total weight.  Every
    step must equal the plain K4 -> K1 chain on the same batch (bit-identical Σw hand-over, gradients to rounding)."""
    from cvcs_b200 import ops
    batches = [make_case(4, 7, 128, 128, seed=30 + i, dtype=dtype, frac_ignored=0.05 * (i + 1)) for i in range(4)]
    w = batches[0][2]
    nxt = [torch.zeros(2, dtype=torch.float64, device=DEV) for _ in range(2)]
    tw = torch.zeros(2, dtype=torch.float64, device=DEV)
    # first step: nothing was scanned ahead -> the kernel's own pre-pass, and it scans batch 1
    for i, (x, t, _) in enumerate(batches):
        t_next = batches[i + 1][1] if i + 1 < len(batches) else None
        local = None if i == 0 else nxt[i % 2][0:1]
        loss, _, d = ops.ce_fused(x, t, w, 255, total_weight="kernel", local_total_weight=local, total_weight_out=tw,
                                  next_target=t_next, next_total_weight_out=nxt[(i + 1) % 2] if t_next is not None else None)
        torch.cuda.synchronize()
        l0, g0, tw0, _ = chain(x, t, w, 255)
        assert abs(float(tw[0]) - tw0[0]) <= 1e-6 * tw0[0]
        assert abs(float(loss) - l0) <= 1e-7 * abs(l0)
        tol = 2e-6 if dtype == torch.float32 else 8e-3
        assert np.abs(d.float().cpu().numpy() - g0).max() <= tol * np.abs(g0).max()
        if t_next is not None:
            # what the next launch will divide by: this rank's sum over the NEXT labels, as K4 computes it
            twn = torch.zeros(2, dtype=torch.float64, device=DEV)
            ops.label_hist(t_next, 7, 255, weight=w, total_weight_out=twn)
            torch.cuda.synchronize()
            assert abs(float(nxt[(i + 1) % 2][0]) - float(twn[0])) <= 1e-6 * float(twn[0])
            assert abs(float(nxt[(i + 1) % 2][1]) * float(nxt[(i + 1) % 2][0]) - 1.0) < 1e-12
    # forward-only call in the middle of a pipelined sequence still produces the next sum (as a launch of its own)
    x, t, _ = batches[0]
    out = torch.zeros(2, dtype=torch.float64, device=DEV)
    ops.ce_fused(x, t, w, 255, want_grad=False, total_weight="kernel", next_target=batches[1][1], next_total_weight_out=out)
    torch.cuda.synchronize()
    twn = torch.zeros(2, dtype=torch.float64, device=DEV)
    ops.label_hist(batches[1][1], 7, 255, weight=w, total_weight_out=twn)
    torch.cuda.synchronize()
    assert float(out[0]) == float(twn[0])


def test_pipelined_sequence_publishes_the_next_sum_one_step_ahead():
    """With an exchange AND next_target, the last CTA of launch i sends this rank's sum for exchange i+1 to the peers as
    the launch ends; launch i+1 must not publish again and must still divide by (own + peers') of ITS batch."""
    from cvcs_b200 import ops
    batches = [make_case(2, 7, 128, 128, seed=40 + i, frac_ignored=0.1 * (i + 1)) for i in range(3)]
    w = batches[0][2]
    own = []
    for x, t, _ in batches:
        twn = torch.zeros(2, dtype=torch.float64, device=DEV)
        ops.label_hist(t, 7, 255, weight=w, total_weight_out=twn)
        torch.cuda.synchronize()
        own.append(float(twn[0]))
    peer = [111.0, 2222.0, 33333.0]
    rk = Ranks(2, 0)
    try:
        nxt = [torch.zeros(2, dtype=torch.float64, device=DEV) for _ in range(2)]
        tw = torch.zeros(2, dtype=torch.float64, device=DEV)
        for i, (x, t, _) in enumerate(batches):
            rk.me.poke(1, i + 1, peer[i])                      # rank 1's value of exchange i+1
            t_next = batches[i + 1][1] if i + 1 < len(batches) else None
            _, _, d = ops.ce_fused(x, t, w, 255, total_weight="kernel", xchg=rk.me, total_weight_out=tw,
                                   local_total_weight=None if i == 0 else nxt[i % 2][0:1],
                                   next_target=t_next, next_total_weight_out=nxt[(i + 1) % 2] if t_next is not None else None)
            torch.cuda.synchronize()
            total = float(tw[0])
            assert abs(total - (own[i] + peer[i])) <= 1e-6 * total
            _, g0, tw0, _ = chain(x, t, w, 255)
            g1 = d.float().cpu().numpy()
            assert np.abs(g1 - g0 * (tw0[0] / total)).max() <= 3e-6 * np.abs(g0).max() * (tw0[0] / total)
            assert rk.me.state() == (i + 1, 0)
            if t_next is not None:
                # rank 1's block (the dummy peer) already holds this rank's value for the NEXT exchange
                blk = rk.others[1]
                assert blk is not None
    finally:
        rk.close()


def test_single_rank_allreduce_is_the_identity():
    from cvcs_b200 import ops
    xc = ops.Exchange(1, 0, DEV)
    try:
        v = torch.arange(100, dtype=torch.float64, device=DEV)
        assert torch.equal(xc.allreduce_(v.clone()), v)
        with pytest.raises(RuntimeError):
            xc.allreduce_(torch.zeros(10, device=DEV))            # float32: refused
    finally:
        xc.close()
