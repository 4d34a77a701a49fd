"""The fused path inside CUDA graphs: the C-ABI entry points take a stream, never synchronise and never allocate, so
the label pre-pass (K4) -> fused CE (K1) chain, the metrics mode and the nn.Module's forward + backward can be
captured once and replayed on new batch contents.  Every replay is checked against the oracle (torch CPU path,
train.py:122-125 / utils.py:88-94), not against an eager run of the same kernels."""
import pytest
import torch

from oracle import torch_path

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _batch(seed, B, C, H, W, dtype=torch.float32, ignore=None):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn((B, C, H, W), generator=g) * 3).to(dtype)
    y = torch.randint(0, C, (B, H, W), generator=g, dtype=torch.uint8)
    if ignore is not None:
        y[torch.rand((B, H, W), generator=g) < 0.1] = ignore
    return x, y


@pytest.mark.parametrize("shape,dtype,weighted,ignore", [
    ((4, 7, 64, 64), torch.float32, False, -100),      # TMA-staged kernel
    ((4, 7, 64, 64), torch.bfloat16, True, 255),
    ((3, 7, 37, 41), torch.float32, True, 0),          # generic kernel (hw not a multiple of 16)
])
def test_prepass_and_fused_ce_replay(shape, dtype, weighted, ignore):
    from cvcs_b200 import ops
    B, C, H, W = shape
    w = torch.linspace(0.5, 2.0, C) if weighted else None
    w_dev = None if w is None else w.to(DEV)
    x_s = torch.zeros(shape, dtype=dtype, device=DEV)
    y_s = torch.zeros((B, H, W), dtype=torch.uint8, device=DEV)
    tw = torch.zeros(2, dtype=torch.float64, device=DEV)
    d_s = torch.zeros_like(x_s)
    am = torch.zeros((B, H, W), dtype=torch.uint8, device=DEV)
    cm = torch.zeros((C, C), dtype=torch.int64, device=DEV)
    sums = torch.zeros(3, dtype=torch.float64, device=DEV)
    loss = torch.zeros(1, dtype=torch.float32, device=DEV)

    def step():
        ops.label_hist(y_s, C, ignore, hist=None, weight=w_dev, total_weight_out=tw)
        ops.ce_fused(x_s, y_s, w_dev, ignore, want_grad=True, inv_total_weight_dev=tw[1:], dlogits=d_s, argmax=am,
                     confmat=cm, loss_sums=sums, loss_out=loss)

    side = torch.cuda.Stream(DEV)
    side.wait_stream(torch.cuda.current_stream(DEV))
    with torch.cuda.stream(side):
        step()                                                            # warm-up outside the capture
    torch.cuda.current_stream(DEV).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    cm.zero_()
    cm_ref = torch.zeros((C, C), dtype=torch.int64)
    for seed in (1, 2, 3):
        x, y = _batch(seed, B, C, H, W, dtype, ignore if 0 <= ignore <= 255 else None)
        x_s.copy_(x)
        y_s.copy_(y)
        graph.replay()
        torch.cuda.synchronize()
        ref_loss, ref_grad = torch_path.ce_loss_and_grad(x.float(), y, w, ignore)
        ref_arg = torch.max(x.float(), dim=1)[1]                           # utils.py:90, all tiles at once
        keep = (y != ignore) if 0 <= ignore <= 255 else torch.ones_like(y, dtype=torch.bool)
        t, p = y[keep].long(), ref_arg[keep].long()                       # K1 drops pixels whose label is ignore_index
        cm_ref += torch.bincount(t * C + p, minlength=C * C).reshape(C, C)
        tol = 1e-5 if dtype == torch.float32 else 1e-2
        assert torch.equal(am.cpu(), ref_arg.to(torch.uint8))
        assert torch.equal(cm.cpu(), cm_ref)
        torch.testing.assert_close(loss.cpu()[0], ref_loss, rtol=tol, atol=tol * 1e-2)
        torch.testing.assert_close(d_s.float().cpu(), ref_grad, rtol=tol, atol=tol * 1e-3)


def test_metrics_mode_replay():
    from cvcs_b200 import ops
    B, C, H, W = 4, 16, 64, 64
    x_s = torch.zeros((B, C, H, W), device=DEV)
    y_s = torch.zeros((B, H, W), dtype=torch.uint8, device=DEV)
    am = torch.zeros((B, H, W), dtype=torch.uint8, device=DEV)
    cm = torch.zeros((C, C), dtype=torch.int64, device=DEV)
    ops.eval_fused(x_s, y_s, 0, argmax=am, confmat=cm)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ops.eval_fused(x_s, y_s, 0, argmax=am, confmat=cm)
    cm.zero_()
    flat_ref = torch.zeros((C, C), dtype=torch.int64)
    for seed in (4, 5):
        x, y = _batch(seed, B, C, H, W)
        x_s.copy_(x)
        y_s.copy_(y)
        graph.replay()
        torch.cuda.synchronize()
        flat, _, preds = torch_path.eval_tiles(x, y, C, True, double_update=False)   # utils.py:85-94, background ignored
        flat_ref += flat.compute()
        assert torch.equal(am.cpu(), preds.to(torch.uint8))
        assert torch.equal(cm.cpu(), flat_ref)


def test_module_forward_backward_replay():
    """``loss = crit(logits, mask.long()); loss.backward()`` (train.py:122-125) captured whole."""
    from cvcs_b200.loss import FusedCrossEntropyLoss
    B, C, H, W = 2, 7, 64, 64
    w = torch.linspace(0.5, 2.0, C)
    crit = FusedCrossEntropyLoss(weight=w.to(DEV), ignore_index=0)
    x_s = torch.zeros((B, C, H, W), device=DEV, requires_grad=True)
    y_s = torch.zeros((B, H, W), dtype=torch.int64, device=DEV)
    side = torch.cuda.Stream(DEV)
    side.wait_stream(torch.cuda.current_stream(DEV))
    with torch.cuda.stream(side):
        for _ in range(3):
            x_s.grad = None
            crit(x_s, y_s).backward()
    torch.cuda.current_stream(DEV).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    x_s.grad = None
    with torch.cuda.graph(graph):
        loss_s = crit(x_s, y_s)
        loss_s.backward()
    for seed in (6, 7):
        x, y = _batch(seed, B, C, H, W)
        with torch.no_grad():
            x_s.copy_(x)
        y_s.copy_(y)
        graph.replay()
        torch.cuda.synchronize()
        ref_loss, ref_grad = torch_path.ce_loss_and_grad(x, y, w, 0)
        torch.testing.assert_close(loss_s.detach().cpu(), ref_loss, rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(x_s.grad.cpu(), ref_grad, rtol=1e-5, atol=1e-8)


def test_one_workspace_per_capture():
    """All the calls captured into one graph share ONE zeroed workspace (keyed by the capture's sequence number): a single
    memset node per graph, consecutive K1 launches directly connected.  Three K1 steps in one graph, replayed twice, give
    the oracle's matrices and losses; a second capture on the same stream gets a workspace of its own."""
    from cvcs_b200 import ops
    B, C, H, W = 2, 7, 64, 64
    batches = [_batch(20 + i, B, C, H, W) for i in range(3)]
    xs = [x.to(DEV) for x, _ in batches]
    ys = [y.to(DEV) for _, y in batches]
    ds = [torch.zeros_like(x) for x in xs]
    am = torch.zeros((B, H, W), dtype=torch.uint8, device=DEV)
    cm = torch.zeros((C, C), dtype=torch.int64, device=DEV)
    sums = torch.zeros((3, 3), dtype=torch.float64, device=DEV)
    loss = torch.zeros(1, dtype=torch.float32, device=DEV)
    inv = 1.0 / (B * H * W)

    def steps():
        for i in range(3):
            ops.ce_fused(xs[i], ys[i], None, -100, want_grad=True, inv_total_weight=inv, dlogits=ds[i], argmax=am, confmat=cm,
                         loss_sums=sums[i], loss_out=loss)

    steps()
    torch.cuda.synchronize()
    before = len(ops._capture_workspaces)
    g1 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g1):
        steps()
    assert len(ops._capture_workspaces) == before + 1
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        steps()
    assert len(ops._capture_workspaces) == before + 2
    cm.zero_()
    cm_ref = torch.zeros((C, C), dtype=torch.int64)
    for graph in (g1, g2, g1):
        graph.replay()
        torch.cuda.synchronize()
        for i, (x, y) in enumerate(batches):
            ref_loss, ref_grad = torch_path.ce_loss_and_grad(x, y, None, -100)
            pred = torch.max(x, dim=1)[1]
            cm_ref += torch.bincount(y.long().reshape(-1) * C + pred.reshape(-1), minlength=C * C).reshape(C, C)
            torch.testing.assert_close(ds[i].cpu(), ref_grad, rtol=1e-5, atol=1e-8)
            torch.testing.assert_close((sums[i, 0] / sums[i, 1]).float().cpu(), ref_loss, rtol=1e-5, atol=1e-7)
        assert torch.equal(cm.cpu(), cm_ref)
