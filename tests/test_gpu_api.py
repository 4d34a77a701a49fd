"""The reference-facing Python layer on the GPU: load_loss / FusedCrossEntropyLoss as a drop-in for
nn.CrossEntropyLoss (utils.py:223-242, train.py:122-125), MulticlassConfusionMatrix + eval_model
(utils.py:59-103), validation_loss (utils.py:106-126), the host-buffer C-ABI entry point."""
import io
import os
import pickle

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import torch_path

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


class TinyNet(nn.Module):
    requires_context = False
    returns_logits = True

    def __init__(self, classes):
        super().__init__()
        self.conv = nn.Conv2d(3, classes, 3, padding=1)

    def forward(self, x, context=None):
        return self.conv(x)


@pytest.mark.parametrize("loss_name,ib", [("CEL", False), ("CEL", True), ("wCEL", False), ("wCEL", True)])
def test_load_loss_dropin_training_step(loss_name, ib):
    """Same config keys as the reference; loss and parameter gradients equal the torch path's."""
    from cvcs_b200.loss import load_loss
    torch.manual_seed(0)
    C = 7
    cfg = {"num_classes": C - 1, "loss": loss_name, "ignore_background": ib}

    class DS:
        def get_class_weights(self, classes, ignore_background):
            counts = torch.tensor([100, 50, 0, 25, 25, 10, 5], dtype=torch.float32)
            return torch_path.class_weights(counts, ignore_background)

    crit = load_loss(cfg, DEV, DS())
    w = DS().get_class_weights(C, ib) if loss_name == "wCEL" else None
    ref_crit = torch_path.make_criterion(w, 0 if ib else -100)

    net = TinyNet(C).to(DEV)
    net_ref = TinyNet(C)
    net_ref.load_state_dict({k: v.cpu() for k, v in net.state_dict().items()})
    img = torch.randint(0, 256, (4, 3, 32, 32), dtype=torch.uint8)
    mask = torch.randint(0, C, (4, 32, 32), dtype=torch.uint8)

    torch.backends.cudnn.allow_tf32 = False                      # compare against fp32 CPU convolutions
    out = net(img.to(DEV).type(torch.float32))
    out.retain_grad()
    loss = crit(out, mask.to(DEV).type(torch.long))              # train.py:122
    val = loss.item()
    loss.backward()                                              # train.py:125
    out_ref = net_ref(img.type(torch.float32))
    out_ref.retain_grad()
    loss_ref = ref_crit(out_ref, mask.type(torch.long))
    loss_ref.backward()
    assert loss.shape == () and loss.dtype == torch.float32
    assert abs(val - loss_ref.item()) <= 1e-5 * abs(loss_ref.item())
    # the hot path proper: loss and dlogits against torch's CE on the SAME logits values
    same = out.detach().cpu().requires_grad_(True)
    loss_same = ref_crit(same, mask.type(torch.long))
    loss_same.backward()
    assert abs(val - loss_same.item()) <= 1e-5 * abs(loss_same.item())
    assert float((out.grad.cpu() - same.grad).abs().max()) <= 1e-5 * float(same.grad.abs().max())
    # and what autograd makes of it upstream (GPU vs CPU convolution backward: summation order differs)
    for p, q in zip(net.parameters(), net_ref.parameters()):
        assert torch.allclose(p.grad.cpu(), q.grad, rtol=1e-3, atol=1e-5 * float(q.grad.abs().max()))
    # uint8 masks are accepted as stored (no .long() copy) and give the identical result
    net.zero_grad()
    loss_u8 = crit(net(img.to(DEV).type(torch.float32)), mask.to(DEV))
    assert loss_u8.item() == val
    # forward-only under no_grad (utils.validation_loss)
    with torch.no_grad():
        assert crit(out.detach(), mask.to(DEV).long()).item() == val


def test_grad_output_scaling_and_modes():
    from cvcs_b200.loss import FusedCrossEntropyLoss
    torch.manual_seed(1)
    x = torch.randn(2, 7, 16, 16, device=DEV, requires_grad=True)
    t = torch.randint(0, 7, (2, 16, 16), device=DEV)
    xr = x.detach().cpu().requires_grad_(True)
    (nn.CrossEntropyLoss()(xr, t.cpu()) * 3.5).backward()
    for mode in ("check", "scale"):
        x.grad = None
        (FusedCrossEntropyLoss(grad_scale_mode=mode)(x, t) * 3.5).backward()
        assert torch.allclose(x.grad.cpu(), xr.grad, rtol=1e-5, atol=1e-9)
    x.grad = None
    FusedCrossEntropyLoss(grad_scale_mode="unit")(x, t).backward()
    assert torch.allclose(x.grad.cpu() * 3.5, xr.grad, rtol=1e-5, atol=1e-9)


def test_input_conventions_and_errors():
    from cvcs_b200.loss import FusedCrossEntropyLoss
    torch.manual_seed(2)
    crit = FusedCrossEntropyLoss()
    # [N, C] with [N]
    x = torch.randn(64, 5, device=DEV, requires_grad=True)
    t = torch.randint(0, 5, (64,), device=DEV)
    l = crit(x, t)
    l.backward()
    xr = x.detach().cpu().requires_grad_(True)
    lr = nn.CrossEntropyLoss()(xr, t.cpu())
    lr.backward()
    assert abs(l.item() - lr.item()) <= 1e-5 * lr.item()
    assert torch.allclose(x.grad.cpu(), xr.grad, rtol=1e-5, atol=1e-8)
    # channels_last logits: same values, gradient keeps the memory format
    y = torch.randn(2, 7, 16, 16, device=DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    ty = torch.randint(0, 7, (2, 16, 16), device=DEV)
    crit(y, ty).backward()
    yr = y.detach().cpu().contiguous().requires_grad_(True)
    nn.CrossEntropyLoss()(yr, ty.cpu()).backward()
    assert torch.allclose(y.grad.cpu(), yr.grad, rtol=1e-5, atol=1e-8)
    # dtype / shape complaints as torch makes them
    with pytest.raises(RuntimeError, match="expected scalar type Long"):
        crit(y, ty.to(torch.int32))
    with pytest.raises(RuntimeError, match="size mismatch"):
        crit(y, ty[:, :8])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        crit(y.detach().cpu(), ty.cpu())
    # out-of-bounds label: NaN-poisoned loss, IndexError on check / in strict mode
    tb = ty.clone()
    tb[0, 0, 0] = 9
    lb = crit(y.detach(), tb)
    assert torch.isnan(lb)
    with pytest.raises(IndexError):
        crit.check_errors()
    with pytest.raises(IndexError):
        FusedCrossEntropyLoss(strict=True)(y.detach(), tb)


def test_bf16_logits_match_fp32_on_same_values():
    from cvcs_b200.loss import FusedCrossEntropyLoss
    torch.manual_seed(3)
    w = torch.rand(7) + 0.5
    x = torch.randn(2, 7, 32, 32).to(torch.bfloat16)
    t = torch.randint(0, 7, (2, 32, 32))
    t[0, :4] = 255
    xd = x.to(DEV).requires_grad_(True)
    crit = FusedCrossEntropyLoss(weight=w, ignore_index=255)
    l = crit(xd, t.to(DEV))
    assert l.dtype == torch.bfloat16
    l.backward()
    l_ref, g_ref = torch_path.ce_loss_and_grad(x.float(), t, w, 255)
    assert abs(float(crit.last_sums[0] / crit.last_sums[1]) - l_ref.item()) <= 1e-5 * l_ref.item()
    assert abs(l.float().item() - l_ref.item()) <= 1e-2 * l_ref.item()
    assert (xd.grad.float().cpu() - g_ref).abs().max() <= 1e-2 * g_ref.abs().max()


def test_confusion_metric_dropin(golden):
    from cvcs_b200.metrics import MulticlassConfusionMatrix, print_metrics
    g = golden("eval_cases")
    x, y = torch.from_numpy(g["logits"]), torch.from_numpy(g["labels"])
    for ib in (0, 1):
        ii = 0 if ib else None
        flat = MulticlassConfusionMatrix(num_classes=16, ignore_index=ii)
        normalized = MulticlassConfusionMatrix(num_classes=16, normalize="true", ignore_index=ii)
        for i in range(x.shape[0]):                              # the reference's loop, CPU index tensors (utils.py:90-94)
            _, pred = torch.max(x[i], dim=0)
            p = pred.unsqueeze(0).type(torch.int64).reshape(1, -1)
            t = y[i:i + 1].type(torch.int64).reshape(1, -1)
            normalized.update(p, t)
            flat.update(p, t)
        cm = flat.compute()
        assert cm.dtype == torch.int64 and np.array_equal(cm.numpy(), g[f"ib{ib}.flat"])
        assert np.array_equal(normalized.compute().numpy(), g[f"ib{ib}.normalized"])
        m = print_metrics(cm, silent=True)
        assert np.array_equal(np.array(m["perclass_IoU"]), g[f"ib{ib}.perclass_IoU"])
        assert [m["mIoU"], m["precision_score"], m["recall_score"], m["dice_score"], m["oa_score"]] == list(g[f"ib{ib}.scalars"])
        # logits straight in (fused argmax + update), and a pickle round trip (checkpoints, utils.py:139-140)
        fused = MulticlassConfusionMatrix(num_classes=16, ignore_index=ii)
        fused.update(x[:3].to(DEV), y[:3].to(DEV))
        buf = io.BytesIO()
        torch.save({"conf_flat": [fused]}, buf)
        buf.seek(0)
        restored = torch.load(buf, weights_only=False)["conf_flat"][0]
        restored.update_from_logits(x[3:].to(DEV), y[3:].to(DEV))
        assert np.array_equal(restored.compute().numpy(), g[f"ib{ib}.flat"])
        assert pickle.loads(pickle.dumps(flat)).compute().equal(cm)


class _Chunk(torch.utils.data.IterableDataset):
    def __init__(self, items):
        self.patches, self.chunk_crops = items, list(range(len(items)))

    def __iter__(self):
        return iter(self.patches)


class _Loader:
    def __init__(self, chunks):
        self.chunks = chunks

    def __len__(self):
        return len(self.chunks)

    def get_iterable_chunk(self, c):
        return _Chunk(self.chunks[c])


class _ReplayNet(nn.Module):
    requires_context, returns_logits = False, True

    def __init__(self, logits):
        super().__init__()
        self.logits, self.i = logits, 0

    def forward(self, x, context=None):
        out = self.logits[self.i:self.i + x.shape[0]]
        self.i += x.shape[0]
        return out


@pytest.mark.parametrize("batch_size", [1, 2, 3])
@pytest.mark.parametrize("ib", [False, True])
def test_eval_model_dropin(golden, batch_size, ib):
    """The reference's eval_model outputs (generated by running it) — here with any batch size."""
    from cvcs_b200.metrics import eval_model, print_metrics, validation_loss
    from cvcs_b200.loss import FusedCrossEntropyLoss
    g = golden("eval_cases")
    logits, labels = torch.from_numpy(g["logits"]).to(DEV), torch.from_numpy(g["labels"])
    items = [(torch.zeros(3, 16, 16, dtype=torch.uint8), labels[i], torch.tensor([0]), torch.tensor([0])) for i in range(6)]
    loader = _Loader([items[:3], items[3:]])
    flat, normalized = eval_model(_ReplayNet(logits), loader, DEV, batch_size=batch_size, ignore_background=ib)
    tag = f"ib{int(ib)}"
    cm = flat.compute()
    assert np.array_equal(cm.numpy(), g[f"{tag}.flat"])
    assert np.array_equal(normalized.compute().numpy(), g[f"{tag}.normalized"])
    m = print_metrics(cm, silent=True)
    assert m["mIoU"] == g[f"{tag}.scalars"][0]                   # bit-exact mIoU
    torch.save({"conf_flat": [flat], "conf_normalized": [normalized]}, io.BytesIO())
    # validation_loss: per-batch values equal the torch criterion's
    crit = FusedCrossEntropyLoss(ignore_index=0 if ib else -100)
    vals = validation_loss(_ReplayNet(logits), loader, crit, DEV, batch_size)
    ref_crit = torch_path.make_criterion(None, 0 if ib else -100)
    ref, lc = [], logits.cpu()
    for chunk in ([0, 1, 2], [3, 4, 5]):
        for k in range(0, 3, batch_size):
            idx = chunk[k:k + batch_size]
            ref.append(ref_crit(lc[idx], labels[idx].long()).item())
    assert len(vals) == len(ref)
    assert np.allclose(vals, ref, rtol=1e-5, atol=0)


def test_fused_training_extras():
    """One pass: loss + dlogits + argmax + confusion update."""
    from cvcs_b200.loss import FusedCrossEntropyLoss
    from cvcs_b200.metrics import MulticlassConfusionMatrix
    torch.manual_seed(5)
    cmx = MulticlassConfusionMatrix(num_classes=7, ignore_index=0)
    crit = FusedCrossEntropyLoss(ignore_index=0, confusion=cmx, return_argmax=True)
    x = torch.randn(3, 7, 32, 32, device=DEV, requires_grad=True)
    t = torch.randint(0, 7, (3, 32, 32), device=DEV, dtype=torch.uint8)
    crit(x, t).backward()
    _, am = torch.max(x.detach().cpu(), dim=1)
    assert torch.equal(crit.last_argmax.cpu().long(), am)
    ref = torch_path.RestatedConfusionMatrix(7, ignore_index=0)
    ref.update(am, t.cpu().long())
    assert torch.equal(cmx.compute(), ref.compute())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("pinned", [True, False])
def test_host_buffer_entry_point(dtype, pinned):
    """cvcs_host_ce_fused: host pointers in, results in host memory (what a non-torch caller binds)."""
    from cvcs_b200 import ops
    g = torch.Generator().manual_seed(9)
    B, C, H, W = 5, 7, 64, 64
    x = (torch.randn(B, C, H, W, generator=g) * 2).to(dtype)
    t = torch.randint(0, C, (B, H, W), generator=g, dtype=torch.uint8)
    t[1, :3] = 255
    w = torch.rand(C, generator=g) + 0.5
    if pinned:
        x, t = x.pin_memory(), t.pin_memory()
    ctx = ops.HostContext(0, B * H * W, C, dtype)
    d = torch.empty_like(x)
    am = torch.empty((B, H, W), dtype=torch.uint8)
    cm = torch.zeros((C, C), dtype=torch.int64)
    loss, sums = ctx.ce_fused(x, t, w, 255, want_grad=True, dlogits=d, argmax=am, confmat=cm)
    l_ref, g_ref = torch_path.ce_loss_and_grad(x.float(), t.long(), w, 255)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert abs(loss.item() - l_ref.item()) <= 1e-5 * l_ref.item()
    assert (d.float() - g_ref).abs().max() <= tol * g_ref.abs().max()
    _, am_ref = torch.max(x.float(), dim=1)
    assert torch.equal(am.long(), am_ref)
    ref = torch_path.RestatedConfusionMatrix(C, ignore_index=255)
    ref.update(am_ref, t.long())
    assert torch.equal(cm, ref.compute())
    # results may stay on the device (training loop): no host buffers for dlogits / argmax
    loss2, _ = ctx.ce_fused(x, t, w, 255, want_grad=True, confmat=cm)
    assert loss2.item() == loss.item() and torch.equal(cm, 2 * ref.compute())
    ctx.close()


def test_prefetched_total_weight_gives_identical_results():
    """crit.prefetch_total_weight(target) moves the int64-label pre-pass (Σw + byte labels) to a side stream ahead of
    time; loss and gradients are bit-identical to the in-line pre-pass, and a target that was modified in place after
    the prefetch (label remap, a reused staging buffer) does not get the stale Σw / stale byte labels."""
    from cvcs_b200.loss import FusedCrossEntropyLoss
    torch.manual_seed(3)
    C = 7
    w = torch.rand(C) + 0.5
    crit = FusedCrossEntropyLoss(weight=w, ignore_index=255).to(DEV)
    x = torch.randn(2, C, 64, 64, device=DEV)
    t = torch.randint(0, C, (2, 64, 64), device=DEV, dtype=torch.int64)
    t[0, :5] = 255
    xa = x.clone().requires_grad_(True)
    la = crit(xa, t)
    la.backward()
    xb = x.clone().requires_grad_(True)
    crit.prefetch_total_weight(t, C)
    assert crit._prefetched is not None
    lb = crit(xb, t)
    lb.backward()
    assert crit._prefetched is None
    assert la.item() == lb.item() and torch.equal(xa.grad, xb.grad)
    ref = torch_path.make_criterion(w, 255)

    def check(xc, lc, tt):
        xr = x.detach().cpu().requires_grad_(True)
        lr = ref(xr, tt.cpu().long())
        lr.backward()
        assert abs(lc.item() - lr.item()) <= 1e-5 * abs(lr.item())
        assert float((xc.grad.cpu() - xr.grad).abs().max()) <= 1e-5 * float(xr.grad.abs().max())

    # a different target tensor does not pick up a stale prefetch
    crit.prefetch_total_weight(t, C)
    t2 = t.clone()
    t2[1, :7] = 255
    xc = x.clone().requires_grad_(True)
    lc = crit(xc, t2)
    lc.backward()
    check(xc, lc, t2)
    # the SAME tensor, modified in place between prefetch and forward: the version counter invalidates the prefetch
    crit.prefetch_total_weight(t, C)
    t[1, 10:30] = 255
    t[0, 40:] = 3
    xd = x.clone().requires_grad_(True)
    ld = crit(xd, t)
    ld.backward()
    check(xd, ld, t)
    # uint8 labels: no launch at prefetch time — the forward of the batch BEFORE sums their weights inside its kernel
    t8a, t8b = t.to(torch.uint8), t2.to(torch.uint8)
    crit.prefetch_total_weight(t8b, C)                 # announce the next batch while the current one is t8a
    assert crit._prefetched is None and crit._next_labels is not None
    xe = x.clone().requires_grad_(True)
    le = crit(xe, t8a)                                 # this launch also scans t8b
    le.backward()
    check(xe, le, t)
    assert crit._scanned is not None and crit._scanned[0] is t8b
    xf = x.clone().requires_grad_(True)
    lf = crit(xf, t8b)                                 # starts from the sum the previous launch left
    lf.backward()
    check(xf, lf, t2)
    assert crit._scanned is None
    # an announced tensor that is modified in place afterwards is not trusted
    crit.prefetch_total_weight(t8a, C)
    xg = x.clone().requires_grad_(True)
    crit(xg, t8b).backward()
    t8a[0, 50:60] = 1
    xh = x.clone().requires_grad_(True)
    lh = crit(xh, t8a)
    lh.backward()
    check(xh, lh, t8a)


def test_module_edge_cases_from_the_review():
    from cvcs_b200.loss import FusedCrossEntropyLoss
    from cvcs_b200.metrics import MulticlassConfusionMatrix
    torch.manual_seed(5)
    C = 5
    x = torch.randn(2, C, 32, 32, device=DEV)
    t = torch.randint(0, C, (2, 32, 32), device=DEV, dtype=torch.uint8)
    # (1) a second backward through the same loss is refused with a message that says why
    xa = x.clone().requires_grad_(True)
    loss = FusedCrossEntropyLoss()(xa, t)
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="second backward"):
        loss.backward()
    # (2) the registered weight buffer is what the kernel uses, also after it was replaced / edited
    crit = FusedCrossEntropyLoss(weight=torch.ones(C)).to(DEV)
    l1 = crit(x, t).item()
    new_w = torch.tensor([0.1, 2.0, 0.5, 1.5, 3.0])
    crit.load_state_dict({"weight": new_w})
    l2 = crit(x, t).item()
    l_ref = torch_path.make_criterion(new_w, -100)(x.cpu(), t.cpu().long()).item()
    assert l1 != l2 and abs(l2 - l_ref) <= 1e-5 * abs(l_ref)
    with torch.no_grad():
        crit.weight[0] = 5.0
    new_w[0] = 5.0
    l3 = crit(x, t).item()
    l_ref = torch_path.make_criterion(new_w, -100)(x.cpu(), t.cpu().long()).item()
    assert abs(l3 - l_ref) <= 1e-5 * abs(l_ref)
    # (3) a metric whose ignore_index differs from the criterion's would silently count other pixels: refused
    cm0 = MulticlassConfusionMatrix(num_classes=C, ignore_index=0)
    with pytest.raises(RuntimeError, match="ignore_index"):
        FusedCrossEntropyLoss(ignore_index=-100, confusion=cm0)(x, t)
    cm1 = MulticlassConfusionMatrix(num_classes=C, ignore_index=0)
    FusedCrossEntropyLoss(ignore_index=0, confusion=cm1)(x, t)           # same filter: fine
    ref = torch_path.RestatedConfusionMatrix(C, None, 0)
    ref.update(x.argmax(1).cpu().reshape(1, -1), t.cpu().long().reshape(1, -1))
    assert torch.equal(cm1.compute(), ref.compute())
    # (4) out-of-range labels seen by the fused pass reach the metric's validate_args status
    tb = t.clone()
    tb[0, 0, :3] = C + 2
    cm2 = MulticlassConfusionMatrix(num_classes=C)
    FusedCrossEntropyLoss(confusion=cm2)(x, tb)
    with pytest.raises(RuntimeError, match="outside"):
        cm2.compute()


def test_functional_wrappers_check_what_crosses_the_abi():
    from cvcs_b200 import ops
    x = torch.randn(1, 5, 16, 16, device=DEV)
    t = torch.randint(0, 5, (1, 16, 16), device=DEV, dtype=torch.uint8)
    with pytest.raises(RuntimeError, match="weight"):
        ops.ce_fused(x, t, torch.ones(5, dtype=torch.float64, device=DEV), want_grad=False)
    with pytest.raises(RuntimeError, match="weight"):
        ops.ce_fused(x, t, torch.ones(4, device=DEV), want_grad=False)
    with pytest.raises(RuntimeError, match="confmat"):
        ops.ce_fused(x, t, want_grad=False, confmat=torch.zeros((4, 4), dtype=torch.int64, device=DEV))
    with pytest.raises(RuntimeError, match="confmat"):
        ops.ce_fused(x, t, want_grad=False, confmat=torch.zeros((5, 5), dtype=torch.int32, device=DEV))
    with pytest.raises(RuntimeError, match="argmax"):
        ops.ce_fused(x, t, want_grad=False, argmax=torch.zeros((1, 16, 8), dtype=torch.uint8, device=DEV))
    with pytest.raises(RuntimeError, match="dlogits"):
        ops.ce_fused(x, t, dlogits=torch.empty_like(x).contiguous(memory_format=torch.channels_last), inv_total_weight=1.0)
    with pytest.raises(RuntimeError, match="loss_sums"):
        ops.ce_fused(x, t, want_grad=False, loss_sums=torch.zeros(3, device=DEV))


def test_c_abi_from_a_plain_c_host(tmp_path):
    """A non-Python caller: a C program (gcc, no CUDA or torch headers) dlopens the library and drives
    cvcs_host_ce_fused on malloc'ed host buffers, checking against its own scalar restatement."""
    import shutil
    import subprocess
    from cvcs_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc on this box")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "c_abi_host_demo")
    subprocess.run(["gcc", "-O2", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "c_abi_host_demo.c"),
                    "-o", exe, "-ldl", "-lm"], check=True)
    r = subprocess.run([exe, _lib.LIB_PATH], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "C-ABI DEMO OK" in r.stdout, r.stdout + r.stderr


def test_converter_and_ensemble_dropins(golden):
    """GID15Converter.iconvert (converters.py:23-36) and Ensemble.forward (utils.py:499-507) on the GPU vs the
    reference-generated fixtures / torch.mode."""
    from cvcs_b200.converters import GID15Converter
    from cvcs_b200.ensemble import Ensemble
    g = golden("misc_cases")
    conv = GID15Converter()
    assert np.array_equal((conv.lut("cpu")).numpy(), g["iconvert.lut"])
    out = conv.iconvert(torch.from_numpy(g["iconvert.in"]))
    assert not out.is_cuda and np.array_equal(out.numpy(), g["iconvert.out"])
    out_d = conv.iconvert(torch.from_numpy(g["iconvert.in"]).to(torch.uint8).to(DEV))
    assert out_d.is_cuda and np.array_equal(out_d.cpu().numpy(), g["iconvert.out"])

    class Fixed(nn.Module):
        def __init__(self, logits):
            super().__init__()
            self.logits = logits

        def forward(self, x, context=None):
            return self.logits

    gen = torch.Generator().manual_seed(2)
    members = [Fixed(torch.randn(1, 6, 24, 40, generator=gen).to(DEV)) for _ in range(4)]
    ens = Ensemble(16, DEV, models=members)
    assert ens.returns_logits is False and ens.requires_context is False
    got = ens(torch.zeros(1, 3, 24, 40, device=DEV))
    preds = [torch.argmax(m.logits.squeeze().permute(1, 2, 0).cpu(), dim=2) for m in members]      # utils.py:504
    want, _ = torch.mode(torch.stack(tuple(preds), dim=0), dim=0)                                   # utils.py:506
    assert got.shape == want.shape and torch.equal(got.cpu(), want)
    with pytest.raises(Exception):
        Ensemble(16, DEV, None)
