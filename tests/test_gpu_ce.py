"""K1 parity (through the C-ABI): fused softmax-CE fwd+bwd + argmax + confusion matrix.

Bars (BASELINE.json): argmax maps and confusion matrices bit-exact; loss and gradients within
1e-5 relative for fp32 logits, 1e-2 for bf16 logits.  "Relative" for the gradient tensor is
max|a-b| <= tol * max|b| (the gradients of one batch share one scale, 1/Σw).
"""
import numpy as np
import pytest
import torch

from oracle import c_oracle, torch_path

pytestmark = pytest.mark.gpu

F32_TOL, BF16_TOL = 1e-5, 1e-2
PATHS = {"tma": 1, "direct": 2, "generic": 3, "auto": 0}


@pytest.fixture()
def lib():
    from cvcs_b200 import _lib
    yield _lib
    _lib.set_option(_lib.OPT_CE_PATH, 0)
    _lib.set_option(_lib.OPT_TMA_STAGES, 0)


def run_k1(logits, target, weight, ignore_index, *, want_grad=True, layout="NCHW", label_dtype=torch.uint8,
           argmax_dtype=torch.uint8, dtype=torch.float32, num_classes=None):
    """Returns loss(float), sums(np f64[3]), grad(np f32, NCHW) | None, argmax(np i64), confmat(np i64)."""
    from cvcs_b200 import ops
    dev = torch.device("cuda", 0)
    x = torch.as_tensor(logits).to(dev).to(dtype)
    if layout == "NHWC":
        x = x.contiguous(memory_format=torch.channels_last)
    t = torch.as_tensor(target).to(dev).to(label_dtype)
    B, C, H, W = x.shape
    w = None if weight is None else torch.as_tensor(weight, dtype=torch.float32).to(dev)
    inv_dev = None
    if want_grad:
        tw = torch.empty(2, dtype=torch.float64, device=dev)
        ops.label_hist(t, C, ignore_index, weight=w, total_weight_out=tw)
        inv_dev = tw[1:]
    am = torch.full((B, H, W), 99, dtype=argmax_dtype, device=dev)
    cm = torch.zeros((C, C), dtype=torch.int64, device=dev)
    loss, sums, d = ops.ce_fused(x, t, w, ignore_index, want_grad=want_grad, inv_total_weight_dev=inv_dev,
                                 argmax=am, confmat=cm)
    torch.cuda.synchronize()
    if d is not None:
        assert d.dtype == x.dtype and d.stride() == x.stride()
        d = d.float().cpu().numpy()
    return float(loss.item()), sums.cpu().numpy(), d, am.cpu().numpy().astype(np.int64), cm.cpu().numpy()


def check_against(loss, grad, am, cm, logits, target, weight, ignore_index, tol):
    C = logits.shape[1]
    l_ref, sums_ref, g_ref = c_oracle.cross_entropy(logits, target, weight, ignore_index)
    if np.isnan(l_ref):
        assert np.isnan(loss)
    else:
        assert abs(loss - l_ref) <= tol * abs(l_ref), (loss, l_ref)
    if grad is not None:
        scale = max(np.abs(g_ref).max(), 1e-30)
        assert np.abs(grad - g_ref).max() <= tol * scale, np.abs(grad - g_ref).max() / scale
        ign = (target == ignore_index)
        assert np.all(grad[np.broadcast_to(ign[:, None], grad.shape)] == 0)   # exact zeros at ignored pixels
    am_ref = c_oracle.argmax(logits)
    assert np.array_equal(am, am_ref)
    cm_ref, _ = c_oracle.confmat(am_ref, target, C, ignore_index)
    assert np.array_equal(cm, cm_ref)


GOLDEN_CASES = ["cel_c7", "cel_c7_ignore0", "wcel_c7", "wcel_c7_ignore0", "cel_c16", "wcel_c16", "cel_c7_bf16vals",
                "cel_c7_allignored", "cel_c3_odd"]


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("path", ["auto", "tma", "direct", "generic"])
def test_golden_cases_fp32(golden, lib, name, path):
    """The reference's own outputs (utils.load_loss criterion + backward) on every K1 variant."""
    g = golden("ce_cases")
    lib.set_option(lib.OPT_CE_PATH, PATHS[path])
    logits, target = g[f"{name}.logits"], g[f"{name}.target"]
    w = g[f"{name}.weight"] if g[f"{name}.weight"].size else None
    ii = int(g[f"{name}.ignore_index"])
    loss, sums, grad, am, cm = run_k1(logits, target, w, ii)
    loss_ref, grad_ref = float(g[f"{name}.loss"]), g[f"{name}.grad"]
    if np.isnan(loss_ref):
        assert np.isnan(loss)
        assert np.all(np.nan_to_num(grad) == 0)
    else:
        assert abs(loss - loss_ref) <= F32_TOL * abs(loss_ref)
        assert np.abs(grad - grad_ref).max() <= F32_TOL * np.abs(grad_ref).max()
    # forward-only mode (validation_loss under no_grad, utils.py:109-120)
    loss2, _, grad2, am2, cm2 = run_k1(logits, target, w, ii, want_grad=False)
    assert grad2 is None and (loss2 == loss or (np.isnan(loss) and np.isnan(loss2)))
    _, am_ref = torch.max(torch.from_numpy(logits), dim=1)
    assert np.array_equal(am, am_ref.numpy()) and np.array_equal(am2, am)
    assert np.array_equal(cm, cm2)
    check_against(loss, grad, am, cm, logits, target, w, ii, F32_TOL)


@pytest.mark.parametrize("layout", ["NCHW", "NHWC"])
@pytest.mark.parametrize("label_dtype", [torch.uint8, torch.int64])
@pytest.mark.parametrize("C", [2, 5, 7, 8, 12, 14, 16, 18, 20, 21, 33])
def test_random_fp32(lib, layout, label_dtype, C):
    g = torch.Generator().manual_seed(C)
    B, H, W = 3, 48, 80
    logits = (torch.randn(B, C, H, W, generator=g) * 3).numpy()
    target = torch.randint(0, C, (B, H, W), generator=g).numpy().astype(np.int64)
    target[0, :5] = 255                                    # LoveDA-style ignore label
    weight = (torch.rand(C, generator=g) + 0.1).numpy()
    weight[C // 2] = 0.0                                   # zero-weight class
    am_dtype = torch.int64 if label_dtype == torch.int64 else torch.uint8
    loss, sums, grad, am, cm = run_k1(logits, target, weight, 255, layout=layout, label_dtype=label_dtype,
                                      argmax_dtype=am_dtype)
    check_against(loss, grad, am, cm, logits, target, weight, 255, F32_TOL)
    assert sums[2] == 0
    # the library call the reference makes, on the same inputs
    l_t, g_t = torch_path.ce_loss_and_grad(torch.from_numpy(logits), torch.from_numpy(target),
                                           torch.from_numpy(weight), 255)
    assert abs(loss - float(l_t)) <= F32_TOL * abs(float(l_t))
    assert np.abs(grad - g_t.numpy()).max() <= F32_TOL * np.abs(g_t.numpy()).max()


@pytest.mark.parametrize("layout", ["NCHW", "NHWC"])
@pytest.mark.parametrize("path", ["auto", "direct", "generic"])
@pytest.mark.parametrize("C", [7, 12, 13, 14, 16, 18, 20])
def test_random_bf16(lib, layout, path, C):
    """cfg3: bf16 logits, class weights, ignore_index=255.  Oracle = fp32 CE on the same bf16 values."""
    lib.set_option(lib.OPT_CE_PATH, PATHS[path])
    g = torch.Generator().manual_seed(100 + C)
    B, H, W = 2, 64, 64
    logits = (torch.randn(B, C, H, W, generator=g) * 3).to(torch.bfloat16).float().numpy()
    target = torch.randint(0, C, (B, H, W), generator=g).numpy().astype(np.int64)
    target[torch.rand(B, H, W, generator=g).numpy() < 0.1] = 255
    weight = (torch.rand(C, generator=g) + 0.1).numpy()
    loss, sums, grad, am, cm = run_k1(logits, target, weight, 255, layout=layout, dtype=torch.bfloat16)
    l_ref, _, g_ref = c_oracle.cross_entropy(logits, target, weight, 255)
    assert abs(loss - l_ref) <= F32_TOL * abs(l_ref)              # the loss itself is computed in fp32
    assert np.abs(grad - g_ref).max() <= BF16_TOL * np.abs(g_ref).max()
    am_ref = c_oracle.argmax(logits)
    assert np.array_equal(am, am_ref)                              # argmax on bf16 values is exact
    assert np.array_equal(cm, c_oracle.confmat(am_ref, target, C, 255)[0])
    # bf16 gradients are the fp32 gradients rounded once (RNE)
    g32 = torch.from_numpy(g_ref).to(torch.bfloat16).float().numpy()
    assert np.abs(grad - g32).max() <= 2.0 ** -7 * np.abs(g_ref).max()


def test_special_values_and_ties(golden, lib):
    g = golden("argmax_cases")
    small = g["small"][None]                                       # [1,C,H,W] with NaN / inf / ties
    target = np.zeros((1, 2, 4), np.int64)
    for path in ("auto", "direct", "generic"):
        lib.set_option(lib.OPT_CE_PATH, PATHS[path])
        loss, _, _, am, cm = run_k1(small, target, None, -100, want_grad=False)
        assert np.array_equal(am[0], g["small_max"])
        assert np.isnan(loss)                                      # NaN logits poison the mean, as in torch
        big = g["big"][None]
        tb = np.zeros((1, 24, 40), np.int64)
        _, _, _, amb, _ = run_k1(big, tb, None, -100, want_grad=False)
        assert np.array_equal(amb[0], g["big_max"])


def test_out_of_bounds_label_poisons_loss(lib):
    logits = np.zeros((1, 7, 16, 16), np.float32)
    target = np.zeros((1, 16, 16), np.int64)
    target[0, 3, 3] = 7
    target[0, 4, 4] = 200
    for dt in (torch.uint8, torch.int64):
        loss, sums, grad, am, cm = run_k1(logits, target, None, -100, label_dtype=dt)
        assert np.isnan(loss) and sums[2] == 2
        assert cm.sum() == 16 * 16 - 2
    target[0, 5, 5] = -1
    loss, sums, *_ = run_k1(logits, target, None, -100, label_dtype=torch.int64)
    assert sums[2] == 3
    # workspace is left clean: the next call is unaffected
    target[:] = 1
    loss, sums, *_ = run_k1(logits, target, None, -100)
    assert sums[2] == 0 and abs(loss - np.log(7)) < 1e-6


@pytest.mark.parametrize("shape", [(1, 7, 1, 16), (2, 7, 3, 16), (1, 7, 50, 50), (2, 7, 225, 225), (1, 3, 1, 1),
                                   (5, 7, 40, 40), (1, 7, 1040, 16)])
def test_ragged_shapes(lib, shape):
    """Chunk tails, plane sizes that are not multiples of 16 (falls back to direct / generic)."""
    B, C, H, W = shape
    g = torch.Generator().manual_seed(H * W)
    logits = torch.randn(B, C, H, W, generator=g).numpy()
    target = torch.randint(0, C, (B, H, W), generator=g).numpy().astype(np.int64)
    for layout in ("NCHW", "NHWC"):
        loss, sums, grad, am, cm = run_k1(logits, target, None, 0, layout=layout)
        check_against(loss, grad, am, cm, logits, target, None, 0, F32_TOL)


@pytest.mark.parametrize("grad", [True, False])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("ctas,stages,vecp_small", [(0, 0, False), (1, 2, False), (1, 4, True), (2, 2, True), (2, 3, False), (3, 3, True)])
def test_tma_pipeline_geometries(lib, ctas, stages, vecp_small, dtype, grad):
    """Every pipeline geometry the tuning knobs can select (CTAs/SM x stages x pixels per thread) gives the
    same results: the knobs only move bytes in flight."""
    lib.set_option(lib.OPT_CE_PATH, PATHS["tma"])
    lib.set_option(lib.OPT_TMA_STAGES, stages)
    lib.set_option(lib.OPT_TMA_CTAS, ctas)
    lib.set_option(lib.OPT_TMA_VECP, (2 if dtype == torch.float32 else 4) if vecp_small else (4 if dtype == torch.float32 else 8))
    try:
        g = torch.Generator().manual_seed(7)
        B, C, H, W = 4, 7, 256, 256                              # 256 chunks of 1024 px
        logits = (torch.randn(B, C, H, W, generator=g) * 3).to(dtype).float().numpy()
        target = torch.randint(0, C, (B, H, W), generator=g).numpy().astype(np.int64)
        target[0, :40] = 0
        w = (torch.rand(C, generator=g) + 0.5).numpy()
        loss, sums, grad_out, am, cm = run_k1(logits, target, w, 0, dtype=dtype, want_grad=grad)
        tol = F32_TOL if dtype == torch.float32 else BF16_TOL
        if grad:
            check_against(loss, grad_out, am, cm, logits, target, w, 0, tol)
        else:
            l_ref, _, _ = c_oracle.cross_entropy(logits, target, w, 0, want_grad=False)
            am_ref = c_oracle.argmax(logits)
            assert grad_out is None and abs(loss - l_ref) <= tol * abs(l_ref)
            assert np.array_equal(am, am_ref) and np.array_equal(cm, c_oracle.confmat(am_ref, target, C, 0)[0])
    finally:
        lib.set_option(lib.OPT_TMA_CTAS, 0)
        lib.set_option(lib.OPT_TMA_VECP, 0)


def test_run_to_run_bit_stable(lib):
    g = torch.Generator().manual_seed(11)
    logits = torch.randn(4, 7, 128, 128, generator=g).numpy()
    target = torch.randint(0, 7, (4, 128, 128), generator=g).numpy().astype(np.int64)
    a = run_k1(logits, target, None, 0)
    b = run_k1(logits, target, None, 0)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


@pytest.mark.parametrize("dtype,C,B", [(torch.float32, 7, 16), (torch.bfloat16, 7, 16), (torch.float32, 20, 4)])
def test_full_size_properties(lib, dtype, C, B):
    """BASELINE.json sizes (B x C x 1024 x 1024): size-independent invariants on the GPU results, and the
    loss / confusion matrix against the oracle (argmax + confusion on the CPU take seconds)."""
    from cvcs_b200 import ops
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(0)
    H = W = 1024
    logits_cpu = (torch.randn(B, C, H, W, generator=g) * 3).to(dtype)
    target_cpu = torch.randint(0, C, (B, H // 32, W // 32), generator=g, dtype=torch.uint8)
    target_cpu = target_cpu.repeat_interleave(32, 1).repeat_interleave(32, 2).contiguous()   # blocky labels
    target_cpu[torch.rand(B, H, W, generator=g) < 0.1] = 255
    weight_cpu = torch.rand(C, generator=g) + 0.5
    x, t, w = logits_cpu.to(dev), target_cpu.to(dev), weight_cpu.to(dev)
    tw = torch.empty(2, dtype=torch.float64, device=dev)
    hist = torch.zeros(C + 2, dtype=torch.int64, device=dev)
    ops.label_hist(t, C, 255, hist=hist, weight=w, total_weight_out=tw)
    am = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    cm = torch.zeros((C, C), dtype=torch.int64, device=dev)
    loss, sums, d = ops.ce_fused(x, t, w, 255, want_grad=True, inv_total_weight_dev=tw[1:], argmax=am, confmat=cm)
    torch.cuda.synchronize()
    hist_c, cm_c, sums_c = hist.cpu(), cm.cpu(), sums.cpu()
    n_valid = int((target_cpu != 255).sum())
    # (1) label histogram == bincount, ignored bucket == #255
    assert torch.equal(hist_c[:C], torch.bincount(target_cpu[target_cpu != 255].long().flatten(), minlength=C))
    assert int(hist_c[C]) == B * H * W - n_valid and int(hist_c[C + 1]) == 0
    # (2) every valid pixel lands in exactly one bin; row sums are the label histogram
    assert int(cm_c.sum()) == n_valid and torch.equal(cm_c.sum(1), hist_c[:C])
    # (3) Σw from the kernel == Σw from the histogram == K4's total weight (K1 adds the weights of a
    # thread's 4-8 pixels in fp32 before the fp64 accumulation: 1e-7; K4 is integer counts x fp64: 1e-12)
    sw = float((hist_c[:C].double() * weight_cpu.double()).sum())
    assert abs(sums_c[1].item() - sw) <= 1e-7 * sw and abs(tw[0].item() - sw) <= 1e-12 * sw
    # (4) per-pixel gradients sum to ~0 over classes, and are exactly 0 at ignored pixels
    ds = d.float().sum(1)
    tol = 1e-6 if dtype == torch.float32 else 2e-2
    assert float(ds.abs().max()) <= tol * float(d.float().abs().max())
    assert float(d.float().abs().amax(1)[t == 255].max()) == 0.0
    # (5) column sums of the confusion matrix == histogram of the argmax map over valid pixels
    amv = am[t != 255].long()
    assert torch.equal(cm_c.sum(0), torch.bincount(amv, minlength=C).cpu())
    # (6) against the reference's library calls on the host
    xf = logits_cpu.float()
    l_ref = torch_path.ce_loss_only(xf, target_cpu.long(), weight_cpu, 255)
    assert abs(loss.item() - l_ref.item()) <= F32_TOL * abs(l_ref.item())
    _, am_ref = torch.max(xf, dim=1)
    assert torch.equal(am.cpu().long(), am_ref)
    ref_cm = torch_path.RestatedConfusionMatrix(C, ignore_index=255)
    ref_cm.update(am_ref, target_cpu.long())
    assert torch.equal(cm_c, ref_cm.compute())
    # (7) gradient parity on the first image (fp32 autograd on the CPU)
    inv = 1.0 / sums_c[1].item()
    x0 = xf[:1].clone().requires_grad_(True)
    l0 = torch.nn.functional.cross_entropy(x0, target_cpu[:1].long(), weight_cpu, ignore_index=255, reduction="sum")
    l0.backward()
    g0 = (x0.grad * inv).numpy()
    gt = d[:1].float().cpu().numpy()
    assert np.abs(gt - g0).max() <= (F32_TOL if dtype == torch.float32 else BF16_TOL) * np.abs(g0).max()


# ---- class counts beyond the register-resident range: the shared-memory class walk (csrc/ce_wide.cu) ----------------
@pytest.mark.parametrize("dtype,label_dtype", [(torch.float32, torch.uint8), (torch.float32, torch.int64), (torch.bfloat16, torch.uint8)])
@pytest.mark.parametrize("C,H,W,weighted,ignore", [(22, 48, 80, False, -100), (32, 64, 64, True, 255), (40, 16, 48, True, 3),
                                                    (64, 48, 48, False, 255), (100, 32, 32, True, -100)])
def test_wide_class_counts(lib, dtype, label_dtype, C, H, W, weighted, ignore):
    """C > 21, NCHW, H*W % 16 == 0: chunks of 256 / 512 pixels with a ragged last chunk per image, weights, ignored
    labels, out-of-range labels excluded here (see the status test), NaN / inf rows and ties.  Against the oracle at the
    path's tolerance, and bit for bit against the generic one-pixel-per-thread kernel (same arithmetic, term for term)."""
    g = torch.Generator().manual_seed(C * 1000 + H + W)
    B = 3
    x = (torch.randn(B, C, H, W, generator=g) * 3).to(dtype).float()
    x[0, :, 0, :8] = 1.0                                        # ties -> first index
    x[0, 5, 1, :4] = float("nan")                               # NaN is maximal, first NaN wins
    x[1, 7, 2, :3] = float("inf")
    x[2, :, 3, :2] = float("-inf")
    t = torch.randint(0, C, (B, H, W), generator=g)
    if 0 <= ignore:
        t[torch.rand(B, H, W, generator=g) < 0.1] = ignore
    w = (torch.rand(C, generator=g) + 0.5).numpy() if weighted else None
    tol = F32_TOL if dtype == torch.float32 else BF16_TOL
    res = {}
    for path in ("auto", "generic"):
        lib.set_option(lib.OPT_CE_PATH, PATHS[path])
        res[path] = run_k1(x.numpy(), t.numpy(), w, ignore, dtype=dtype, label_dtype=label_dtype,
                           argmax_dtype=torch.uint8 if label_dtype == torch.uint8 else torch.int64)
        res[path + "/fwd"] = run_k1(x.numpy(), t.numpy(), w, ignore, want_grad=False, dtype=dtype, label_dtype=label_dtype)
    loss, sums, grad, am, cm = res["auto"]
    clean = np.isfinite(x.numpy()).all(axis=1)                   # the gradient comparison skips the rows with NaN / inf
    xs, ts = x.numpy().copy(), t.numpy().copy()
    am_ref = c_oracle.argmax(xs)
    assert np.array_equal(am, am_ref)
    cm_ref, _ = c_oracle.confmat(am_ref, ts, C, ignore if 0 <= ignore else None)
    assert np.array_equal(cm, cm_ref)
    # bit-identical to the generic kernel: gradients, argmax, matrix — with and without gradients (the f64 loss sums
    # are folded in a different pixel order: equal to rounding)
    for a, b in ((res["auto"], res["generic"]), (res["auto/fwd"], res["generic/fwd"])):
        np.testing.assert_allclose(a[1], b[1], rtol=1e-12, equal_nan=True)
        assert (a[2] is None and b[2] is None) or np.array_equal(a[2], b[2], equal_nan=True)
        assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])
    assert np.array_equal(res["auto/fwd"][3], am_ref) and np.array_equal(res["auto/fwd"][4], cm_ref)
    # finite rows against the oracle's fp64 gradient: per-pixel gradients of clean pixels do not depend on the others
    # except through 1/Σw, which both sides take from the same labels
    l2, s2, g2 = c_oracle.cross_entropy(xs, ts, w, ignore)
    mask = np.broadcast_to(clean[:, None], g2.shape)
    scale = max(np.abs(g2[mask]).max(), 1e-30)
    assert np.abs(grad[mask] - g2[mask]).max() <= tol * scale


def test_wide_metrics_mode_and_status(lib):
    """cvcs_eval_fused (no loss) and the out-of-range label count on the wide path."""
    from cvcs_b200 import ops
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(5)
    B, C, H, W = 2, 48, 32, 48
    x = torch.randn(B, C, H, W, generator=g)
    t = torch.randint(0, C, (B, H, W), generator=g, dtype=torch.uint8)
    t[0, 0, :5] = 200                                           # out of range, not the ignore value
    t[1, 1, :7] = 255
    am = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    cm = torch.zeros((C, C), dtype=torch.int64, device=dev)
    ops.eval_fused(x.to(dev), t.to(dev), 255, argmax=am, confmat=cm)
    am_ref = c_oracle.argmax(x.numpy())
    keep = (t.numpy() < C)
    cm_ref, _ = c_oracle.confmat(am_ref[keep], t.numpy()[keep], C, 255)
    assert np.array_equal(am.cpu().numpy().astype(np.int64), am_ref)
    assert np.array_equal(cm.cpu().numpy(), cm_ref)
    loss, sums, _ = ops.ce_fused(x.to(dev), t.to(dev), None, 255, want_grad=False)
    assert int(sums[2].item()) == 5 and np.isnan(float(loss.item()))     # torch would raise; the loss is poisoned
