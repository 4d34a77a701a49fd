"""Maximum sizes: batches whose logit tensor (and, in the last case, whose pixel count) exceeds 2^31 elements, so every
index on the path has to be 64-bit.  Whole-tensor results are checked through size-independent properties and against
torch's own CUDA kernels evaluated image block by image block (an independent implementation of the same library calls
the reference makes: utils.py:90,230,238); the images at both ends of the batch and the ones that straddle element /
pixel 2^31 are checked against the host path (torch CPU, train.py:122-125) like every other parity test."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)
IGN = 255


@pytest.mark.parametrize("B,C,dtype,layout", [
    (328, 7, torch.float32, "NCHW"),        # 2.41e9 logit elements (9.6 GB + 9.6 GB of gradients)
    (328, 7, torch.float32, "NHWC"),
    (2112, 2, torch.bfloat16, "NCHW"),      # 2.21e9 PIXELS (2^31 = image 2048), 4.4e9 logit elements
])
def test_batches_beyond_2_31(B, C, dtype, layout):
    from cvcs_b200 import ops
    H = W = 1024
    free, _ = torch.cuda.mem_get_info(DEV)
    if free < 60 * 2 ** 30:
        pytest.skip("needs 60 GB of free device memory")
    torch.manual_seed(5)
    x = torch.empty((B, C, H, W), dtype=dtype, device=DEV)
    if layout == "NHWC":
        x = x.contiguous(memory_format=torch.channels_last)
    t = torch.empty((B, H, W), dtype=torch.uint8, device=DEV)
    blk = 64
    for b0 in range(0, B, blk):                                           # fill in place, image block by image block
        nb = min(blk, B - b0)
        x[b0:b0 + nb] = (torch.randn((nb, C, H, W), device=DEV) * 3).to(dtype)
        t[b0:b0 + nb] = torch.randint(0, C, (nb, H, W), dtype=torch.uint8, device=DEV)
        t[b0:b0 + nb].masked_fill_(torch.rand((nb, H, W), device=DEV) < 0.1, IGN)
    assert x.numel() > 2 ** 31
    w_cpu = torch.linspace(0.5, 1.5, C)
    w = w_cpu.to(DEV)

    hist = torch.zeros(C + 2, dtype=torch.int64, device=DEV)
    tw = torch.empty(2, dtype=torch.float64, device=DEV)
    ops.label_hist(t, C, IGN, hist=hist, weight=w, total_weight_out=tw)
    am = torch.full((B, H, W), 99, dtype=torch.uint8, device=DEV)
    cm = torch.zeros((C, C), dtype=torch.int64, device=DEV)
    loss, sums, d = ops.ce_fused(x, t, w, IGN, want_grad=True, inv_total_weight_dev=tw[1:], argmax=am, confmat=cm)
    am2 = ops.argmax(x, out_dtype=torch.uint8)                                                # K2 and K3 on the same batch
    cm2 = torch.zeros((C, C), dtype=torch.int64, device=DEV)
    ops.confmat_update(cm2, am2, t, C, IGN)
    torch.cuda.synchronize()
    assert d.stride() == x.stride()

    # ---- whole batch, block by block, against torch's CUDA kernels
    nll_sum = torch.zeros((), dtype=torch.float64, device=DEV)
    cm_ref = torch.zeros(C * C, dtype=torch.int64, device=DEV)
    hist_ref = torch.zeros(C, dtype=torch.int64, device=DEV)
    gmax = 0.0
    for b0 in range(0, B, blk):
        xb, tb = x[b0:b0 + blk].float(), t[b0:b0 + blk]
        ab = xb.argmax(1)
        assert torch.equal(am[b0:b0 + blk].long(), ab), b0
        assert torch.equal(am2[b0:b0 + blk].long(), ab), b0
        keep = tb != IGN
        tl = tb[keep].long()
        cm_ref += torch.bincount(tl * C + ab[keep], minlength=C * C)
        hist_ref += torch.bincount(tl, minlength=C)
        nll_sum += F.cross_entropy(xb, tb.long(), w, ignore_index=IGN, reduction="none").double().sum()   # fp64 fold
        db = d[b0:b0 + blk].float()
        assert float(db.abs().amax(1)[~keep].max()) == 0.0, b0            # exact zeros at ignored pixels
        gmax = max(gmax, float(db.abs().max()))
        del xb, ab, db
    assert torch.equal(hist[:C], hist_ref) and int(hist[C]) == B * H * W - int(hist_ref.sum()) and int(hist[C + 1]) == 0
    assert torch.equal(cm.flatten(), cm_ref) and torch.equal(cm2.flatten(), cm_ref)
    sw = float((hist_ref.double().cpu() * w_cpu.double()).sum())
    assert abs(float(tw[0]) - sw) <= 1e-12 * sw and abs(float(sums[1]) - sw) <= 1e-7 * sw and float(sums[2]) == 0.0
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    l_ref = float(nll_sum) / sw
    assert abs(float(loss) - l_ref) <= tol * abs(l_ref), (float(loss), l_ref)

    # ---- the ends of the batch and the images around element / pixel 2^31, against the host path
    per_image = C * H * W
    picks = {0, B - 1, (2 ** 31) // per_image, min(B - 1, (2 ** 31) // per_image + 1)}
    if B * H * W > 2 ** 31:
        picks |= {2 ** 31 // (H * W) - 1, 2 ** 31 // (H * W)}
    for b in sorted(picks):
        xb, tb = x[b:b + 1].float().cpu().contiguous(), t[b:b + 1].cpu()
        x0 = xb.clone().requires_grad_(True)
        F.cross_entropy(x0, tb.long(), w_cpu, ignore_index=IGN, reduction="sum").backward()   # train.py:122-125, unscaled
        g_ref = x0.grad / sw
        g = d[b:b + 1].float().cpu()
        assert float((g - g_ref).abs().max()) <= tol * float(g_ref.abs().max()), b
        assert torch.equal(am[b:b + 1].cpu().long(), torch.max(xb, dim=1)[1]), b                 # utils.py:90
    assert gmax > 0.0
    del x, d, t, am, am2
    torch.cuda.empty_cache()


def test_tiler_scene_beyond_2_31_bytes():
    """K5 on a 13-band scene of 2.2e9 bytes (source offsets past 2^31): every grid tile against torch's own slicing +
    (v - mean) / std on the device (IEEE fp32, the same arithmetic), and the tiles of the last rows / columns plus an
    overhanging one against the oracle run on the scene's bottom-right corner (dataset.py:29-31, nets.py:339-342)."""
    import numpy as np
    from cvcs_b200 import ops
    from oracle import c_oracle
    Cb, H, W, p = 13, 13056, 13056, 512
    free, _ = torch.cuda.mem_get_info(DEV)
    if free < 60 * 2 ** 30:
        pytest.skip("needs 60 GB of free device memory")
    torch.manual_seed(9)
    scene = torch.empty((Cb, H, W), dtype=torch.uint8, device=DEV)
    for c in range(Cb):
        scene[c] = torch.randint(0, 256, (H, W), dtype=torch.uint8, device=DEV)
    assert scene.numel() > 2 ** 31
    lab = torch.randint(0, 16, (H, W), dtype=torch.uint8, device=DEV)
    mean = torch.rand(Cb, device=DEV) * 100
    std = torch.rand(Cb, device=DEV) * 50 + 1
    rows, cols = H // p, W // p
    grid = [(r * p, c * p) for r in range(rows) for c in range(cols)]
    extra = [(H - p // 2, W - p // 2), (H - p, W - p - 3), (H - p - 7, W - p)]            # overhanging / unaligned, far corner
    yx = torch.tensor(grid + extra, dtype=torch.int32, device=DEV)
    hist = torch.zeros(18, dtype=torch.int64, device=DEV)
    out, lo = ops.tile_normalize(scene, yx, (p, p), mean, std, label=lab, hist=hist, hist_classes=16)
    torch.cuda.synchronize()
    n = len(grid)
    ref_u8 = scene[:, :rows * p, :cols * p].reshape(Cb, rows, p, cols, p).permute(1, 3, 0, 2, 4).reshape(n, Cb, p, p)
    ref = (ref_u8.float() - mean[None, :, None, None]) / std[None, :, None, None]
    assert torch.equal(out[:n], ref)
    del ref, ref_u8
    ref_lab = lab[:rows * p, :cols * p].reshape(rows, p, cols, p).permute(0, 2, 1, 3).reshape(n, p, p)
    assert torch.equal(lo[:n], ref_lab)
    # the far corner on the host: a 1536-pixel square holds the last grid tile and the three extra ones
    k = 1536
    sub = scene[:, H - k:, W - k:].cpu().numpy()
    sub_lab = lab[H - k:, W - k:].cpu().numpy()
    sel = [n - 1, n, n + 1, n + 2]
    yx_sub = (yx[sel].cpu().numpy() - np.array([[H - k, W - k]], dtype=np.int32)).astype(np.int32)
    assert yx_sub.min() >= 0
    r, rl = c_oracle.tile(sub, yx_sub, p, p, mean.cpu().numpy(), std.cpu().numpy(), labels=sub_lab)
    assert np.array_equal(out[sel].cpu().numpy(), r)
    assert np.array_equal(lo[sel].cpu().numpy(), rl)
    # label histogram of all gathered tiles (K5's fused class count, dataset.py:360-384)
    assert torch.equal(hist[:16], torch.bincount(lo.flatten().long(), minlength=16)[:16])
