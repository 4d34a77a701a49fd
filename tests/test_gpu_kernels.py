"""K2 / K3 / K4 / K5 and the N2-N4 kernels through the C-ABI: bit-exact against the oracle and
the reference-generated golden vectors."""
import numpy as np
import pytest
import torch

from oracle import c_oracle, torch_path

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("dtype", [torch.uint8, torch.int64])
@pytest.mark.parametrize("C,ignore", [(7, -100), (7, 255), (7, 0), (16, 0), (20, 255), (100, 255)])
@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 4099, 1 << 20])
def test_label_hist(dtype, C, ignore, n):
    from cvcs_b200 import ops
    g = torch.Generator().manual_seed(n + C)
    lab = torch.randint(0, C, (n,), generator=g, dtype=torch.uint8)
    if n > 3:
        lab[::7] = 255
        lab[3] = C  # out of bounds (C <= 255 here)
    ref = c_oracle.label_hist(lab.numpy(), C, ignore)
    w = torch.rand(C, generator=g) + 0.1
    hist = torch.ones(C + 2, dtype=torch.int64, device=DEV)     # accumulated INTO
    tw = torch.zeros(2, dtype=torch.float64, device=DEV)
    ops.label_hist(lab.to(DEV).to(dtype), C, ignore, hist=hist, weight=w.to(DEV), total_weight_out=tw)
    assert np.array_equal(hist.cpu().numpy() - 1, ref)
    sw = sum(float(ref[c]) * float(w[c]) for c in range(C) if c != ignore)
    assert abs(tw[0].item() - sw) <= 1e-12 * max(sw, 1)
    tw2 = ops.total_weight(hist - 1, w.to(DEV), C, ignore)
    assert tw2[0].item() == pytest.approx(sw, rel=1e-12) and (sw == 0 or tw2[1].item() == pytest.approx(1 / sw, rel=1e-12))


def test_label_hist_matches_reference_class_counts(golden):
    from cvcs_b200 import ops
    from cvcs_b200.loss import class_weights_from_counts
    g = golden("dataset_cases")
    hist = torch.zeros(18, dtype=torch.int64, device=DEV)
    for k in (0, 1):
        ops.label_hist(torch.from_numpy(g[f"scene{k}.label"]).to(DEV), 16, -100, hist=hist)
    counts = hist[:16].cpu()
    assert np.array_equal(counts.numpy().astype(np.float32), g["counts"])
    assert np.array_equal(class_weights_from_counts(counts, False).numpy(), g["weights_ib0"])
    assert np.array_equal(class_weights_from_counts(counts, True).numpy(), g["weights_ib1"])


@pytest.mark.parametrize("layout", ["NCHW", "NHWC"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 7, 64, 64), (1, 16, 33, 17), (3, 20, 32, 48), (1, 1, 8, 8), (1, 300, 4, 4)])
def test_argmax(layout, dtype, shape):
    from cvcs_b200 import ops
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g).to(dtype)
    x[:, :, ::2, ::3] = x[:, :1, ::2, ::3]                       # ties across all classes -> index 0
    xd = x.to(DEV)
    if layout == "NHWC":
        xd = xd.contiguous(memory_format=torch.channels_last)
    _, ref = torch.max(x.float(), dim=1)                          # utils.py:90 semantics
    out64 = ops.argmax(xd, torch.int64)
    assert torch.equal(out64.cpu(), ref)
    if shape[1] <= 256:
        assert torch.equal(ops.argmax(xd, torch.uint8).cpu().long(), ref)


def test_argmax_special_values(golden):
    from cvcs_b200 import ops
    g = golden("argmax_cases")
    x = torch.from_numpy(g["small"])[None].to(DEV)
    assert np.array_equal(ops.argmax(x)[0].cpu().numpy(), g["small_max"])
    hwc = x.contiguous(memory_format=torch.channels_last)         # utils.py:158 reads HWC
    assert np.array_equal(ops.argmax(hwc)[0].cpu().numpy(), g["small_argmax_hwc"])
    assert np.array_equal(ops.argmax(torch.from_numpy(g["big"])[None].to(DEV))[0].cpu().numpy(), g["big_max"])


@pytest.mark.parametrize("pd,td", [(torch.uint8, torch.uint8), (torch.int64, torch.int64), (torch.uint8, torch.int64)])
@pytest.mark.parametrize("C,ignore", [(2, None), (7, None), (7, 0), (16, 0), (20, 255), (150, None)])
@pytest.mark.parametrize("n", [0, 5, 4096, 100003])
def test_confmat(pd, td, C, ignore, n):
    from cvcs_b200 import ops
    g = torch.Generator().manual_seed(n + C)
    hi = min(C, 255)
    p = torch.randint(0, hi, (n,), generator=g)
    t = torch.randint(0, hi, (n,), generator=g)
    t = (t // 3) * 3 % hi                                          # runs of identical labels
    if ignore == 255 and n:
        t[::5] = 255
    ref, _ = c_oracle.confmat(p.numpy(), t.numpy(), C, ignore)
    cm = torch.zeros((C, C), dtype=torch.int64, device=DEV)
    st = torch.zeros(1, dtype=torch.int64, device=DEV)
    ops.confmat_update(cm, p.to(pd).to(DEV), t.to(td).to(DEV), C, ignore, status=st)
    ops.confmat_update(cm, p.to(pd).to(DEV), t.to(td).to(DEV), C, ignore, status=st)   # accumulates
    assert np.array_equal(cm.cpu().numpy(), 2 * ref)
    assert int(st.item()) == 0


@pytest.mark.parametrize("C,ignore", [(7, None), (7, 255), (7, 3), (8, 200), (15, 0), (16, 255), (16, None), (17, 255)])
def test_confmat_u8_byte_parallel_path(C, ignore):
    """u8 / u8 maps: words of four in-range pairs take the one-multiply key path (C <= 16), words holding an ignored
    label, an out-of-range label or an out-of-range prediction the pixel path — mixed at every lane position, with runs
    of identical pairs (the +4 update) — all equal to the oracle, and out-of-range pixels are counted in `status`."""
    from cvcs_b200 import ops
    g = torch.Generator().manual_seed(C * 7 + (ignore or 0))
    n = 64 * 1024 + 7
    t = torch.randint(0, C, (n,), generator=g, dtype=torch.uint8)
    p = torch.randint(0, C, (n,), generator=g, dtype=torch.uint8)
    t[4096:8192] = 2                                                   # long runs: whole words agree
    p[4096:8192] = 1
    p[6000:6100] = torch.randint(0, C, (100,), generator=g, dtype=torch.uint8)
    if ignore is not None:
        t[torch.rand(n, generator=g) < 0.05] = ignore
    bad_t = torch.rand(n, generator=g) < 0.01
    bad_p = torch.rand(n, generator=g) < 0.01
    t[bad_t] = 250
    p[bad_p] = 251
    ref, n_bad = c_oracle.confmat(p.numpy(), t.numpy(), C, ignore)
    cm = torch.zeros((C, C), dtype=torch.int64, device=DEV)
    st = torch.zeros(1, dtype=torch.int64, device=DEV)
    ops.confmat_update(cm, p.to(DEV), t.to(DEV), C, ignore, status=st)
    assert np.array_equal(cm.cpu().numpy(), ref)
    assert int(st.item()) == int(n_bad)
    cm.zero_()
    ops.confmat_update(cm, p[1:].to(DEV)[:-2], t[1:].to(DEV)[:-2], C, ignore)   # contiguous copies, different phase
    assert np.array_equal(cm.cpu().numpy(), c_oracle.confmat(p[1:-2].numpy(), t[1:-2].numpy(), C, ignore)[0])


def test_confmat_flags_out_of_range():
    from cvcs_b200 import ops
    cm = torch.zeros((7, 7), dtype=torch.int64, device=DEV)
    st = torch.zeros(1, dtype=torch.int64, device=DEV)
    p = torch.tensor([0, 7, 3, 2], dtype=torch.int64, device=DEV)
    t = torch.tensor([0, 1, 9, 2], dtype=torch.int64, device=DEV)
    ops.confmat_update(cm, p, t, 7, None, status=st)
    assert int(st.item()) == 2 and int(cm.sum()) == 2


def test_eval_golden_through_k2_k3(golden):
    """utils.eval_model's matrices (reference-generated) via argmax + confusion kernels."""
    from cvcs_b200 import ops
    g = golden("eval_cases")
    x = torch.from_numpy(g["logits"]).to(DEV)
    y = torch.from_numpy(g["labels"]).to(DEV)
    for ib in (0, 1):
        cm = torch.zeros((16, 16), dtype=torch.int64, device=DEV)
        for i in range(x.shape[0]):                               # tile by tile, as the reference loops
            pred = ops.argmax(x[i:i + 1])
            ops.confmat_update(cm, pred, y[i:i + 1], 16, 0 if ib else None)
        assert np.array_equal(cm.cpu().numpy(), g[f"ib{ib}.flat"])


# ---- K5 --------------------------------------------------------------------------------------------
def test_tile_reference_crops(golden):
    from cvcs_b200 import ops
    g = golden("dataset_cases")
    img = torch.from_numpy(g["crop.image"]).to(DEV)
    msk = torch.from_numpy(g["crop.mask"][0]).to(DEV)
    for i, (tly, tlx, q) in enumerate(g["crop.cases"]):
        yx = torch.tensor([[tly, tlx]], dtype=torch.int32, device=DEV)
        out, lab = ops.tile_normalize(img, yx, (int(q), int(q)), out_dtype=torch.uint8, label=msk)
        assert np.array_equal(out[0].cpu().numpy(), g[f"crop.{i}.patch"])
        assert np.array_equal(lab[0].cpu().numpy(), g[f"crop.{i}.mask"][0])
    out, _ = ops.tile_normalize(img, torch.tensor([[0, 0], [-2, -2]], dtype=torch.int32, device=DEV), (6, 6),
                                out_dtype=torch.uint8)
    assert np.array_equal(out[0].cpu().numpy(), g["padded.patch"])      # _get_padded_patch(img, 2, 2, (4,4), 6)
    assert np.array_equal(out[1].cpu().numpy(), g["padded.corner"])


def test_tile_loader_order_cast_and_normalize(golden):
    from cvcs_b200 import ops
    g = golden("dataset_cases")
    p, cols, tpi = 224, 2, 2
    for tag in ("shift0",):
        for k, idx in enumerate(g[f"{tag}.chunk_crops"]):
            im, tly, tlx = torch_path.tile_top_left(int(idx), tpi, cols, p)
            scene = torch.from_numpy(g[f"scene{im}.image"]).to(DEV)
            lab = torch.from_numpy(g[f"scene{im}.label"]).to(DEV)
            yx = torch.tensor([[tly, tlx]], dtype=torch.int32, device=DEV)
            out, lo = ops.tile_normalize(scene, yx, (p, p), out_dtype=torch.uint8, label=lab)
            assert np.array_equal(out[0].cpu().numpy(), g[f"{tag}.patches"][k])
            assert np.array_equal(lo[0].cpu().numpy(), g[f"{tag}.index_masks"][k])
    allv = torch.from_numpy(g["normalize.in"]).to(DEV)
    yx = torch.zeros((1, 2), dtype=torch.int32, device=DEV)
    mean = torch.tensor([0.485, 0.456, 0.406], device=DEV)
    std = torch.tensor([0.229, 0.224, 0.225], device=DEV)
    out, _ = ops.tile_normalize(allv, yx, (16, 16), mean, std)
    assert np.array_equal(out[0].cpu().numpy(), g["normalize.out"])      # bit-exact fp32 (nets.py:339-342)
    out, _ = ops.tile_normalize(allv, yx, (16, 16))
    assert np.array_equal(out[0].cpu().numpy(), g["cast.out"])           # train.py:121
    outb, _ = ops.tile_normalize(allv, yx, (16, 16), mean, std, out_dtype=torch.bfloat16)
    assert torch.equal(outb[0].cpu(), torch.from_numpy(g["normalize.out"]).to(torch.bfloat16))


@pytest.mark.parametrize("Cb,H,W,p", [(3, 300, 500, 224), (13, 520, 1030, 256), (4, 100, 100, 37), (1, 64, 64, 64)])
def test_tile_random_multiband(Cb, H, W, p):
    from cvcs_b200 import ops
    g = torch.Generator().manual_seed(Cb * H)
    scene = torch.randint(0, 256, (Cb, H, W), generator=g, dtype=torch.uint8)
    lab = torch.randint(0, 20, (H, W), generator=g, dtype=torch.uint8)
    rows, cols = torch_path.tiles_in_image(H, W, p)
    yx = [(r * p, c * p) for r in range(rows) for c in range(cols)]
    yx += [(-5, -7), (H - p // 2, W - p // 2), (3, 1)]               # shifted / overhanging tiles
    yx = np.array(yx, dtype=np.int32)
    mean = (torch.rand(Cb, generator=g) * 100).numpy().astype(np.float32)
    std = (torch.rand(Cb, generator=g) * 50 + 1).numpy().astype(np.float32)
    ref, ref_lab = c_oracle.tile(scene.numpy(), yx, p, p, mean, std, labels=lab.numpy())
    hist = torch.zeros(22, dtype=torch.int64, device=DEV)
    out, lo = ops.tile_normalize(scene.to(DEV), torch.from_numpy(yx).to(DEV), (p, p), torch.from_numpy(mean).to(DEV),
                                 torch.from_numpy(std).to(DEV), label=lab.to(DEV), hist=hist, hist_classes=20)
    assert np.array_equal(out.cpu().numpy(), ref)
    assert np.array_equal(lo.cpu().numpy(), ref_lab)
    assert np.array_equal(hist.cpu().numpy(), c_oracle.label_hist(ref_lab, 20))
    out64, lo64 = ops.tile_normalize(scene.to(DEV), torch.from_numpy(yx).to(DEV), (p, p), label=lab.to(DEV),
                                     label_out_dtype=torch.int64)
    assert np.array_equal(lo64.cpu().numpy(), ref_lab.astype(np.int64))
    assert np.array_equal(out64.cpu().numpy(), c_oracle.tile(scene.numpy(), yx, p, p)[0])


# ---- N2 / N3 / N4 ----------------------------------------------------------------------------------
def test_vote_colorize_stitch(golden):
    from cvcs_b200 import ops
    g = golden("misc_cases")
    for n in (2, 3, 4, 5):
        maps = torch.from_numpy(g[f"vote{n}.in"]).to(DEV)
        assert np.array_equal(ops.vote(maps).cpu().numpy(), g[f"vote{n}.out"])          # torch.mode (utils.py:506)
        assert np.array_equal(ops.vote(maps.to(torch.uint8)).cpu().numpy(), g[f"vote{n}.out"])
    out = ops.colorize(torch.from_numpy(g["iconvert.in"]).to(DEV), torch.from_numpy(g["iconvert.lut"]).to(DEV))
    assert np.array_equal(out.cpu().numpy(), g["iconvert.out"])                        # converters.py:23-36
    tiles = torch.arange(2 * 6 * 6, dtype=torch.uint8).reshape(2, 6, 6)
    yx = torch.tensor([[0, 0], [0, 4]], dtype=torch.int32)
    scene = ops.stitch(tiles.to(DEV), yx.to(DEV), (4, 8), crop_hw=(4, 4))
    assert np.array_equal(scene.cpu().numpy(), c_oracle.stitch(tiles.numpy(), yx.numpy(), 4, 8, crop=(4, 4)))


def test_tile_argmax_stitch_roundtrip():
    """Size-independent property: tiling a label scene and stitching the tiles back is the identity
    on the covered area (remainder rows/cols dropped, dataset.py:63,125)."""
    from cvcs_b200 import ops
    g = torch.Generator().manual_seed(3)
    H, W, p = 1000, 1300, 224
    lab = torch.randint(0, 16, (H, W), generator=g, dtype=torch.uint8).to(DEV)
    rows, cols = H // p, W // p
    yx = torch.tensor([(r * p, c * p) for r in range(rows) for c in range(cols)], dtype=torch.int32, device=DEV)
    tiles, _ = ops.tile_normalize(lab[None], yx, (p, p), out_dtype=torch.uint8)
    back = ops.stitch(tiles[:, 0].contiguous(), yx, (H, W))
    assert torch.equal(back[:rows * p, :cols * p], lab[:rows * p, :cols * p])
    assert int(back[rows * p:].sum()) == 0 and int(back[:, cols * p:].sum()) == 0


@pytest.mark.parametrize("n_maps", [1, 2, 3, 5, 8, 9])
@pytest.mark.parametrize("n", [4096, 4099])
def test_vote_vector_and_scalar_paths(n_maps, n):
    """Majority vote (torch.mode tie rule) on u8 maps: the 4-pixel register path (n % 4 == 0, <= 8 maps) and
    the generic path agree with the oracle."""
    from cvcs_b200 import ops
    g = torch.Generator().manual_seed(n_maps * 31 + n)
    maps = torch.randint(0, 6, (n_maps, n), generator=g, dtype=torch.uint8)
    ref = c_oracle.vote(maps.numpy().astype(np.int64))
    assert np.array_equal(ops.vote(maps.to(DEV)).cpu().numpy().astype(np.int64), ref)
    assert np.array_equal(ops.vote(maps.long().to(DEV)).cpu().numpy(), ref)
    assert np.array_equal(torch.mode(maps.long(), dim=0).values.numpy(), ref)     # the oracle is torch.mode (utils.py:506)


def test_stitch_vector_path_with_center_crop():
    """Border-corrected stitching (CenterCrop of the padded tile, utils.py:146,154): 4-pixel path vs oracle."""
    from cvcs_b200 import ops
    g = torch.Generator().manual_seed(11)
    n, th, ch, H, W = 6, 32, 24, 80, 104
    tiles = torch.randint(0, 16, (n, th, th), generator=g, dtype=torch.uint8)
    yx = torch.tensor([[0, 0], [0, 24], [24, 48], [56, 80], [-8, 100], [72, -4]], dtype=torch.int32)
    out = ops.stitch(tiles.to(DEV), yx.to(DEV), (H, W), crop_hw=(ch, ch))
    assert np.array_equal(out.cpu().numpy(), c_oracle.stitch(tiles.numpy(), yx.numpy(), H, W, crop=(ch, ch)))


@pytest.mark.parametrize("n,offset", [(4096, 0), (4099, 0), (3, 0), (4100, 1), (64, 4)])
def test_colorize_paths(n, offset):
    """iconvert (converters.py:23-36) on u8 and int64 maps of odd lengths and alignments; indices outside the table
    keep torch.ones' white."""
    from cvcs_b200 import ops
    g = torch.Generator().manual_seed(n + offset)
    C = 16
    lut = torch.rand(C, 3, generator=g)
    base = torch.randint(0, C + 3, (n + offset,), generator=g, dtype=torch.uint8).to(DEV)
    idx = base[offset:]                                                   # offset != 0: not 4-byte aligned
    ref = c_oracle.colorize(idx.cpu().numpy(), lut.numpy())
    assert np.array_equal(ops.colorize(idx, lut.to(DEV)).cpu().numpy(), ref)
    assert np.array_equal(ops.colorize(idx.long(), lut.to(DEV)).cpu().numpy(), ref)


@pytest.mark.parametrize("th,ch", [(64, 64), (64, 32), (48, 16), (32, 22)])
def test_stitch_16_pixel_path(th, ch):
    """The 128-bit stitch path (crop, offset, widths multiples of 16): tiles on the grid, off the 16-pixel grid, partly
    and wholly outside the scene, ch not a multiple of the rows a thread walks — all equal to the oracle."""
    from cvcs_b200 import ops
    g = torch.Generator().manual_seed(th * 100 + ch)
    H, W = 300, 416                                   # the placed crops never overlap (overlaps are write races by contract)
    yx = torch.tensor([[0, 0], [0, 64], [64, 128], [128, 192], [-16, 272], [280, -16], [71, 19], [150, 333], [900, 0], [200, 400]], dtype=torch.int32)
    tiles = torch.randint(1, 16, (yx.shape[0], th, th), generator=g, dtype=torch.uint8)
    out = ops.stitch(tiles.to(DEV), yx.to(DEV), (H, W), crop_hw=(ch, ch))
    assert np.array_equal(out.cpu().numpy(), c_oracle.stitch(tiles.numpy(), yx.numpy(), H, W, crop=(ch, ch)))


@pytest.mark.parametrize("layout", ["NCHW", "NHWC"])
@pytest.mark.parametrize("dtype,label_dtype", [(torch.float32, torch.uint8), (torch.float32, torch.int64), (torch.bfloat16, torch.uint8)])
@pytest.mark.parametrize("C,H,W", [(7, 64, 128), (16, 48, 80), (20, 32, 64), (5, 37, 41)])
def test_eval_fused_metrics_mode(layout, dtype, label_dtype, C, H, W):
    """cvcs_eval_fused (K1 without softmax / loss): argmax map + confusion matrix in one read of the logits,
    bit-exact vs the oracle, including rows with NaN / inf, ties, ignored and out-of-range labels."""
    from cvcs_b200 import ops
    g = torch.Generator().manual_seed(C * H + W)
    B = 3
    x = (torch.randn(B, C, H, W, generator=g) * 3).to(dtype).float()
    x[0, :, 0, :8] = 1.0                                         # ties -> first index
    x[0, 2, 1, :4] = float("nan")                                # NaN is maximal, first NaN wins
    x[0, 4, 1, 2:6] = float("nan")
    x[1, 1, 2, :3] = float("inf")
    x[1, 3, 2, 1:5] = float("-inf")
    x[2, :, 3, :2] = float("-inf")
    t = torch.randint(0, C, (B, H, W), generator=g)
    t[0, 5] = 255                                                # ignored
    t[1, 6, :3] = C + 1                                          # out of range -> status
    xd = x.to(dtype).to(DEV)
    if layout == "NHWC":
        xd = xd.contiguous(memory_format=torch.channels_last)
    td = t.to(label_dtype).to(DEV)
    am = torch.full((B, H, W), 99, dtype=torch.uint8, device=DEV)
    cm = torch.zeros((C, C), dtype=torch.int64, device=DEV)
    st = torch.zeros(1, dtype=torch.int64, device=DEV)
    ops.eval_fused(xd, td, 255, argmax=am, confmat=cm, status=st)
    am_ref = c_oracle.argmax(x.numpy())
    assert np.array_equal(am.cpu().numpy().astype(np.int64), am_ref)
    keep = (t.numpy() != C + 1)
    cm_ref, _ = c_oracle.confmat(am_ref[keep], t.numpy()[keep], C, 255)
    assert np.array_equal(cm.cpu().numpy(), cm_ref)
    assert int(st.item()) == 3
    # confusion only / argmax only
    cm2 = torch.zeros_like(cm)
    ops.eval_fused(xd, td, 255, confmat=cm2)
    assert torch.equal(cm2, cm)
    am64 = torch.empty((B, H, W), dtype=torch.int64, device=DEV)
    ops.eval_fused(xd, td, None, argmax=am64)
    assert np.array_equal(am64.cpu().numpy(), am_ref)


@pytest.mark.parametrize("n", [0, 1, 2, 7, 4096, 100003])
@pytest.mark.parametrize("C,ignore", [(7, 255), (7, 0), (16, -100), (254, 3)])
def test_labels_prepare_i64(n, C, ignore):
    """cvcs_labels_prepare: int64 labels -> {Σ v·w[y], 1/Σ} and the byte labels K1 reads (255 ignored, 254 out of range)."""
    from cvcs_b200 import ops
    g = torch.Generator().manual_seed(n + C)
    t = torch.randint(0, C, (n,), generator=g, dtype=torch.int64)
    if n >= 7:
        t[1] = ignore
        t[2] = C + 5            # out of range
        t[3] = -7               # out of range (negative), unless it is the ignore value
        t[5] = 1 << 40          # out of range, does not fit 32 bits
    w = torch.rand(C, generator=g) + 0.5
    tw, t8 = ops.labels_prepare(t.to(DEV), C, ignore, w.to(DEV))
    tn = t.numpy()
    valid = (tn >= 0) & (tn < C) & (tn != ignore)
    want = np.where(tn == ignore, 255, np.where((tn >= 0) & (tn < C), tn, 254)).astype(np.uint8)
    assert np.array_equal(t8.cpu().numpy(), want)
    sw = float(w.double().numpy()[tn[valid]].sum()) if n else 0.0
    if sw > 0:
        assert abs(float(tw[0]) - sw) <= 1e-6 * sw and abs(float(tw[1]) * sw - 1.0) <= 1e-6
