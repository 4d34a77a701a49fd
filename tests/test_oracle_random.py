"""The C restatement against the library calls the reference makes (torch CPU: nn.CrossEntropyLoss, torch.max,
torch.mode, torchvision crop / Normalize), on seeded random inputs beyond the committed goldens: ragged shapes, class
weights with zeros, ignore_index inside and outside the class range, NaN / inf / tied logits, bf16-valued logits."""
import zlib

import numpy as np
import pytest
import torch

from oracle import c_oracle, torch_path

SHAPES = [(1, 2, 1, 1), (2, 3, 5, 7), (3, 7, 16, 16), (1, 16, 9, 33), (2, 20, 8, 8), (1, 150, 4, 4)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("mode", ["plain", "weights", "ignore_in", "ignore_out", "weights_ignore_zero_class", "bf16vals"])
def test_cross_entropy_argmax_confusion(shape, mode):
    B, C, H, W = shape
    g = torch.Generator().manual_seed(zlib.crc32(repr((shape, mode)).encode()))
    x = torch.randn(shape, generator=g) * 4
    if mode == "bf16vals":
        x = x.to(torch.bfloat16).float()                     # many exact ties
    t = torch.randint(0, C, (B, H, W), generator=g)
    w, ii = None, -100
    if mode in ("weights", "weights_ignore_zero_class"):
        w = torch.rand(C, generator=g) + 0.25
    if mode == "ignore_in":
        ii = C - 1
    if mode == "ignore_out":
        ii = 255
        t[torch.rand((B, H, W), generator=g) < 0.3] = 255
    if mode == "weights_ignore_zero_class":
        w[0] = 0.0                                            # wCEL with an absent class (dataset.py:376-380)
        ii = 0
    loss_ref, grad_ref = torch_path.ce_loss_and_grad(x, t, w, ii)
    loss, sums, grad = c_oracle.cross_entropy(x.numpy(), t.numpy(), None if w is None else w.numpy(), ii)
    if torch.isnan(loss_ref):
        assert np.isnan(loss)
    else:
        # the library result is fp32: 1e-5 relative, plus its absolute rounding of lse - x_t when the loss is ~0
        assert abs(loss - float(loss_ref)) <= 1e-5 * abs(float(loss_ref)) + 1e-6
    gr = np.nan_to_num(grad_ref.numpy())
    assert np.abs(grad - gr).max() <= 1e-5 * max(np.abs(gr).max(), 1e-30)
    assert sums[2] == 0
    am = c_oracle.argmax(x.numpy())
    assert np.array_equal(am, torch.max(x, dim=1)[1].numpy())          # first maximal index (utils.py:90)
    for ign in (None, 0, ii if 0 <= ii < 256 else None):
        ref = torch_path.RestatedConfusionMatrix(C, ignore_index=ign)
        keep = t < C
        ref.update(torch.from_numpy(am)[keep], t[keep])
        cm, bad = c_oracle.confmat(am, t.numpy(), C, ign)
        assert np.array_equal(cm, ref.compute().numpy())
        assert bad == int(((t >= C) & (t != (ign if ign is not None else -1))).sum())


def test_argmax_special_values():
    """torch.max's rule: NaN is maximal and the first NaN wins; otherwise the first maximal value; -inf rows give 0."""
    nan, inf = float("nan"), float("inf")
    rows = [[1, 2, 3], [3, 3, 1], [nan, 1, 2], [1, nan, nan], [inf, nan, inf], [-inf, -inf, -inf], [inf, inf, 0],
            [-0.0, 0.0, -1], [0.0, -0.0, -1], [-inf, 5, inf]]
    x = torch.tensor(rows, dtype=torch.float32).T.reshape(1, 3, 1, len(rows)).contiguous()
    assert np.array_equal(c_oracle.argmax(x.numpy())[0, 0], torch.max(x, dim=1)[1][0, 0].numpy())
    nhwc = np.ascontiguousarray(np.moveaxis(x.numpy(), 1, -1))
    assert np.array_equal(c_oracle.argmax(nhwc, "NHWC")[0, 0], torch.max(x, dim=1)[1][0, 0].numpy())


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8])
def test_vote_is_torch_mode(n):
    g = torch.Generator().manual_seed(n)
    maps = torch.randint(0, 4, (n, 13, 17), generator=g)
    assert np.array_equal(c_oracle.vote(maps.numpy()), torch.mode(maps, dim=0)[0].numpy())     # utils.py:506


@pytest.mark.parametrize("Cb,H,W,p", [(1, 9, 9, 4), (3, 50, 70, 32), (13, 40, 40, 16)])
def test_tile_is_torchvision_crop_and_normalize(Cb, H, W, p):
    g = torch.Generator().manual_seed(Cb + H)
    scene = torch.randint(0, 256, (Cb, H, W), generator=g, dtype=torch.uint8)
    lab = torch.randint(0, 16, (H, W), generator=g, dtype=torch.uint8)
    mean = (torch.rand(Cb, generator=g) * 100).tolist()
    std = (torch.rand(Cb, generator=g) * 50 + 1).tolist()
    yx = np.array([(0, 0), (H - p, W - p), (H - p // 2, W - p // 2), (-3, -1), (5, 2)], dtype=np.int32)
    out, lo = c_oracle.tile(scene.numpy(), yx, p, p, np.array(mean, np.float32), np.array(std, np.float32), labels=lab.numpy())
    raw, _ = c_oracle.tile(scene.numpy(), yx, p, p)
    for i, (y, x) in enumerate(yx):
        crop = torch_path.crop(scene, int(y), int(x), p, p)                                    # dataset.py:29-31
        assert np.array_equal(raw[i], torch_path.cast_normalize(crop).numpy())                 # train.py:121
        assert np.array_equal(out[i], torch_path.cast_normalize(crop, mean, std).numpy())      # nets.py:339-342
        assert np.array_equal(lo[i], torch_path.crop(lab[None], int(y), int(x), p, p)[0].numpy())
