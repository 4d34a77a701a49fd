"""Pin the oracle: the C restatement and the torch-call restatement against the golden vectors the
reference's own Python produced (tests/golden/make_golden.py), plus sklearn as an independent
check of the restated torchmetrics update."""
import numpy as np
import pytest
import torch

from oracle import c_oracle, torch_path

CE_CASES = ["cel_c7", "cel_c7_ignore0", "wcel_c7", "wcel_c7_ignore0", "cel_c16", "wcel_c16", "cel_c7_bf16vals",
            "cel_c3_odd"]


def _case(g, name):
    w = g[f"{name}.weight"]
    return (g[f"{name}.logits"], g[f"{name}.target"], (w if w.size else None), int(g[f"{name}.ignore_index"]),
            g[f"{name}.loss"], g[f"{name}.grad"])


@pytest.mark.parametrize("name", CE_CASES)
def test_c_oracle_cross_entropy_matches_reference(golden, name):
    logits, target, w, ii, loss_ref, grad_ref = _case(golden("ce_cases"), name)
    loss, sums, grad = c_oracle.cross_entropy(logits, target, w, ii, "NCHW")
    # fp64 definition vs the reference's fp32 library result: 1e-5 relative (BASELINE.json)
    assert abs(loss - float(loss_ref)) <= 1e-5 * abs(float(loss_ref))
    scale = np.abs(grad_ref).max()
    assert np.abs(grad - grad_ref).max() <= 1e-5 * scale
    assert np.array_equal(grad == 0, grad_ref == 0) or np.abs(grad[grad_ref == 0]).max() < 1e-12
    assert sums[2] == 0


@pytest.mark.parametrize("name", CE_CASES)
def test_c_oracle_nhwc_equals_nchw(golden, name):
    logits, target, w, ii, _, _ = _case(golden("ce_cases"), name)
    l0, _, g0 = c_oracle.cross_entropy(logits, target, w, ii, "NCHW")
    nhwc = np.ascontiguousarray(np.moveaxis(logits, 1, -1))
    l1, _, g1 = c_oracle.cross_entropy(nhwc, target, w, ii, "NHWC")
    assert l0 == l1
    assert np.array_equal(np.moveaxis(g1, -1, 1), g0)


def test_c_oracle_all_ignored_is_nan(golden):
    g = golden("ce_cases")
    logits, target, w, ii, loss_ref, grad_ref = _case(g, "cel_c7_allignored")
    loss, sums, grad = c_oracle.cross_entropy(logits, target, w, ii)
    assert np.isnan(loss_ref) and np.isnan(loss)
    assert sums[1] == 0
    # torch's backward of the 0/0 mean leaves zeros at ignored pixels
    assert np.all(grad == 0) and np.all(np.nan_to_num(grad_ref) == 0)


def test_c_oracle_out_of_bounds_label_is_flagged():
    logits = np.zeros((1, 3, 2, 2), np.float32)
    target = np.array([[[0, 3], [1, -1]]], np.int64)
    _, sums, _ = c_oracle.cross_entropy(logits, target, None, -100)
    assert sums[2] == 2
    with pytest.raises(IndexError):
        torch_path.ce_loss_only(torch.from_numpy(logits), torch.from_numpy(target), None, -100)


@pytest.mark.parametrize("name", CE_CASES + ["cel_c7_allignored"])
def test_torch_path_matches_reference_bitwise(golden, name):
    logits, target, w, ii, loss_ref, grad_ref = _case(golden("ce_cases"), name)
    wt = None if w is None else torch.from_numpy(w)
    loss, grad = torch_path.ce_loss_and_grad(torch.from_numpy(logits), torch.from_numpy(target), wt, ii)
    assert np.array_equal(loss.numpy(), loss_ref, equal_nan=True)
    assert np.array_equal(grad.numpy(), grad_ref, equal_nan=True)
    lo = torch_path.ce_loss_only(torch.from_numpy(logits), torch.from_numpy(target), wt, ii)
    assert np.array_equal(lo.numpy(), golden("ce_cases")[f"{name}.loss_nograd"], equal_nan=True)


def test_argmax_oracle_matches_torch_max(golden):
    g = golden("argmax_cases")
    small = g["small"]  # [C,H,W]
    assert np.array_equal(c_oracle.argmax(small[None])[0], g["small_max"])
    hwc = np.ascontiguousarray(np.moveaxis(small, 0, -1))
    assert np.array_equal(c_oracle.argmax(hwc[None], "NHWC")[0], g["small_argmax_hwc"])
    assert np.array_equal(g["small_max"], g["small_argmax_hwc"])
    assert np.array_equal(c_oracle.argmax(g["big"][None])[0], g["big_max"])


@pytest.mark.parametrize("ib", [0, 1])
def test_eval_oracle_matches_reference_eval_model(golden, ib):
    g = golden("eval_cases")
    logits, labels = g["logits"], g["labels"]
    pred = c_oracle.argmax(logits)
    cm, bad = c_oracle.confmat(pred, labels, 16, 0 if ib else None)
    assert bad == 0
    assert np.array_equal(cm, g[f"ib{ib}.flat"])
    flat, normalized, _ = torch_path.eval_tiles(torch.from_numpy(logits), torch.from_numpy(labels), 16, bool(ib))
    assert np.array_equal(flat.compute().numpy(), g[f"ib{ib}.flat"])
    assert np.array_equal(normalized.compute().numpy(), g[f"ib{ib}.normalized"])
    if ib:
        assert cm[0].sum() == 0 and cm[:, 0].sum() > 0   # row 0 empty, column 0 not


def test_restated_confusion_matrix_equals_sklearn():
    from sklearn.metrics import confusion_matrix
    g = torch.Generator().manual_seed(0)
    for C in (2, 7, 16, 20):
        t = torch.randint(0, C, (5000,), generator=g)
        p = torch.randint(0, C, (5000,), generator=g)
        m = torch_path.RestatedConfusionMatrix(C)
        m.update(p[:2000].reshape(1, -1), t[:2000].reshape(1, -1))
        m.update(p[2000:].reshape(1, -1), t[2000:].reshape(1, -1))
        ref = confusion_matrix(t.numpy(), p.numpy(), labels=list(range(C)))
        assert np.array_equal(m.compute().numpy(), ref)
        cm, bad = c_oracle.confmat(p.numpy(), t.numpy(), C)
        assert bad == 0 and np.array_equal(cm, ref)
        mi = torch_path.RestatedConfusionMatrix(C, ignore_index=0)
        mi.update(p, t)
        keep = t.numpy() != 0
        assert np.array_equal(mi.compute().numpy(), confusion_matrix(t.numpy()[keep], p.numpy()[keep], labels=list(range(C))))


def test_out_of_range_labels_against_sklearn():
    """torchmetrics' validate_args path: values outside [0, C) that are not ignore_index.  sklearn with labels=range(C)
    drops exactly those pairs; the oracle must count them (the product raises at compute()) and leave the same matrix —
    with and without ignore_index=0, and with the out-of-range value on either side."""
    from sklearn.metrics import confusion_matrix
    g = torch.Generator().manual_seed(1)
    for C in (3, 7, 16):
        t = torch.randint(0, C, (4000,), generator=g).numpy()
        p = torch.randint(0, C, (4000,), generator=g).numpy()
        t[::97] = C + 3          # out-of-range targets
        p[5::101] = C            # out-of-range predictions
        t[7::89] = 255           # a LoveDA-style ignore label
        for ign in (None, 0, 255):
            keep = np.ones_like(t, dtype=bool) if ign is None else (t != ign)
            ref = confusion_matrix(t[keep], p[keep], labels=list(range(C)))
            n_bad = int(((t[keep] >= C) | (p[keep] >= C)).sum())
            cm, bad = c_oracle.confmat(p, t, C, ign)
            assert np.array_equal(cm, ref) and bad == n_bad and bad > 0


@pytest.mark.parametrize("name", ["dense7", "absent_row", "absent_col", "row0_empty_col0_not", "big_counts", "diag"])
def test_metric_formula_restatement(golden, name):
    g = golden("metrics_cases")
    cm = torch.from_numpy(g[f"{name}.cm"])
    for kind in ("iou", "f1", "precision", "recall"):
        scores, excluded = torch_path.class_scores(cm, kind)
        assert np.array_equal(torch.tensor(scores).numpy(), g[f"{name}.{kind}.scores"])
        assert list(excluded) == list(g[f"{name}.{kind}.excluded"])
        m = torch_path.macro_mean(scores, excluded)
        assert m == float(g[f"{name}.{kind}.mean"]) or (np.isnan(m) and np.isnan(g[f"{name}.{kind}.mean"]))
    assert torch_path.overall_accuracy(cm) == float(g[f"{name}.accuracy"])


def test_label_hist_and_class_weights(golden):
    g = golden("dataset_cases")
    labs = [g["scene0.label"], g["scene1.label"]]
    hist = sum(c_oracle.label_hist(l, 16) for l in labs)
    assert hist[16] == 0 and hist[17] == 0
    assert np.array_equal(hist[:16].astype(np.float32), g["counts"])
    counts = torch_path.class_count([torch.from_numpy(l)[None] for l in labs], 16)
    assert np.array_equal(counts.numpy(), g["counts"])
    assert np.array_equal(torch_path.class_weights(counts, False).numpy(), g["weights_ib0"])
    assert np.array_equal(torch_path.class_weights(counts, True).numpy(), g["weights_ib1"])
    h2 = c_oracle.label_hist(np.array([0, 1, 255, 9, 255], np.uint8), 7, 255)
    assert h2[0] == 1 and h2[1] == 1 and h2[7] == 2 and h2[8] == 1


def test_tile_oracle_matches_reference_crops(golden):
    g = golden("dataset_cases")
    img, msk = g["crop.image"], g["crop.mask"]
    for i, (tly, tlx, q) in enumerate(g["crop.cases"]):
        out, lab = c_oracle.tile(img, np.array([[tly, tlx]]), int(q), int(q), labels=msk[0])
        assert np.array_equal(out[0].astype(np.uint8), g[f"crop.{i}.patch"])
        assert np.array_equal(lab[0], g[f"crop.{i}.mask"][0])
        ref = torch_path.crop(torch.from_numpy(img), int(tly), int(tlx), int(q), int(q))
        assert np.array_equal(ref.numpy(), g[f"crop.{i}.patch"])
    # padded patch: margin = bc - p (dataset.py:19)
    out, _ = c_oracle.tile(img, np.array([[2 - 2, 2 - 2]]), 6, 6)
    assert np.array_equal(out[0].astype(np.uint8), g["padded.patch"])
    out, _ = c_oracle.tile(img, np.array([[0 - 2, 0 - 2]]), 6, 6)
    assert np.array_equal(out[0].astype(np.uint8), g["padded.corner"])


def test_tile_oracle_loader_order_and_normalize(golden):
    g = golden("dataset_cases")
    p = 224
    tpi = int(g["shift0.tpi"])
    rows, cols = torch_path.tiles_in_image(230, 460, p)
    assert rows * cols == tpi == 2
    yx, scene_of = [], []
    for idx in g["shift0.chunk_crops"]:
        im, tly, tlx = torch_path.tile_top_left(int(idx), tpi, cols, p)
        yx.append((tly, tlx))
        scene_of.append(im)
    for k, (im, (tly, tlx)) in enumerate(zip(scene_of, yx)):
        out, lab = c_oracle.tile(g[f"scene{im}.image"], np.array([[tly, tlx]]), p, p, labels=g[f"scene{im}.label"])
        assert np.array_equal(out[0].astype(np.uint8), g["shift0.patches"][k])
        assert np.array_equal(lab[0], g["shift0.index_masks"][k])
    # cast and ImageNet normalisation, bit-exact in fp32
    allv = g["normalize.in"]
    mean, std = np.array([0.485, 0.456, 0.406], np.float32), np.array([0.229, 0.224, 0.225], np.float32)
    out, _ = c_oracle.tile(allv, np.array([[0, 0]]), 16, 16, mean, std)
    assert np.array_equal(out[0], g["normalize.out"])
    out, _ = c_oracle.tile(allv, np.array([[0, 0]]), 16, 16)
    assert np.array_equal(out[0], g["cast.out"])
    ref = torch_path.cast_normalize(torch.from_numpy(allv), [0.485, 0.456, 0.406], [0.229, 0.224, 0.225])
    assert np.array_equal(ref.numpy(), g["normalize.out"])


def test_vote_colorize_stitch(golden):
    g = golden("misc_cases")
    for n in (2, 3, 4, 5):
        assert np.array_equal(c_oracle.vote(g[f"vote{n}.in"]), g[f"vote{n}.out"])
    assert np.array_equal(c_oracle.colorize(g["iconvert.in"], g["iconvert.lut"]), g["iconvert.out"])
    tiles = np.arange(2 * 6 * 6, dtype=np.uint8).reshape(2, 6, 6)
    scene = c_oracle.stitch(tiles, np.array([[0, 0], [0, 4]]), 4, 8, crop=(4, 4))
    assert np.array_equal(scene[:, :4], tiles[0, 1:5, 1:5]) and np.array_equal(scene[:, 4:], tiles[1, 1:5, 1:5])


def test_stitch_center_offset_is_torchvisions_centercrop():
    """oracle.stitch's centred window == torchvision CenterCrop for every parity of (tile - crop), including the
    round-half-to-even cases (utils.py:146,154)."""
    import torch
    import torchvision.transforms as T
    for th, ch in [(8, 8), (9, 8), (10, 8), (11, 8), (12, 8), (13, 8), (15, 8), (40, 32), (39, 28), (35, 32)]:
        tile = (torch.arange(th * th, dtype=torch.int64) % 251).to(torch.uint8).reshape(1, th, th)
        want = T.CenterCrop(ch)(tile)[0].numpy()
        got = c_oracle.stitch(tile.numpy(), np.array([[0, 0]], dtype=np.int32), ch, ch, crop=(ch, ch))
        assert np.array_equal(got, want), (th, ch)
