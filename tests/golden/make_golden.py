"""Generate the golden vectors under tests/golden/ by running the REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Every output below comes from the reference's unmodified Python (imported through
oracle/ref_shim.py) or from the torch / torchvision call the cited reference line makes.  The
.npz files are committed; tests read them on machines where /root/reference does not exist.

  ce_cases.npz       utils.load_loss(...)(logits, target.long()) + .backward()  (utils.py:223-242,
                     train.py:122-125), CEL / CEL+ignore_background / wCEL, fp32 and bf16-valued
  argmax_cases.npz   torch.max(y_pred, dim=0) (utils.py:90) incl. ties, NaN, +-inf
  eval_cases.npz     utils.eval_model(fake net, fake loader) (utils.py:59-103) -> flat/normalised
                     matrices, utils.print_metrics (utils.py:375-403)
  metrics_cases.npz  IoU/F1/precision/recall/accuracy on hand-made matrices (utils.py:301-373)
  metrics_random.npz the same functions + print_metrics on 24 seeded random matrices
  weight_cases.npz   Loader.get_class_weights on 16 seeded random class-count vectors (dataset.py:360-384)
  dataset_cases.npz  dataset.Loader on a tiny on-disk GID-like tree: tile order, crops, class
                     counts and weights (dataset.py:28-32,105-221,241-384); crop helpers with
                     out-of-bounds offsets (dataset.py:11-32); Normalize (nets.py:339-342)
  misc_cases.npz     GID15Converter.iconvert (converters.py:23-36); torch.mode vote (utils.py:504-507)
  context_cases.npz  dataset._get_context (dataset.py:11-16) with the reference's own resizer (v2.Resize(p), dataset.py:65,131)
                     on uint8 tv_tensors.Image scenes: interior, scene-border and partly-outside patch origins, p = 8 / 32 / 224
"""
from __future__ import annotations

import json
import os
import random
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

utils, dataset, converters = ref_shim.load()


def blocky_labels(gen, shape, n_classes, block):
    """labels constant over block x block squares (real masks have long runs)."""
    *lead, H, W = shape
    small = torch.randint(0, n_classes, (*lead, (H + block - 1) // block, (W + block - 1) // block), generator=gen)
    return small.repeat_interleave(block, -2).repeat_interleave(block, -1)[..., :H, :W].contiguous()


def ce_cases():
    out = {}
    g = torch.Generator().manual_seed(0)
    specs = [
        # name, B, C, H, W, loss, ignore_background, extra
        ("cel_c7", 2, 7, 16, 16, "CEL", False, {}),
        ("cel_c7_ignore0", 2, 7, 16, 16, "CEL", True, {}),
        ("wcel_c7", 2, 7, 16, 16, "wCEL", False, {}),
        ("wcel_c7_ignore0", 2, 7, 16, 16, "wCEL", True, {}),
        ("cel_c16", 1, 16, 32, 32, "CEL", True, {}),
        ("wcel_c16", 1, 16, 16, 32, "wCEL", False, {}),   # load_loss prints labels[i]: C <= 16 (utils.py:236)
        ("cel_c7_bf16vals", 2, 7, 16, 16, "CEL", False, {"bf16": True}),
        ("cel_c7_allignored", 1, 7, 16, 16, "CEL", True, {"all_ignored": True}),
        ("cel_c3_odd", 3, 3, 5, 7, "CEL", False, {}),
    ]

    class CountsOnly:  # the only thing load_loss('wCEL') needs from the Loader
        def __init__(self, counts):
            self.counts = counts

        get_class_weights = dataset.Loader.get_class_weights

        def _get_class_count(self, classes):
            return self.counts

    for name, B, C, H, W, loss, ib, extra in specs:
        logits = torch.randn(B, C, H, W, generator=g) * 3
        if extra.get("bf16"):
            logits = logits.to(torch.bfloat16).to(torch.float32)
        target = blocky_labels(g, (B, H, W), C, 4)
        if extra.get("all_ignored"):
            target.zero_()
        cfg = {"num_classes": C - 1, "loss": loss, "ignore_background": ib}
        ds = None
        if loss == "wCEL":
            counts = torch.bincount(target.flatten(), minlength=C).to(torch.float32)
            counts[C - 1] = 0  # a class that never occurs -> weight 0
            target[target == C - 1] = 1
            ds = CountsOnly(counts)
        crit = utils.load_loss(cfg, "cpu", ds)
        x = logits.clone().requires_grad_(True)
        lossv = crit(x, target.type(torch.long))
        lossv.backward()
        out[f"{name}.logits"] = logits.numpy()
        out[f"{name}.target"] = target.numpy().astype(np.int64)
        out[f"{name}.ignore_index"] = np.int64(crit.ignore_index)
        out[f"{name}.weight"] = crit.weight.numpy() if crit.weight is not None else np.zeros(0, np.float32)
        out[f"{name}.loss"] = lossv.detach().numpy()
        out[f"{name}.grad"] = x.grad.numpy()
        # forward-only call, as utils.validation_loss makes it (utils.py:109,120)
        with torch.no_grad():
            out[f"{name}.loss_nograd"] = crit(logits, target.type(torch.long)).numpy()
    np.savez_compressed(os.path.join(HERE, "ce_cases.npz"), **out)
    return [s[0] for s in specs]


def argmax_cases():
    nan, inf = float("nan"), float("inf")
    rows = torch.tensor([
        [1.0, 3.0, 3.0, 2.0],      # tie -> first
        [nan, 1.0, 2.0, nan],      # NaN is maximal, first NaN wins
        [-inf, -inf, -inf, -inf],  # all -inf -> 0
        [inf, 1.0, inf, 0.0],      # +inf tie -> first
        [0.0, -0.0, 0.0, -0.0],    # signed zeros compare equal -> 0
        [1.0, nan, inf, 2.0],      # NaN beats +inf
        [5.0, 4.0, 3.0, 6.0],
        [2.0, 2.0, 2.0, 2.0],
    ])
    y_pred = rows.t().contiguous().reshape(4, 2, 4)      # [C, H, W] as after .squeeze() (utils.py:88)
    _, pred = torch.max(y_pred, dim=0)                   # utils.py:90
    hwc = y_pred.permute(1, 2, 0)
    pred2 = torch.argmax(hwc, dim=2)                     # utils.py:158
    g = torch.Generator().manual_seed(1)
    big = torch.randn(7, 24, 40, generator=g)
    big[:, ::3, ::5] = big[0:1, ::3, ::5]                # plant exact ties across all classes
    _, predbig = torch.max(big, dim=0)
    np.savez_compressed(os.path.join(HERE, "argmax_cases.npz"), small=y_pred.numpy(), small_max=pred.numpy(),
                        small_argmax_hwc=pred2.numpy(), big=big.numpy(), big_max=predbig.numpy())


class _FakeNet(torch.nn.Module):
    """Returns pre-computed logits tile by tile (the segmenter itself is off the path)."""
    requires_context = False
    returns_logits = True

    def __init__(self, logits):
        super().__init__()
        self.logits = logits
        self.i = 0

    def forward(self, x, context=None):
        out = self.logits[self.i:self.i + x.shape[0]]
        self.i += x.shape[0]
        return out


class _FakeChunk(torch.utils.data.IterableDataset):
    def __init__(self, items):
        self.patches = items
        self.chunk_crops = list(range(len(items)))

    def __iter__(self):
        return iter(self.patches)


class _FakeLoader:
    def __init__(self, chunks):
        self.chunks = chunks

    def __len__(self):
        return len(self.chunks)

    def get_iterable_chunk(self, c):
        return _FakeChunk(self.chunks[c])


def eval_cases():
    out = {}
    g = torch.Generator().manual_seed(2)
    N, C, H, W = 6, 16, 16, 16   # eval_model hard-codes 16 classes (utils.py:77-78)
    logits = torch.randn(N, C, H, W, generator=g) * 2
    labels = blocky_labels(g, (N, H, W), 12, 4).to(torch.uint8)   # classes 12..15 never occur
    items = [(torch.zeros(3, H, W, dtype=torch.uint8), labels[i], torch.tensor([0]), torch.tensor([0])) for i in range(N)]
    loader = _FakeLoader([items[:3], items[3:]])
    out["logits"], out["labels"] = logits.numpy(), labels.numpy()
    for ib in (False, True):
        net = _FakeNet(logits)
        flat, normalized = utils.eval_model(net, loader, "cpu", batch_size=1, show_progress=False, ignore_background=ib)
        cm = flat.compute()
        tag = f"ib{int(ib)}"
        out[f"{tag}.flat"] = cm.numpy()
        out[f"{tag}.normalized"] = normalized.compute().numpy()
        m = utils.print_metrics(cm, silent=True)
        out[f"{tag}.perclass_IoU"] = np.array(m["perclass_IoU"], dtype=np.float64)
        out[f"{tag}.scalars"] = np.array([m["mIoU"], m["precision_score"], m["recall_score"], m["dice_score"],
                                          m["oa_score"]], dtype=np.float64)
        _, excluded = utils.IoU(cm, mean=False, return_excluded=True)
        out[f"{tag}.excluded"] = np.array(excluded, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "eval_cases.npz"), **out)


def metrics_cases():
    out = {}
    g = torch.Generator().manual_seed(3)
    mats = {
        "dense7": torch.randint(0, 1000, (7, 7), generator=g),
        "absent_row": torch.tensor([[5, 1, 0], [2, 3, 0], [0, 0, 0]]),
        "absent_col": torch.tensor([[5, 1, 0], [2, 3, 0], [4, 0, 0]]),
        "row0_empty_col0_not": torch.tensor([[0, 0, 0], [3, 9, 1], [2, 0, 7]]),   # ignore_index=0 shape
        "big_counts": torch.randint(0, 2 ** 40, (16, 16), generator=g),
        "diag": torch.diag(torch.arange(1, 8)),
    }
    for name, cm in mats.items():
        out[f"{name}.cm"] = cm.numpy().astype(np.int64)
        for kind, fn, kw in (("iou", utils.IoU, "mean"), ("f1", utils.F1, "mean"),
                             ("precision", utils.precision, "macro"), ("recall", utils.recall, "macro")):
            scores, excluded = fn(cm, **{kw: False}, return_excluded=True)
            out[f"{name}.{kind}.scores"] = scores.numpy()
            out[f"{name}.{kind}.excluded"] = np.array(excluded, dtype=np.int64)
            out[f"{name}.{kind}.mean"] = np.float64(fn(cm, **{kw: True}))
        out[f"{name}.accuracy"] = np.float64(utils.accuracy(cm))
    np.savez_compressed(os.path.join(HERE, "metrics_cases.npz"), **out)


def metrics_random_cases(n=24):
    """Seeded random confusion matrices (2..16 classes, empty rows / columns, counts up to 2^36) through the
    reference's IoU / F1 / precision / recall / accuracy and print_metrics (utils.py:301-403)."""
    out = {}
    g = torch.Generator().manual_seed(17)
    for i in range(n):
        C = int(torch.randint(2, 17, (1,), generator=g))
        hi = int(2 ** int(torch.randint(3, 37, (1,), generator=g)))
        cm = torch.randint(0, hi, (C, C), generator=g)
        for axis in (0, 1):                                  # knock out some rows / columns (absent classes)
            dead = torch.rand(C, generator=g) < 0.2
            if axis == 0:
                cm[dead, :] = 0
            else:
                cm[:, dead] = 0
        if i % 6 == 0:
            cm[0, :] = 0                                     # the ignore_background shape: row 0 empty
        out[f"r{i}.cm"] = cm.numpy().astype(np.int64)
        for kind, fn, kw in (("iou", utils.IoU, "mean"), ("f1", utils.F1, "mean"),
                             ("precision", utils.precision, "macro"), ("recall", utils.recall, "macro")):
            scores, excluded = fn(cm, **{kw: False}, return_excluded=True)
            out[f"r{i}.{kind}.scores"] = scores.numpy()
            out[f"r{i}.{kind}.excluded"] = np.array(excluded, dtype=np.int64)
            out[f"r{i}.{kind}.mean"] = np.float64(fn(cm, **{kw: True}))
        out[f"r{i}.accuracy"] = np.float64(utils.accuracy(cm))
        if cm.sum() > 0:
            m = utils.print_metrics(cm, silent=True)
            out[f"r{i}.print.keys"] = np.array(sorted(m.keys()))
            for k, v in m.items():
                out[f"r{i}.print.{k}"] = np.asarray(v, dtype=np.float64)
    out["n"] = np.int64(n)
    np.savez_compressed(os.path.join(HERE, "metrics_random.npz"), **out)


def weight_cases(n=16):
    """dataset.Loader.get_class_weights (dataset.py:360-384) on seeded random class counts: empty classes, huge and
    tiny counts, with and without ignore_background."""
    class CountsOnly:
        def __init__(self, counts):
            self.counts = counts

        get_class_weights = dataset.Loader.get_class_weights

        def _get_class_count(self, classes):
            return self.counts

    out = {}
    g = torch.Generator().manual_seed(23)
    for i in range(n):
        C = int(torch.randint(2, 21, (1,), generator=g))
        hi = int(2 ** int(torch.randint(2, 34, (1,), generator=g)))
        counts = torch.randint(0, hi, (C,), generator=g).to(torch.float32)     # _get_class_count returns float32
        counts[torch.rand(C, generator=g) < 0.25] = 0
        if i == 0:
            counts[:] = 0
            counts[1] = 5
        out[f"w{i}.counts"] = counts.numpy()
        for ib in (False, True):
            out[f"w{i}.ib{int(ib)}"] = CountsOnly(counts.clone()).get_class_weights(C, ib).numpy()
    out["n"] = np.int64(n)
    np.savez_compressed(os.path.join(HERE, "weight_cases.npz"), **out)


def dataset_cases():
    from PIL import Image
    out = {}
    H, W, p = 230, 460, 224
    rng = np.random.RandomState(4)
    with tempfile.TemporaryDirectory() as root:
        for sub in ("Image__8bit_NirRGB", "Annotation__index", "Annotation__color"):
            os.makedirs(os.path.join(root, sub))
        names = ["scene_b", "scene_a"]   # Loader sorts the listing
        for si, stem in enumerate(sorted(names)):
            yy, xx = np.mgrid[0:H, 0:W]
            img = np.stack([(yy * (k + 1) + xx * (si + 2) + 17 * k) % 256 for k in range(4)], axis=-1).astype(np.uint8)
            lab = (rng.randint(0, 6, ((H + 31) // 32, (W + 31) // 32)).repeat(32, 0).repeat(32, 1)[:H, :W]).astype(np.uint8)
            col = np.stack([lab * 10, lab * 20, 255 - lab * 5], axis=-1).astype(np.uint8)
            Image.fromarray(img, "RGBA").save(os.path.join(root, "Image__8bit_NirRGB", stem + ".png"))
            Image.fromarray(lab, "L").save(os.path.join(root, "Annotation__index", stem + "_15label.png"))
            Image.fromarray(col, "RGB").save(os.path.join(root, "Annotation__color", stem + "_15label.tif"))
            out[f"scene{si}.image"] = img.transpose(2, 0, 1).copy()   # CHW, what tv_tensors.Image yields
            out[f"scene{si}.label"] = lab
            out[f"scene{si}.color"] = col.transpose(2, 0, 1).copy()
        for shift in (False, True):
            random.seed(1234)
            L = dataset.Loader(root, chunk_size=2, random_shift=shift, patch_size=p, load_context=False,
                               load_color_mask=True)
            chunk = L.get_iterable_chunk(0)
            tag = f"shift{int(shift)}"
            out[f"{tag}.len"] = np.int64(len(L))
            out[f"{tag}.tpi"] = np.int64(chunk.tpi)
            out[f"{tag}.chunk_crops"] = np.array(chunk.chunk_crops, dtype=np.int64)
            out[f"{tag}.patches"] = np.stack([t[0].numpy() for t in chunk.patches])
            out[f"{tag}.index_masks"] = np.stack([t[1].numpy() for t in chunk.patches])
            out[f"{tag}.color_masks"] = np.stack([t[2].numpy() for t in chunk.patches])
        L = dataset.Loader(root, chunk_size=1, patch_size=p, load_context=False, load_color_mask=False)
        out["weights_ib0"] = L.get_class_weights(16, False).numpy()
        out["counts"] = L.count.numpy()
        L2 = dataset.Loader(root, chunk_size=1, patch_size=p, load_context=False, load_color_mask=False)
        out["weights_ib1"] = L2.get_class_weights(16, True).numpy()
    # crop helpers with out-of-bounds offsets (zero padding)
    img = torch.arange(3 * 10 * 12, dtype=torch.uint8).reshape(3, 10, 12)
    msk = (torch.arange(10 * 12, dtype=torch.uint8) % 7).reshape(1, 10, 12)
    out["crop.image"], out["crop.mask"] = img.numpy(), msk.numpy()
    cases = [(0, 0, 4), (-2, -3, 6), (7, 9, 6), (4, 4, 4), (-5, 8, 8)]
    out["crop.cases"] = np.array(cases, dtype=np.int64)
    for i, (tly, tlx, q) in enumerate(cases):
        a, b, c = dataset._get_cropped_data(img, msk, msk, tly, tlx, q)
        out[f"crop.{i}.patch"], out[f"crop.{i}.mask"] = a.numpy(), b.numpy()
    out["padded.patch"] = dataset._get_padded_patch(img, 2, 2, (4, 4), 6).numpy()   # margin = bc - p = 2
    out["padded.corner"] = dataset._get_padded_patch(img, 0, 0, (4, 4), 6).numpy()
    # the SegFormer preprocessor (nets.py:339-342): ToDtype(float32) then Normalize(ImageNet stats)
    from torchvision.transforms import v2
    pre = v2.Compose([v2.ToDtype(torch.float32),
                      v2.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    allv = torch.arange(256, dtype=torch.uint8).reshape(1, 16, 16).repeat(3, 1, 1)
    out["normalize.in"] = allv.numpy()
    out["normalize.out"] = pre(allv).numpy()
    out["cast.out"] = allv.type(torch.float32).numpy()   # train.py:121
    np.savez_compressed(os.path.join(HERE, "dataset_cases.npz"), **out)


def misc_cases():
    out = {}
    conv = converters.GID15Converter()
    g = torch.Generator().manual_seed(5)
    idx = torch.randint(0, 16, (12, 20), generator=g)
    idx[0, 0] = 200  # no colour for it: stays (1,1,1)
    out["iconvert.in"] = idx.numpy().astype(np.int64)
    out["iconvert.out"] = conv.iconvert(idx).numpy()
    out["iconvert.lut"] = (torch.tensor(list(conv.color_to_label.keys())).type(torch.float32) / 255).numpy()
    for n_maps in (2, 3, 4, 5):
        stack = torch.randint(0, 6, (n_maps, 16, 24), generator=g)
        values, _ = torch.mode(stack, dim=0)      # utils.py:506
        out[f"vote{n_maps}.in"] = stack.numpy().astype(np.int64)
        out[f"vote{n_maps}.out"] = values.numpy().astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "misc_cases.npz"), **out)


def context_cases():
    """The context view of a patch, by the reference's own function and resizer objects (only written when run
    separately with --context, so that regenerating it does not touch the other committed files)."""
    import torchvision.transforms as v2            # the module dataset.py imports as v2
    from torchvision import tv_tensors
    out = {}
    g = torch.Generator().manual_seed(11)
    for name, cb, H, W, p in (("p8", 4, 50, 70, 8), ("p32", 3, 160, 130, 32), ("p224", 4, 700, 900, 224)):
        scene = tv_tensors.Image(torch.randint(0, 256, (cb, H, W), generator=g, dtype=torch.uint8))
        # smooth structure as well as noise: half of the scene is a gradient + blocks
        yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
        scene[:, :, : W // 2] = ((yy * 3 + xx * 5)[:, : W // 2] % 256).to(torch.uint8)
        resizer = v2.Resize(p, interpolation=v2.InterpolationMode.BILINEAR)     # dataset.py:131 (and :65 default)
        origins = [(0, 0), (p, p), (H - p, W - p), (H // 2, W // 3), (-p // 2, 5), (H - 3, W - 3), (-p, 3 - p),
                   (3, W - p - 1), (2 * p + 1, 0)]
        ctx = [dataset._get_context(scene, tly, tlx, p, resizer) for tly, tlx in origins]
        assert all(c.dtype == torch.uint8 and tuple(c.shape) == (cb, p, p) for c in ctx)
        out[f"{name}.scene"] = scene.numpy()
        out[f"{name}.yx"] = np.array(origins, dtype=np.int32)
        out[f"{name}.p"] = np.int64(p)
        out[f"{name}.context"] = torch.stack([torch.as_tensor(c) for c in ctx]).numpy()
    np.savez_compressed(os.path.join(HERE, "context_cases.npz"), **out)
    return sorted({k.split(".")[0] for k in out})


if __name__ == "__main__":
    if "--context" in sys.argv:
        print("context cases:", context_cases())
        sys.exit(0)
    names = ce_cases()
    argmax_cases()
    eval_cases()
    metrics_cases()
    metrics_random_cases()
    weight_cases()
    dataset_cases()
    misc_cases()
    manifest = {
        "generator": "tests/golden/make_golden.py",
        "reference": "theElandor/CVCS @ /root/reference (unmodified, imported via oracle/ref_shim.py)",
        "torch": torch.__version__,
        "ce_case_names": names,
        "files": sorted(f for f in os.listdir(HERE) if f.endswith(".npz")),
    }
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    for fn in sorted(os.listdir(HERE)):
        print(fn, os.path.getsize(os.path.join(HERE, fn)))
