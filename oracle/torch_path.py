"""The reference's own CPU execution of the hot path, restated call for call.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  theElandor/CVCS is pure Python; on this path
it makes the library calls below (file:line are into /root/reference/source/scripts/).  This
module makes the same calls on CPU tensors, so it is both an oracle (the "repo's own PyTorch
path" BASELINE.json asks parity against) and the timed CPU baseline of bench.py.

  crit = nn.CrossEntropyLoss(weight=w, ignore_index=ii)         utils.py:230,238
  loss = crit(mask_pred, mask.type(torch.long)); loss.backward() train.py:122,125
  _, pred_mask = torch.max(y_pred, dim=0)   (per tile, CPU)      utils.py:88-90
  MulticlassConfusionMatrix.update(p, t) x2 (flat + normalised)  utils.py:91-94
  Loader._get_class_count / get_class_weights                    dataset.py:346-384
  v2.functional.crop(image, tly, tlx, p, p)                      dataset.py:28-32
  IoU / F1 / precision / recall / accuracy                       utils.py:301-373

torchmetrics is not installable here (no network), so its ``MulticlassConfusionMatrix`` is
restated from its published algorithm (torchmetrics/functional/classification/confusion_matrix.py,
unpinned in README.MD:28): flatten; keep = target != ignore_index; bincount(target*C + preds,
minlength=C*C).reshape(C, C); normalize='true' divides rows by their sum with NaN -> 0.
tests/test_oracle.py cross-checks it against sklearn.metrics.confusion_matrix.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn


class RestatedConfusionMatrix:
    """torchmetrics.classification.MulticlassConfusionMatrix, restated (index inputs only)."""

    def __init__(self, num_classes: int, normalize: Optional[str] = None, ignore_index: Optional[int] = None):
        self.num_classes = num_classes
        self.normalize = normalize
        self.ignore_index = ignore_index
        self.confmat = torch.zeros(num_classes, num_classes, dtype=torch.long)

    def update(self, preds: torch.Tensor, target: torch.Tensor) -> None:
        preds, target = preds.flatten(), target.flatten()
        if self.ignore_index is not None:
            keep = target != self.ignore_index
            preds, target = preds[keep], target[keep]
        mapping = (target * self.num_classes + preds).to(torch.long)
        bins = torch.bincount(mapping, minlength=self.num_classes ** 2)
        self.confmat += bins.reshape(self.num_classes, self.num_classes)

    def compute(self) -> torch.Tensor:
        cm = self.confmat
        if self.normalize in (None, "none"):
            return cm
        cm = cm.float()
        if self.normalize == "true":
            cm = cm / cm.sum(dim=-1, keepdim=True)
        elif self.normalize == "pred":
            cm = cm / cm.sum(dim=-2, keepdim=True)
        elif self.normalize == "all":
            cm = cm / cm.sum(dim=[-2, -1], keepdim=True)
        cm[torch.isnan(cm)] = 0
        return cm


def make_criterion(weight: Optional[torch.Tensor], ignore_index: int) -> nn.Module:
    """utils.load_loss's two CE branches (utils.py:229-238)."""
    if weight is None:
        return nn.CrossEntropyLoss(ignore_index=ignore_index)
    return nn.CrossEntropyLoss(weight=weight, ignore_index=ignore_index)


def ce_loss_and_grad(logits: torch.Tensor, target: torch.Tensor, weight: Optional[torch.Tensor],
                     ignore_index: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """train.py:122-125 on a leaf logits tensor: (loss, dloss/dlogits)."""
    x = logits.detach().clone().requires_grad_(True)
    crit = make_criterion(weight, ignore_index)
    loss = crit(x, target.type(torch.long))
    loss.backward()
    return loss.detach(), x.grad


def ce_loss_only(logits: torch.Tensor, target: torch.Tensor, weight: Optional[torch.Tensor],
                 ignore_index: int) -> torch.Tensor:
    """utils.validation_loss's call under no_grad (utils.py:109,120)."""
    with torch.no_grad():
        return make_criterion(weight, ignore_index)(logits, target.type(torch.long))


def eval_tiles(logits: torch.Tensor, labels: torch.Tensor, num_classes: int, ignore_background: bool,
               double_update: bool = True):
    """utils.eval_model's inner loop (utils.py:85-94) for logits [B,C,H,W] the net would have
    produced: one tile at a time, argmax on CPU, two metric updates on identical data."""
    ignored_index = 0 if ignore_background else None
    normalized = RestatedConfusionMatrix(num_classes, "true", ignored_index)
    flat = RestatedConfusionMatrix(num_classes, None, ignored_index)
    preds = []
    for b in range(logits.shape[0]):
        y_pred = logits[b:b + 1].squeeze().cpu()
        _, pred_mask = torch.max(y_pred, dim=0)
        p = pred_mask.unsqueeze(0).type(torch.int64).reshape(1, -1)
        t = labels[b:b + 1].cpu().type(torch.int64).reshape(1, -1)
        if double_update:
            normalized.update(p, t)
        flat.update(p, t)
        preds.append(pred_mask)
    return flat, normalized, torch.stack(preds)


def hot_path_step(logits: torch.Tensor, labels: torch.Tensor, weight: Optional[torch.Tensor], ignore_index: int,
                  num_classes: int, ignore_background_eval: Optional[bool] = None):
    """One pass of the whole path the way the reference executes it on the host: CE forward +
    backward over the batch, then per-tile argmax + confusion updates.  Returns
    (loss, grad, flat confusion tensor, argmax maps)."""
    loss, grad = ce_loss_and_grad(logits, labels, weight, ignore_index)
    ib = (ignore_index == 0) if ignore_background_eval is None else ignore_background_eval
    flat, _, preds = eval_tiles(logits, labels, num_classes, ib)
    return loss, grad, flat.compute(), preds


# ---- dataset side -----------------------------------------------------------------------------------
def tiles_in_image(H: int, W: int, p: int) -> Tuple[int, int]:
    """dataset.py:63,125 — remainder pixels are dropped."""
    return H // p, W // p


def tile_top_left(idx: int, tpi: int, cols: int, p: int) -> Tuple[int, int, int]:
    """dataset.py:137-140 / :82-85 — global tile index -> (image, tly, tlx), row-major."""
    image = idx // tpi
    t = idx % tpi
    return image, (t // cols) * p, (t % cols) * p


def crop(image: torch.Tensor, tly: int, tlx: int, h: int, w: int) -> torch.Tensor:
    """dataset.py:29-31 — torchvision's crop (zero padding outside the image)."""
    import torchvision.transforms as v2
    return v2.functional.crop(image, tly, tlx, h, w)


def cast_normalize(tiles_u8: torch.Tensor, mean: Optional[List[float]] = None,
                   std: Optional[List[float]] = None) -> torch.Tensor:
    """train.py:121 `.type(torch.float32)`; with mean/std: SegformerMod.preprocessor (nets.py:339-342)."""
    x = tiles_u8.type(torch.float32)
    if mean is None:
        return x
    from torchvision.transforms import v2
    return v2.Normalize(mean=mean, std=std)(x)


def class_count(masks: List[torch.Tensor], classes: int) -> torch.Tensor:
    """dataset.py:352-358 — float32 accumulators, one `torch.sum(mask == cl)` per class per scene."""
    count = torch.zeros(classes, dtype=torch.float32)
    for mask in masks:
        for cl in range(classes):
            count[cl] += torch.sum(mask == cl)
    return count


def class_weights(counts: torch.Tensor, ignore_background: bool = False) -> torch.Tensor:
    """dataset.py:371-384."""
    w = []
    if ignore_background:
        counts = counts[1:]
    numerator = torch.sum(counts)
    bincount = len(counts)
    for class_count_ in counts:
        if class_count_.item() == 0:
            w.append(0)
        else:
            w.append(numerator / (bincount * class_count_.item()))
    if ignore_background:
        return torch.concat((torch.tensor([0]), torch.tensor(w)), dim=0)
    return torch.tensor(w)


# ---- metric formulas (utils.py:301-373), plain-Python restatement ----------------------------------
def class_scores(confusion: torch.Tensor, kind: str):
    """Per-class score list (0 for excluded classes) and the excluded class indices."""
    C = confusion.shape[1]
    scores, excluded = [], []
    for i in range(C):
        tp = confusion[i, i].item()
        fp = (torch.sum(confusion[:, i]) - tp).item()
        fn = (torch.sum(confusion[i, :]) - tp).item()
        if kind == "precision":
            skip, val = tp + fp == 0, (tp / (tp + fp) if tp + fp else 0)
        elif kind == "recall":
            skip, val = tp + fn == 0, (tp / (tp + fn) if tp + fn else 0)
        elif kind == "iou":
            skip, val = tp + fn == 0, (tp / (tp + fn + fp) if tp + fn + fp else 0)
        elif kind == "f1":
            skip, val = tp + fn == 0, ((2 * tp) / (2 * tp + fn + fp) if 2 * tp + fn + fp else 0)
        else:
            raise ValueError(kind)
        if skip:
            scores.append(0)
            excluded.append(i)
        else:
            scores.append(val)
    return scores, excluded


def macro_mean(scores, excluded) -> float:
    """utils.py:343-346 — float32 mean over the classes that are present."""
    st = torch.tensor(scores)
    return torch.mean(torch.tensor([x for i, x in enumerate(st) if i not in excluded])).item()


def overall_accuracy(confusion: torch.Tensor) -> float:
    C = confusion.shape[1]
    return sum(confusion[i, i].item() for i in range(C)) / torch.sum(confusion).item()


# ---- BASELINE.json configs[0]: the reference's own CPU-runnable case ------------------------------------------
def unet_like(num_classes: int) -> nn.Module:
    """A restatement of the reference's U-Net variant (nets.py:117-199 `Urnetv2`, blocks.py:8-50): five encoder
    levels of two 3x3 conv + BatchNorm + ReLU (64..1024 channels, 2x2 max-pool between levels), four
    transposed-conv upsamplings each followed by two 3x3 conv + ReLU + BatchNorm on the concatenated skip, and a
    final 1x1 conv.  Same layer sequence and widths, hence the same parameter count (31.04 M for 7 classes); used
    only as the timed CPU workload of cfg1 — the network itself is out of this repo's scope."""

    def enc(cin, cout):
        return [nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU()]

    def dec(cin, cout):
        return nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1), nn.ReLU(), nn.BatchNorm2d(cout),
                             nn.Conv2d(cout, cout, 3, padding=1), nn.ReLU(), nn.BatchNorm2d(cout))

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            widths = [64, 128, 256, 512, 1024]
            self.down = nn.ModuleList()
            cin = 3
            for i, wd in enumerate(widths):
                layers = ([nn.MaxPool2d(2, 2)] if i else []) + enc(cin, wd) + enc(wd, wd)
                self.down.append(nn.Sequential(*layers))
                cin = wd
            self.up = nn.ModuleList(nn.ConvTranspose2d(w2, w1, 2, stride=2) for w1, w2 in zip(widths[-2::-1], widths[:0:-1]))
            self.mix = nn.ModuleList(dec(2 * w1, w1) for w1 in widths[-2::-1])
            self.head = nn.Conv2d(widths[0], num_classes, 1)

        def forward(self, x):
            skips = []
            for blk in self.down:
                x = blk(x)
                skips.append(x)
            skips.pop()
            for up, mix in zip(self.up, self.mix):
                x = mix(torch.cat((skips.pop(), up(x)), 1))
            return self.head(x)

    return Net()


def cfg1_step(seed: int = 0, num_classes: int = 7, batch: int = 2, size: int = 512):
    """One step of cfg1 the way train.py / utils.eval_model run it on the host: u8 RGB tiles -> float32 (train.py:121),
    segmenter forward, CrossEntropyLoss(ignore_index=0) (utils.py:223-230 with ignore_background), per-tile
    torch.max + confusion update (utils.py:88-94), mean IoU (utils.py:301-346)."""
    torch.manual_seed(seed)
    net = unet_like(num_classes).eval()
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, 256, (batch, 3, size, size), generator=g, dtype=torch.uint8)
    y = torch.randint(0, num_classes, (batch, size, size), generator=g, dtype=torch.uint8)
    with torch.no_grad():
        logits = net(x.type(torch.float32))
        loss = ce_loss_only(logits, y, None, 0)
    flat, _, _ = eval_tiles(logits, y, num_classes, ignore_background=True)
    cm = flat.compute()
    scores, excluded = class_scores(cm, "iou")
    return {"model": "U-Net variant (restated nets.py Urnetv2)", "params": sum(p.numel() for p in net.parameters()),
            "loss": float(loss), "miou": macro_mean(scores, excluded)}
