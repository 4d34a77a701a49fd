"""Drop-ins for the reference's metric boundary.

* ``MulticlassConfusionMatrix`` — the torchmetrics class the reference instantiates twice per
  evaluation (utils.py:77-78) and updates per tile (utils.py:93-94).  Same constructor keywords
  (``num_classes, normalize, ignore_index``), ``update(preds, target)``, ``compute()``, picklable
  (the reference stores the metric objects in its checkpoints, utils.py:139-140).  The state
  lives on the GPU as int64[C,C] and is updated by K3 (index maps) or K1 (logits).
* ``eval_model`` / ``validation_loss`` — same signatures as utils.py:59-103 / :106-126, but the
  logits never leave the GPU, batches may be larger than 1, and both returned views share a
  single accumulated state.
* ``IoU / F1 / precision / recall / accuracy / print_metrics`` — the reference's formulas
  (utils.py:301-403) evaluated with the same arithmetic (Python ints -> float64 per class,
  float32 ``torch.mean`` over the classes present), so that a bit-exact matrix gives a
  bit-exact mIoU.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import ops
from .loss import GID15_LABELS

_NORMALIZE = (None, "none", "true", "pred", "all")


class MulticlassConfusionMatrix:
    """GPU confusion matrix with torchmetrics' interface (rows = target, cols = prediction)."""

    def __init__(self, num_classes: int, normalize: Optional[str] = None, ignore_index: Optional[int] = None,
                 validate_args: bool = True, device=None, _shared_state: Optional[dict] = None):
        if normalize not in _NORMALIZE:
            raise ValueError(f"Argument `normalize` needs to one of the following: {_NORMALIZE[1:]}")
        if not isinstance(num_classes, int) or num_classes < 2:
            raise ValueError(f"Argument `num_classes` to be an integer larger than 1, but got {num_classes}")
        self.num_classes = num_classes
        self.normalize = normalize
        self.ignore_index = ignore_index
        self.validate_args = validate_args
        self._device = torch.device(device) if device is not None else None
        # state shared between views: {'confmat': int64[C,C] cuda tensor | None, 'status': int64[1] | None,
        #                              'host': int64[C,C] cpu tensor (restored from a pickle) | None}
        self._s = _shared_state if _shared_state is not None else {"confmat": None, "status": None, "host": None}

    # -- state -----------------------------------------------------------------------------------
    def _state_for(self, dev: torch.device, num_classes: Optional[int] = None) -> torch.Tensor:
        if num_classes is not None and num_classes != self.num_classes:
            raise RuntimeError(f"confusion matrix was built for {self.num_classes} classes, got logits with {num_classes}")
        cm = self._s["confmat"]
        if cm is None or cm.device != dev:
            new = torch.zeros((self.num_classes, self.num_classes), dtype=torch.int64, device=dev)
            if cm is not None:
                new += cm.to(dev)
            if self._s["host"] is not None:
                new += self._s["host"].to(dev)
                self._s["host"] = None
            self._s["confmat"] = new
            self._s["status"] = torch.zeros(1, dtype=torch.int64, device=dev)
        return self._s["confmat"]

    def view(self, normalize: Optional[str]) -> "MulticlassConfusionMatrix":
        """Another metric object over the SAME accumulated state (e.g. the row-normalised view)."""
        return MulticlassConfusionMatrix(self.num_classes, normalize, self.ignore_index, self.validate_args,
                                         self._device, _shared_state=self._s)

    def reset(self) -> None:
        if self._s["confmat"] is not None:
            self._s["confmat"].zero_()
            self._s["status"].zero_()
        self._s["host"] = None

    def to(self, device) -> "MulticlassConfusionMatrix":
        self._device = torch.device(device)
        return self

    # -- update ----------------------------------------------------------------------------------
    def _target_device(self, *ts: torch.Tensor) -> torch.device:
        for t in ts:
            if t.is_cuda:
                return t.device
        if self._device is not None and self._device.type == "cuda":
            return self._device
        return torch.device("cuda", torch.cuda.current_device())

    def update(self, preds: torch.Tensor, target: torch.Tensor) -> None:
        """preds: index map with target's shape, or logits/probabilities [N, C, ...] (argmax over
        dim 1, as torchmetrics does for floating point input)."""
        dev = self._target_device(preds, target)
        preds, target = preds.to(dev, non_blocking=True), target.to(dev, non_blocking=True)
        if target.dtype not in (torch.uint8, torch.int64):
            target = target.to(torch.int64)
        if preds.is_floating_point():
            if preds.dim() != target.dim() + 1:
                raise ValueError("floating point `preds` must have one more dimension than `target`")
            self.update_from_logits(preds, target)
            return
        if preds.dtype not in (torch.uint8, torch.int64):
            preds = preds.to(torch.int64)
        if preds.numel() != target.numel():
            raise ValueError("The `preds` and `target` should have the same shape")
        cm = self._state_for(dev)
        ops.confmat_update(cm, preds, target, self.num_classes, self.ignore_index, status=self._s["status"])

    def update_from_logits(self, logits: torch.Tensor, target: torch.Tensor) -> None:
        """Fused argmax + confusion update straight from [B,C,H,W] (or [C,H,W]) logits (K1, metrics mode:
        one read of the logits, no softmax)."""
        dev = self._target_device(logits, target)
        logits, target = logits.to(dev), target.to(dev)
        if logits.dim() == 3:
            logits = logits.unsqueeze(0)
        if logits.dim() != 4:
            b, c = logits.shape[:2]
            logits = logits.reshape(b, c, -1, 1)
        if logits.dtype not in (torch.float32, torch.bfloat16):
            logits = logits.float()
        B, C, H, W = logits.shape
        target = target.reshape(B, H, W)
        if target.dtype not in (torch.uint8, torch.int64):
            target = target.to(torch.int64)
        cm = self._state_for(dev, C)
        ops.eval_fused(logits, target, self.ignore_index, confmat=cm, status=self._s["status"])

    __call__ = update

    # -- compute ---------------------------------------------------------------------------------
    def _note_bad_labels(self, loss_sums: torch.Tensor) -> None:
        """The fused loss kernel counts labels outside [0, C) that are not ignore_index in loss_sums[2]; feed them into
        this metric's validate_args status so that compute() raises as update() would have (no host read here)."""
        if self._s["status"] is not None:
            self._s["status"].add_(loss_sums[2:3].to(torch.int64))

    def sync(self, group=None) -> None:
        """COLLECTIVE: sum the state over all ranks of ``group`` (one all-reduce of C*C int64 plus the status word).
        Every rank of the group must call it exactly once per evaluation pass, and only when the ranks evaluated
        DISJOINT tile sets (``cvcs_b200.shard.local_tiles``): ranks that each evaluated the full validation set would
        count every pixel world_size times.  A rank-0-only ``validate()`` must not call it (the others never arrive)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            return
        cm = self._s["confmat"]
        if cm is None and (self._s["host"] is not None or not torch.cuda.is_available()):
            # state restored from a checkpoint and not touched on a GPU since (or a CPU-only process, e.g. the gloo
            # tests): reduce the host copy
            if self._s["host"] is None:
                self._s["host"] = torch.zeros((self.num_classes, self.num_classes), dtype=torch.int64)
            dist.all_reduce(self._s["host"], op=dist.ReduceOp.SUM, group=group)
            return
        if cm is None:
            dev = self._device or torch.device("cuda", torch.cuda.current_device())
            cm = self._state_for(dev)
        dist.all_reduce(cm, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(self._s["status"], op=dist.ReduceOp.SUM, group=group)

    def compute(self) -> torch.Tensor:
        """int64[C,C] on the CPU (normalize None) or the float32 normalised matrix (NaN -> 0)."""
        if self._s["confmat"] is not None:
            cm = self._s["confmat"].cpu()  # the one device->host read of an evaluation pass
            if self.validate_args and int(self._s["status"].item()) > 0:
                raise RuntimeError(f"Detected {int(self._s['status'].item())} prediction/target values outside "
                                   f"[0, {self.num_classes}) (and not equal to ignore_index)")
        elif self._s["host"] is not None:
            cm = self._s["host"].clone()
        else:
            cm = torch.zeros((self.num_classes, self.num_classes), dtype=torch.int64)
        return normalize_confmat(cm, self.normalize)

    # -- pickling (checkpoints hold the metric objects, utils.py:128-142) ---------------------------
    def __getstate__(self):
        cm = self._s["confmat"].cpu() if self._s["confmat"] is not None else self._s["host"]
        return {"num_classes": self.num_classes, "normalize": self.normalize, "ignore_index": self.ignore_index,
                "validate_args": self.validate_args, "confmat": cm}

    def __setstate__(self, st):
        self.num_classes = st["num_classes"]
        self.normalize = st["normalize"]
        self.ignore_index = st["ignore_index"]
        self.validate_args = st.get("validate_args", True)
        self._device = None
        self._s = {"confmat": None, "status": None, "host": st["confmat"]}


def normalize_confmat(cm: torch.Tensor, normalize: Optional[str]) -> torch.Tensor:
    """torchmetrics' reduction: 'true' rows, 'pred' columns, 'all' everything; NaN -> 0."""
    if normalize in (None, "none"):
        return cm
    cm = cm.float() if not cm.is_floating_point() else cm
    if normalize == "true":
        cm = cm / cm.sum(dim=-1, keepdim=True)
    elif normalize == "pred":
        cm = cm / cm.sum(dim=-2, keepdim=True)
    else:
        cm = cm / cm.sum(dim=[-2, -1], keepdim=True)
    cm[torch.isnan(cm)] = 0
    return cm


# ---- the reference's metric formulas (utils.py:301-403) ------------------------------------------
def _tp_fp_fn(confusion: torch.Tensor) -> List[Tuple[int, int, int]]:
    cm = confusion.detach().to("cpu")
    diag = cm.diagonal()
    col = cm.sum(dim=0)
    row = cm.sum(dim=1)
    return [(diag[i].item(), (col[i] - diag[i]).item(), (row[i] - diag[i]).item()) for i in range(cm.shape[1])]


_FORMULAS = {
    # name: (score(tp, fp, fn), excluded-when(tp, fp, fn))
    "precision": (lambda tp, fp, fn: tp / (tp + fp), lambda tp, fp, fn: tp + fp == 0),
    "recall": (lambda tp, fp, fn: tp / (tp + fn), lambda tp, fp, fn: tp + fn == 0),
    "iou": (lambda tp, fp, fn: tp / (tp + fn + fp), lambda tp, fp, fn: tp + fn == 0),
    "f1": (lambda tp, fp, fn: (2 * tp) / (2 * tp + fn + fp), lambda tp, fp, fn: tp + fn == 0),
}


def _score(confusion: torch.Tensor, kind: str, macro: bool, return_excluded: bool):
    formula, skip = _FORMULAS[kind]
    per_class, excluded = [], []
    for i, (tp, fp, fn) in enumerate(_tp_fp_fn(confusion)):
        if skip(tp, fp, fn):
            per_class.append(0)
            excluded.append(i)
        else:
            per_class.append(formula(tp, fp, fn))
    scores = torch.tensor(per_class)  # float32 (int64 if every class is excluded), as in the reference
    kept = torch.tensor([x for i, x in enumerate(scores) if i not in excluded])
    mean = torch.mean(kept).item()    # float32 mean over the classes present in the target
    if macro:
        return (mean, excluded) if return_excluded else mean
    return (scores, excluded) if return_excluded else mean  # (sic) the reference returns the mean here too


def precision(confusion, macro=False, return_excluded=False):
    return _score(confusion, "precision", macro, return_excluded)


def recall(confusion, macro=False, return_excluded=False):
    return _score(confusion, "recall", macro, return_excluded)


def IoU(confusion, mean=False, return_excluded=False):
    return _score(confusion, "iou", mean, return_excluded)


def F1(confusion, mean=False, return_excluded=False):
    return _score(confusion, "f1", mean, return_excluded)


def accuracy(confusion):
    cm = confusion.detach().to("cpu")
    return cm.diagonal().sum().item() / cm.sum().item()


def print_metrics(confusion, silent=False, labels=GID15_LABELS):
    out = {
        "perclass_IoU": None,
        "mIoU": IoU(confusion, mean=True),
        "precision_score": precision(confusion, macro=True),
        "recall_score": recall(confusion, macro=True),
        "dice_score": F1(confusion, mean=True),
        "oa_score": accuracy(confusion),
    }
    values, excluded = IoU(confusion, mean=False, return_excluded=True)
    out["perclass_IoU"] = values.tolist()
    if not silent:
        rows = [("mIoU", out["mIoU"]), ("mPrec", out["precision_score"]), ("mRec", out["recall_score"]),
                ("Dice", out["dice_score"]), ("OA", out["oa_score"])]
        width = max(len(str(labels.get(i, i))) for i in range(len(out["perclass_IoU"])))
        for name, val in rows:
            print(f"{name:>{width}} | {val}")
        print(f"Excluded classes (not in target): {list(excluded)}")
        for i, v in enumerate(out["perclass_IoU"]):
            print(f"{str(labels.get(i, i)):>{width}} | {v}", flush=True)
    return out


# ---- evaluation loops (utils.py:59-126) -----------------------------------------------------------
def eval_model(net, Loader_validation, device, batch_size=1, show_progress=False, ignore_background=False,
               num_classes: int = 16, sync_ranks: bool = False):
    """Same contract as the reference's ``utils.eval_model`` (utils.py:59-103): returns
    ``(flat_confusion_metric, normalized_confusion_metric)``.  ``num_classes`` defaults to the 16
    the reference hard-codes (utils.py:77-78).  Differences, all internal: logits stay on the
    GPU, argmax + confusion run fused (K1, forward only), any ``batch_size`` works, and the two
    returned metrics are views of one state.

    ``sync_ranks`` (default False, the reference is single-process): pass True ONLY when torch.distributed is initialised,
    EVERY rank calls eval_model, and ``Loader_validation`` hands each rank a disjoint share of the tiles — then the
    returned metrics hold the global counts (``MulticlassConfusionMatrix.sync``, a collective).  With the default each
    rank returns what it evaluated itself, which for an unsharded loader is exactly the single-process result."""
    net.eval()
    ignored_index = 0 if ignore_background else None
    flat = MulticlassConfusionMatrix(num_classes=num_classes, ignore_index=ignored_index, device=device)
    normalized = flat.view("true")
    with torch.no_grad():
        for c in range(len(Loader_validation)):
            dataset = Loader_validation.get_iterable_chunk(c)
            dl = torch.utils.data.DataLoader(dataset, batch_size=batch_size)
            for i, (x, y, _, context) in enumerate(dl):
                x, y = x.to(device), y.to(device)
                if net.requires_context:
                    context = context.to(device)
                y_pred = net(x.type(torch.float32), context.type(torch.float32))
                if net.returns_logits:
                    flat.update_from_logits(y_pred, y.reshape(y_pred.shape[0], *y_pred.shape[-2:]))
                else:  # the net already returns class indices (Ensemble, utils.py:89)
                    flat.update(y_pred.reshape(-1), y.reshape(-1))
                if show_progress:
                    print(f"chunk {c + 1} batch {i + 1}", end="\r")
            print("Updating confusion matrix...")
    if sync_ranks:
        flat.sync()
    return flat, normalized


def validation_loss(net, Loader_validation, crit, device, bs, show_progress=False):
    """Same contract as utils.py:106-126: list of per-batch loss values (one host sync at the end
    instead of one ``.item()`` per batch)."""
    losses = []
    net.eval()
    with torch.no_grad():
        for c in range(len(Loader_validation)):
            dataset = Loader_validation.get_iterable_chunk(c)
            dl = torch.utils.data.DataLoader(dataset, batch_size=bs)
            for image, index_mask, _, context in dl:
                image, mask = image.to(device), index_mask.to(device)
                if net.requires_context:
                    context = context.to(device)
                mask_pred = net(image.type(torch.float32), context.type(torch.float32)).to(device)
                if mask.dim() == 4:
                    mask = mask.squeeze(1)
                if mask.dtype not in (torch.uint8, torch.int64):
                    mask = mask.type(torch.long)
                losses.append(crit(mask_pred, mask).detach().reshape(1))
    if not losses:
        return []
    return torch.cat(losses).float().cpu().tolist()
