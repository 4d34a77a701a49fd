"""Drop-in for the reference's loss boundary.

Reference: ``utils.load_loss`` (utils.py:223-242) returns ``nn.CrossEntropyLoss(weight,
ignore_index)``; the training loop calls ``loss = crit(mask_pred, mask.type(torch.long))``
followed by ``loss.item()`` and ``loss.backward()`` (train.py:122-125); validation calls the
same criterion under ``torch.no_grad()`` (utils.py:120).

``FusedCrossEntropyLoss`` keeps that call signature.  One CUDA pass (K1) produces the loss AND
the logit gradients (and, on request, the argmax map and the confusion-matrix update);
``backward`` only hands the stashed gradient to autograd.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops

# GID-15 class names, as printed by the reference's weight table (utils.py:23-40)
GID15_LABELS = {
    0: "unlabeled", 1: "industrial land", 2: "urban residential", 3: "rural residential", 4: "traffic land",
    5: "paddy field", 6: "irrigated cropland", 7: "dry cropland", 8: "garden plot", 9: "arbor forest",
    10: "shrub land", 11: "natural grassland", 12: "artificial grassland", 13: "river", 14: "lake", 15: "pond",
}


def _as_bchw(logits: torch.Tensor, target: torch.Tensor):
    """Normalise nn.CrossEntropyLoss input conventions to [B,C,H,W] / [B,H,W] views.
    Returns (logits4, target3, restore) where restore maps a gradient laid out like logits4
    back to the caller's input shape."""
    shape = logits.shape
    if logits.dim() == 4:
        return logits, target, (lambda d: d)
    if logits.dim() == 2:  # [N, C] with target [N]: N pixels, classes innermost (an NHWC image)
        n, c = shape
        x = logits.contiguous().view(1, n, 1, c).permute(0, 3, 1, 2)
        return x, target.reshape(1, n, 1), (lambda d: d.permute(0, 2, 3, 1).reshape(n, c))
    if logits.dim() == 3:  # [B, C, L]
        b, c, l = shape
        return logits.reshape(b, c, l, 1), target.reshape(b, l, 1), (lambda d: d.reshape(shape))
    if logits.dim() > 4:   # [B, C, d1, d2, ...]
        b, c = shape[:2]
        return logits.reshape(b, c, -1, 1), target.reshape(b, -1, 1), (lambda d: d.reshape(shape))
    raise RuntimeError(f"FusedCrossEntropyLoss: unsupported input shape {tuple(shape)}")


class _FusedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, module, want_grad):
        # grad mode is switched off inside Function.forward, so the caller decides want_grad
        loss, dlogits, restore = module._run(logits, target, want_grad)
        ctx.module = module
        ctx.restore = restore
        ctx.dlogits = dlogits
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        d = ctx.dlogits
        if d is None:
            if getattr(ctx, "consumed", False):
                raise RuntimeError("FusedCrossEntropyLoss: the gradient buffer of this forward was already handed to autograd "
                                   "(a second backward through the same loss, e.g. retain_graph=True, is not supported: "
                                   "call the criterion again)")
            raise RuntimeError("FusedCrossEntropyLoss: backward called but the forward ran without grad")
        ctx.dlogits = None
        ctx.consumed = True
        if ctx.module.grad_scale_mode != "unit":
            # dlogits *= grad_output on the device; the kernel returns at once when grad_output == 1.0, which is what
            # `loss.backward()` feeds (train.py:125) — no host read, no synchronisation, graph-capturable
            ops.scale_inplace(d, grad_output.to(torch.float32).reshape(1))
        # mode == "unit": trust the caller that grad_output == 1 (saves the ~3 us launch)
        return ctx.restore(d), None, None, None


class FusedCrossEntropyLoss(nn.Module):
    """``nn.CrossEntropyLoss(weight=..., ignore_index=..., reduction='mean')`` on the fused kernel.

    Extras over the reference signature (all optional):
      confusion      a ``cvcs_b200.metrics.MulticlassConfusionMatrix`` to update in the same pass
      return_argmax  keep the argmax map of the last call in ``self.last_argmax`` (uint8)
      grad_scale_mode  'check' (default) = 'scale': dlogits *= grad_output by a kernel that exits immediately when
                     grad_output == 1 (no host read); 'unit': skip even that launch (caller guarantees 1.0)
      exchange       a ``cvcs_b200.shard.WeightExchange``: the 'mean' divides by Σ v·w[y] over ALL ranks' batches, exchanged
                     inside the fused kernel over NVLink (no all-reduce, no extra launch)
      strict         synchronise after every call and raise IndexError on out-of-bounds labels
                     (default: the loss is NaN-poisoned and ``check_errors()`` raises)
    Targets may be int64 (as the reference passes) or uint8 (as the masks are stored: skips the
    8x wider label read).
    """

    def __init__(self, weight: Optional[torch.Tensor] = None, ignore_index: int = -100, reduction: str = "mean",
                 label_smoothing: float = 0.0, confusion=None, return_argmax: bool = False,
                 grad_scale_mode: str = "check", strict: bool = False, exchange=None):
        super().__init__()
        if reduction != "mean" or label_smoothing != 0.0:
            raise NotImplementedError("the reference only uses reduction='mean', label_smoothing=0 (utils.py:230,238)")
        if grad_scale_mode not in ("check", "unit", "scale"):
            raise ValueError("grad_scale_mode must be 'check', 'unit' or 'scale'")
        self.register_buffer("weight", None if weight is None else weight.detach().clone())
        self.ignore_index = int(ignore_index)
        self.confusion = confusion
        self.return_argmax = return_argmax
        self.grad_scale_mode = grad_scale_mode
        self.strict = strict
        self.exchange = exchange
        self.last_argmax: Optional[torch.Tensor] = None
        self.last_sums: Optional[torch.Tensor] = None  # f64[3] {Σ w·nll, Σ w, #out-of-bounds}
        self.last_total_weight: Optional[torch.Tensor] = None  # f64[2] {Σ v·w[y] (global), 1/Σ} when data dependent
        self._w32: Optional[torch.Tensor] = None
        self._w32_key = None
        self._prefetched = None   # (target tensor, its version, num_classes, f64[2] total weight, ready event, u8 labels)
        self._next_labels = None  # uint8 labels announced for the next batch: (tensor, version, num_classes)
        self._scanned = None      # (tensor, version, num_classes, f64[2]) summed by the previous launch for THIS batch
        self._side: Optional[torch.cuda.Stream] = None

    # -- helpers -------------------------------------------------------------------------------------
    def _weight_f32(self, dev: torch.device) -> Optional[torch.Tensor]:
        if self.weight is None:
            return None
        # `weight` is a registered buffer: load_state_dict / in-place edits must reach the kernel
        key = (dev, self.weight.data_ptr(), self.weight._version)
        if self._w32 is None or self._w32_key != key:
            self._w32 = self.weight.to(device=dev, dtype=torch.float32).contiguous()
            self._w32_key = key
        return self._w32

    def prefetch_total_weight(self, target: torch.Tensor, num_classes: int) -> None:
        """Optional: announce the labels of the NEXT batch as soon as they are on the device.
        int64 labels: the label pre-pass (Σ v·w[y] + the byte copy of the labels) starts NOW on a side stream — e.g.
        while the model's forward pass runs — and the ``forward`` that receives this same, unmodified tensor finds it done.
        uint8 labels: the next ``forward`` (of the CURRENT batch) sums the weights over them inside its own kernel launch
        (``next_target``), so that the forward of this batch starts from a finished sum: the pre-pass is pipelined
        across launches and never sits on the critical path."""
        if target.is_cuda and target.dtype == torch.uint8 and target.is_contiguous():
            # uint8 labels: no launch now — the NEXT forward call sums the weights over these labels inside its own
            # kernel (staged with its chunks) and the call after that, on this very tensor, starts from that sum
            self._next_labels = (target, target._version, num_classes)
            return
        if not target.is_cuda or target.dtype != torch.int64 or num_classes > 254:
            return
        dev = target.device
        t = target.reshape(target.shape[0], -1, 1) if target.dim() != 3 else target
        if self._side is None or self._side.device != dev:
            self._side = torch.cuda.Stream(device=dev)
        self._side.wait_stream(torch.cuda.current_stream(dev))     # the labels must have been produced
        with torch.cuda.stream(self._side):
            tw, t8 = ops.labels_prepare(t.contiguous(), num_classes, self.ignore_index, self._weight_f32(dev))
            ev = torch.cuda.Event()
            ev.record(self._side)
        self._prefetched = (target, target._version, num_classes, tw, ev, t8)

    def _run(self, logits: torch.Tensor, target: torch.Tensor, want_grad: bool):
        if not logits.is_cuda:
            raise RuntimeError("FusedCrossEntropyLoss needs CUDA tensors (cvcs_b200 has no CPU fallback)")
        if logits.dtype not in (torch.float32, torch.bfloat16):
            raise RuntimeError(f"FusedCrossEntropyLoss: logits must be float32 or bfloat16, got {logits.dtype}")
        if target.dtype not in (torch.int64, torch.uint8):
            # same complaint torch makes
            raise RuntimeError(f"expected scalar type Long but found {str(target.dtype).replace('torch.', '')}")
        x, t, restore = _as_bchw(logits, target)
        B, C, H, W = x.shape
        dev = x.device
        w = self._weight_f32(dev)
        if w is not None and w.numel() != C:
            raise RuntimeError(f"weight tensor should be defined either for all {C} classes or no classes "
                               f"but got weight tensor of shape: {list(w.shape)}")
        conf = None
        if self.confusion is not None:
            # the fused update filters labels with THIS criterion's ignore_index: a metric built with another one
            # would silently count different pixels than its own update() does
            m_ign = self.confusion.ignore_index
            mine = self.ignore_index if 0 <= self.ignore_index < C else None
            if (m_ign if (m_ign is not None and 0 <= m_ign < C) else None) != mine:
                raise RuntimeError(f"FusedCrossEntropyLoss(ignore_index={self.ignore_index}) cannot update a confusion matrix "
                                   f"built with ignore_index={m_ign}: the fused pass drops the loss's ignored pixels")
            conf = self.confusion._state_for(dev, C)
        xchg = self.exchange.handle_for(dev) if self.exchange is not None else None
        ii = self.ignore_index
        mode, inv_tw, tw = "given", 0.0, None
        if t.dtype == torch.int64 and C <= 254 and want_grad:
            # the reference passes mask.type(torch.long): ONE pass over the 8-byte labels yields Σ v·w[y] and a byte copy
            # (255 = ignored, 254 = out of range) that the fused kernel reads instead (1 B/px instead of 8 B/px)
            pf = self._prefetched
            if pf is not None and pf[0] is target and pf[1] == target._version and pf[2] == C:
                _, _, _, tw, ev, t8 = pf                          # already running / done on the side stream
                cur = torch.cuda.current_stream(dev)
                cur.wait_event(ev)
                tw.record_stream(cur)                             # consumed on this stream: keep the allocator away
                t8.record_stream(cur)
                t8 = t8.reshape(t.shape)
            else:
                tw, t8 = ops.labels_prepare(t, C, self.ignore_index, w)
            self._prefetched = None
            t, ii = t8, 255
        if want_grad:
            if xchg is not None:
                mode = "kernel"                                   # global Σw: exchanged inside the kernel
            elif tw is not None:
                pass                                              # int64 labels, single GPU: 1/Σ from the pre-pass
            elif w is None and not (0 <= ii <= 255):
                inv_tw = 1.0 / float(B * H * W)                   # nothing can be ignored: Σ v·w is the pixel count
            else:
                mode = "kernel"                                   # uint8 labels: the kernel sums the weights itself
        argmax = torch.empty((B, H, W), dtype=torch.uint8, device=dev) if (self.return_argmax and C <= 256) else None
        if mode == "kernel":
            tw = torch.empty(2, dtype=torch.float64, device=dev)
            local = None
            sc = self._scanned
            if sc is not None and sc[0] is target and sc[1] == target._version and sc[2] == C and t.dtype == torch.uint8:
                local = sc[3][0:1]                                # the previous launch already summed this batch's weights
            self._scanned = None
            nt, nbuf = None, None
            nl = self._next_labels
            if nl is not None and nl[2] == C and nl[0]._version == nl[1] and nl[0].device == dev:
                nt, nbuf = nl[0], torch.empty(2, dtype=torch.float64, device=dev)
                self._scanned = (nl[0], nl[1], C, nbuf)
            self._next_labels = None
            loss_out, sums, dlogits = ops.ce_fused(x, t, w, ii, want_grad=True, total_weight="kernel", xchg=xchg,
                                                   total_weight_out=tw, local_total_weight=local, next_target=nt,
                                                   next_total_weight_out=nbuf, argmax=argmax, confmat=conf)
        else:
            loss_out, sums, dlogits = ops.ce_fused(x, t, w, ii, want_grad=want_grad, inv_total_weight=inv_tw,
                                                   inv_total_weight_dev=tw[1:] if (want_grad and tw is not None) else None,
                                                   argmax=argmax, confmat=conf)
        self.last_argmax = argmax
        self.last_sums = sums
        self.last_total_weight = tw if want_grad else None
        if self.confusion is not None:
            self.confusion._note_bad_labels(sums)
        if self.strict:
            self.check_errors()
        loss = loss_out.reshape(())
        if logits.dtype != torch.float32:
            loss = loss.to(logits.dtype)  # torch returns the loss in the logits dtype
        return loss, dlogits, restore

    def check_errors(self) -> None:
        """Raise what torch raises on an out-of-bounds label (synchronises)."""
        if self.last_sums is not None and float(self.last_sums[2]) > 0:
            raise IndexError("Target is out of bounds.")

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        want_grad = torch.is_grad_enabled() and input.requires_grad
        return _FusedCE.apply(input, target, self, want_grad)


def class_weights_from_counts(counts: torch.Tensor, ignore_background: bool = False) -> torch.Tensor:
    """Reference weighting rule ``w_c = Σn / (bins · n_c)``, 0 for empty classes, background
    forced to 0 when ignored (dataset.py:360-384), evaluated exactly as the reference does:
    float32 counts, Python-float division of a float32 tensor numerator."""
    counts = counts.detach().to("cpu", torch.float32)
    if ignore_background:
        counts = counts[1:]
    numerator = torch.sum(counts)
    bins = len(counts)
    w = []
    for cc in counts:
        n = cc.item()
        w.append(0 if n == 0 else numerator / (bins * n))
    wt = torch.tensor(w)
    if ignore_background:
        return torch.concat((torch.tensor([0]), wt), dim=0)
    return wt


def load_loss(config: dict, device, dataset=None):
    """Same signature and config keys as the reference's ``utils.load_loss`` (utils.py:223-242):
    ``loss`` in {CEL, wCEL, MSE}, ``num_classes`` (+1 for background), ``ignore_background``."""
    classes = config["num_classes"] + 1
    name = config["loss"]
    ignore_background = config.get("ignore_background", False)
    ignore_index = 0 if ignore_background else -100
    if name == "CEL":
        return FusedCrossEntropyLoss(ignore_index=ignore_index).to(device)
    if name == "wCEL":
        print("Computing class weights, it might take several minutes...", flush=True)
        weights = dataset.get_class_weights(classes, ignore_background).to(device)
        for i, score in enumerate(weights):
            print(f"{GID15_LABELS.get(i, i)!s:>22}  {score.item():.6f}", flush=True)
        return FusedCrossEntropyLoss(weight=weights, ignore_index=ignore_index).to(device)
    if name == "MSE":
        return nn.MSELoss()  # not on the hot path (no shipped net regresses); kept for config parity
    raise Exception
