"""Functional wrappers: torch CUDA tensors -> raw pointers -> the C-ABI kernels.

PyTorch is only plumbing here (device memory, streams); every computation below happens in
libcvcs_b200.so.  CPU tensors are rejected: there is no CPU path.
"""
from __future__ import annotations

import ctypes
import collections
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import lib, check

_DTYPE_TAG = {
    torch.float32: _lib.F32,
    torch.bfloat16: _lib.BF16,
    torch.uint8: _lib.U8,
    torch.int64: _lib.I64,
    torch.int32: _lib.I32,
}

_workspaces: dict = {}
_capture_workspaces: "collections.OrderedDict" = collections.OrderedDict()   # see workspace()


def _tag(t: torch.Tensor) -> int:
    try:
        return _DTYPE_TAG[t.dtype]
    except KeyError:
        raise RuntimeError(f"cvcs_b200: unsupported dtype {t.dtype}") from None


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*ts: Optional[torch.Tensor]) -> torch.device:
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("cvcs_b200 kernels need CUDA tensors (there is no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"cvcs_b200: tensors on different devices ({dev} vs {t.device})")
    assert dev is not None
    return dev


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def workspace(dev: torch.device) -> torch.Tensor:
    """Zero-initialised scratch buffer, one per (device, stream); kernels leave it zeroed.

    While the current stream is being captured into a CUDA graph the buffer comes from the graph's own memory pool (its
    zeroing becomes a memset node, so every replay starts from zeros) and is shared by all the calls of THAT capture on
    that stream — keyed by the capture's sequence number, because the capture stream is reused by later captures, whose
    kernels must not inherit a buffer that belongs to an earlier, possibly destroyed, graph.  One memset node per graph:
    consecutive K1 launches stay directly connected, so their programmatic overlap survives the capture.  The last 256
    such buffers (128 KB each) are kept alive."""
    if torch.cuda.is_current_stream_capturing():
        stream = _stream(dev)
        key = (dev.index if dev.index is not None else torch.cuda.current_device(), stream, _lib.stream_capture_id(stream))
        ws = _capture_workspaces.get(key)
        if ws is None:
            ws = torch.zeros(_lib.workspace_bytes(), dtype=torch.uint8, device=dev)
            _capture_workspaces[key] = ws
            while len(_capture_workspaces) > 256:
                _capture_workspaces.popitem(last=False)
        return ws
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), _stream(dev))
    ws = _workspaces.get(key)
    if ws is None:
        ws = torch.zeros(_lib.workspace_bytes(), dtype=torch.uint8, device=dev)
        _workspaces[key] = ws
    return ws


def logits_layout(logits: torch.Tensor) -> Tuple[torch.Tensor, int]:
    """[B,C,H,W] logits -> (tensor whose memory the kernel may read as-is, layout tag)."""
    if logits.is_contiguous():
        return logits, _lib.NCHW
    if logits.dim() == 4 and logits.is_contiguous(memory_format=torch.channels_last):
        return logits, _lib.NHWC
    return logits.contiguous(), _lib.NCHW


# ---- K4 ------------------------------------------------------------------------------------------
def label_hist(target: torch.Tensor, num_classes: int, ignore_index: int = -100,
               hist: Optional[torch.Tensor] = None, weight: Optional[torch.Tensor] = None,
               total_weight_out: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """Accumulate the label histogram (int64[C+2]: classes, ignored-outside, out-of-bounds)."""
    dev = _need_cuda(target, hist, weight, total_weight_out)
    target = target.contiguous()
    if target.dtype not in (torch.uint8, torch.int64):
        raise RuntimeError(f"expected scalar type Long or Byte but found {target.dtype}")
    if hist is None and total_weight_out is None:
        hist = torch.zeros(num_classes + 2, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(lib.cvcs_label_hist(target.data_ptr(), _tag(target), target.numel(), num_classes, ignore_index,
                                  _ptr(hist), _ptr(weight), _ptr(total_weight_out), workspace(dev).data_ptr(),
                                  _stream(dev)))
    return hist


def labels_prepare(target: torch.Tensor, num_classes: int, ignore_index: int = -100,
                   weight: Optional[torch.Tensor] = None, total_weight_out: Optional[torch.Tensor] = None,
                   labels_u8_out: Optional[torch.Tensor] = None):
    """int64 labels -> (f64[2] {Σ v·w[y], 1/Σ}, u8 labels with 255 = ignored / 254 = out of range) in one pass."""
    dev = _need_cuda(target, weight, total_weight_out, labels_u8_out)
    if target.dtype != torch.int64:
        raise RuntimeError(f"labels_prepare expects int64 labels, got {target.dtype}")
    target = target.contiguous()
    if total_weight_out is None:
        total_weight_out = torch.empty(2, dtype=torch.float64, device=dev)
    if labels_u8_out is None:
        labels_u8_out = torch.empty(target.shape, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.cvcs_labels_prepare(target.data_ptr(), target.numel(), num_classes, ignore_index, _ptr(weight),
                                      total_weight_out.data_ptr(), labels_u8_out.data_ptr(), workspace(dev).data_ptr(),
                                      _stream(dev)))
    return total_weight_out, labels_u8_out


def total_weight(hist: torch.Tensor, weight: Optional[torch.Tensor], num_classes: int, ignore_index: int,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    dev = _need_cuda(hist, weight, out)
    if out is None:
        out = torch.empty(2, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(lib.cvcs_total_weight(hist.data_ptr(), _ptr(weight), num_classes, ignore_index, out.data_ptr(),
                                    _stream(dev)))
    return out


# ---- K1 ------------------------------------------------------------------------------------------
class Exchange:
    """cvcs_xchg: this rank's Σw exchange block plus the peers' blocks (CUDA IPC mappings).  Built by
    ``cvcs_b200.shard.WeightExchange`` for one-process-per-GPU jobs; tests wire several handles of one process together
    with ``set_peer`` / play a peer with ``poke``."""

    def __init__(self, world: int, rank: int, device: Optional[torch.device] = None):
        self._h = ctypes.c_void_p()
        self.world, self.rank = int(world), int(rank)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        with torch.cuda.device(self.device):
            check(lib.cvcs_xchg_create(ctypes.byref(self._h), self.world, self.rank))

    @property
    def handle(self) -> ctypes.c_void_p:
        if not self._h:
            raise RuntimeError("cvcs_b200.Exchange: already closed")
        return self._h

    def local_handle(self) -> bytes:
        buf = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(self.device):
            check(lib.cvcs_xchg_local_handle(self.handle, buf))
        return bytes(buf)

    def open_peer(self, peer_rank: int, handle: bytes) -> None:
        buf = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
        with torch.cuda.device(self.device):
            check(lib.cvcs_xchg_open_peer(self.handle, int(peer_rank), buf))

    def set_peer(self, peer_rank: int, other: "Exchange") -> None:
        check(lib.cvcs_xchg_set_peer(self.handle, int(peer_rank), lib.cvcs_xchg_local_block(other.handle)))

    def state(self) -> Tuple[int, int]:
        """(exchanges completed by this rank, time-outs / overruns its kernels saw) — synchronises the device."""
        seq, err = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
        with torch.cuda.device(self.device):
            check(lib.cvcs_xchg_state(self.handle, ctypes.byref(seq), ctypes.byref(err)))
        return int(seq.value), int(err.value)

    def poke(self, as_rank: int, seq: int, value: float) -> None:
        with torch.cuda.device(self.device):
            check(lib.cvcs_xchg_poke(self.handle, int(as_rank), int(seq), float(value), _stream(self.device)))

    def allreduce_(self, buf: torch.Tensor) -> torch.Tensor:
        """In-place sum over the ranks of a short float64 device vector (<= 2048 elements), added in rank order — the
        pass-end sums (confusion matrix, loss table) without an NCCL launch.  Every rank must call it equally often."""
        if not buf.is_cuda or buf.dtype != torch.float64 or not buf.is_contiguous() or buf.numel() > 2048:
            raise RuntimeError("Exchange.allreduce_: a contiguous float64 CUDA tensor of at most 2048 elements")
        with torch.cuda.device(self.device):
            check(lib.cvcs_xchg_allreduce_f64(self.handle, buf.data_ptr(), buf.numel(), _stream(self.device)))
        return buf

    def close(self) -> None:
        if self._h:
            lib.cvcs_xchg_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def _check_k1_buffers(logits, Cc, weight, dlogits, argmax, confmat, loss_sums, loss_out, B, H, W):
    """Cheap host-side checks of everything that crosses the C-ABI as a raw pointer (a wrong dtype or size would be a
    silent out-of-bounds access on the GPU)."""
    if weight is not None and (weight.dtype != torch.float32 or weight.numel() != Cc or not weight.is_contiguous()):
        raise RuntimeError(f"weight tensor should be a contiguous float32 tensor of size {Cc}, got {weight.dtype} {tuple(weight.shape)}")
    if confmat is not None and (confmat.dtype != torch.int64 or confmat.numel() != Cc * Cc or not confmat.is_contiguous()):
        raise RuntimeError(f"confmat must be a contiguous int64 [{Cc},{Cc}] tensor, got {confmat.dtype} {tuple(confmat.shape)}")
    if argmax is not None:
        if argmax.dtype not in (torch.uint8, torch.int64) or argmax.numel() != B * H * W or not argmax.is_contiguous():
            raise RuntimeError(f"argmax must be a contiguous uint8 / int64 tensor of {B * H * W} elements, got {argmax.dtype} {tuple(argmax.shape)}")
    if dlogits is not None:
        if dlogits.dtype != logits.dtype or dlogits.shape != logits.shape or dlogits.stride() != logits.stride():
            raise RuntimeError("dlogits must have the dtype, shape and memory format of logits")
    if loss_sums is not None and (loss_sums.dtype != torch.float64 or loss_sums.numel() != 3 or not loss_sums.is_contiguous()):
        raise RuntimeError("loss_sums must be a contiguous float64 tensor of 3 elements")
    if loss_out is not None and (loss_out.dtype != torch.float32 or loss_out.numel() != 1):
        raise RuntimeError("loss_out must be a float32 tensor of 1 element")


def ce_fused(logits: torch.Tensor, target: torch.Tensor, weight: Optional[torch.Tensor] = None,
             ignore_index: int = -100, *, want_grad: bool = True, inv_total_weight: float = 0.0,
             inv_total_weight_dev: Optional[torch.Tensor] = None, dlogits: Optional[torch.Tensor] = None,
             argmax: Optional[torch.Tensor] = None, confmat: Optional[torch.Tensor] = None,
             loss_sums: Optional[torch.Tensor] = None, loss_out: Optional[torch.Tensor] = None,
             total_weight: str = "given", xchg: Optional[Exchange] = None,
             total_weight_out: Optional[torch.Tensor] = None, local_total_weight: Optional[torch.Tensor] = None,
             next_target: Optional[torch.Tensor] = None, next_total_weight_out: Optional[torch.Tensor] = None):
    """One fused pass. logits [B,C,H,W] f32/bf16 (contiguous or channels_last), target [B,H,W]
    u8/i64.  Returns (loss_out f32[1], loss_sums f64[3], dlogits or None).

    total_weight="given": the 'mean' divisor comes from inv_total_weight / inv_total_weight_dev (cvcs_ce_fused).
    total_weight="kernel": the kernel computes it from the labels itself — and, with ``xchg``, over all ranks' labels —
    before it writes the first gradient (cvcs_ce_fused_tw); ``total_weight_out`` f64[2] receives {Σ, 1/Σ}.  With
    ``local_total_weight`` (f64[1] on the device, e.g. ``label_hist``'s total_weight_out[0:1] from a launch one step
    ahead) the kernel skips its own label pre-pass and only exchanges.  ``next_target`` (uint8 labels of the NEXT batch)
    makes this launch also sum the weights over them into ``next_total_weight_out`` f64[2] — pass that as
    ``local_total_weight`` of the next call and the pre-pass never sits on the critical path."""
    dev = _need_cuda(logits, target, weight, inv_total_weight_dev, dlogits, argmax, confmat, loss_sums, loss_out, total_weight_out,
                     local_total_weight)
    if local_total_weight is not None and (local_total_weight.dtype != torch.float64 or local_total_weight.numel() < 1):
        raise RuntimeError("local_total_weight must be a float64 device tensor")
    if next_target is not None:
        _need_cuda(next_target, next_total_weight_out)
        if next_target.dtype != torch.uint8 or not next_target.is_contiguous():
            raise RuntimeError("next_target must be a contiguous uint8 label tensor")
        if next_total_weight_out is None or next_total_weight_out.dtype != torch.float64 or next_total_weight_out.numel() != 2:
            raise RuntimeError("next_target needs next_total_weight_out: a float64 tensor of 2 elements")
        if total_weight != "kernel":
            raise RuntimeError("next_target is served by total_weight='kernel' (cvcs_ce_fused_tw)")
    if logits.dim() != 4:
        raise RuntimeError(f"cvcs_b200.ce_fused expects [B,C,H,W] logits, got {tuple(logits.shape)}")
    B, Cc, H, W = logits.shape
    if tuple(target.shape) != (B, H, W):
        raise RuntimeError(f"size mismatch (got input: {list(logits.shape)} , target: {list(target.shape)}")
    logits, layout = logits_layout(logits)
    target = target.contiguous()
    if want_grad and dlogits is None:
        dlogits = torch.empty_like(logits)  # preserves the memory format
    if loss_sums is None:
        loss_sums = torch.empty(3, dtype=torch.float64, device=dev)
    if loss_out is None:
        loss_out = torch.empty(1, dtype=torch.float32, device=dev)
    _check_k1_buffers(logits, Cc, weight, dlogits if want_grad else None, argmax, confmat, loss_sums, loss_out, B, H, W)
    if inv_total_weight_dev is not None and inv_total_weight_dev.dtype != torch.float64:
        raise RuntimeError("inv_total_weight_dev must be float64")
    with torch.cuda.device(dev):
        if total_weight == "kernel":
            if total_weight_out is None:
                total_weight_out = torch.empty(2, dtype=torch.float64, device=dev)
            elif total_weight_out.dtype != torch.float64 or total_weight_out.numel() != 2:
                raise RuntimeError("total_weight_out must be a float64 tensor of 2 elements")
            check(lib.cvcs_ce_fused_tw(logits.data_ptr(), _tag(logits), layout, target.data_ptr(), _tag(target),
                                       _ptr(weight), ignore_index, B, Cc, H, W, xchg.handle if xchg is not None else None,
                                       _ptr(local_total_weight), total_weight_out.data_ptr(), _ptr(next_target),
                                       next_target.numel() if next_target is not None else 0, _ptr(next_total_weight_out), _ptr(dlogits) if want_grad else None, _ptr(argmax),
                                       _tag(argmax) if argmax is not None else _lib.U8, _ptr(confmat),
                                       loss_sums.data_ptr(), loss_out.data_ptr(), workspace(dev).data_ptr(), _stream(dev)))
        elif total_weight == "given":
            check(lib.cvcs_ce_fused(logits.data_ptr(), _tag(logits), layout, target.data_ptr(), _tag(target),
                                    _ptr(weight), ignore_index, B, Cc, H, W, float(inv_total_weight),
                                    _ptr(inv_total_weight_dev), _ptr(dlogits) if want_grad else None, _ptr(argmax),
                                    _tag(argmax) if argmax is not None else _lib.U8, _ptr(confmat),
                                    loss_sums.data_ptr(), loss_out.data_ptr(), workspace(dev).data_ptr(), _stream(dev)))
        else:
            raise ValueError(f"total_weight must be 'given' or 'kernel', got {total_weight!r}")
    return loss_out, loss_sums, (dlogits if want_grad else None)


def eval_fused(logits: torch.Tensor, target: torch.Tensor, ignore_index: Optional[int] = None, *,
               argmax: Optional[torch.Tensor] = None, confmat: Optional[torch.Tensor] = None,
               status: Optional[torch.Tensor] = None) -> None:
    """K1 in metrics mode: argmax map and / or confusion-matrix update from [B,C,H,W] logits in one read
    (no softmax, no loss) — what ``utils.eval_model`` does per tile (utils.py:88-94)."""
    dev = _need_cuda(logits, target, argmax, confmat, status)
    if logits.dim() != 4:
        raise RuntimeError(f"cvcs_b200.eval_fused expects [B,C,H,W] logits, got {tuple(logits.shape)}")
    B, Cc, H, W = logits.shape
    if tuple(target.shape) != (B, H, W):
        raise RuntimeError(f"size mismatch (got input: {list(logits.shape)} , target: {list(target.shape)}")
    logits, layout = logits_layout(logits)
    target = target.contiguous()
    ign = -(1 << 62) if ignore_index is None else int(ignore_index)
    with torch.cuda.device(dev):
        check(lib.cvcs_eval_fused(logits.data_ptr(), _tag(logits), layout, target.data_ptr(), _tag(target), ign, B, Cc, H,
                                  W, _ptr(argmax), _tag(argmax) if argmax is not None else _lib.U8, _ptr(confmat),
                                  _ptr(status), workspace(dev).data_ptr(), _stream(dev)))


def scale_inplace(x: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    dev = _need_cuda(x, scale)
    assert scale.dtype == torch.float32 and scale.numel() == 1
    with torch.cuda.device(dev):
        check(lib.cvcs_scale_inplace(x.data_ptr(), _tag(x), x.numel(), scale.data_ptr(), _stream(dev)))
    return x


# ---- K2 / K3 ---------------------------------------------------------------------------------------
def argmax(logits: torch.Tensor, out_dtype: torch.dtype = torch.int64, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """First-maximal class index over dim 1 of [B,C,H,W] logits (torch.max / argmax semantics)."""
    dev = _need_cuda(logits, out)
    if logits.dim() != 4:
        raise RuntimeError(f"cvcs_b200.argmax expects [B,C,H,W], got {tuple(logits.shape)}")
    logits, layout = logits_layout(logits)
    B, Cc, H, W = logits.shape
    if out is None:
        out = torch.empty((B, H, W), dtype=out_dtype, device=dev)
    elif out.numel() != B * H * W or not out.is_contiguous() or out.dtype not in (torch.uint8, torch.int64):
        raise RuntimeError("argmax: `out` must be a contiguous uint8 / int64 tensor of B*H*W elements")
    with torch.cuda.device(dev):
        check(lib.cvcs_argmax(logits.data_ptr(), _tag(logits), layout, B, Cc, H, W, out.data_ptr(), _tag(out),
                              _stream(dev)))
    return out


def confmat_update(confmat: torch.Tensor, preds: torch.Tensor, target: torch.Tensor, num_classes: int,
                   ignore_index: Optional[int], status: Optional[torch.Tensor] = None) -> torch.Tensor:
    """confmat[t, p] += 1 over index maps (any shape, same numel); int64[C,C] accumulated."""
    dev = _need_cuda(confmat, preds, target, status)
    preds, target = preds.contiguous(), target.contiguous()
    if preds.numel() != target.numel():
        raise RuntimeError("preds and target must have the same number of elements")
    ign = -(1 << 62) if ignore_index is None else int(ignore_index)
    with torch.cuda.device(dev):
        check(lib.cvcs_confmat(preds.data_ptr(), _tag(preds), target.data_ptr(), _tag(target), target.numel(),
                               num_classes, ign, confmat.data_ptr(), _ptr(status), workspace(dev).data_ptr(),
                               _stream(dev)))
    return confmat


# ---- K5 ------------------------------------------------------------------------------------------
def tile_normalize(scene: torch.Tensor, tile_yx: torch.Tensor, tile_hw: Tuple[int, int],
                   mean: Optional[torch.Tensor] = None, std: Optional[torch.Tensor] = None,
                   out_dtype: torch.dtype = torch.float32, label: Optional[torch.Tensor] = None,
                   label_out_dtype: torch.dtype = torch.uint8, hist: Optional[torch.Tensor] = None,
                   hist_classes: int = 0, hist_ignore_index: int = -100, *,
                   slots: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                   label_out: Optional[torch.Tensor] = None):
    """scene u8 [Cb,H,W] -> tiles [n,Cb,th,tw] (cast / normalised), label u8 [H,W] -> [n,th,tw].

    ``slots`` (int32 [n]) scatters tile i to ``out[slots[i]]`` of a caller-provided batch (``out`` /
    ``label_out``), so the tiles of several scenes can be laid down in one batch in any order."""
    dev = _need_cuda(scene, tile_yx, mean, std, label, hist, slots, out, label_out)
    assert scene.dtype == torch.uint8 and scene.dim() == 3 and scene.is_contiguous()
    assert tile_yx.dtype == torch.int32 and tile_yx.dim() == 2 and tile_yx.shape[1] == 2 and tile_yx.is_contiguous()
    Cb, H, W = scene.shape
    th, tw = tile_hw
    n = tile_yx.shape[0]
    if slots is not None:
        assert slots.dtype == torch.int32 and slots.numel() == n and slots.is_contiguous()
        assert out is not None, "slots need a caller-provided output batch"
    if out is None:
        out = torch.empty((n, Cb, th, tw), dtype=out_dtype, device=dev)
    else:
        assert out.is_contiguous() and tuple(out.shape[1:]) == (Cb, th, tw) and (slots is not None or out.shape[0] >= n)
    if label is not None:
        assert label.dtype == torch.uint8 and tuple(label.shape[-2:]) == (H, W) and label.is_contiguous()
        if label_out is None:
            assert slots is None, "slots need a caller-provided label batch"
            label_out = torch.empty((n, th, tw), dtype=label_out_dtype, device=dev)
        else:
            assert label_out.is_contiguous() and tuple(label_out.shape[1:]) == (th, tw)
    else:
        label_out = None
    with torch.cuda.device(dev):
        check(lib.cvcs_tile_normalize(scene.data_ptr(), Cb, H, W, tile_yx.data_ptr(), _ptr(slots), n, th, tw,
                                      _ptr(mean), _ptr(std), out.data_ptr(), _tag(out), _ptr(label), _ptr(label_out),
                                      _tag(label_out) if label_out is not None else _lib.U8, _ptr(hist),
                                      hist_classes, hist_ignore_index, workspace(dev).data_ptr(), _stream(dev)))
    return out, label_out


def tile_context(scene: torch.Tensor, tile_yx: torch.Tensor, p: int, *, slots: Optional[torch.Tensor] = None,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``_get_context`` (dataset.py:11-16) for a batch of patch origins: scene u8 [Cb,H,W], tile_yx i32 [n,2] ->
    u8 [n,Cb,p,p], the 3p x 3p neighbourhoods (zeros outside the scene) reduced to p x p with the reference's
    antialiased bilinear uint8 resize, bit-identical."""
    dev = _need_cuda(scene, tile_yx, slots, out)
    assert scene.dtype == torch.uint8 and scene.dim() == 3 and scene.is_contiguous()
    assert tile_yx.dtype == torch.int32 and tile_yx.dim() == 2 and tile_yx.shape[1] == 2 and tile_yx.is_contiguous()
    Cb, H, W = scene.shape
    n = tile_yx.shape[0]
    if slots is not None:
        assert slots.dtype == torch.int32 and slots.numel() == n and slots.is_contiguous() and out is not None
    if out is None:
        out = torch.empty((n, Cb, p, p), dtype=torch.uint8, device=dev)
    else:
        assert out.dtype == torch.uint8 and out.is_contiguous() and tuple(out.shape[1:]) == (Cb, p, p)
    step = max(1, 65535 // Cb)
    with torch.cuda.device(dev):
        for k in range(0, n, step):
            m = min(step, n - k)
            check(lib.cvcs_tile_context(scene.data_ptr(), Cb, H, W, tile_yx[k:k + m].data_ptr(),
                                        None if slots is None else slots[k:k + m].data_ptr(), m, p,
                                        out.data_ptr() if slots is not None else out[k:k + m].data_ptr(), _stream(dev)))
    return out


# ---- N2 / N3 / N4 ----------------------------------------------------------------------------------
def vote(maps: torch.Tensor, num_classes: int = 0, out_dtype: Optional[torch.dtype] = None,
         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-pixel majority vote over dim 0 (ties -> smallest index, as torch.mode)."""
    dev = _need_cuda(maps, out)
    maps = maps.contiguous()
    if out is None:
        out = torch.empty(maps.shape[1:], dtype=out_dtype or maps.dtype, device=dev)
    elif out.numel() != maps[0].numel() or not out.is_contiguous() or out.dtype not in (torch.uint8, torch.int64):
        raise RuntimeError("vote: `out` must be a contiguous uint8 / int64 tensor with one element per pixel")
    with torch.cuda.device(dev):
        check(lib.cvcs_vote(maps.data_ptr(), _tag(maps), maps.shape[0], out.numel(), num_classes, out.data_ptr(),
                            _tag(out), _stream(dev)))
    return out


def colorize(index_map: torch.Tensor, lut: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[H,W] class indices -> [H,W,3] f32 colours (GID15Converter.iconvert)."""
    dev = _need_cuda(index_map, lut, out)
    index_map = index_map.contiguous()
    lut = lut.contiguous().to(torch.float32)
    if out is None:
        out = torch.empty((*index_map.shape, 3), dtype=torch.float32, device=dev)
    elif out.dtype != torch.float32 or out.numel() != index_map.numel() * 3 or not out.is_contiguous():
        raise RuntimeError("colorize: `out` must be a contiguous float32 tensor of 3 values per pixel")
    with torch.cuda.device(dev):
        check(lib.cvcs_colorize(index_map.data_ptr(), _tag(index_map), index_map.numel(), lut.data_ptr(),
                                lut.shape[0], out.data_ptr(), _stream(dev)))
    return out


def stitch(tiles: torch.Tensor, tile_yx: torch.Tensor, scene_hw: Tuple[int, int],
           crop_hw: Optional[Tuple[int, int]] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Paste u8 tiles [n,th,tw] (optionally their centred crop) into a scene-sized u8 map."""
    dev = _need_cuda(tiles, tile_yx, out)
    assert tiles.dtype == torch.uint8 and tiles.dim() == 3
    tiles = tiles.contiguous()
    n, th, tw = tiles.shape
    ch, cw = crop_hw if crop_hw is not None else (th, tw)
    H, W = scene_hw
    if out is None:
        out = torch.zeros((H, W), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.cvcs_stitch(tiles.data_ptr(), n, th, tw, tile_yx.data_ptr(), ch, cw, out.data_ptr(), H, W,
                              _stream(dev)))
    return out


# ---- host-buffer context -------------------------------------------------------------------------
class HostContext:
    """cvcs_host_* : the C-ABI call a non-torch caller makes, on HOST buffers (copies inside)."""

    def __init__(self, device: int, max_pixels: int, max_classes: int, logits_dtype: torch.dtype = torch.float32):
        self._h = ctypes.c_void_p()
        check(lib.cvcs_host_ctx_create(ctypes.byref(self._h), device, max_pixels, max_classes,
                                       _DTYPE_TAG[logits_dtype]))
        self.device = device

    def close(self) -> None:
        if self._h:
            lib.cvcs_host_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def ce_fused(self, logits: torch.Tensor, target: torch.Tensor, weight: Optional[torch.Tensor], ignore_index: int,
                 want_grad: bool = True, dlogits: Optional[torch.Tensor] = None,
                 argmax: Optional[torch.Tensor] = None, confmat: Optional[torch.Tensor] = None):
        """All tensors are CPU tensors (ideally pinned).  Returns (loss float32 tensor[1], sums f64[3])."""
        for t in (logits, target, weight, dlogits, argmax, confmat):
            if t is not None and t.is_cuda:
                raise RuntimeError("HostContext.ce_fused takes host tensors")
        B, Cc, H, W = logits.shape
        logits, layout = logits_layout(logits)
        loss = torch.empty(1, dtype=torch.float32)
        sums = torch.empty(3, dtype=torch.float64)
        check(lib.cvcs_host_ce_fused(self._h, logits.data_ptr(), _tag(logits), layout, target.data_ptr(), _tag(target),
                                     _ptr(weight), ignore_index, B, Cc, H, W, 1 if want_grad else 0, _ptr(dlogits),
                                     _ptr(argmax), _tag(argmax) if argmax is not None else _lib.U8, _ptr(confmat),
                                     loss.data_ptr(), sums.data_ptr()))
        return loss, sums
