"""numpy-facing wrappers around cvcs_oracle.c (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import clib

_i64, _vp = C.c_int64, C.c_void_p


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_vp)


def _strides(layout: str, Cc: int, HW: int) -> Tuple[int, int, int]:
    """(image_stride, class_stride, pixel_stride) in elements."""
    if layout == "NCHW":
        return Cc * HW, HW, 1
    if layout == "NHWC":
        return Cc * HW, 1, Cc
    raise ValueError(layout)


def cross_entropy(logits: np.ndarray, target: np.ndarray, weight: Optional[np.ndarray] = None,
                  ignore_index: int = -100, layout: str = "NCHW", want_grad: bool = True):
    """logits: float32 array holding B images of C x HW (NCHW) or HW x C (NHWC) values;
    target int64 [B, HW...].  Returns (loss float64, sums float64[3], dlogits float32 | None)."""
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    target = np.ascontiguousarray(target, dtype=np.int64)
    B = logits.shape[0]
    Cc = logits.shape[1] if layout == "NCHW" else logits.shape[-1]
    HW = logits[0].size // Cc
    assert target.size == B * HW
    w = None if weight is None else np.ascontiguousarray(weight, dtype=np.float32)
    sums = np.zeros(3, dtype=np.float64)
    d = np.empty_like(logits) if want_grad else None
    ist, cst, pst = _strides(layout, Cc, HW)
    f = clib().oracle_cross_entropy
    f.restype = C.c_int
    f.argtypes = [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp]
    f(_p(logits), _p(target), _p(w), ignore_index, B, Cc, HW, ist, cst, pst, _p(sums), _p(d))
    with np.errstate(invalid="ignore", divide="ignore"):
        loss = sums[0] / sums[1]
    return loss, sums, d


def argmax(logits: np.ndarray, layout: str = "NCHW") -> np.ndarray:
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    B = logits.shape[0]
    Cc = logits.shape[1] if layout == "NCHW" else logits.shape[-1]
    HW = logits[0].size // Cc
    out = np.empty(B * HW, dtype=np.int64)
    ist, cst, pst = _strides(layout, Cc, HW)
    f = clib().oracle_argmax
    f.restype = None
    f.argtypes = [_vp, _i64, _i64, _i64, _i64, _i64, _i64, _vp]
    f(_p(logits), B, Cc, HW, ist, cst, pst, _p(out))
    spatial = logits.shape[2:] if layout == "NCHW" else logits.shape[1:-1]
    return out.reshape((B, *spatial))


def confmat(preds: np.ndarray, target: np.ndarray, num_classes: int, ignore_index: Optional[int] = None,
            into: Optional[np.ndarray] = None) -> Tuple[np.ndarray, int]:
    preds = np.ascontiguousarray(preds, dtype=np.int64).reshape(-1)
    target = np.ascontiguousarray(target, dtype=np.int64).reshape(-1)
    cm = into if into is not None else np.zeros((num_classes, num_classes), dtype=np.int64)
    f = clib().oracle_confmat
    f.restype = _i64
    f.argtypes = [_vp, _vp, _i64, _i64, C.c_int, _i64, _vp]
    bad = f(_p(preds), _p(target), preds.size, num_classes, 0 if ignore_index is None else 1,
            0 if ignore_index is None else ignore_index, _p(cm))
    return cm, int(bad)


def label_hist(labels: np.ndarray, num_classes: int, ignore_index: int = -100) -> np.ndarray:
    labels = np.ascontiguousarray(labels, dtype=np.uint8).reshape(-1)
    hist = np.zeros(num_classes + 2, dtype=np.int64)
    f = clib().oracle_label_hist
    f.restype = None
    f.argtypes = [_vp, _i64, _i64, _i64, _vp]
    f(_p(labels), labels.size, num_classes, ignore_index, _p(hist))
    return hist


def tile(scene: np.ndarray, yx: np.ndarray, th: int, tw: int, mean: Optional[np.ndarray] = None,
         std: Optional[np.ndarray] = None, labels: Optional[np.ndarray] = None):
    scene = np.ascontiguousarray(scene, dtype=np.uint8)
    yx = np.ascontiguousarray(yx, dtype=np.int32)
    Cb, H, W = scene.shape
    n = yx.shape[0]
    out = np.empty((n, Cb, th, tw), dtype=np.float32)
    m = None if mean is None else np.ascontiguousarray(mean, dtype=np.float32)
    s = None if std is None else np.ascontiguousarray(std, dtype=np.float32)
    lab = None if labels is None else np.ascontiguousarray(labels, dtype=np.uint8).reshape(H, W)
    lab_out = None if labels is None else np.empty((n, th, tw), dtype=np.uint8)
    f = clib().oracle_tile
    f.restype = None
    f.argtypes = [_vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]
    f(_p(scene), Cb, H, W, _p(yx), n, th, tw, _p(m), _p(s), _p(out), _p(lab), _p(lab_out))
    return out, lab_out


def vote(maps: np.ndarray) -> np.ndarray:
    maps = np.ascontiguousarray(maps, dtype=np.int64)
    n_maps = maps.shape[0]
    out = np.empty(maps[0].size, dtype=np.int64)
    f = clib().oracle_vote
    f.restype = None
    f.argtypes = [_vp, _i64, _i64, _vp]
    f(_p(maps), n_maps, out.size, _p(out))
    return out.reshape(maps.shape[1:])


def colorize(idx: np.ndarray, lut: np.ndarray) -> np.ndarray:
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    lut = np.ascontiguousarray(lut, dtype=np.float32)
    out = np.empty((*idx.shape, 3), dtype=np.float32)
    f = clib().oracle_colorize
    f.restype = None
    f.argtypes = [_vp, _i64, _vp, _i64, _vp]
    f(_p(idx), idx.size, _p(lut), lut.shape[0], _p(out))
    return out


def stitch(tiles: np.ndarray, yx: np.ndarray, H: int, W: int, crop: Optional[Tuple[int, int]] = None) -> np.ndarray:
    tiles = np.ascontiguousarray(tiles, dtype=np.uint8)
    yx = np.ascontiguousarray(yx, dtype=np.int32)
    n, th, tw = tiles.shape
    ch, cw = crop if crop is not None else (th, tw)
    scene = np.zeros((H, W), dtype=np.uint8)
    f = clib().oracle_stitch
    f.restype = None
    f.argtypes = [_vp, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64]
    f(_p(tiles), n, th, tw, _p(yx), ch, cw, _p(scene), H, W)
    return scene


def context(scene: np.ndarray, yx: np.ndarray, p: int, return_pre: bool = False):
    """dataset.py:11-16 _get_context for a batch of patch origins: u8 [n,Cb,p,p] (and, on request, the float32 values
    before rounding — the reference's own byte is only defined up to its FMA behaviour where they sit on a .5 tie)."""
    scene = np.ascontiguousarray(scene, dtype=np.uint8)
    yx = np.ascontiguousarray(yx, dtype=np.int32)
    Cb, H, W = scene.shape
    n = yx.shape[0]
    out = np.empty((n, Cb, p, p), dtype=np.uint8)
    pre = np.empty((n, Cb, p, p), dtype=np.float32) if return_pre else None
    f = clib().oracle_context
    f.restype = C.c_int
    f.argtypes = [_vp, _i64, _i64, _i64, _vp, _i64, _i64, _vp, _vp]
    rc = f(_p(scene), Cb, H, W, _p(yx), n, p, _p(out), _p(pre))
    assert rc == 0, rc
    return (out, pre) if return_pre else out
