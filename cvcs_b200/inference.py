"""Drop-in for the reference's inference boundary.

Reference: ``dataset.GID15`` (dataset.py:35-105) serves one tile at a time — the crop, the mask crop, a
context crop and, with ``border_correction=bc``, a *padded patch* ``crop(image, tly-(bc-p), tlx-(bc-p), bc, bc)``
(dataset.py:18-23: the margin is ``bc - p`` on the top/left only, zero filled outside the scene);
``utils.inference`` (utils.py:145-171) runs the net on it, takes ``CenterCrop(p)`` of the output and the argmax,
writes one PNG per tile, and ``inference.py:40-57`` re-assembles the scene by concatenating the tile images at
their grid positions.

Here the padded patches of a whole batch are gathered by K5, the argmax is K2, and the centre crops are pasted
straight into a scene-sized index map on the GPU (``cvcs_stitch``, same CenterCrop offset rule as torchvision) —
no PNG round trip, no O(n²) ``torch.concat``.  ``GID15`` keeps the reference's indexing and item layout.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import ops
from .dataset import ArrayScenes, DirectoryScenes, tile_origin, tiles_in_image, tiles_per_image


def padded_origin(tly: int, tlx: int, p: int, border_correction: int) -> Tuple[int, int]:
    """Top-left corner of the padded patch (dataset.py:18-23): margin = bc - p, applied up/left only."""
    margin = border_correction - p
    return tly - margin, tlx - margin


class GID15(torch.utils.data.Dataset):
    """Same constructor, length and item layout as the reference's ``GID15`` (dataset.py:35-105):
    ``ds[idx] -> (tif_img u8[Cb,p,p], mask_img u8[..,p,p], context u8[Cb,p,p], padded_patch u8[Cb,bc,bc] | tensor([0]))``.
    Tiles are cut on the GPU; the scene of the last index stays resident, as the reference caches ``last_image``."""

    def __init__(self, root, patch_shape=(224, 224), color_masks=False, random_shift=False, border_correction=None, *,
                 device=None):
        self.source = DirectoryScenes(root) if isinstance(root, str) else root
        self.files = list(self.source.images)
        self.color_masks = color_masks
        self.border_correction = border_correction
        self.patch_shape = tuple(patch_shape)
        self.random_shift = random_shift
        self.image_shape = tuple(self.source.shape()) if not isinstance(root, str) else (6800, 7200)  # dataset.py:62
        self.tiles_in_img_shape = tiles_in_image(self.image_shape, self.patch_shape[0])
        self.tiles_per_img = tiles_per_image(self.image_shape, self.patch_shape[0])
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.last_image = self.last_target = None
        self.last_image_idx = -1

    def __len__(self):
        return len(self.files) * self.tiles_per_img

    def _load(self, image_idx: int) -> None:
        if image_idx != self.last_image_idx:
            img, idx, col = self.source.load(self.files[image_idx], self.color_masks)
            self.last_image = img.to(self.device).contiguous()
            tgt = col if self.color_masks else idx
            self.last_target = tgt.to(self.device).contiguous()
            self.last_image_idx = image_idx

    def __getitem__(self, idx):
        if self.random_shift:
            # the reference calls _random_shift(tly, tlx) without its third argument (dataset.py:86 vs :25)
            raise TypeError("_random_shift() missing 1 required positional argument: 'offset'")
        p = self.patch_shape[0]
        image_idx, tly, tlx = tile_origin(idx, self.tiles_per_img, self.tiles_in_img_shape[1], p)
        self._load(image_idx)
        yx = torch.tensor([[tly, tlx]], dtype=torch.int32, device=self.device)
        tif_img, _ = ops.tile_normalize(self.last_image, yx, (p, p), out_dtype=torch.uint8)
        mask_img, _ = ops.tile_normalize(self.last_target, yx, (p, p), out_dtype=torch.uint8)
        if self.border_correction:
            bc = self.border_correction
            pyx = torch.tensor([padded_origin(tly, tlx, p, bc)], dtype=torch.int32, device=self.device)
            padded, _ = ops.tile_normalize(self.last_image, pyx, (bc, bc), out_dtype=torch.uint8)
            padded_patch = padded[0]
        else:
            padded_patch = torch.tensor([0])
        # context: the 3p x 3p neighbourhood reduced to p (dataset.py:11-16, resizer dataset.py:65) — one kernel,
        # byte-identical to the reference's crop + Resize
        context = ops.tile_context(self.last_image, yx, p)[0]
        return tif_img[0], mask_img[0], context, padded_patch


def inference_scene(net, scene: torch.Tensor, patch_size: int, border_correction: Optional[int] = None,
                    batch_size: int = 16, indexes: Optional[Sequence[int]] = None,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """What ``utils.inference`` + the re-assembly loop of ``inference.py`` compute for one scene, as a u8 index
    map ``[rows*p, cols*p]`` on the GPU: per tile ``argmax(CenterCrop(p)(net(padded_patch)))`` (or
    ``argmax(net(tile))`` without border correction) pasted at the tile's grid position.  ``indexes`` restricts
    the tiles (the reference's ``range`` option); untouched tiles stay 0."""
    if scene.dtype != torch.uint8 or scene.dim() != 3 or not scene.is_cuda:
        raise RuntimeError("inference_scene expects a CUDA uint8 scene [Cb, H, W]")
    dev = scene.device
    p = patch_size
    rows, cols = tiles_in_image(scene.shape[1:], p)
    tpi = rows * cols
    ids = list(range(tpi)) if indexes is None else list(indexes)
    if out is None:
        out = torch.zeros((rows * p, cols * p), dtype=torch.uint8, device=dev)
    bc = border_correction if border_correction else p
    net.eval()
    with torch.no_grad():
        for k in range(0, len(ids), batch_size):
            part = ids[k:k + batch_size]
            origins = [tile_origin(i, tpi, cols, p)[1:] for i in part]
            src = [padded_origin(y, x, p, bc) if border_correction else (y, x) for y, x in origins]
            src_yx = torch.tensor(src, dtype=torch.int32, device=dev)
            patches, _ = ops.tile_normalize(scene.contiguous(), src_yx, (bc, bc), out_dtype=torch.float32)  # .type(float32)
            output = net(patches)
            if getattr(net, "returns_logits", True):
                pred = ops.argmax(output, torch.uint8)               # [b, bc, bc]
            else:
                pred = output.to(torch.uint8).reshape(len(part), bc, bc).contiguous()
            dst_yx = torch.tensor(origins, dtype=torch.int32, device=dev)
            ops.stitch(pred, dst_yx, tuple(out.shape), crop_hw=(p, p), out=out)     # CenterCrop(p) + paste
    return out
