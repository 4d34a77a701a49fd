"""Drop-in for the reference's dataset boundary: ``Loader`` / ``IterableChunk``.

Reference: ``dataset.Loader`` (dataset.py:241-345) pre-loads ``chunk_size`` full scenes, and
``IterableChunk`` (dataset.py:108-240) cuts every ``patch_size`` tile out of them on the CPU with
``torchvision.transforms.v2.functional.crop`` (``_get_cropped_data``, dataset.py:28-32), one tile at
a time, into a Python list.  The training loop then casts each batch with
``image.type(torch.float32)`` (train.py:121).

Here the scenes are uploaded to the GPU once per chunk and ONE kernel launch per scene (K5,
``cvcs_tile_normalize``) gathers all of its tiles — image, index mask and colour mask — straight into
batch tensors laid out in the chunk's shuffled order.  The tile index arithmetic, the shuffle, the
random-shift draws and the chunk bookkeeping are the reference's, call for call, so that under the
same ``random.seed`` both produce the same tiles in the same order:

    tiles_in_img = (H // p, W // p)                  dataset.py:125      (remainder pixels dropped)
    x -> image x // tpi, tile x % tpi                dataset.py:137-138
    tile -> (row, col) = (t // cols, t % cols)       dataset.py:139
    (tly, tlx) = (row * p, col * p)                  dataset.py:140
    random shift: +randint(-20, 20) on each axis     dataset.py:25-26,143
    out-of-bounds crops are zero filled              torchvision crop

Extras over the reference (all optional): ``device``; scenes may come from memory
(``ArrayScenes``) instead of the GID-15 directory tree; ``IterableChunk.batch`` exposes the
contiguous batch tensors; ``IterableChunk.float_tiles`` fuses the float cast / per-band
normalisation (train.py:121, nets.py:339-342) into the gather; ``Loader(...,
strict_patch_size=False)`` lifts the reference's {224, 256, 512} assertion for 1024-pixel tiles.
"""
from __future__ import annotations

import os
import random
from pathlib import Path
from typing import List, Optional, Sequence, Tuple

import torch

from . import ops
from .loss import class_weights_from_counts

SHIFT_OFFSET = 20  # dataset.py:143


# ---- tile index arithmetic (pure host logic, shared with the sharded tiler) ---------------------
def tiles_in_image(image_shape: Sequence[int], p: int) -> Tuple[int, int]:
    """(rows, cols) of whole tiles; remainder pixels are dropped (dataset.py:63,125)."""
    return image_shape[0] // p, image_shape[1] // p


def tiles_per_image(image_shape: Sequence[int], p: int) -> int:
    """``Loader.__get_tpi`` (dataset.py:280-287)."""
    h, w = image_shape
    return (h // p) * (w // p)


def tile_origin(x: int, tpi: int, cols: int, p: int) -> Tuple[int, int, int]:
    """crop id -> (scene index inside the chunk, tly, tlx) (dataset.py:136-140; GID15.__getitem__ :82-85)."""
    target_image = x // tpi
    tile_idx = x % tpi
    row, col = tile_idx // cols, tile_idx % cols
    return target_image, row * p, col * p


def _random_shift(tly: int, tlx: int, offset: int) -> Tuple[int, int]:
    # same two draws, in the same order, as dataset.py:25-26
    return tly + random.randint(-offset, offset), tlx + random.randint(-offset, offset)


# ---- scene sources ---------------------------------------------------------------------------------
class DirectoryScenes:
    """The GID-15 directory tree the reference reads (dataset.py:261-266)."""

    def __init__(self, root: str):
        self.root = root
        self.imdir = os.path.join(root, "Image__8bit_NirRGB")
        self.indexdir = os.path.join(root, "Annotation__index")
        self.maskdir = os.path.join(root, "Annotation__color")
        self.images = sorted(os.path.join(self.imdir, f) for f in os.listdir(self.imdir))
        self.index_masks = sorted(os.path.join(self.indexdir, f) for f in os.listdir(self.indexdir))

    @staticmethod
    def _open(path: str) -> torch.Tensor:
        import numpy as np
        from PIL import Image
        a = np.array(Image.open(path))          # a writable copy (torch tensors must own writable memory)
        t = torch.from_numpy(np.ascontiguousarray(a))
        if t.dim() == 2:
            return t.unsqueeze(0)          # tv_tensors.Image / Mask of a single-band file: [1, H, W]
        return t.permute(2, 0, 1).contiguous()

    def shape(self) -> List[int]:
        return list(self._open(self.images[0]).shape)[1:]

    def load(self, name: str, want_color: bool):
        img = self._open(name)
        idx = self._open(os.path.join(self.indexdir, Path(name).stem + "_15label.png"))
        col = self._open(os.path.join(self.maskdir, Path(name).stem + "_15label.tif")) if want_color else None
        return img, idx, col

    def load_index_mask(self, path: str) -> torch.Tensor:
        return self._open(path)


class ArrayScenes:
    """In-memory scenes (synthetic benchmarks, tests): u8 images [Cb,H,W], index masks [H,W] or
    [1,H,W], optional colour masks [3,H,W].  Tensors may live on the host or already on the GPU."""

    def __init__(self, images: Sequence[torch.Tensor], index_masks: Sequence[torch.Tensor],
                 color_masks: Optional[Sequence[torch.Tensor]] = None):
        assert len(images) == len(index_masks) and len(images) > 0
        self.root = "<memory>"
        self.imdir = self.indexdir = self.maskdir = "<memory>"
        self._img = {f"scene_{i:05d}": t for i, t in enumerate(images)}
        self._idx = {f"scene_{i:05d}": (m if m.dim() == 3 else m.unsqueeze(0)) for i, m in enumerate(index_masks)}
        self._col = None if color_masks is None else {f"scene_{i:05d}": t for i, t in enumerate(color_masks)}
        self.images = sorted(self._img)
        self.index_masks = list(self.images)

    def shape(self) -> List[int]:
        return list(self._img[self.images[0]].shape)[1:]

    def load(self, name: str, want_color: bool):
        col = None
        if want_color:
            col = self._col[name] if self._col is not None else self._idx[name].expand(3, -1, -1).contiguous()
        return self._img[name], self._idx[name], col

    def load_index_mask(self, name: str) -> torch.Tensor:
        return self._idx[name]


def _to_device_u8(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    if t.dtype != torch.uint8:
        raise RuntimeError(f"cvcs_b200.dataset: scenes must be uint8 (8-bit GID imagery), got {t.dtype}")
    return t.to(dev, non_blocking=True).contiguous()


# ---- IterableChunk -----------------------------------------------------------------------------------
class IterableChunk(torch.utils.data.IterableDataset):
    """Same constructor arguments, attributes (``patches``, ``chunk_crops``, ``chunk_size``,
    ``tiles_in_img_shape``) and iteration protocol as the reference's (dataset.py:108-240); the
    tiles are cut by K5 on the GPU and ``patches`` holds views into the batch tensors."""

    def __init__(self, chunk, images, indexdir, maskdir, image_shape, tpi, patch_size=224, random_shift=False,
                 random_tps=None, iT=None, mT=None, load_context=True, load_color_mask=True, *, source=None,
                 device=None):
        super().__init__()
        self.indexdir, self.maskdir = indexdir, maskdir
        self.p = patch_size
        self.iT, self.mT = iT, mT
        self.image_shape = image_shape
        self.tpi = tpi
        self.random_shift = random_shift
        self.random_tps = random_tps
        self.load_context = load_context
        self.load_color_mask = load_color_mask
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.source = source
        p = self.p

        self.tiles_in_img_shape = tiles_in_image(image_shape, p)
        self.to_load = [images[idx] for idx in chunk]
        self.chunk_size = len(chunk)
        self.chunk_crops = list(range(self.tpi * self.chunk_size))
        random.shuffle(self.chunk_crops)                                   # dataset.py:128
        self.images, self.index_masks, self.color_masks = self.load_images(self.to_load)

        # tile origins in chunk order (consumes `random` exactly as the reference loop does)
        n = len(self.chunk_crops)
        scene_of, yx = [], []
        for x in self.chunk_crops:
            s, tly, tlx = tile_origin(x, self.tpi, self.tiles_in_img_shape[1], p)
            if self.random_shift:
                tly, tlx = _random_shift(tly, tlx, SHIFT_OFFSET)
            scene_of.append(s)
            yx.append((tly, tlx))
        self.tile_scene = scene_of
        self.tile_yx = yx

        cb = self.images[0].shape[0] if self.images else 3
        dev = self.device
        self.batch_image = torch.empty((n, cb, p, p), dtype=torch.uint8, device=dev)
        self.batch_index = torch.empty((n, p, p), dtype=torch.uint8, device=dev)
        self.batch_color = torch.empty((n, 3, p, p), dtype=torch.uint8, device=dev) if load_color_mask else None
        self.batch_context = None
        self._gather(self.batch_image, label_out=self.batch_index, color_out=self.batch_color)
        if self.load_context:
            self.batch_context = self._contexts()

        none = torch.tensor([0])
        self.patches = []
        for i in range(n):
            patch, index_mask = self.batch_image[i], self.batch_index[i]
            color_mask = self.batch_color[i] if load_color_mask else none
            context = self.batch_context[i] if self.load_context else none
            if self.iT is not None:                                        # dataset.py:160-161
                patch = self.iT(patch)
            if self.mT is not None:                                        # dataset.py:163-168
                cm3 = color_mask if load_color_mask else torch.zeros((0, p, p), dtype=torch.uint8, device=dev)
                cat = self.mT(torch.concat((patch, index_mask.unsqueeze(0), cm3), dim=0))
                patch, index_mask = cat[:cb], cat[cb]
                color_mask = cat[cb + 1:] if load_color_mask else none
            self.patches.append((patch, index_mask, color_mask, context))
        if self.iT is not None or self.mT is not None:
            self.batch_image = self.batch_index = self.batch_color = None  # no longer what `patches` holds
        if self.random_tps:
            self._append_random_tps()

    # -- reference API -------------------------------------------------------------------------------
    def load_images(self, names):
        """dataset.py:206-224: returns (images, index_masks, color_masks), here as GPU u8 tensors."""
        print("Loading chunk:")
        for i, name in enumerate(names):
            print(name, flush=True)
            if i == 5:
                print("...\nOutput is collapsed. More then 5 images are being loaded!")
                break
        images, index_masks, color_masks = [], [], []
        for name in names:
            img, idx, col = self.source.load(name, self.load_color_mask)
            images.append(_to_device_u8(img, self.device))
            index_masks.append(_to_device_u8(idx, self.device))
            color_masks.append(_to_device_u8(col, self.device) if col is not None else None)
        return images, index_masks, color_masks

    def __iter__(self):
        return iter(self.patches)

    def __len__(self):
        return len(self.patches)

    # -- GPU-native access ---------------------------------------------------------------------------
    @property
    def batch(self):
        """(images u8 [N,Cb,p,p], index masks u8 [N,p,p], colour masks u8 [N,3,p,p] | None) in chunk order."""
        if self.batch_image is None:
            raise RuntimeError("per-patch transforms were applied; the contiguous batch is not available")
        return self.batch_image, self.batch_index, self.batch_color

    def float_tiles(self, mean: Optional[torch.Tensor] = None, std: Optional[torch.Tensor] = None,
                    dtype: torch.dtype = torch.float32, hist: Optional[torch.Tensor] = None, hist_classes: int = 0,
                    hist_ignore_index: int = -100):
        """Gather + cast (+ per-band normalise) in the same pass: what ``image.type(torch.float32)``
        (train.py:121) / ``SegformerMod.preprocessor`` (nets.py:339-342) compute after the crop.
        Optionally accumulates the label histogram of the emitted tiles (K4 fused into K5).
        Returns (tiles [N,Cb,p,p] dtype, index masks u8 [N,p,p])."""
        n, p = len(self.chunk_crops), self.p
        cb = self.images[0].shape[0]
        out = torch.empty((n, cb, p, p), dtype=dtype, device=self.device)
        lab = torch.empty((n, p, p), dtype=torch.uint8, device=self.device)
        self._gather(out, label_out=lab, color_out=None, mean=mean, std=std, hist=hist, hist_classes=hist_classes,
                     hist_ignore_index=hist_ignore_index)
        return out, lab

    # -- internals ---------------------------------------------------------------------------------------
    def _gather(self, out, label_out, color_out, mean=None, std=None, hist=None, hist_classes=0,
                hist_ignore_index=-100):
        dev = self.device
        for s in range(self.chunk_size):
            slots = [i for i, so in enumerate(self.tile_scene) if so == s]
            if not slots:
                continue
            yx = torch.tensor([self.tile_yx[i] for i in slots], dtype=torch.int32).to(dev, non_blocking=True)
            sl = torch.tensor(slots, dtype=torch.int32).to(dev, non_blocking=True)
            ops.tile_normalize(self.images[s], yx, (self.p, self.p), mean, std, out_dtype=out.dtype,
                               label=self.index_masks[s][0], hist=hist, hist_classes=hist_classes,
                               hist_ignore_index=hist_ignore_index, slots=sl, out=out, label_out=label_out)
            if color_out is not None:
                ops.tile_normalize(self.color_masks[s], yx, (self.p, self.p), out_dtype=torch.uint8, slots=sl,
                                   out=color_out)

    def _contexts(self):
        """``_get_context`` (dataset.py:11-16): the 3p x 3p neighbourhood of every tile of the chunk, reduced to p x p
        by ``cvcs_tile_context`` (crop + the reference's antialiased bilinear Resize in one kernel, byte-identical;
        one launch per scene, tiles scattered to their slots of the chunk batch)."""
        p, dev = self.p, self.device
        n = len(self.chunk_crops)
        cb = self.images[0].shape[0]
        out = torch.empty((n, cb, p, p), dtype=torch.uint8, device=dev)
        for s in range(self.chunk_size):
            slots = [i for i, so in enumerate(self.tile_scene) if so == s]
            if not slots:
                continue
            yx = torch.tensor([self.tile_yx[i] for i in slots], dtype=torch.int32).to(dev)
            ops.tile_context(self.images[s], yx, p, slots=torch.tensor(slots, dtype=torch.int32).to(dev), out=out)
        return out

    def _append_random_tps(self):
        """Random rescaled crops (dataset.py:173-203), same draws in the same order; the crops are K5,
        the resizes torchvision's on the GPU."""
        import torchvision.transforms.v2 as v2
        p, dev = self.p, self.device
        image_resizer = v2.Resize(p, interpolation=v2.InterpolationMode.BILINEAR)
        mask_resizer = v2.Resize(p, interpolation=v2.InterpolationMode.NEAREST_EXACT)
        none = torch.tensor([0])
        for aug_size, percentage in self.random_tps:
            h, w = self.image_shape
            for _ in range(int(percentage * len(self.chunk_crops))):
                rand_index = random.randint(0, len(self.images) - 1)
                random_y = random.randint(0, h - 1 - aug_size)
                random_x = random.randint(0, w - 1 - aug_size)
                yx = torch.tensor([[random_y, random_x]], dtype=torch.int32, device=dev)
                patch, index_mask = ops.tile_normalize(self.images[rand_index], yx, (aug_size, aug_size),
                                                       out_dtype=torch.uint8, label=self.index_masks[rand_index][0])
                if self.load_context:
                    context = ops.tile_context(self.images[rand_index], yx, p)[0]      # _get_context(…, random_y, random_x, p, …)
                else:
                    context = none
                patch = image_resizer(patch)[0]
                index_mask = mask_resizer(index_mask)[0]
                if self.load_color_mask:
                    color_mask, _ = ops.tile_normalize(self.color_masks[rand_index], yx, (aug_size, aug_size),
                                                       out_dtype=torch.uint8)
                    color_mask = mask_resizer(color_mask)[0]
                else:
                    color_mask = none
                self.patches.append((patch, index_mask, color_mask, context))
                random.shuffle(self.patches)                               # (sic) dataset.py:203, inside the loop


# ---- Loader ------------------------------------------------------------------------------------------------
class Loader:
    """Same signature and behaviour as the reference's ``Loader`` (dataset.py:241-345):
    ``__len__``, ``shuffle()``, ``specify()``, ``get_iterable_chunk(idx, random_tps=None)``,
    ``get_chunk``, ``print_chunk``, ``get_class_weights``, ``get_class_priors``.

    ``root`` is the dataset directory, or a scene source object (``ArrayScenes``)."""

    def __init__(self, root, chunk_size=2, random_shift=False, patch_size=224, image_transforms=None,
                 mask_transforms=None, load_context=True, load_color_mask=True, *, device=None,
                 strict_patch_size=True):
        self.source = DirectoryScenes(root) if isinstance(root, (str, os.PathLike)) else root
        self.root = self.source.root
        self.patch_size = patch_size
        self.chunk_size = chunk_size
        self.random_shift = random_shift
        self.image_transforms = image_transforms
        self.mask_transforms = mask_transforms
        self.count = None
        self.load_context = load_context
        self.load_color_mask = load_color_mask
        self.device = device

        self.imdir, self.indexdir, self.maskdir = self.source.imdir, self.source.indexdir, self.source.maskdir
        self.images = list(self.source.images)
        self.index_masks = list(self.source.index_masks)
        self.image_shape = self.source.shape()
        self.tpi = self.__get_tpi()
        self.idxs = list(range(len(self.images)))
        self.chunks = None
        if strict_patch_size:
            assert patch_size in [224, 256, 512], "Patch size either not supported or not recommended"
        assert len(self.images) % self.chunk_size == 0, (
            f"Number of images not divisible by chunk size. images:{len(self.images)}, cs:{self.chunk_size}")
        self.__generate_chunks()

    def __get_tpi(self, p=None):
        return tiles_per_image(self.image_shape, self.patch_size if p is None else p)

    def shuffle(self):
        random.shuffle(self.idxs)
        self.__generate_chunks()

    def get_iterable_chunk(self, idx, random_tps=None):
        return IterableChunk(self.chunks[idx], self.images, self.indexdir, self.maskdir, image_shape=self.image_shape,
                             tpi=self.__get_tpi(self.patch_size), random_shift=self.random_shift,
                             patch_size=self.patch_size, random_tps=random_tps, iT=self.image_transforms,
                             mT=self.mask_transforms, load_context=self.load_context,
                             load_color_mask=self.load_color_mask, source=self.source, device=self.device)

    def get_chunk(self, idx):
        return [self.images[i] for i in self.chunks[idx]]

    def print_chunk(self, idx):
        for im in self.get_chunk(idx):
            print(im)

    def __generate_chunks(self):
        cs = self.chunk_size
        self.chunks = [[self.idxs[i + cs * offset] for i in range(cs)] for offset in range(len(self.idxs) // cs)]

    def __len__(self):
        return len(self.chunks)

    def specify(self, targets):
        self.idxs = [self.idxs[i] for i in targets]
        self.__generate_chunks()

    # -- class statistics (dataset.py:346-387) ---------------------------------------------------------------
    def _get_class_count(self, classes):
        """Per-class pixel counts over every index mask.  One K4 launch per scene replaces the
        reference's ``classes`` full passes (``torch.sum(mask == cl)`` per class, dataset.py:356-357);
        the per-scene integer counts are then accumulated into a float32 vector in scene order,
        exactly as the reference's ``self.count[cl] += ...`` does (so counts above 2^24 round the same)."""
        if self.count is None:
            dev = torch.device(self.device) if self.device is not None else torch.device("cuda", torch.cuda.current_device())
            self.count = torch.zeros(classes, dtype=torch.float32)
            for name in self.index_masks:
                mask = _to_device_u8(self.source.load_index_mask(name), dev)
                hist = ops.label_hist(mask, classes, ignore_index=-100)
                per_scene = hist[:classes].cpu()
                for cl in range(classes):
                    self.count[cl] += per_scene[cl]
        return self.count

    def get_class_weights(self, classes: int, ignore_background=False):
        """w_j = Σn / (bins · n_j), 0 for empty classes (dataset.py:360-384)."""
        return class_weights_from_counts(self._get_class_count(classes), ignore_background)

    def get_class_priors(self, classes):
        counts = self._get_class_count(classes)
        return torch.sum(counts) / counts   # (sic) the reference returns the inverse priors, dataset.py:386-387
