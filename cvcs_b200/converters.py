"""Drop-in for ``converters.GID15Converter`` (converters.py:3-36): class-index mask -> RGB mask.

The reference loops over the 16 classes on the CPU (``output[mask == label] = color``); here it is one
look-up kernel (``cvcs_colorize``).  Indices outside the table keep the initial value 1.0, as in the reference
(``torch.ones``)."""
from __future__ import annotations

import torch

from . import ops

# GID-15 palette (RGB, 0-255) by class index — the dataset's published colour coding (converters.py:5-22)
GID15_COLORS = (
    (0, 0, 0), (200, 0, 0), (250, 0, 150), (200, 150, 150), (250, 150, 150), (0, 200, 0), (150, 250, 0),
    (150, 200, 150), (200, 0, 200), (150, 0, 250), (150, 150, 250), (250, 200, 0), (200, 200, 0), (0, 0, 200),
    (0, 150, 200), (0, 200, 250),
)


class GID15Converter:
    def __init__(self):
        self.color_to_label = {c: i for i, c in enumerate(GID15_COLORS)}
        self._lut = {}

    def lut(self, device) -> torch.Tensor:
        key = str(device)
        if key not in self._lut:
            self._lut[key] = (torch.tensor(GID15_COLORS).type(torch.float32) / 255).to(device)   # same arithmetic as the reference
        return self._lut[key]

    def iconvert(self, mask: torch.Tensor) -> torch.Tensor:
        """class label mask [H,W] (uint8 / int64; CPU tensors are moved to the current GPU) -> float32 [H,W,3]
        in 0..1 on the mask's device (CPU in -> CPU out, as the reference returns)."""
        was_cpu = not mask.is_cuda
        m = mask.to("cuda") if was_cpu else mask
        if m.dtype not in (torch.uint8, torch.int64):
            m = m.to(torch.int64)
        out = ops.colorize(m, self.lut(m.device))
        return out.cpu() if was_cpu else out
