"""Data-parallel tile sharding: one process per GPU, every tile independent.

The reference is single-process; its hot path shards because every stage is per-pixel and the two
reductions (loss, confusion matrix) are commutative sums.  This module holds the host logic:

  * which rank owns which tile — global tile id ``g = scene * tpi + row * cols + col`` in the
    reference's own ordering (dataset.py:137-140), dealt round-robin ``g mod R`` (default) or by
    whole scenes;
  * the three tiny collectives that make an R-rank run produce the single-process result:
      (1) before K1, training only: all-reduce of the label histogram u64[C+2] -> global Σ v·w[y],
          so that every rank divides its gradients by the same total weight;
      (2) after K1: all-reduce of f64[3] {Σ w·nll, Σ w, #out-of-bounds} -> global mean loss;
      (3) once per evaluation pass: all-reduce of the C×C int64 confusion matrix.
    Integer sums are exact in any order; the fp64 loss sums agree to ~1e-15.

``torch.distributed`` is plumbing (NCCL over NVLink on the GPU box, gloo in the CPU tests); the
payloads are at most C*C*8 bytes, so the collectives are latency-bound and are issued once per
step / per pass, never per tile.
"""
from __future__ import annotations

from typing import Iterator, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .dataset import tile_origin, tiles_in_image


# ---- ownership ---------------------------------------------------------------------------------------
def world_info(group=None) -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def tile_owner(g: int, world: int, tpi: int = 0, policy: str = "round_robin") -> int:
    """Rank that processes global tile ``g``."""
    if policy == "round_robin":
        return g % world
    if policy == "scene":
        return (g // tpi) % world
    raise ValueError(f"unknown sharding policy {policy!r}")


def local_tiles(n_scenes: int, image_shape: Sequence[int], p: int, rank: int, world: int,
                policy: str = "round_robin") -> List[Tuple[int, int, int, int]]:
    """This rank's tiles as (global id, scene, tly, tlx), ascending global id."""
    rows, cols = tiles_in_image(image_shape, p)
    tpi = rows * cols
    out = []
    for g in range(n_scenes * tpi):
        if tile_owner(g, world, tpi, policy) == rank:
            s, tly, tlx = tile_origin(g, tpi, cols, p)
            out.append((g, s, tly, tlx))
    return out


def batches(items: Sequence, batch_size: int) -> Iterator[Sequence]:
    for i in range(0, len(items), batch_size):
        yield items[i:i + batch_size]


# ---- collectives ----------------------------------------------------------------------------------------
def all_reduce_sum_(t: torch.Tensor, group=None, async_op: bool = False):
    """In-place sum over ranks (no-op for a single process).  int64 / float64 payloads only: the
    results must not depend on the reduction order beyond fp64 rounding."""
    if t.dtype not in (torch.int64, torch.float64):
        raise TypeError(f"all_reduce_sum_: int64 or float64 expected, got {t.dtype}")
    _, world = world_info(group)
    if world == 1:
        return None
    return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def global_total_weight(hist: torch.Tensor, weight: Optional[torch.Tensor], num_classes: int, ignore_index: int,
                        group=None) -> torch.Tensor:
    """(1): all-reduce the label histogram (int64[C+2], cvcs_label_hist layout) in place and return
    f64[2] {Σ v·w[y], 1/Σ} — on the GPU through K4's companion kernel, on the CPU (gloo tests) with
    the same fp64 arithmetic."""
    all_reduce_sum_(hist, group)
    if hist.is_cuda:
        from . import ops
        return ops.total_weight(hist, weight, num_classes, ignore_index)
    h = hist[:num_classes].to(torch.float64)
    w = torch.ones(num_classes, dtype=torch.float64) if weight is None else weight.to(torch.float32).to(torch.float64)
    keep = torch.ones(num_classes, dtype=torch.bool)
    if 0 <= ignore_index < num_classes:
        keep[ignore_index] = False
    s = torch.zeros((), dtype=torch.float64)
    for c in range(num_classes):                         # fixed order, as the one-warp kernel's lanes fold
        if keep[c]:
            s = s + h[c] * w[c]
    return torch.stack((s, 1.0 / s))


def global_loss(sums: torch.Tensor, group=None) -> torch.Tensor:
    """(2): f64[3] {Σ w·nll, Σ w, #oob} per rank -> the mean loss over ALL ranks' pixels (float32,
    NaN if any rank saw an out-of-bounds label, or if every pixel everywhere was ignored)."""
    s = sums.clone()
    all_reduce_sum_(s, group)
    loss = (s[0] / s[1]).to(torch.float32)
    return torch.where(s[2] > 0, torch.full_like(loss, float("nan")), loss)


def global_confmat(confmat: torch.Tensor, group=None) -> torch.Tensor:
    """(3): in-place all-reduce of the int64[C,C] confusion matrix."""
    all_reduce_sum_(confmat, group)
    return confmat


class WeightExchange:
    """Collective (1) without a collective call: the per-step global Σ v·w[y] is exchanged INSIDE the fused loss kernel
    (``ops.ce_fused(..., total_weight="kernel", xchg=...)`` / ``FusedCrossEntropyLoss(exchange=...)``).  Each rank owns a
    small device block; here the blocks' CUDA IPC handles travel once through ``all_gather_object`` and every rank maps
    its peers' blocks (NVLink peer access).  Afterwards a step costs one kernel launch: rank r's kernel stores its sum
    into every rank's block and adds the N values it finds in its own, in rank order.

    Every rank must run the same sequence of exchanging launches (as with any collective).  COLLECTIVE constructor."""

    def __init__(self, group=None, device=None):
        from . import ops
        self.group = group
        self.rank, self.world = world_info(group)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.x = ops.Exchange(self.world, self.rank, self.device)
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, self.x.local_handle(), group=group)
            for q, h in enumerate(handles):
                if q != self.rank:
                    self.x.open_peer(q, h)
            dist.barrier(group)            # nobody launches before every mapping exists

    def handle_for(self, dev: torch.device):
        if torch.device(dev) != self.device:
            raise RuntimeError(f"WeightExchange lives on {self.device}, tensors on {dev}")
        return self.x

    def all_reduce_(self, t: torch.Tensor) -> torch.Tensor:
        """Collectives (2) / (3) without NCCL: in-place sum over the ranks of a short float64 or int64 device tensor
        (<= 2048 elements; int64 counts travel as float64, exact below 2^53), added in rank order.  COLLECTIVE."""
        if self.world == 1:
            return t
        if t.dtype == torch.float64 and t.is_contiguous():
            self.x.allreduce_(t.view(-1))
            return t
        if t.dtype != torch.int64:
            raise TypeError(f"WeightExchange.all_reduce_: int64 or float64 expected, got {t.dtype}")
        f = t.to(torch.float64).contiguous().view(-1)
        self.x.allreduce_(f)
        t.copy_(f.view(t.shape).to(torch.int64))
        return t

    def state(self):
        """(exchanges completed, time-outs / overruns seen) on this rank — synchronises."""
        return self.x.state()

    def close(self):
        if self.world > 1:
            torch.cuda.synchronize(self.device)
            dist.barrier(self.group)       # peers may still be reading / writing this rank's block
        self.x.close()


# ---- a sharded evaluation / loss pass over scenes ---------------------------------------------------------------
class ShardedScenePass:
    """Tiles a list of scenes, runs ``logits_fn`` on this rank's tile batches and accumulates the
    fused loss / argmax / confusion-matrix path; ``finish()`` returns the global results.

    scenes      sequence of (image u8 [Cb,H,W], label u8 [H,W]) tensors (host or device)
    logits_fn   (tiles float [b,Cb,p,p], labels u8 [b,p,p]) -> logits [b,C,p,p] (the segmenter, or
                a stub in benchmarks — it is not part of the hot path)
    """

    def __init__(self, scenes, p: int, num_classes: int, logits_fn, *, weight: Optional[torch.Tensor] = None,
                 ignore_index: int = -100, batch_size: int = 16, device=None, group=None, policy: str = "round_robin",
                 mean: Optional[torch.Tensor] = None, std: Optional[torch.Tensor] = None,
                 tile_dtype: torch.dtype = torch.float32, want_grad: bool = False, single_process: bool = False):
        from . import ops
        self.ops = ops
        self.scenes, self.p, self.C, self.logits_fn = scenes, p, num_classes, logits_fn
        self.weight, self.ignore_index, self.batch_size = weight, ignore_index, batch_size
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.group, self.policy = group, policy
        self.mean, self.std, self.tile_dtype, self.want_grad = mean, std, tile_dtype, want_grad
        # single_process: ignore torch.distributed (this process takes every tile, no collectives)
        self.single = single_process
        self.rank, self.world = (0, 1) if single_process else world_info(group)
        self.image_shape = list(scenes[0][0].shape[1:])
        self.tiles = local_tiles(len(scenes), self.image_shape, p, self.rank, self.world, policy)
        dev = self.device
        self.confmat = torch.zeros((num_classes, num_classes), dtype=torch.int64, device=dev)
        self.sums = torch.zeros(3, dtype=torch.float64, device=dev)
        self.hist = torch.zeros(num_classes + 2, dtype=torch.int64, device=dev)
        self.n_tiles_done = 0
        self.last_dlogits = None

    def _label_hist_pass(self):
        """K4 over this rank's label tiles (needed before K1 only when gradients are wanted)."""
        ops = self.ops
        by_scene = {}
        for g, s, tly, tlx in self.tiles:
            by_scene.setdefault(s, []).append((tly, tlx))
        for s, yx in by_scene.items():
            lab = self.scenes[s][1].to(self.device, non_blocking=True).contiguous()
            yx_t = torch.tensor(yx, dtype=torch.int32).to(self.device)
            tiles, _ = ops.tile_normalize(lab[None], yx_t, (self.p, self.p), out_dtype=torch.uint8)
            ops.label_hist(tiles, self.C, self.ignore_index, hist=self.hist)

    def run(self):
        ops, dev, p = self.ops, self.device, self.p
        inv_tw_dev = None
        if self.want_grad:
            self._label_hist_pass()
            if self.single:
                tw = ops.total_weight(self.hist, self.weight, self.C, self.ignore_index)
            else:
                tw = global_total_weight(self.hist, self.weight, self.C, self.ignore_index, self.group)   # collective (1)
            inv_tw_dev = tw[1:]
        step_sums = torch.empty(3, dtype=torch.float64, device=dev)
        cache = {}
        for part in batches(self.tiles, self.batch_size):
            n = len(part)
            x = torch.empty((n, self.scenes[0][0].shape[0], p, p), dtype=self.tile_dtype, device=dev)
            y = torch.empty((n, p, p), dtype=torch.uint8, device=dev)
            for s in sorted({t[1] for t in part}):
                if s not in cache:
                    cache.clear()                                          # one scene resident at a time
                    cache[s] = (self.scenes[s][0].to(dev, non_blocking=True).contiguous(),
                                self.scenes[s][1].to(dev, non_blocking=True).contiguous())
                img, lab = cache[s]
                slots = [i for i, t in enumerate(part) if t[1] == s]
                yx = torch.tensor([(part[i][2], part[i][3]) for i in slots], dtype=torch.int32).to(dev)
                ops.tile_normalize(img, yx, (p, p), self.mean, self.std, out_dtype=self.tile_dtype, label=lab,
                                   slots=torch.tensor(slots, dtype=torch.int32).to(dev), out=x, label_out=y)
            logits = self.logits_fn(x, y)
            _, _, d = ops.ce_fused(logits, y, self.weight, self.ignore_index, want_grad=self.want_grad,
                                   inv_total_weight_dev=inv_tw_dev, confmat=self.confmat, loss_sums=step_sums)
            self.sums += step_sums
            self.last_dlogits = d
            self.n_tiles_done += n
        return self

    def finish(self):
        """Collectives (2) and (3).  Returns (global loss f32 0-dim, global confusion int64[C,C] on the host)."""
        if self.single:
            loss = (self.sums[0] / self.sums[1]).to(torch.float32)
            loss = torch.where(self.sums[2] > 0, torch.full_like(loss, float("nan")), loss)
            return loss, self.confmat.cpu()
        loss = global_loss(self.sums, self.group)
        cm = global_confmat(self.confmat, self.group)
        return loss, cm.cpu()
