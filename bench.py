#!/usr/bin/env python
"""bench.py — throughput of the fused segmentation loss + metric path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch of synthetic tiles per GPU.  Rank 0 prints ONE JSON line.

Workloads (BASELINE.json configs; the default line is cfg2 and carries the others in `secondary`):
  cfg2   K1 on B16 x C7 x 1024^2 fp32 logits: softmax-CE fwd+bwd + argmax + confusion matrix
  cfg3   the same with bf16 logits, class weights and ignore_index 255 (K4 pre-pass for the data dependent Σw)
  cfg4   10 000^2 x 3-band u8 scenes -> 81 tiles of 1024^2 each, dealt round-robin by global tile id to the ranks:
         K5 (tile + cast) -> [segmenter, not timed: logits are synthetic] -> K1; ONE CxC all-reduce per pass
  cfg5   13-band u8 scene -> K5 (tile + per-band normalise) -> C=20 K1, batch 64 of 1024^2 per GPU
  cfg5head / c16 / ref / tile13 / tile3   parts and reference-shaped variants (see WORKLOADS)

  value     whole-job Gpixel/s, inputs resident in HBM, CUDA-event timed, max over ranks
  roofline  the dominant kernel (K1): algorithmic bytes per launch / its average launch duration against
            MEASURED_PEAKS.json (events around every launch, or region / steps when launches overlap under --pdl 1)
  e2e       the same metric through the host-buffer C-ABI call (cvcs_host_ce_fused): pinned host logits + labels
            copied in, loss + confusion matrix read back, every step; e2e_eval also returns the u8 argmax map
  cpu_baseline   the reference's own CPU path (oracle/torch_path.py: the torch calls the reference makes) timed on
            this box's host cores on a bounded sample (rank 0, N=1)
  torch_cuda_baseline  stock torch CUDA ops for the same step on the same GPU (what the reference runs with
            device: gpu — nn.CrossEntropyLoss fwd+bwd, argmax, bincount), rank 0, N=1
  --impl reference   the CPU path as the measured arm (no GPU work at all)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gpixel/s fused seg loss+metric"
UNIT = "Gpixel/s"

WORKLOADS = {
    # K1 workloads: per-GPU batch, classes, H, W, logits dtype, weights, ignore_index
    "cfg2": dict(kind="ce", B=16, C=7, H=1024, W=1024, dtype="f32", weighted=False, ignore_index=-100,
                 desc="1xB200: fused softmax-CE fwd+bwd + argmax + confusion matrix, batch 16 of 1024x1024, 7 classes, fp32 logits"),
    "cfg3": dict(kind="ce", B=16, C=7, H=1024, W=1024, dtype="bf16", weighted=True, ignore_index=255,
                 desc="same path with bf16 logits, class weights and ignore_index=255 (LoveDA-style labels)"),
    "bf16c7": dict(kind="ce", B=16, C=7, H=1024, W=1024, dtype="bf16", weighted=False, ignore_index=-100,
                   desc="cfg2 with bf16 logits (no class weights, nothing ignored: K1 alone, no pre-pass)"),
    "cfg5head": dict(kind="ce", B=16, C=20, H=1024, W=1024, dtype="f32", weighted=False, ignore_index=-100,
                     desc="20-class head alone, fp32 logits, batch 16 per GPU"),
    "c16": dict(kind="ce", B=16, C=16, H=1024, W=1024, dtype="f32", weighted=False, ignore_index=-100,
                desc="16 classes (what utils.py:77-78 hard-codes), fp32 logits, batch 16 of 1024x1024"),
    "c32": dict(kind="ce", B=4, C=32, H=1024, W=1024, dtype="f32", weighted=False, ignore_index=-100,
                desc="32 classes (beyond the register-resident range C <= 21: the generic one-pixel-per-thread kernel), fp32, batch 4"),
    "c64": dict(kind="ce", B=4, C=64, H=1024, W=1024, dtype="f32", weighted=False, ignore_index=-100,
                desc="64 classes (generic kernel), fp32 logits, batch 4 of 1024x1024"),
    "ref": dict(kind="ce", B=10, C=16, H=224, W=224, dtype="f32", weighted=False, ignore_index=0, label_dtype="i64",
                desc="the reference's own training shape (configs/train/server.yaml:23-36): batch 10 of 224x224, 16 classes, "
                     "int64 labels, ignore_index 0"),
    # chains: K5 -> (segmenter, not timed) -> K1
    "cfg4": dict(kind="chain", B=16, C=7, Cb=3, H=1024, W=1024, dtype="f32", weighted=False, ignore_index=-100,
                 scene=10000, scenes_per_gpu=2, normalise=False,
                 desc="2/4/8xB200: 10000x10000 3-band u8 scenes, 81 tiles of 1024x1024 each dealt round-robin by global tile "
                      "id; K5 tile+cast -> K1 (fp32, 7 classes); one CxC all-reduce per pass"),
    "cfg5": dict(kind="chain", B=64, C=20, Cb=13, H=1024, W=1024, dtype="f32", weighted=False, ignore_index=-100,
                 scene=8192, scenes_per_gpu=1, normalise=True,
                 desc="8xB200: 13-band u8 scene -> K5 tile + per-band normalise -> 20-class head K1, batch 64 of 1024x1024 per GPU"),
    # K5 alone: the tiler / normaliser either side of the model (Cb + Cb*s_out + 2 bytes/px)
    "tile13": dict(kind="tile", B=64, Cb=13, H=1024, W=1024, dtype="f32", scene=8192,
                   desc="cfg5 tiler: 13-band u8 scene -> 64 normalised fp32 tiles of 1024x1024 + label tiles"),
    "tile3": dict(kind="tile", B=64, Cb=3, H=1024, W=1024, dtype="f32", scene=8192,
                  desc="RGB tiler: u8 scene -> 64 fp32 tiles of 1024x1024 (train.py:121 cast) + label tiles"),
}
SECONDARY_DEFAULT = ["cfg3", "cfg4", "cfg5", "cfg5head", "c16", "tile13", "eval_only", "metrics_only", "i64_labels", "ref"]


def algorithmic_bytes_per_pixel(C: int, esize: int, grad: bool = True) -> int:
    """read logits + write dlogits + read label (u8) + write argmax (u8)."""
    return C * esize * (2 if grad else 1) + 2


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(key: str):
    """DRAM bytes per launch from the committed ncu captures (profiles/traffic.json), keyed by
    workload[/nograd|/metrics][/nhwc] so that a forward-only line never reports the gradient kernel's bytes."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(key)
    return None


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML, ~1 ms period: the driver's 20-step region lasts
    3 ms)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons = index, [], set()
        self.max_mhz, self._stop, self._t, self.ok = None, threading.Event(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    _NAMES = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
              0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self._NAMES.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.001)

    def start(self):
        if self.ok:
            self._stop.clear()
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        if self._t is not None:
            self._stop.set()
            self._t.join()
            self._t = None

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---- synthetic inputs ---------------------------------------------------------------------------------------------
def blocky_labels(torch, shape, C, g, dev, block=32):
    """Constant block x block label blocks (real masks have long runs)."""
    *lead, H, W = shape
    hb, wb = -(-H // block), -(-W // block)
    t = torch.randint(0, C, (*lead, hb, wb), generator=g, device=dev, dtype=torch.uint8)
    t = t.repeat_interleave(block, -2).repeat_interleave(block, -1)[..., :H, :W]
    return t.contiguous()


def synth_inputs(torch, wl, dev, seed, n_sets, label_dtype="u8", layout="nchw", block=32):
    """Seeded synthetic logits (randn*3) and blocky labels; cfg3 adds 10% ignore pixels and histogram-derived class
    weights; ignore_index 0 (the reference's ignore_background) simply makes class 0 the ignored one."""
    B, C, H, W = wl["B"], wl["C"], wl["H"], wl["W"]
    dt = torch.float32 if wl["dtype"] == "f32" else torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(seed)
    sets = []
    for _ in range(n_sets):
        x = (torch.randn(B, C, H, W, generator=g, device=dev, dtype=torch.float32) * 3).to(dt)
        t = blocky_labels(torch, (B, H, W), C, g, dev, block=block)
        if wl["ignore_index"] == 255:
            t[torch.rand(B, H, W, generator=g, device=dev) < 0.1] = 255
        if label_dtype == "i64":
            t = t.long()
        if layout == "nhwc":
            x = x.contiguous(memory_format=torch.channels_last)
        sets.append((x, t))
    weight = None
    if wl["weighted"]:
        t0 = sets[0][1]
        counts = torch.bincount(t0[t0 != 255].flatten().long(), minlength=C).float()
        weight = (counts.sum() / (C * counts.clamp(min=1))).to(torch.float32)
    return sets, weight


# ---- the CPU arm ----------------------------------------------------------------------------------------------------
def cpu_reference_rate(wl, steps, warmup, budget_s):
    """Times the reference's CPU path (oracle/torch_path.hot_path_step) on a bounded sample of the
    workload.  Returns (Gpixel/s, seconds per step, sample description, cores)."""
    import torch
    from oracle import torch_path
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    C, W, H = wl["C"], wl["W"], wl["H"]
    g = torch.Generator().manual_seed(0)

    def make(rows, tiles):
        x = torch.randn(tiles, C, rows, W, generator=g) * 3
        if wl["dtype"] == "bf16":
            x = x.to(torch.bfloat16).float()   # torch rejects fp32 weights with bf16 logits: fp32 on the same values
        t = blocky_labels(torch, (tiles, rows, W), C, g, torch.device("cpu"))
        if wl["ignore_index"] == 255:
            t[torch.rand(tiles, rows, W, generator=g) < 0.1] = 255
        w = (torch.rand(C, generator=g) + 0.5) if wl["weighted"] else None
        return x, t, w

    def one(x, t, w):
        t0 = time.perf_counter()
        torch_path.hot_path_step(x, t, w, wl["ignore_index"], C, ignore_background_eval=False)
        return time.perf_counter() - t0

    # calibrate on a strip, then size the per-step sample to the time budget
    strip = min(128, H)
    x, t, w = make(strip, 1)
    one(x, t, w)
    per_px = min(one(x, t, w) for _ in range(2)) / (strip * W)
    px_budget = budget_s / max(steps + warmup, 1) / per_px
    rows = int(min(H, max(32, (px_budget // W) // 32 * 32)))
    tiles = int(min(wl["B"], max(1, px_budget // (rows * W)))) if rows == H else 1   # up to the full per-GPU batch
    x, t, w = make(rows, tiles)
    for _ in range(warmup):
        one(x, t, w)
    times = [one(x, t, w) for _ in range(steps)]
    sec = sum(times) / len(times)
    px = tiles * rows * W
    sample = f"{tiles} tile(s) of {rows}x{W} px, {C} classes per step ({px} px); {steps} steps after {warmup} warm-up"
    return px / sec / 1e9, sec, sample, cores


def cfg1_cpu_line(budget_s=60.0):
    """BASELINE.json configs[0], the reference's own CPU-runnable case: a U-Net-style segmenter forward + CE loss +
    per-tile argmax + confusion matrix + mIoU on batch 2 of 512x512 RGB tiles, 7 classes (nets.py:117-199 is
    restated by oracle/torch_path.unet_like when the reference's nets.py cannot be imported on this box)."""
    import torch
    from oracle import torch_path
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    out = torch_path.cfg1_step(seed=0)
    first = time.perf_counter() - t0
    reps = [first]
    while sum(reps) < budget_s and len(reps) < 3:
        t0 = time.perf_counter()
        torch_path.cfg1_step(seed=0)
        reps.append(time.perf_counter() - t0)
    sec = statistics.median(reps)
    px = 2 * 512 * 512
    return {"value": px / sec / 1e9, "unit": UNIT, "seconds_per_step": sec, "cores": torch.get_num_threads(),
            "cpu": cpu_model(), "kind": "port", "model": out["model"], "params": out["params"], "miou": out["miou"],
            "loss": out["loss"], "sample": f"batch 2 of 512x512 RGB u8 tiles, 7 classes, ignore_index 0; median of {len(reps)}"}


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, sec, sample, cores = cpu_reference_rate(wl, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "sample": sample,
                   "arm": "reference CPU path (torch CPU calls as at utils.py:230,90,93-94; train.py:122-125) on a bounded sample"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "cpu": cpu_model(), "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.cfg1:
        line["cfg1_cpu"] = cfg1_cpu_line()
    print(json.dumps(line), flush=True)


# ---- GPU measurement plumbing -----------------------------------------------------------------------------------------
class Ctx:
    """torch / torch.distributed handles and the per-process options shared by all measurements."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from cvcs_b200 import _lib, ops
        self.torch, self.dist, self.lib, self.ops, self.args = torch, dist, _lib, ops, args
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (cvcs_b200 has no CPU path; use --impl reference for the CPU arm)")
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.sampler = ClockSampler(self.local)       # NVML init takes milliseconds: done once, outside timed regions
        self.peak, self.peak_src = load_peaks()

    def fence(self, *streams):
        torch = self.torch
        for s in streams:
            if s is not None:
                torch.cuda.current_stream(self.dev).wait_stream(s)
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize(self.dev)

    def align_start(self):
        """N > 1: the ranks leave a barrier up to ~100 us apart, and the max-over-ranks region would count that skew once
        per pass.  All ranks share this host's clock: agree on a start instant a few milliseconds ahead and spin to it."""
        if self.world == 1:
            return
        torch = self.torch
        t0 = torch.tensor([time.time() + 0.0015], dtype=torch.float64, device=self.dev)
        self.dist.all_reduce(t0, op=self.dist.ReduceOp.MAX)
        target = float(t0.item())
        if not self.args.nccl_pass_end:
            # ... and the GPUs meet once more ON THE DEVICE: a one-element exchange over the peer-mapped region is a
            # barrier kernel (every rank waits for every rank's flag).  It is enqueued in front of the start event, so the
            # hosts have long queued the timed work when the kernels let go — the regions then start within microseconds
            # of each other whatever the hosts' enqueue jitter.
            if getattr(self, "_align_buf", None) is None:
                self._align_buf = torch.zeros(1, dtype=torch.float64, device=self.dev)
            xchg = self.exchange()
        while time.time() < target:
            pass
        if not self.args.nccl_pass_end:
            # a short device-side delay in front of the rendezvous absorbs a host hiccup between these enqueues and the
            # timed work's (everything is queued long before the delay ends)
            torch.cuda._sleep(400_000)                               # ~0.2 ms of SM clocks
            xchg.allreduce_(self._align_buf)

    def max_over_ranks(self, *vals):
        """Element-wise max over ranks of a few host floats (device-timed regions are compared on the device)."""
        if self.world == 1:
            return list(vals), None
        torch = self.torch
        mine = torch.tensor(list(vals), dtype=torch.float64, device=self.dev)
        allt = [torch.zeros_like(mine) for _ in range(self.world)]
        self.dist.all_gather(allt, mine)
        table = [[float(v) for v in t] for t in allt]
        return [max(row[i] for row in table) for i in range(len(vals))], table

    def exchange(self):
        """The Σw exchange (cvcs_b200.shard.WeightExchange): one per process, created on first use (a collective)."""
        if getattr(self, "_xchg", None) is None:
            from cvcs_b200 import shard
            self._xchg = shard.WeightExchange(device=self.dev)
        return self._xchg.handle_for(self.dev)

    def allreduce_pass_end(self, buf):
        """Pass-end sums (loss table + C x C counts as f64): the one-shot exchange over peer-mapped memory when the vector is
        short enough, else NCCL."""
        if self.world == 1:
            return
        if buf.numel() <= 2048 and not self.args.nccl_pass_end:
            self.exchange().allreduce_(buf)
        else:
            self.dist.all_reduce(buf)

    def set_k1_options(self, path="auto", stages=0, no_wait_hint=False, vecp=0, ctas=0, pdl=1, reserve=0, l2_hint=0):
        L = self.lib
        L.set_option(L.OPT_CE_PATH, {"auto": 0, "tma": 1, "direct": 2, "generic": 3}[path])
        L.set_option(L.OPT_TMA_STAGES, stages)
        L.set_option(L.OPT_TMA_WAIT_HINT, 1 if no_wait_hint else 0)
        L.set_option(L.OPT_TMA_VECP, vecp)
        L.set_option(L.OPT_TMA_CTAS, ctas)
        L.set_option(L.OPT_PDL, 1 if pdl else 2)
        L.set_option(L.OPT_RESERVE_SMS, reserve)
        L.set_option(L.OPT_L2_HINT, l2_hint)


def roofline_dict(ctx, achieved_gbs, bpp, px, k_ms, kernel, traffic_key, source, extra=None):
    d = {"bound": "hbm", "achieved": achieved_gbs, "peak": ctx.peak, "unit": "GB/s", "frac": achieved_gbs / ctx.peak,
         "traffic": load_traffic(traffic_key), "traffic_key": traffic_key, "kernel": kernel, "bytes_per_pixel": bpp,
         "pixels_per_launch": px, "avg_launch_ms": k_ms, "avg_launch_ms_source": source, "peak_source": ctx.peak_src,
         "frac_of_8TBps_nominal": achieved_gbs / 8000.0}
    if extra:
        d.update(extra)
    return d


def measure_ce(ctx, name, wl, steps, warmup, *, grad=True, metrics_only=False, label_dtype=None, layout="nchw",
               per_launch_events=True, copy_ref=False, want_clocks=True, tw_mode="auto", label_block=32, graph=False):
    """K1 on rotating buffer sets larger than L2.  When Σ v·w[y] is data dependent (class weights / ignore_index):
      tw_mode "chain"   a K4 launch one step ahead on a side stream (the next batch's labels are known while the current
                        K1 runs); at N > 1 followed by an NCCL all-reduce of the 8-byte sum
      tw_mode "xchg"    the same K4, but the ranks' sums are exchanged INSIDE K1 over NVLink peer memory (no NCCL)
      tw_mode "kernel"  no K4 at all: K1 sums the weights in its own label pre-pass (grid barrier), then exchanges
      tw_mode "pipe"    no K4 and no pre-pass on the critical path: launch i also sums the weights over batch i+1's labels
                        (in its prologue, while its pipeline fills) and launch i+1 starts from that sum; across GPUs the
                        ranks' sums are exchanged inside K1.  One launch per step, one stream.
      tw_mode "auto"    "pipe" (u8 labels; int64 labels: "chain" on one GPU, "xchg" on several)
    graph=True (stateless launch sequences only: constant Σw, no side stream, no per-launch events): the K timed steps
    and the pass-end exchange are captured once into ONE CUDA graph and the timed region is one replay of it — the
    first launch of a 20-step pass then starts ~10 us after the start event instead of after a Python call, and the
    launches follow each other without the host in between.
    Returns a result dict."""
    torch, dist, ops, dev, world = ctx.torch, ctx.dist, ctx.ops, ctx.dev, ctx.world
    label_dtype = label_dtype or wl.get("label_dtype", "u8")
    B, C, H, W = wl["B"], wl["C"], wl["H"], wl["W"]
    esize = 4 if wl["dtype"] == "f32" else 2
    px_per_gpu = B * H * W
    grad = grad and not metrics_only
    set_bytes = px_per_gpu * C * esize * (2 if grad else 1)
    n_sets = max(3, min(16, -(-400_000_000 // set_bytes)))     # rotate >= 400 MB (L2 is 126 MB) through the caches
    ii = wl["ignore_index"]
    data_dependent_tw = wl["weighted"] or (0 <= ii <= 255) or label_dtype == "i64"
    if tw_mode == "auto":
        tw_mode = "pipe" if label_dtype == "u8" else ("xchg" if world > 1 else "chain")
    if graph and grad and data_dependent_tw and tw_mode == "pipe" and label_dtype == "u8" and world > 1:
        # a replayed pass must end where it began (the last launch pre-publishes the first one's sum to the peers):
        # rotate over a number of sets that divides the pass
        for n in range(n_sets, 9):
            if steps % n == 0:
                n_sets = n
                break
    sets, weight = synth_inputs(torch, wl, dev, seed=1234 + ctx.rank, n_sets=n_sets, label_dtype=label_dtype, layout=layout,
                                block=label_block)
    dl = [torch.empty_like(x) for x, _ in sets] if grad else [None] * n_sets
    am = [torch.empty((B, H, W), dtype=torch.uint8, device=dev) for _ in sets]
    confmat = torch.zeros((C, C), dtype=torch.int64, device=dev)
    loss_out = torch.zeros(1, dtype=torch.float32, device=dev)
    tw_kernel = grad and data_dependent_tw and tw_mode == "kernel" and label_dtype == "u8"
    tw_pipe = grad and data_dependent_tw and tw_mode == "pipe" and label_dtype == "u8"
    prepass_on = grad and data_dependent_tw and not tw_kernel and not tw_pipe
    tw_xchg = prepass_on and tw_mode in ("xchg", "kernel") and world > 1      # K4 locally, exchange inside K1
    xchg = ctx.exchange() if ((tw_kernel or tw_xchg or tw_pipe) and world > 1) else None
    nxt = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in range(2)] if tw_pipe else None
    gstep = {"n": 0}                                             # pipe mode: batches are consumed in one global order
    twg = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in range(n_sets)] if tw_xchg else None
    # K4 pre-pass (Σ v·w[y] must be known before the first dlogit is written) runs ONE STEP AHEAD on its own
    # stream: the labels of the next batch are known while the current K1 runs (as in a training loop with a
    # prefetching loader), so the pre-pass — and at N > 1 its all-reduce — overlaps K1.
    pre = torch.cuda.Stream(device=dev) if prepass_on else None
    tws = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in range(n_sets)]
    t8s = [torch.empty((B, H, W), dtype=torch.uint8, device=dev) for _ in range(n_sets)] if label_dtype == "i64" else None
    tw_sum = [t_[0:1] for t_ in tws]
    tw_inv = [t_[1:2] for t_ in tws]
    pre_ready, k1_done = [None] * n_sets, [None] * n_sets
    issued = {"upto": -1}
    # per-step loss sums f64[3] land in a [rows, 3] table; at N > 1 the table and the C x C matrix travel in ONE
    # all-reduce per pass (counts < 2^53 are exact in f64)
    sums_rows = max(steps, warmup, 1)
    pass_buf = torch.zeros(sums_rows * 3 + C * C, dtype=torch.float64, device=dev)
    sums_table = pass_buf[:sums_rows * 3].view(sums_rows, 3)
    launches = {"n": 0}
    k1_events = []

    def prepass(i):
        j = i % n_sets
        _, t = sets[j]
        if k1_done[j] is not None:
            pre.wait_event(k1_done[j])
        with torch.cuda.stream(pre):
            if world > 1 and not tw_xchg:
                ops.label_hist(t, C, ii, weight=weight, total_weight_out=tws[j])    # this rank's Σ v·w[y] (fp64)
                dist.all_reduce(tw_sum[j])                      # global Σw: every rank divides by the same total
                torch.reciprocal(tw_sum[j], out=tw_inv[j])
            elif t.dtype == torch.int64:
                # the reference's .long() labels: one pass gives Σ v·w[y] and the byte labels K1 then reads
                ops.labels_prepare(t, C, ii, weight, tws[j], t8s[j])
            else:
                ops.label_hist(t, C, ii, weight=weight, total_weight_out=tws[j])
            launches["n"] += 1
            ev = torch.cuda.Event()
            ev.record(pre)
        pre_ready[j] = ev
        issued["upto"] = i

    def step(i, timed, last=False):
        j = i % n_sets
        if tw_pipe:
            g_ = gstep["n"]
            gstep["n"] += 1
            j = g_ % n_sets
        x, t = sets[j]
        inv, inv_dev = 0.0, None
        if grad:
            if tw_kernel or tw_pipe:
                pass
            elif prepass_on:
                if issued["upto"] < i:
                    prepass(i)
                torch.cuda.current_stream(dev).wait_event(pre_ready[j])
                inv_dev = tw_inv[j]
            else:
                inv = 1.0 / float(px_per_gpu * world)           # nothing can be ignored: Σw = global pixel count
        timed = timed and per_launch_events
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        t_k1, ii_k1 = t, ii
        if prepass_on and (world == 1 or tw_xchg) and t.dtype == torch.int64:
            t_k1, ii_k1 = t8s[j], 255                         # byte labels written by cvcs_labels_prepare
        if metrics_only:
            ops.eval_fused(x, t, ii, argmax=am[j], confmat=confmat)
        elif tw_kernel:
            ops.ce_fused(x, t, weight, ii, want_grad=True, total_weight="kernel", xchg=xchg, total_weight_out=tws[j],
                         dlogits=dl[j], argmax=am[j], confmat=confmat, loss_sums=sums_table[i % sums_rows], loss_out=loss_out)
        elif tw_pipe:
            # this launch divides by the sum launch g-1 computed for it (exchanged across ranks inside the kernel) and
            # computes the next batch's sum for launch g+1
            ops.ce_fused(x, t, weight, ii, want_grad=True, total_weight="kernel", xchg=xchg, local_total_weight=nxt[g_ % 2][0:1],
                         total_weight_out=tws[j], next_target=sets[(g_ + 1) % n_sets][1], next_total_weight_out=nxt[(g_ + 1) % 2],
                         dlogits=dl[j], argmax=am[j], confmat=confmat, loss_sums=sums_table[i % sums_rows], loss_out=loss_out)
        elif tw_xchg:
            ops.ce_fused(x, t_k1, weight, ii_k1, want_grad=True, total_weight="kernel", xchg=xchg, local_total_weight=tw_sum[j],
                         total_weight_out=twg[j], dlogits=dl[j], argmax=am[j], confmat=confmat,
                         loss_sums=sums_table[i % sums_rows], loss_out=loss_out)
        else:
            ops.ce_fused(x, t_k1, weight, ii_k1, want_grad=grad, inv_total_weight=inv, inv_total_weight_dev=inv_dev,
                         dlogits=dl[j], argmax=am[j], confmat=confmat, loss_sums=sums_table[i % sums_rows], loss_out=loss_out)
        launches["n"] += 1
        if timed:
            e1.record()
            k1_events.append((e0, e1))
        if prepass_on:
            k1_done[j] = torch.cuda.Event()
            k1_done[j].record()
            if not last:
                prepass(i + 1)

    def pass_end():
        if world > 1:
            pass_buf[sums_rows * 3:].copy_(confmat.view(-1))     # exact: counts << 2^53
            ctx.allreduce_pass_end(pass_buf)                      # every step's loss sums + the C x C matrix, one exchange

    if tw_pipe:
        ops.label_hist(sets[0][1], C, ii, weight=weight, total_weight_out=nxt[0])    # the very first batch: one K4, untimed
    for i in range(warmup):
        step(i, False, last=(i == warmup - 1))
    pass_end()                                                    # warm the pass-end collective up too
    ctx.fence(pre)
    # pipelined Σw is capturable too: the only state a replay needs is the first batch's sum, which an untimed K4 launch
    # supplies before each replay; across GPUs the previous launch has also pre-published its successor's sum to the
    # peers, so the pass must be a whole number of rotations (and of the two-slot next-sum ring) to be replayed
    use_graph = bool(graph) and not prepass_on and not tw_kernel and not per_launch_events \
        and not (tw_pipe and world > 1 and (steps % n_sets or steps % 2)) and not (world > 1 and ctx.args.nccl_pass_end)
    cuda_graph = None
    g0 = gstep["n"]

    def prime():
        if tw_pipe:
            ops.label_hist(sets[g0 % n_sets][1], C, ii, weight=weight, total_weight_out=nxt[g0 % 2])

    graph_note = None
    if use_graph:
        g_before = gstep["n"]
        try:
            cap = torch.cuda.Stream(device=dev)
            cap.wait_stream(torch.cuda.current_stream(dev))
            cuda_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cuda_graph, stream=cap):
                for i in range(steps):
                    step(i, False, last=(i == steps - 1))
                pass_end()
            torch.cuda.current_stream(dev).wait_stream(cap)
        except Exception as e:                                    # never lose the line to a capture problem: plain launches
            cuda_graph = None
            gstep["n"] = g_before
            graph_note = f"graph capture failed ({type(e).__name__}); "
            torch.cuda.synchronize(dev)
        if world > 1:                                             # all ranks replay, or none does
            ok = torch.tensor([1.0 if cuda_graph is not None else 0.0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok.item()) < 1.0:
                cuda_graph = None
                gstep["n"] = g_before
        if cuda_graph is not None:
            prime()
            cuda_graph.replay()                                   # untimed: instantiation / upload costs land here
            ctx.fence(pre)
            prime()
    confmat.zero_()
    launches["n"] = 0
    issued["upto"] = -1
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.fence(pre)                                                # line the ranks up right before the timed region
    if want_clocks:
        ctx.sampler.start()
    ctx.align_start()
    start.record()
    host_t0 = time.perf_counter()
    if cuda_graph is not None:
        cuda_graph.replay()
        launches["n"] = steps + (1 if world > 1 else 0)            # K1 per step (+ the pass-end exchange kernel)
    else:
        for i in range(steps):
            step(i, True, last=(i == steps - 1))
    host_ms = (time.perf_counter() - host_t0) * 1e3 / steps     # enqueue cost; must stay below ms_per_step
    if pre is not None:
        torch.cuda.current_stream(dev).wait_stream(pre)
    if cuda_graph is None:
        pass_end()
    end.record()
    ctx.fence(pre)
    if want_clocks:
        ctx.sampler.stop()
    ms_total = start.elapsed_time(end)
    tw_rel_err = None
    if tw_pipe or tw_kernel or tw_xchg:
        # the divisor the last timed launch used (its total_weight_out) against K4 over the same labels, summed over ranks
        j_last = (gstep["n"] - 1) % n_sets if tw_pipe else (steps - 1) % n_sets
        got = float((twg if tw_xchg else tws)[j_last][0].item())
        exp = torch.zeros(2, dtype=torch.float64, device=dev)
        ops.label_hist(sets[j_last][1], C, ii, weight=weight, total_weight_out=exp)
        if world > 1:
            dist.all_reduce(exp[0:1])
        tw_rel_err = abs(got - float(exp[0].item())) / max(float(exp[0].item()), 1e-30)
    k1_ms = [a.elapsed_time(b) for a, b in k1_events] or [ms_total / steps]
    k1_avg = sum(k1_ms) / len(k1_ms)
    (ms_max,), table = ctx.max_over_ranks(ms_total)
    per_rank = None
    if table is not None:
        _, kt = ctx.max_over_ranks(k1_avg)
        per_rank = {"region_ms": [round(r[0], 3) for r in table], "k1_avg_ms": [round(r[0], 4) for r in kt]}
    ms_per_step = ms_max / steps
    bpp = algorithmic_bytes_per_pixel(C, esize, grad)
    achieved = bpp * px_per_gpu / (k1_avg * 1e-3) / 1e9

    copy_gbs = None
    if grad and copy_ref:
        # same-size copy, same harness (events around every launch, rotating buffers): context for roofline.frac
        cp = []
        for i in range(10 + 50):
            x, _ = sets[i % n_sets]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dl[i % n_sets].copy_(x)
            e1.record()
            if i >= 10:
                cp.append((e0, e1))
        torch.cuda.synchronize(dev)
        cp_ms = sum(a.elapsed_time(b) for a, b in cp) / len(cp)
        copy_gbs = 2 * px_per_gpu * C * esize / (cp_ms * 1e-3) / 1e9

    tkey = name + ("" if grad else ("/metrics" if metrics_only else "/nograd")) + ("/nhwc" if layout == "nhwc" else "")
    src = "CUDA events around every K1 launch" if per_launch_events else \
        "timed region / steps (programmatic dependent launches overlap; per-launch events would serialise them)"
    res = {
        "value": world * px_per_gpu / (ms_per_step * 1e-3) / 1e9, "ms_per_step": ms_per_step, "px_per_gpu": px_per_gpu,
        "roofline": roofline_dict(ctx, achieved, bpp, px_per_gpu, k1_avg, "cvcs K1 ce_fused" if not metrics_only else
                                  "cvcs K1 eval_fused (metrics mode)", tkey, src,
                                  {"same_size_copy_gbs_in_this_harness": copy_gbs}),
        "gpu_launches": launches["n"], "host_enqueue_ms_per_step": host_ms, "per_rank": per_rank,
        "check": {"confusion_total": int(confmat.sum().item()), "loss": float(loss_out.item()), "total_weight_rel_err": tw_rel_err},
        "config": {"per_gpu_batch": B, "classes": C, "tile": [H, W],
                   "labels": label_dtype + (f" (blocky {label_block}x{label_block})" if label_block > 1 else " (i.i.d.)"), "grad": grad,
                   "metrics_only": metrics_only, "layout": layout,
                   "l2": f"inputs larger than L2: {n_sets} rotating sets of {set_bytes / 1e6:.0f} MB",
                   "launch": (f"the {steps} steps" + (" + the pass-end exchange" if world > 1 else "") + " captured in one CUDA graph, replayed once in the timed region")
                   if cuda_graph is not None else (graph_note or "") + "one C-ABI call per step from Python on the current stream",
                   "total_weight": ("computed inside K1 (label pre-pass + grid barrier" + (", exchanged across ranks over NVLink inside the kernel)" if xchg is not None else ")"))
                   if tw_kernel else "pipelined across launches: K1 of step i sums the weights over step i+1's labels in its prologue" + (
                       "; the ranks' sums are exchanged inside K1 over NVLink peer memory" if xchg is not None else "") if tw_pipe
                   else (("K4 pre-pass one step ahead on a side stream" + (", the ranks' sums exchanged inside K1 over NVLink peer memory" if tw_xchg
                                                                                        else (" + NCCL all-reduce" if world > 1 else "")))
                                      if prepass_on else "constant (nothing can be ignored)")},
        "_state": (sets, weight),
    }
    return res


def scene_tiles(scene_hw, p):
    rows, cols = scene_hw // p, scene_hw // p
    return [(r * p, c * p) for r in range(rows) for c in range(cols)]


def measure_chain(ctx, name, wl, steps, warmup, *, per_launch_events=True, want_clocks=True):
    """K5 (tile gather + cast / normalise + label tiles) -> [segmenter: not timed, its logits are synthetic] -> K1, with
    the label tiles flowing from K5 into K1.  Scenes are dealt to the ranks tile by tile (global tile id
    g = scene * tpi + row * cols + col, rank g mod R — dataset.py:137-140 ordering); one C x C all-reduce per pass."""
    torch, dist, ops, dev, world, rank = ctx.torch, ctx.dist, ctx.ops, ctx.dev, ctx.world, ctx.rank
    from cvcs_b200 import shard
    B, C, Cb, p, S = wl["B"], wl["C"], wl["Cb"], wl["H"], wl["scene"]
    n_scenes = wl["scenes_per_gpu"] * world
    esize = 4
    tiles_all = shard.local_tiles(n_scenes, [S, S], p, rank, world, "round_robin")     # (g, scene, tly, tlx)
    batches = [tiles_all[i:i + B] for i in range(0, len(tiles_all), B)]
    # every rank holds every scene (round-robin by tile id touches all of them); seeded per scene, so all ranks agree
    scenes = []
    for s in range(n_scenes):
        g = torch.Generator(device=dev).manual_seed(3 + s)
        img = torch.randint(0, 256, (Cb, S, S), generator=g, device=dev, dtype=torch.uint8)
        lab = blocky_labels(torch, (S, S), C, g, dev)
        scenes.append((img, lab))
    mean = std = None
    if wl["normalise"]:
        mean = torch.arange(Cb, device=dev, dtype=torch.float32) * 7 + 90
        std = torch.arange(Cb, device=dev, dtype=torch.float32) * 3 + 40
    # per batch: tile origins / output slots grouped by scene (device tensors made once)
    plans = []
    for part in batches:
        by_scene = {}
        for slot, (_, s, tly, tlx) in enumerate(part):
            by_scene.setdefault(s, []).append((slot, tly, tlx))
        plan = []
        for s, items in sorted(by_scene.items()):
            yx = torch.tensor([(y, x) for _, y, x in items], dtype=torch.int32, device=dev)
            slots = torch.tensor([sl for sl, _, _ in items], dtype=torch.int32, device=dev)
            plan.append((s, yx, slots))
        plans.append((len(part), plan))
    n_sets = 2 if B * C >= 640 else 3
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    logits = [torch.randn(B, C, p, p, generator=g, device=dev, dtype=torch.float32) * 3 for _ in range(n_sets)]
    dl = [torch.empty_like(x) for x in logits]
    tiles = [torch.empty((B, Cb, p, p), dtype=torch.float32, device=dev) for _ in range(2)]
    labs = [torch.empty((B, p, p), dtype=torch.uint8, device=dev) for _ in range(n_sets)]
    am = torch.empty((B, p, p), dtype=torch.uint8, device=dev)
    confmat = torch.zeros((C, C), dtype=torch.int64, device=dev)
    loss_out = torch.zeros(1, dtype=torch.float32, device=dev)
    sums_rows = max(steps, warmup, 1)
    pass_buf = torch.zeros(sums_rows * 3 + C * C, dtype=torch.float64, device=dev)
    sums_table = pass_buf[:sums_rows * 3].view(sums_rows, 3)
    total_px_global = float(n_scenes * len(scene_tiles(S, p)) * p * p)
    launches = {"n": 0}
    ev_k1, ev_k5 = [], []
    px_done = {"n": 0}

    def step(i, timed):
        nb, plan = plans[i % len(plans)]
        j = i % n_sets
        timed_ev = timed and per_launch_events
        if timed_ev:
            a0, a1, b1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a0.record()
        for s, yx, slots in plan:
            img, lab = scenes[s]
            ops.tile_normalize(img, yx, (p, p), mean, std, label=lab, slots=slots, out=tiles[i % 2], label_out=labs[j])
            launches["n"] += 1
        if timed_ev:
            a1.record()
        # the segmenter would run here on tiles[i % 2]; its output is replaced by synthetic logits of the same shape
        ops.ce_fused(logits[j][:nb], labs[j][:nb], None, wl["ignore_index"], want_grad=True,
                     inv_total_weight=1.0 / total_px_global, dlogits=dl[j][:nb], argmax=am[:nb], confmat=confmat,
                     loss_sums=sums_table[i % sums_rows], loss_out=loss_out)
        launches["n"] += 1
        if timed_ev:
            b1.record()
            ev_k5.append((a0, a1, nb))
            ev_k1.append((a1, b1, nb))
        if timed:
            px_done["n"] += nb * p * p

    def pass_end():
        if world > 1:
            pass_buf[sums_rows * 3:].copy_(confmat.view(-1))
            ctx.allreduce_pass_end(pass_buf)

    for i in range(warmup):
        step(i, False)
    pass_end()
    ctx.fence()
    confmat.zero_()
    launches["n"] = 0
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.fence()
    if want_clocks:
        ctx.sampler.start()
    ctx.align_start()
    start.record()
    host_t0 = time.perf_counter()
    for i in range(steps):
        step(i, True)
    host_ms = (time.perf_counter() - host_t0) * 1e3 / steps
    pass_end()
    end.record()
    ctx.fence()
    if want_clocks:
        ctx.sampler.stop()
    ms_total = start.elapsed_time(end)
    (ms_max, px_sum), table = ctx.max_over_ranks(ms_total, float(px_done["n"]))
    if table is not None:
        px_sum = sum(r[1] for r in table)
    ms_per_step = ms_max / steps
    # roofline of the dominant kernel (K1) over the FULL batches only (ragged last batches are launched, not averaged in)
    full = [(a, b) for a, b, nb in ev_k1 if nb == B]
    bpp = algorithmic_bytes_per_pixel(C, esize, True)
    bpp5 = Cb + Cb * 4 + 2
    if full:
        k1_avg = sum(a.elapsed_time(b) for a, b in full) / len(full)
        k5_full = [(a, b) for a, b, nb in ev_k5 if nb == B]
        k5_avg = sum(a.elapsed_time(b) for a, b in k5_full) / len(k5_full)
        src = "CUDA events around every K1 launch (full batches)"
    else:
        k1_avg = ms_total / steps * (bpp / (bpp + bpp5))
        k5_avg = ms_total / steps - k1_avg
        src = "timed region / steps, split by algorithmic bytes (no per-launch events)"
    achieved = bpp * B * p * p / (k1_avg * 1e-3) / 1e9
    res = {
        "value": px_sum / (ms_max * 1e-3) / 1e9, "ms_per_step": ms_per_step, "px_per_gpu": B * p * p,
        "roofline": roofline_dict(ctx, achieved, bpp, B * p * p, k1_avg, "cvcs K1 ce_fused", name, src,
                                  {"k5": {"kernel": "cvcs K5 tile_normalize", "bytes_per_pixel": bpp5, "avg_ms_per_batch": k5_avg,
                                          "achieved": bpp5 * B * p * p / (k5_avg * 1e-3) / 1e9,
                                          "frac": bpp5 * B * p * p / (k5_avg * 1e-3) / 1e9 / ctx.peak}}),
        "gpu_launches": launches["n"], "host_enqueue_ms_per_step": host_ms,
        "per_rank": None if table is None else {"region_ms": [round(r[0], 3) for r in table]},
        "check": {"confusion_total": int(confmat.sum().item()), "loss": float(loss_out.item())},
        "config": {"per_gpu_batch": B, "classes": C, "bands": Cb, "tile": [p, p], "scene": [Cb, S, S], "scenes": n_scenes,
                   "tiles_per_scene": len(scene_tiles(S, p)), "tiles_this_rank": len(tiles_all), "batches_per_pass": len(plans),
                   "normalise": bool(wl["normalise"]), "sharding": "global tile id mod world (round robin)",
                   "segmenter": "not timed: K1 reads synthetic logits; the label tiles flow K5 -> K1",
                   "l2": f"K1 inputs larger than L2: {n_sets} rotating logit sets of {B * C * p * p * 8 / 1e6:.0f} MB (logits + dlogits)"},
    }
    return res


def measure_tile(ctx, name, wl, steps, warmup, want_clocks=True):
    """K5 alone (tile gather + cast + normalise + label tiles)."""
    torch, ops, dev = ctx.torch, ctx.ops, ctx.dev
    B, Cb, p, S = wl["B"], wl["Cb"], wl["H"], wl["scene"]
    g = torch.Generator(device=dev).manual_seed(7)
    scene = torch.randint(0, 256, (Cb, S, S), generator=g, device=dev, dtype=torch.uint8)
    label = blocky_labels(torch, (S, S), 20, g, dev)
    cols = S // p
    yx = torch.tensor([((i // cols) * p, (i % cols) * p) for i in range(B)], dtype=torch.int32, device=dev)
    normalise = Cb != 3
    mean = (torch.arange(Cb, device=dev, dtype=torch.float32) * 7 + 90) if normalise else None
    std = (torch.arange(Cb, device=dev, dtype=torch.float32) * 3 + 40) if normalise else None
    outs = [torch.empty((B, Cb, p, p), dtype=torch.float32, device=dev) for _ in range(2)]
    labs = [torch.empty((B, p, p), dtype=torch.uint8, device=dev) for _ in range(2)]
    ev = []

    def step(i, timed):
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        ops.tile_normalize(scene, yx, (p, p), mean, std, out=outs[i % 2], label=label, label_out=labs[i % 2])
        if timed:
            e1.record()
            ev.append((e0, e1))

    for i in range(warmup):
        step(i, False)
    ctx.fence()
    if want_clocks:
        ctx.sampler.start()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(steps):
        step(i, True)
    end.record()
    ctx.fence()
    if want_clocks:
        ctx.sampler.stop()
    (ms_max,), _ = ctx.max_over_ranks(start.elapsed_time(end))
    ms = ms_max / steps
    k_ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
    px = B * p * p
    bpp = Cb + Cb * 4 + 2
    achieved = bpp * px / (k_ms * 1e-3) / 1e9
    return {
        "value": ctx.world * px / (ms * 1e-3) / 1e9, "ms_per_step": ms, "px_per_gpu": px,
        "roofline": roofline_dict(ctx, achieved, bpp, px, k_ms, "cvcs K5 tile_normalize", name, "CUDA events around every launch"),
        "gpu_launches": steps, "host_enqueue_ms_per_step": None, "per_rank": None, "check": None,
        "config": {"tiles_per_step": B, "bands": Cb, "tile": [p, p], "scene": [Cb, S, S], "normalise": normalise,
                   "l2": f"outputs larger than L2: 2 rotating sets of {px * Cb * 4 / 1e6:.0f} MB"},
        "_state": (scene, label, mean, std, normalise),
    }


def torch_cuda_baseline(ctx, wl, state, steps=5, warmup=2):
    """What the reference runs for this step with `device: gpu` (utils.py:276, train.py:115-125, utils.py:88-94): stock
    torch CUDA ops — nn.CrossEntropyLoss(weight, ignore_index) fwd + backward on the logits, argmax over the classes,
    confusion matrix as bincount(t*C + p) — on the same GPU and inputs.  Not our kernels; a comparator."""
    torch, dev = ctx.torch, ctx.dev
    (sets, weight) = state
    C, ii = wl["C"], wl["ignore_index"]
    x0, t0 = sets[0]
    t = t0.long()
    w = None if weight is None else weight.to(x0.dtype)
    crit = torch.nn.CrossEntropyLoss(weight=w, ignore_index=ii)
    cm = torch.zeros(C * C, dtype=torch.int64, device=dev)
    times = []
    for i in range(warmup + steps):
        x = sets[i % len(sets)][0].detach().requires_grad_(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = crit(x, t)
        loss.backward()
        pred = x.detach().argmax(1)
        keep = t != ii
        cm += torch.bincount(t[keep] * C + pred[keep], minlength=C * C)
        e1.record()
        torch.cuda.synchronize(dev)
        if i >= warmup:
            times.append(e0.elapsed_time(e1))
        del x, loss, pred, keep
    ms = statistics.median(times)
    px = wl["B"] * wl["H"] * wl["W"]
    return {"value": px / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "ops": "torch.nn.CrossEntropyLoss fwd+bwd (int64 labels) + argmax(1) + bincount(t*C+p), torch " + torch.__version__}


def graph_replay_rate(ctx, wl, state, steps=200):
    """The small reference-shaped step is bound by the host's launch rate, not by the GPU: the same chain
    (cvcs_labels_prepare on the int64 labels -> cvcs_ce_fused) captured ONCE into a CUDA graph — one rotation over the
    input sets — and replayed shows what the device itself needs per step (every entry point takes
    the caller's stream and neither synchronises nor allocates)."""
    torch, ops, dev = ctx.torch, ctx.ops, ctx.dev
    (sets, weight) = state
    C, ii = wl["C"], wl["ignore_index"]
    B, _, H, W = sets[0][0].shape
    n = len(sets)
    dl = [torch.empty_like(x) for x, _ in sets]
    am = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
    t8 = [torch.empty((B, H, W), dtype=torch.uint8, device=dev) for _ in sets]
    tws = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in sets]
    cm = torch.zeros((C, C), dtype=torch.int64, device=dev)
    sums = torch.zeros((n, 3), dtype=torch.float64, device=dev)
    loss = torch.zeros(1, dtype=torch.float32, device=dev)

    def rotation():
        for j, (x, t) in enumerate(sets):
            if t.dtype == torch.int64:
                ops.labels_prepare(t, C, ii, weight, tws[j], t8[j])
                tk, ik = t8[j], 255
            else:
                ops.label_hist(t, C, ii, weight=weight, total_weight_out=tws[j])
                tk, ik = t, ii
            ops.ce_fused(x, tk, weight, ik, want_grad=True, inv_total_weight_dev=tws[j][1:2], dlogits=dl[j], argmax=am,
                         confmat=cm, loss_sums=sums[j], loss_out=loss)

    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        rotation()
    torch.cuda.current_stream(dev).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        rotation()
    replays = max(2, steps // n)
    graph.replay()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        graph.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / (replays * n)
    return {"value": B * H * W / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms, "steps": replays * n, "gpu_launches_per_step": 2,
            "how": f"one CUDA graph of {n} steps (labels_prepare + ce_fused each) replayed {replays} times"}


def measure_e2e(ctx, wl, state, steps, grad=True, return_argmax=False):
    """The same step through the host-buffer C-ABI call: pinned host logits + labels in, loss + C x C (and optionally the
    u8 argmax map) back on the host, every step."""
    torch, ops, dev = ctx.torch, ctx.ops, ctx.dev
    (sets, weight) = state
    x0, t0 = sets[0]
    B, C, H, W = x0.shape
    hx = torch.empty(x0.shape, dtype=x0.dtype, pin_memory=True)
    ht = torch.empty(t0.shape, dtype=t0.dtype, pin_memory=True)
    hx.copy_(x0)
    ht.copy_(t0)
    hw = None if weight is None else weight.cpu()
    hcm = torch.zeros((C, C), dtype=torch.int64)
    ham = torch.empty((B, H, W), dtype=torch.uint8, pin_memory=True) if return_argmax else None
    hctx = ops.HostContext(ctx.local, B * H * W, C, x0.dtype)
    ii = wl["ignore_index"]
    for _ in range(2):
        hctx.ce_fused(hx, ht, hw, ii, want_grad=grad, confmat=hcm, argmax=ham)
    ctx.fence()
    t_0 = time.perf_counter()
    for _ in range(steps):
        hctx.ce_fused(hx, ht, hw, ii, want_grad=grad, confmat=hcm, argmax=ham)   # returns with results on the host
    dt = time.perf_counter() - t_0
    (dt,), _ = ctx.max_over_ranks(dt)
    hctx.close()
    h2d = hx.numel() * hx.element_size() + ht.numel() * ht.element_size() + (0 if hw is None else hw.numel() * 4)
    d2h = C * C * 8 + 3 * 8 * min(B, 64) + (0 if ham is None else ham.numel())
    return {"value": ctx.world * B * H * W * steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "steps": steps, "ms_per_step": dt / steps * 1e3, "h2d_gbs_per_gpu": h2d / (dt / steps) / 1e9,
            "api": "cvcs_host_ce_fused (pinned host logits+labels in; loss + confusion matrix"
                   + (" + u8 argmax map" if return_argmax else "") + " out; dlogits stay on the device for the model backward)"}


def run_secondary(ctx, names, steps, warmup):
    """The other BASELINE configs and variants, measured in the same process (about a second each): value + roofline."""
    torch = ctx.torch
    out = {}
    for nm in names:
        try:
            ctx.set_k1_options()                                   # defaults
            if nm in WORKLOADS:
                wl = dict(WORKLOADS[nm])
                kind = wl["kind"]
                if kind == "ce":
                    r = measure_ce(ctx, nm, wl, steps, warmup, want_clocks=False, per_launch_events=False, graph=bool(ctx.args.graph))
                elif kind == "chain":
                    st = max(4, min(steps, 12)) if nm == "cfg5" else steps
                    r = measure_chain(ctx, nm, wl, st, min(warmup, 3), want_clocks=False)
                else:
                    r = measure_tile(ctx, nm, wl, min(steps, 20), min(warmup, 3), want_clocks=False)
            elif nm == "eval_only":
                r = measure_ce(ctx, "cfg2", dict(WORKLOADS["cfg2"]), steps, warmup, grad=False, want_clocks=False, per_launch_events=False, graph=bool(ctx.args.graph))
            elif nm == "metrics_only":
                r = measure_ce(ctx, "cfg2", dict(WORKLOADS["cfg2"]), steps, warmup, metrics_only=True, want_clocks=False, per_launch_events=False, graph=bool(ctx.args.graph))
            elif nm == "i64_labels":
                r = measure_ce(ctx, "cfg2", dict(WORKLOADS["cfg2"]), steps, warmup, label_dtype="i64", want_clocks=False, per_launch_events=False)
            else:
                continue
            rf = r["roofline"]
            out[nm] = {"value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "n_gpus": ctx.world,
                       "roofline": {"frac": rf["frac"], "achieved": rf["achieved"], "bytes_per_pixel": rf["bytes_per_pixel"],
                                    "avg_launch_ms": rf["avg_launch_ms"], "kernel": rf["kernel"]},
                       "gpu_launches": r["gpu_launches"], "check": r["check"], "per_rank": r.get("per_rank")}
            if "k5" in rf:
                out[nm]["roofline"]["k5"] = rf["k5"]
            if nm == "ref" and ctx.world == 1:
                out[nm]["torch_cuda_baseline"] = torch_cuda_baseline(ctx, WORKLOADS["ref"], r["_state"], steps=20, warmup=5)
                out[nm]["graph_replay"] = graph_replay_rate(ctx, WORKLOADS["ref"], r["_state"])
            del r
        except Exception as e:                                   # a secondary must never take the primary line down
            out[nm] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch override")
    ap.add_argument("--path", default="auto", choices=["auto", "tma", "direct", "generic"])
    ap.add_argument("--stages", type=int, default=0)
    ap.add_argument("--no-wait-hint", action="store_true", help="A/B: producer mbarrier waits without the suspend-time hint")
    ap.add_argument("--reserve-sms", type=int, default=-1,
                    help="SMs K1 leaves free for concurrent collectives (-1: 2 when a per-step all-reduce must overlap K1, else 0)")
    ap.add_argument("--ctas", type=int, default=0, help="A/B: CTAs per SM the TMA variant sizes its stages for")
    ap.add_argument("--vecp", type=int, default=0, help="A/B: pixels per consumer thread per stage of the TMA variant (f32: 2|4, bf16: 4|8)")
    ap.add_argument("--label-dtype", default=None, choices=["u8", "i64"])
    ap.add_argument("--layout", default="nchw", choices=["nchw", "nhwc"], help="logits memory format (nhwc = torch channels_last)")
    ap.add_argument("--no-grad", action="store_true", help="forward/eval only (no dlogits)")
    ap.add_argument("--metrics-only", action="store_true", help="K1 metrics mode: argmax + confusion matrix, no loss (cvcs_eval_fused)")
    ap.add_argument("--pdl", type=int, default=1, choices=[0, 1],
                    help="1 (default, the library's default): K1 launches with programmatic stream serialization (its prologue overlaps "
                         "the previous kernel's tail); per-launch events would serialise the launches, so K1's average launch time "
                         "is then timed region / steps.  0: plain launches with CUDA events around every one")
    ap.add_argument("--graph", type=int, default=1, choices=[0, 1],
                    help="1 (default): a stateless step sequence (constant total weight) is captured in one CUDA graph and the timed "
                         "region replays it; 0: one Python call per step")
    ap.add_argument("--l2-hint", type=int, default=0, help="A/B: K1 L2 eviction hints (CVCS_OPT_L2_HINT: 0 default, v = bit mask v - 1)")
    ap.add_argument("--label-block", type=int, default=32, help="side of the constant label blocks (1 = i.i.d. labels)")
    ap.add_argument("--tw-mode", default="auto", choices=["auto", "pipe", "kernel", "chain", "xchg"],
                    help="data dependent total weight (see measure_ce): auto = pipe for u8 labels")
    ap.add_argument("--nccl-pass-end", action="store_true", help="A/B: all-reduce the pass-end sums with NCCL instead of the one-shot peer-memory exchange")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-copy-ref", action="store_true", help="skip the same-size torch copy reference measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary workloads of the default line")
    ap.add_argument("--secondary", default=None, help="comma-separated secondary workloads (default: all, only with --workload cfg2)")
    ap.add_argument("--no-torch-cuda-baseline", action="store_true")
    ap.add_argument("--cfg1", action="store_true", help="reference arm: also time BASELINE configs[0] (U-Net-style CPU step)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["B"] = args.batch
    if args.impl == "reference":
        if wl["kind"] != "ce":
            raise SystemExit("the reference arm times the loss/metric path (cfg2, cfg3, cfg5head, c16, ref)")
        run_reference_arm(args, wl)
        return

    ctx = Ctx(args)
    torch, dist, world, rank = ctx.torch, ctx.dist, ctx.world, ctx.rank
    kind = wl["kind"]
    grad = not (args.no_grad or args.metrics_only)
    label_dtype = args.label_dtype or wl.get("label_dtype", "u8")
    data_dependent_tw = kind == "ce" and grad and (wl["weighted"] or (0 <= wl["ignore_index"] <= 255) or label_dtype == "i64")
    # a per-step collective (global Σw) has to run WHILE K1 runs: leave it two SMs (K1 claims chunks dynamically)
    reserve = args.reserve_sms if args.reserve_sms >= 0 else (2 if (world > 1 and data_dependent_tw and args.tw_mode == "chain") else 0)
    opts = dict(path=args.path, stages=args.stages, no_wait_hint=args.no_wait_hint, vecp=args.vecp, ctas=args.ctas,
                pdl=args.pdl, reserve=reserve, l2_hint=args.l2_hint)
    ctx.set_k1_options(**opts)
    if kind == "tile":
        ctx.lib.set_option(ctx.lib.OPT_TILE_CTAS, args.ctas)
        ctx.lib.set_option(ctx.lib.OPT_TMA_CTAS, 0)
        res = measure_tile(ctx, args.workload, wl, args.steps, args.warmup)
    elif kind == "chain":
        res = measure_chain(ctx, args.workload, wl, args.steps, args.warmup)
    else:
        res = measure_ce(ctx, args.workload, wl, args.steps, args.warmup, grad=grad, metrics_only=args.metrics_only,
                         label_dtype=label_dtype, layout=args.layout, per_launch_events=not args.pdl,
                         copy_ref=not args.no_copy_ref, tw_mode=args.tw_mode, label_block=args.label_block, graph=bool(args.graph))
    clocks = ctx.sampler.summary()

    e2e = e2e_eval = tcb = cpu = None
    state = res.pop("_state", None)
    if kind == "ce":
        if not args.no_e2e:
            e2e = measure_e2e(ctx, wl, state, args.e2e_steps, grad=grad)
            e2e_eval = measure_e2e(ctx, wl, state, max(3, args.e2e_steps // 2), grad=False, return_argmax=True)
        if world == 1 and not args.no_torch_cuda_baseline and not args.metrics_only:
            try:
                tcb = torch_cuda_baseline(ctx, wl, state)
            except Exception as e:
                tcb = {"error": f"{type(e).__name__}: {e}"[:300]}
    del state
    torch.cuda.empty_cache()

    secondary = None
    want_secondary = (args.workload == "cfg2" and not args.no_secondary and not args.batch and grad
                      and args.layout == "nchw" and label_dtype == "u8") or args.secondary
    if want_secondary:
        names = args.secondary.split(",") if args.secondary else SECONDARY_DEFAULT
        secondary = run_secondary(ctx, names, steps=min(args.steps, 30), warmup=min(args.warmup, 5))
        ctx.set_k1_options(**opts)

    if world == 1 and not args.no_cpu_baseline and kind == "ce":
        rate, sec, sample, cores = cpu_reference_rate(wl, steps=3, warmup=1, budget_s=20.0)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "cpu": cpu_model(), "kind": "port", "sample": sample}

    if rank == 0:
        cfg = {"workload": f"{args.workload}: {wl['desc']}"}
        cfg.update(res["config"])
        cfg.update({"parallelism": (f"dp{world}: tiles sharded per GPU; one all-reduce per pass carrying the [steps,3] f64 loss-sum "
                                    "table and the CxC confusion matrix (" + ("NCCL" if args.nccl_pass_end else
                                    "one-shot exchange over IPC-mapped peer memory, NVLink") + ")") if world > 1 else "single GPU",
                    "path": args.path, "pdl": args.pdl, "l2_hint": args.l2_hint, "sms_reserved_for_collectives": reserve})
        line = {
            "metric": METRIC if kind != "tile" else "Gpixel/s tile gather + cast/normalise (K5)", "value": res["value"], "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": wl["dtype"] if kind != "tile" else "u8->f32", "data": "synthetic", "config": cfg,
            "roofline": res["roofline"], "cpu_baseline": cpu, "torch_cuda_baseline": tcb, "e2e": e2e, "e2e_eval": e2e_eval,
            "gpu_launches": res["gpu_launches"], "host_enqueue_ms_per_step": res["host_enqueue_ms_per_step"],
            "per_rank": res["per_rank"], "clocks": clocks, "check": res["check"], "secondary": secondary,
        }
        print(json.dumps(line), flush=True)
    if getattr(ctx, "_xchg", None) is not None:
        ctx._xchg.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
