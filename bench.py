#!/usr/bin/env python
"""bench.py — throughput of the fused segmentation loss + metric path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg5] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (K1: softmax-CE fwd+bwd + argmax + confusion matrix, preceded by
the K4 label pre-pass when class weights / ignore_index make Σw data dependent) over one batch of
synthetic tiles per GPU.  Rank 0 prints ONE JSON line.

  value     whole-job Gpixel/s, inputs resident in HBM, CUDA-event timed, max over ranks
  roofline  the dominant kernel (K1): algorithmic bytes per launch / its average launch duration
            (events around every launch inside the timed region) against MEASURED_PEAKS.json
  e2e       the same metric through the host-buffer C-ABI call (cvcs_host_ce_fused): pinned host
            logits + labels copied in, loss + confusion matrix read back, every step
  cpu_baseline   the reference's own CPU path (oracle/torch_path.py: the torch calls the reference
            makes) timed on this box's host cores on a bounded sample (rank 0, N=1)
  --impl reference   that CPU path as the measured arm (no GPU work at all)
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
    # name: per-GPU batch, classes, H, W, logits dtype, weights, ignore_index, label dtype
    "cfg2": dict(B=16, C=7, H=1024, W=1024, dtype="f32", weighted=False, ignore_index=-100,
                 desc="1xB200: fused softmax-CE fwd+bwd + argmax + confusion matrix, batch 16 of 1024x1024, 7 classes, fp32 logits"),
    "cfg3": dict(B=16, C=7, H=1024, W=1024, dtype="bf16", weighted=True, ignore_index=255,
                 desc="same path with bf16 logits, class weights and ignore_index=255 (LoveDA-style labels)"),
    "cfg5": dict(B=16, C=20, H=1024, W=1024, dtype="f32", weighted=False, ignore_index=-100,
                 desc="20-class head, fp32 logits (the cfg5 head at batch 16 per GPU)"),
    # K5 alone: the tiler / normaliser either side of the model (Cb + Cb*s_out + 2 bytes/px)
    "tile13": dict(kind="tile", B=64, Cb=13, H=1024, W=1024, dtype="f32", scene=8192,
                   desc="cfg5 tiler: 13-band u8 scene -> 64 normalised fp32 tiles of 1024x1024 + label tiles"),
    "tile3": dict(kind="tile", B=64, Cb=3, H=1024, W=1024, dtype="f32", scene=8192,
                  desc="RGB tiler: u8 scene -> 64 fp32 tiles of 1024x1024 (train.py:121 cast) + label tiles"),
}


def algorithmic_bytes_per_pixel(C: int, esize: int, grad: bool = True) -> int:
    """read logits + write dlogits + read label (u8) + write argmax (u8)."""
    return C * esize * (2 if grad else 1) + 2


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(workload: str):
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(workload)
    return None


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (NVML, ~5 ms period)."""

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
            time.sleep(0.004)

    def start(self):
        if self.ok:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        if self._t is not None:
            self._stop.set()
            self._t.join()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def synth_inputs(torch, wl, dev, seed, n_sets):
    """Seeded synthetic logits (randn*3) and blocky labels (32x32 constant blocks, real masks have
    long runs); cfg3 adds 10% ignore pixels and histogram-derived class weights."""
    B, C, H, W = wl["B"], wl["C"], wl["H"], wl["W"]
    dt = torch.float32 if wl["dtype"] == "f32" else torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(seed)
    sets = []
    for _ in range(n_sets):
        x = (torch.randn(B, C, H, W, generator=g, device=dev, dtype=torch.float32) * 3).to(dt)
        t = torch.randint(0, C, (B, H // 32, W // 32), generator=g, device=dev, dtype=torch.uint8)
        t = t.repeat_interleave(32, 1).repeat_interleave(32, 2).contiguous()
        if wl["ignore_index"] == 255:
            t[torch.rand(B, H, W, generator=g, device=dev) < 0.1] = 255
        sets.append((x, t))
    weight = None
    if wl["weighted"]:
        counts = torch.bincount(sets[0][1][sets[0][1] != 255].flatten().long(), minlength=C).float()
        weight = (counts.sum() / (C * counts.clamp(min=1))).to(torch.float32)
    return sets, weight


def cpu_reference_rate(wl, steps, warmup, budget_s, log=None):
    """Times the reference's CPU path (oracle/torch_path.hot_path_step) on a bounded sample of the
    workload.  Returns (Gpixel/s, seconds per step, sample description, cores)."""
    import torch
    from oracle import torch_path
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    C, W = wl["C"], wl["W"]
    g = torch.Generator().manual_seed(0)

    def make(rows, tiles):
        x = torch.randn(tiles, C, rows, W, generator=g) * 3
        if wl["dtype"] == "bf16":
            x = x.to(torch.bfloat16).float()   # torch rejects fp32 weights with bf16 logits: fp32 on the same values
        t = torch.randint(0, C, (tiles, max(rows // 32, 1), W // 32), generator=g, dtype=torch.uint8)
        t = t.repeat_interleave(32, 1).repeat_interleave(32, 2)[:, :rows].contiguous()
        if wl["ignore_index"] == 255:
            t[torch.rand(tiles, rows, W, generator=g) < 0.1] = 255
        w = (torch.rand(C, generator=g) + 0.5) if wl["weighted"] else None
        return x, t, w

    def one(x, t, w):
        t0 = time.perf_counter()
        torch_path.hot_path_step(x, t, w, wl["ignore_index"], C, ignore_background_eval=False)
        return time.perf_counter() - t0

    # calibrate on a 128-row strip, then size the per-step sample to the time budget
    x, t, w = make(128, 1)
    one(x, t, w)
    per_px = min(one(x, t, w) for _ in range(2)) / (128 * W)
    px_budget = budget_s / max(steps + warmup, 1) / per_px
    rows = int(min(wl["H"], max(32, (px_budget // W) // 32 * 32)))
    tiles = int(min(wl["B"], max(1, px_budget // (rows * W)))) if rows == wl["H"] else 1   # up to the full per-GPU batch
    x, t, w = make(rows, tiles)
    for _ in range(warmup):
        one(x, t, w)
    times = [one(x, t, w) for _ in range(steps)]
    sec = sum(times) / len(times)
    px = tiles * rows * W
    sample = f"{tiles} tile(s) of {rows}x{W} px, {C} classes per step ({px} px); {steps} steps after {warmup} warm-up"
    return px / sec / 1e9, sec, sample, cores


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, sec, sample, cores = cpu_reference_rate(wl, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']} — reference CPU path (torch CPU calls as at "
                               "utils.py:230,90,93-94; train.py:122-125) on a bounded sample", "sample": sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_tile_bench(args, wl):
    """Secondary workload: K5 (tile gather + cast + normalise + label tiles) on one GPU."""
    import torch
    from cvcs_b200 import ops
    from oracle import torch_path
    from cvcs_b200 import _lib
    _lib.set_option(_lib.OPT_TILE_CTAS, args.ctas)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, Cb, p, S = wl["B"], wl["Cb"], wl["H"], wl["scene"]
    g = torch.Generator(device=dev).manual_seed(7)
    scene = torch.randint(0, 256, (Cb, S, S), generator=g, device=dev, dtype=torch.uint8)
    label = torch.randint(0, 20, (S // 32, S // 32), generator=g, device=dev, dtype=torch.uint8)
    label = label.repeat_interleave(32, 0).repeat_interleave(32, 1).contiguous()
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

    for i in range(args.warmup):
        step(i, False)
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(0)
    sampler.start()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(args.steps):
        step(i, True)
    end.record()
    torch.cuda.synchronize(dev)
    sampler.stop()
    ms = start.elapsed_time(end) / args.steps
    k_ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
    px = B * p * p
    bpp = Cb + Cb * 4 + 2
    peak, peak_src = load_peaks()
    achieved = bpp * px / (k_ms * 1e-3) / 1e9
    cpu = None
    if not args.no_cpu_baseline:
        hs, hl = scene[:, :2 * p, :2 * p].cpu(), label[:2 * p, :2 * p].cpu()
        torch.set_num_threads(os.cpu_count() or 1)

        def ref_once():
            t0 = time.perf_counter()
            for ty, tx in ((0, 0), (0, p), (p, 0), (p, p)):
                t = torch_path.crop(hs, ty, tx, p, p)
                torch_path.crop(hl[None], ty, tx, p, p)
                t = t.type(torch.float32)
                if normalise:
                    t = (t - mean.cpu()[:, None, None]) / std.cpu()[:, None, None]
            return time.perf_counter() - t0
        ref_once()
        sec = min(ref_once() for _ in range(3))
        cpu = {"value": 4 * p * p / sec / 1e9, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"4 tiles of {p}x{p}, {Cb} bands per step; best of 3"}
    line = {
        "metric": "Gpixel/s tile gather + cast/normalise (K5)", "value": px / (ms * 1e-3) / 1e9, "unit": UNIT,
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8->f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "tiles_per_step": B, "bands": Cb, "tile": [p, p],
                   "scene": [Cb, S, S], "normalise": normalise,
                   "l2": f"outputs larger than L2: 2 rotating sets of {px * Cb * 4 / 1e6:.0f} MB"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": load_traffic(args.workload), "kernel": "cvcs K5 tile_normalize", "bytes_per_pixel": bpp,
                     "pixels_per_launch": px, "avg_launch_ms": k_ms, "peak_source": peak_src},
        "cpu_baseline": cpu, "e2e": None, "gpu_launches": args.steps, "clocks": sampler.summary(),
    }
    print(json.dumps(line), flush=True)


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
    ap.add_argument("--no-wait-hint", action="store_true", help="A/B: mbarrier waits without the suspend-time hint")
    ap.add_argument("--reserve-sms", type=int, default=-1,
                    help="SMs K1 leaves free for concurrent collectives (-1: 2 when a per-step all-reduce must overlap K1, else 0)")
    ap.add_argument("--ctas", type=int, default=0, help="A/B: CTAs per SM the TMA variant sizes its stages for")
    ap.add_argument("--vecp", type=int, default=0, help="A/B: pixels per consumer thread of the TMA variant (f32: 2|4, bf16: 4|8)")
    ap.add_argument("--label-dtype", default="u8", choices=["u8", "i64"])
    ap.add_argument("--layout", default="nchw", choices=["nchw", "nhwc"], help="logits memory format (nhwc = torch channels_last)")
    ap.add_argument("--no-grad", action="store_true", help="forward/eval only (no dlogits)")
    ap.add_argument("--metrics-only", action="store_true", help="K1 metrics mode: argmax + confusion matrix, no loss (cvcs_eval_fused)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-copy-ref", action="store_true", help="skip the same-size torch copy reference measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=10)
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["B"] = args.batch
    if wl.get("kind") == "tile":
        if args.impl == "reference" or int(os.environ.get("WORLD_SIZE", "1")) > 1:
            raise SystemExit("the tile workloads are single-GPU secondary measurements of the b200 arm")
        run_tile_bench(args, wl)
        return
    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    import torch
    import torch.distributed as dist
    from cvcs_b200 import _lib, ops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (cvcs_b200 has no CPU path; use --impl reference for the CPU arm)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    _lib.set_option(_lib.OPT_CE_PATH, {"auto": 0, "tma": 1, "direct": 2, "generic": 3}[args.path])
    _lib.set_option(_lib.OPT_TMA_STAGES, args.stages)
    _lib.set_option(_lib.OPT_TMA_WAIT_HINT, 1 if args.no_wait_hint else 0)
    _lib.set_option(_lib.OPT_TMA_VECP, args.vecp)
    _lib.set_option(_lib.OPT_TMA_CTAS, args.ctas)

    B, C, H, W = wl["B"], wl["C"], wl["H"], wl["W"]
    esize = 4 if wl["dtype"] == "f32" else 2
    px_per_gpu = B * H * W
    grad = not (args.no_grad or args.metrics_only)
    n_sets = 3   # rotate buffer sets; each set (logits + dlogits) is far larger than the 126 MB L2 anyway
    sets, weight = synth_inputs(torch, wl, dev, seed=1234 + rank, n_sets=n_sets)
    if args.label_dtype == "i64":
        sets = [(x, t.long()) for x, t in sets]
    if args.layout == "nhwc":
        sets = [(x.contiguous(memory_format=torch.channels_last), t) for x, t in sets]
    dl = [torch.empty_like(x) for x, _ in sets] if grad else [None] * n_sets
    am = [torch.empty((B, H, W), dtype=torch.uint8, device=dev) for _ in sets]
    confmat = torch.zeros((C, C), dtype=torch.int64, device=dev)
    loss_out = torch.zeros(1, dtype=torch.float32, device=dev)
    sums = torch.zeros(3, dtype=torch.float64, device=dev)
    ii = wl["ignore_index"]
    data_dependent_tw = wl["weighted"] or (0 <= ii <= 255) or args.label_dtype == "i64"
    prepass_on = grad and data_dependent_tw
    # a per-step collective (global Σw) has to run WHILE K1 runs: leave it two SMs (K1 claims chunks dynamically)
    reserve = args.reserve_sms if args.reserve_sms >= 0 else (2 if (world > 1 and prepass_on) else 0)
    _lib.set_option(_lib.OPT_RESERVE_SMS, reserve)
    # K4 pre-pass (Σ v·w[y] must be known before the first dlogit is written) runs ONE STEP AHEAD on its own
    # stream: the labels of the next batch are known while the current K1 runs (as in a training loop with a
    # prefetching loader), so the pre-pass — and at N > 1 its label-histogram all-reduce — overlaps K1.
    pre = torch.cuda.Stream(device=dev) if prepass_on else None
    tws = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in range(n_sets)]
    t8s = [torch.empty((B, H, W), dtype=torch.uint8, device=dev) for _ in range(n_sets)] if args.label_dtype == "i64" else None
    tw_sum = [t_[0:1] for t_ in tws]     # views made once: the per-step Python path is the bottleneck at N > 1
    tw_inv = [t_[1:2] for t_ in tws]
    pre_ready = [None] * n_sets      # event on `pre`: tws[j] holds this step's total weight
    k1_done = [None] * n_sets        # event on the main stream: the K1 that read tws[j] has run
    issued = {"upto": -1}
    # per-step loss sums f64[3] land in a [rows, 3] table; at N > 1 the table is all-reduced ONCE per pass
    # (the global per-step losses are what a training loop logs — nothing on the data path waits for them, and a
    # collective per step would have to squeeze its kernel in between back-to-back persistent K1 launches)
    sums_rows = max(args.steps, args.warmup, 1)
    sums_table = torch.zeros((sums_rows, 3), dtype=torch.float64, device=dev)
    launches = {"n": 0}
    k1_events = []

    def prepass(i):
        j = i % n_sets
        _, t = sets[j]
        if k1_done[j] is not None:
            pre.wait_event(k1_done[j])
        with torch.cuda.stream(pre):
            if world > 1:
                ops.label_hist(t, C, ii, weight=weight, total_weight_out=tws[j])    # this rank's Σ v·w[y] (fp64)
                dist.all_reduce(tw_sum[j])                      # global Σw: every rank divides by the same total
                torch.reciprocal(tw_sum[j], out=tw_inv[j])
                launches["n"] += 1
            elif t.dtype == torch.int64:
                # the reference's .long() labels: one pass gives Σ v·w[y] and the byte labels K1 then reads
                ops.labels_prepare(t, C, ii, weight, tws[j], t8s[j])
                launches["n"] += 1
            else:
                ops.label_hist(t, C, ii, weight=weight, total_weight_out=tws[j])
                launches["n"] += 1
            ev = torch.cuda.Event()
            ev.record(pre)
        pre_ready[j] = ev
        issued["upto"] = i

    def step(i, timed, last=False):
        j = i % n_sets
        x, t = sets[j]
        inv, inv_dev = 0.0, None
        if grad:
            if prepass_on:
                if issued["upto"] < i:
                    prepass(i)
                torch.cuda.current_stream(dev).wait_event(pre_ready[j])
                inv_dev = tw_inv[j]
            else:
                inv = 1.0 / float(px_per_gpu * world)           # nothing can be ignored: Σw = global pixel count
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        t_k1, ii_k1 = t, ii
        if prepass_on and world == 1 and t.dtype == torch.int64:
            t_k1, ii_k1 = t8s[j], 255                         # byte labels written by cvcs_labels_prepare
        if args.metrics_only:
            ops.eval_fused(x, t, ii, argmax=am[j], confmat=confmat)
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
    def fence():
        if pre is not None:
            torch.cuda.current_stream(dev).wait_stream(pre)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step(i, False, last=(i == args.warmup - 1))
    if world > 1:                                               # warm the pass-end collectives up too
        dist.all_reduce(sums_table)
        dist.all_reduce(confmat)
    fence()
    confmat.zero_()
    launches["n"] = 0
    issued["upto"] = -1
    sampler = ClockSampler(local)          # NVML init takes milliseconds and differs per rank ...
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()                                # ... so line the ranks up again right before the timed region
    sampler.start()
    start.record()
    host_t0 = time.perf_counter()
    for i in range(args.steps):
        step(i, True, last=(i == args.steps - 1))
    host_ms_per_step = (time.perf_counter() - host_t0) * 1e3 / args.steps   # enqueue cost; must stay below ms_per_step
    if pre is not None:
        torch.cuda.current_stream(dev).wait_stream(pre)
    if world > 1:
        dist.all_reduce(sums_table)                             # every step's global loss sums, one collective
        dist.all_reduce(confmat)                                # one C*C all-reduce per pass
    end.record()
    fence()
    sampler.stop()
    ms_total = start.elapsed_time(end)
    k1_ms = [a.elapsed_time(b) for a, b in k1_events]
    per_rank = None
    if world > 1:
        # max over ranks of the device-timed region; the per-rank K1 averages say which GPU is the slow one
        mine_t = torch.tensor([ms_total, sum(k1_ms) / max(len(k1_ms), 1)], dtype=torch.float64, device=dev)
        allt = [torch.zeros_like(mine_t) for _ in range(world)]
        dist.all_gather(allt, mine_t)
        ms_total = max(float(t_[0]) for t_ in allt)
        per_rank = {"region_ms": [round(float(t_[0]), 3) for t_ in allt], "k1_avg_ms": [round(float(t_[1]), 4) for t_ in allt]}
    ms_per_step = ms_total / args.steps
    value = world * px_per_gpu / (ms_per_step * 1e-3) / 1e9
    gpu_launches = launches["n"]
    total_cm = int(confmat.sum().item())

    # ---- same-size copy, same harness (events around every launch, rotating buffers): what a plain
    # device-to-device copy of K1's logits -> dlogits bytes reaches here; context for roofline.frac
    copy_gbs = None
    if grad and not args.no_copy_ref:
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

    # ---- roofline of the dominant kernel (K1) -----------------------------------------------------
    peak, peak_src = load_peaks()
    bpp = algorithmic_bytes_per_pixel(C, esize, grad)
    k1_avg_ms = sum(k1_ms) / len(k1_ms)
    achieved = bpp * px_per_gpu / (k1_avg_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": load_traffic(args.workload), "kernel": "cvcs K1 ce_fused", "bytes_per_pixel": bpp,
                "pixels_per_launch": px_per_gpu, "avg_launch_ms": k1_avg_ms, "peak_source": peak_src,
                "frac_of_8TBps_nominal": achieved / 8000.0, "same_size_copy_gbs_in_this_harness": copy_gbs}

    # ---- e2e through the host-buffer C-ABI call ----------------------------------------------------
    e2e = None
    if not args.no_e2e:
        x0, t0 = sets[0]
        hx = torch.empty(x0.shape, dtype=x0.dtype, pin_memory=True)
        ht = torch.empty(t0.shape, dtype=t0.dtype, pin_memory=True)
        hx.copy_(x0)
        ht.copy_(t0)
        hw = None if weight is None else weight.cpu()
        hcm = torch.zeros((C, C), dtype=torch.int64)
        ctx = ops.HostContext(local, px_per_gpu, C, x0.dtype)
        for _ in range(2):
            ctx.ce_fused(hx, ht, hw, ii, want_grad=grad, confmat=hcm)
        fence()
        t_0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            l_e2e, _ = ctx.ce_fused(hx, ht, hw, ii, want_grad=grad, confmat=hcm)   # returns with results on the host
        dt = time.perf_counter() - t_0
        if world > 1:
            tm = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dt = float(tm.item())
        ctx.close()
        e2e = {"value": world * px_per_gpu * args.e2e_steps / dt / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": hx.numel() * hx.element_size() + ht.numel() * ht.element_size()
               + (0 if hw is None else hw.numel() * 4),
               "d2h_bytes_per_step": C * C * 8 + 3 * 8 * min(B, 64),
               "steps": args.e2e_steps, "ms_per_step": dt / args.e2e_steps * 1e3,
               "api": "cvcs_host_ce_fused (pinned host logits+labels in; loss + confusion matrix out; dlogits/argmax stay "
                      "on the device for the model backward)"}

    # ---- CPU baseline (rank 0, N=1 only) -----------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, sec, sample, cores = cpu_reference_rate(wl, steps=3, warmup=1, budget_s=20.0)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": wl["dtype"], "data": "synthetic",
            "config": {"workload": f"{args.workload}: {wl['desc']}", "per_gpu_batch": B, "classes": C,
                       "tile": [H, W], "labels": args.label_dtype + " (blocky 32x32)", "grad": grad, "metrics_only": args.metrics_only, "layout": args.layout,
                       "l2": f"inputs larger than L2: {n_sets} rotating sets of {px_per_gpu * C * esize * (2 if grad else 1) / 1e6:.0f} MB",
                       "parallelism": f"dp{world}: tiles sharded per GPU; one all-reduce of the [steps,3] f64 loss-sum table + "
                                      "one CxC confusion all-reduce per pass" if world > 1 else "single GPU",
                       "k4_prepass": "one step ahead on a side stream (overlaps K1)" if prepass_on else False, "path": args.path,
                       "sms_reserved_for_collectives": reserve},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": gpu_launches, "host_enqueue_ms_per_step": host_ms_per_step, "per_rank": per_rank,
            "clocks": sampler.summary(),
            "check": {"confusion_total": total_cm, "loss": float(loss_out.item())},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
