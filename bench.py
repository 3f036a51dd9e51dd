#!/usr/bin/env python
"""Benchmark driver contract: ``python bench.py --gpus N --steps K --warmup W [--impl reference]``.

Prints ONE JSON line on rank 0.  Workloads (``--workload``):

  fusion  ProbEn late fusion of saved per-model detections (BASELINE.json configs[0] at scale): one step =
          one ``pe_fuse_batch`` pass over a batch of synthetic RGB+thermal detection pairs.
  pairs   (default once the detector path is built) dual detector -> ProbEn over a batch of image pairs.

The CPU oracle (``oracle/``) is only used for the ``cpu_baseline`` leg and for ``--impl reference``.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ----------------------------------------------------------------------------------------------- utils
class ClockSampler:
    """Samples nvidia-smi SM clocks + throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------- fusion workload (GPU)
def fusion_bytes(packed, n_out):
    """Algorithmic HBM bytes of one pass: (7+K)*4 B per input detection, 24 B per output
    detection, 4 B per offset entry, 4 B per per-image count."""
    N = int(packed["offsets"][-1])
    K = packed["K"]
    return (7 + K) * 4 * N + 24 * int(n_out) + 4 * len(packed["offsets"]) + 4 * packed["B"]


def run_fusion(args):
    import torch
    from probenb200 import fusion, synth
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback exists)")
    torch.cuda.set_device(local_rank)
    dev_name = "cuda:%d" % local_rank
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(dev_name))
    method = (args.score_fusion, args.box_fusion)
    B, M = args.images, args.models
    packed = synth.synth_packed(B, num_models=M, mean_dets=args.mean_dets, seed=1234 + rank)
    N = int(packed["offsets"][-1])
    dev = fusion.to_device(packed, dev_name)
    buf = fusion.FuseBuffers(N, B, dev["boxes"].device)
    # pinned host mirrors for the end-to-end leg
    host_in = {k: torch.from_numpy(packed[k]).pin_memory() for k in ("boxes", "scores", "classes", "probs", "vars", "offsets")}
    e2e_dev = {k: torch.empty_like(v, device=dev_name) for k, v in host_in.items()}
    e2e_dev.update(B=B, M=M, K=packed["K"])
    host_out = {"boxes": torch.empty((N, 4), dtype=torch.float32).pin_memory(), "scores": torch.empty(N).pin_memory(),
                "classes": torch.empty(N, dtype=torch.int32).pin_memory(), "counts": torch.empty(B, dtype=torch.int32).pin_memory()}
    h2d = sum(v.numel() * v.element_size() for v in host_in.values())
    d2h = sum(v.numel() * v.element_size() for v in host_out.values())

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        fusion.fuse_packed(dev, method, buffers=buf)

    def step_e2e():
        for k, v in host_in.items():
            e2e_dev[k].copy_(v, non_blocking=True)
        fusion.fuse_packed(e2e_dev, method, buffers=buf)
        host_out["boxes"].copy_(buf.out_boxes[:N], non_blocking=True)
        host_out["scores"].copy_(buf.out_scores[:N], non_blocking=True)
        host_out["classes"].copy_(buf.out_classes[:N], non_blocking=True)
        host_out["counts"].copy_(buf.out_counts[:B], non_blocking=True)

    def timed(step, steps, warmup):
        for _ in range(warmup):
            step()
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        t0 = time.time()
        ev[0].record()
        for i in range(steps):
            step()
            ev[i + 1].record()
        barrier()
        t1 = time.time()
        ms = ev[0].elapsed_time(ev[steps])
        per = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev_name)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, per, t0, t1

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ms, per, t0, t1 = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    n_out = int(buf.out_counts[:B].clamp(min=0).sum().item())
    ms_e2e, _, _, _ = timed(step_e2e, args.steps, args.warmup)

    if rank != 0:
        return None
    peaks, peak_kind = measured_peaks()
    alg = fusion_bytes(packed, n_out)
    kernel_ms = float(np.mean(per))
    achieved = alg / (kernel_ms * 1e-3) / 1e9
    value = world * B * args.steps / (ms * 1e-3)
    out = {
        "metric": "RGB+thermal image-pairs/sec (ProbEn late-fusion stage on saved detections)",
        "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (iou/log) + f64 (box fusion, borderline iou)", "data": "synthetic",
        "config": {"workload": "proben_fusion: %d pairs/GPU/step, M=%d models, %.1f detections/pair, %s/%s, K=3, 640x512"
                               % (B, M, N / B, method[0], method[1]),
                   "l2": "inputs %.0f MB > 126 MB L2, no flush needed" % (alg / 1e6), "parallelism": "dp%d" % world},
        "e2e": {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": "pairs/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": 2 * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_kind": peak_kind,
                     "kernel": "fuse_warp_kernel<3>", "algorithmic_bytes_per_launch": alg},
    }
    out["cpu_baseline"] = cpu_fusion_baseline(method, M, args.mean_dets, budget_s=args.cpu_seconds, procs=1)
    return out


# -------------------------------------------------------------------------------- CPU legs (oracle port)
def _cpu_fuse_shard(a):
    method, images = a
    from oracle import proben_oracle as O
    t = time.perf_counter()
    for infos in images:
        O.late_fusion_dispatch(method, infos)
    return time.perf_counter() - t


def cpu_fusion_baseline(method, M, mean_dets, budget_s=10.0, procs=1, n_images=None):
    """Times the oracle port of demo_probEn.py's fusion loop on host cores over a bounded sample."""
    from probenb200 import synth
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import proben_cases as pc
    n = n_images or max(512, int(budget_s / 0.35e-3) * max(1, procs) // 4)
    packed = synth.synth_packed(n, num_models=M, mean_dets=mean_dets, seed=99)
    images = pc.packed_to_images(packed)
    if procs <= 1:
        dt = _cpu_fuse_shard((method, images))
    else:
        import multiprocessing as mp
        shards = [images[i::procs] for i in range(procs)]
        with mp.get_context("fork").Pool(procs) as pool:
            t = time.perf_counter()
            pool.map(_cpu_fuse_shard, [(method, s) for s in shards])
            dt = time.perf_counter() - t
    return {"value": n / dt, "unit": "pairs/s", "cores": procs, "kind": "port",
            "sample": "%d synthetic pairs (same generator as the GPU arm), oracle/proben_oracle.py late_fusion_dispatch, "
                      "fusion only (no JSON parse / imread / evaluator)" % n}


def run_reference_fusion(args):
    rank, _, world = dist_env()
    if rank != 0:
        return None
    method = (args.score_fusion, args.box_fusion)
    procs = os.cpu_count() or 1
    vals = []
    n = max(2048, int(args.cpu_seconds / 0.35e-3) * procs // 4)
    for i in range(args.warmup + args.steps):
        r = cpu_fusion_baseline(method, args.models, args.mean_dets, procs=procs, n_images=n)
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    r = vals[-1]
    r["value"] = v
    return {
        "impl": "reference", "metric": "RGB+thermal image-pairs/sec (ProbEn late-fusion stage on saved detections)",
        "value": v, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * n / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "proben_fusion: %d pairs/step (bounded sample), M=%d, %s/%s" % (n, args.models, *method)},
        "cpu_baseline": r, "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="fusion", choices=["fusion"])
    ap.add_argument("--images", type=int, default=1 << 20)
    ap.add_argument("--models", type=int, default=2)
    ap.add_argument("--mean-dets", type=float, default=7.5, dest="mean_dets")
    ap.add_argument("--score_fusion", default="probEn")
    ap.add_argument("--box_fusion", default="v-avg")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, dest="cpu_seconds")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    out = run_reference_fusion(args) if args.impl == "reference" else run_fusion(args)
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
