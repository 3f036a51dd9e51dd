#!/usr/bin/env python
"""Benchmark driver contract: ``python bench.py --gpus N --steps K --warmup W [--impl reference]``.

Prints ONE JSON line on rank 0.  The line's ``metric`` / ``value`` / ``e2e`` / ``roofline`` describe the headline workload
(``--workload pairs`` = BASELINE.json configs[2]: FLIR RGB+thermal dual R50-FPN detector -> ProbEn, batch 16 pairs per GPU);
the default run also appends one compact sub-record per remaining BASELINE config under ``extra``:

  fusion      configs[0] at scale: the ProbEn kernel alone on saved detections (HBM roofline, three detection-count regimes)
  thermal8    configs[1]: thermal_only R50-FPN, batch 8, one GPU
  kaist32     configs[3]: KAIST dual detector (K = 1) -> binary ProbEn, 32 pairs STRONG-scaled over the ranks (32 / N per GPU)
  ensemble64  configs[4]: thermal_only + early_fusion + middle_fusion -> ProbEn over M = 3 models, 64 pairs strong-scaled over 8 GPUs
              (at N < 8 the 8-pair shard one GPU of the 8-GPU job would run)

A step = ``batches_per_step`` back-to-back passes of the hot path over one batch (the batch of the config); that multiplier
is chosen so that the timed region of K steps lasts >= ~3 s (thermal / power steady state) and is stated in ``config``.
Every step is ONE CUDA-graph submission per batch (``ProbEnPipeline.capture``); at N > 1 the single all-gather of a batch's
detections runs on a side stream while the next batch computes (``pipeline.AsyncGather``).

The CPU oracle (``oracle/``) is only used for the ``cpu_baseline`` leg (rank 0, N = 1) and for ``--impl reference``.
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
        sm, pw, mx, reasons = [], [], None, set()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1]); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_median": float(np.median(pw)) if pw else None}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class Dist:
    """Process-group plumbing shared by all workloads of one bench run."""

    def __init__(self):
        import torch
        self.rank, self.local_rank, self.world = dist_env()
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device (no CPU fallback exists)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda:%d" % self.local_rank)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        import torch
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        import torch
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms


# ------------------------------------------------------------------------------- fusion workload (GPU)
def fusion_bytes(packed, n_out):
    """Algorithmic HBM bytes of one pass: (7+K)*4 B per input detection, 24 B per output
    detection, 4 B per offset entry, 4 B per per-image count."""
    N = int(packed["offsets"][-1])
    K = packed["K"]
    return (7 + K) * 4 * N + 24 * int(n_out) + 4 * len(packed["offsets"]) + 4 * packed["B"]


def fusion_regime(D, images, models, mean_dets, method, steps, warmup, e2e=True, force_count=None):
    """One ProbEn-kernel measurement: device-resident pass, end-to-end pass (pinned host in / out), HBM roofline."""
    import torch
    from probenb200 import fusion, synth
    packed = synth.synth_packed(images, num_models=models, mean_dets=mean_dets, seed=1234 + D.rank) if force_count is None else \
        synth.synth_packed(images, num_models=models, seed=1234 + D.rank, force_count=force_count)
    B, N = images, int(packed["offsets"][-1])
    dev = fusion.to_device(packed, str(D.dev))
    buf = fusion.FuseBuffers(N, B, dev["boxes"].device)

    def timed(step):
        for _ in range(warmup):
            step()
        D.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        t0 = time.time()
        ev[0].record()
        for i in range(steps):
            step()
            ev[i + 1].record()
        D.barrier()
        t1 = time.time()
        per = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
        return D.max_over_ranks(ev[0].elapsed_time(ev[steps])), per, t0, t1

    ms, per, t0, t1 = timed(lambda: fusion.fuse_packed(dev, method, buffers=buf))
    n_out = int(buf.out_counts[:B].clamp(min=0).sum().item())
    alg = fusion_bytes(packed, n_out)
    peaks, peak_kind = measured_peaks()
    achieved = alg / (float(np.mean(per)) * 1e-3) / 1e9
    rec = {"value": D.world * B * steps / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms / steps, "pairs_per_step_per_gpu": B,
           "detections_per_pair": N / B, "models": models,
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                        "traffic": None, "peak_kind": peak_kind, "algorithmic_bytes_per_launch": alg,
                        "kernel": "fuse_block_kernel" if N / B > 256 else ("fuse_mid_kernel" if N / B > 32 else "fuse_packed_kernel"),
                        "note": ("%.0f detections per pair: n/4 = %.0f pair-flops per byte, the ALU-bound regime of SURVEY.md §8d (> 40 "
                                 "detections); the HBM fraction is reported for completeness" % (N / B, N / B / 4)) if N / B > 40 else
                                "HBM-bound regime by arithmetic intensity (SURVEY.md §8d); measured instruction-issue bound"}}
    if N / B > 40:  # the ALU-bound regime: IoU decisions per second (every unordered pair once)
        rec["pair_evaluations_per_s"] = rec["value"] * (N / B) * (N / B - 1) / 2
    if e2e:
        host_in = {k: torch.from_numpy(packed[k]).pin_memory() for k in ("boxes", "scores", "classes", "probs", "vars", "offsets")}
        e2e_dev = {k: torch.empty_like(v, device=D.dev) for k, v in host_in.items()}
        e2e_dev.update(B=B, M=models, K=packed["K"])
        host_out = {"boxes": torch.empty((N, 4), dtype=torch.float32).pin_memory(), "scores": torch.empty(N).pin_memory(),
                    "classes": torch.empty(N, dtype=torch.int32).pin_memory(), "counts": torch.empty(B, dtype=torch.int32).pin_memory()}

        def step_e2e():
            for k, v in host_in.items():
                e2e_dev[k].copy_(v, non_blocking=True)
            fusion.fuse_packed(e2e_dev, method, buffers=buf)
            host_out["boxes"].copy_(buf.out_boxes[:N], non_blocking=True)
            host_out["scores"].copy_(buf.out_scores[:N], non_blocking=True)
            host_out["classes"].copy_(buf.out_classes[:N], non_blocking=True)
            host_out["counts"].copy_(buf.out_counts[:B], non_blocking=True)

        ms_e2e, _, _, _ = timed(step_e2e)
        rec["e2e"] = {"value": D.world * B * steps / (ms_e2e * 1e-3), "unit": "pairs/s",
                      "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in host_in.values()),
                      "d2h_bytes_per_step": sum(v.numel() * v.element_size() for v in host_out.values())}
    rec["_window"] = (t0, t1)
    return rec


def run_fusion(args, D):
    method = (args.score_fusion, args.box_fusion)
    sampler = ClockSampler(D.local_rank)
    if D.rank == 0:
        sampler.start()
        time.sleep(0.3)
    rec = fusion_regime(D, args.images, args.models, args.mean_dets, method, args.steps, args.warmup)
    t0, t1 = rec.pop("_window")
    clocks = sampler.stop(t0, t1) if D.rank == 0 else None
    if D.rank != 0:
        return None
    out = {
        "metric": "RGB+thermal image-pairs/sec (ProbEn late-fusion stage on saved detections)",
        "value": rec["value"], "unit": "pairs/s", "n_gpus": D.world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (iou/log) + f64 (box fusion, borderline iou)", "data": "synthetic",
        "config": {"workload": "proben_fusion: %d pairs/GPU/step, M=%d models, %.1f detections/pair, %s/%s, K=3, 640x512"
                               % (args.images, args.models, rec["detections_per_pair"], method[0], method[1]),
                   "l2": "inputs %.0f MB > 126 MB L2, no flush needed" % (rec["roofline"]["algorithmic_bytes_per_launch"] / 1e6),
                   "parallelism": "dp%d" % D.world},
        "e2e": rec["e2e"], "gpu_launches": 2 * args.steps, "clocks": clocks, "roofline": rec["roofline"],
    }
    if D.world == 1 and args.cpu_seconds > 0:
        out["cpu_baseline"] = cpu_fusion_baseline(method, args.models, args.mean_dets, budget_s=args.cpu_seconds, procs=1)
    else:
        out["cpu_baseline"] = {"value": None, "unit": "pairs/s", "cores": 0, "kind": "port", "sample": "N = 1 only"}
    return out


# ------------------------------------------------------------------------------- detector workloads (GPU)
def detector_gflop(depth=50, H=800, W=1024, K=3, cin=3, props=1000, fc=256, backbone_passes=1):
    """Algorithmic FLOPs (2 x MAC) of one detector forward, the BASELINE.md §3 accounting (middle fusion: two backbone
    passes, 512-channel RPN conv and fc1)."""
    blocks = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}[depth]
    mac = (H // 2) * (W // 2) * 64 * 49 * cin
    c_in, h, w = 64, H // 4, W // 4
    for s, nb in enumerate(blocks):
        mid, cout = 64 << s, 256 << s
        for b in range(nb):
            if b == 0 and s > 0:
                ho, wo = h // 2, w // 2
            else:
                ho, wo = h, w
            if b == 0:
                mac += ho * wo * cout * c_in
            mac += ho * wo * mid * c_in + ho * wo * mid * mid * 9 + ho * wo * cout * mid
            c_in, h, w = cout, ho, wo
    conv = mac
    for l, c in zip((2, 3, 4, 5), (256, 512, 1024, 2048)):
        hh, ww = H >> l, W >> l
        conv += hh * ww * 256 * c + hh * ww * 256 * 256 * 9
    conv *= backbone_passes
    for l in (2, 3, 4, 5, 6):
        hh, ww = (H >> l, W >> l) if l < 6 else (((H >> 5) - 1) // 2 + 1, ((W >> 5) - 1) // 2 + 1)
        conv += hh * ww * (fc * fc * 9 + fc * 15)
    head = props * (49 * fc * 1024 + 1024 * 1024 + 1024 * (K + 1 + 4 * K + 1))
    return 2e-9 * conv, 2e-9 * head


def synth_frames(B, seed, channels=(3, 3)):
    """Synthetic 640x512 frames per model: RGB = U{0..255}^3, thermal = one plane replicated x3 (what cv2.imread returns
    for the 8-bit thermal JPEGs), early fusion = BGR+T (4 ch), middle fusion = BGR+TTT (6 ch)."""
    rng = np.random.default_rng(seed)
    rgb = rng.integers(0, 256, size=(B, 512, 640, 3), dtype=np.uint8)
    t1 = rng.integers(0, 256, size=(B, 512, 640, 1), dtype=np.uint8)
    th = np.repeat(t1, 3, axis=3)
    out = []
    for i, c in enumerate(channels):
        if c == 3:
            out.append(rgb if (i == 0 and len(channels) == 2) else th)
        elif c == 4:
            out.append(np.concatenate([rgb, t1], axis=3))
        else:
            out.append(np.concatenate([rgb, th], axis=3))
    return out


WORKLOADS = {
    # name: models = (fusion_method, input channels, middle_fusion), K, per-GPU batch rule
    "pairs": dict(config=2, label="FLIR RGB+thermal dual detector -> ProbEn", models=[("rgb_only", 3, False), ("thermal_only", 3, False)],
                  K=3, scaling="weak", batch=16, unit="pairs/s"),
    "thermal8": dict(config=1, label="FLIR thermal_only detector (no fusion)", models=[("thermal_only", 3, False)], K=3, scaling="weak",
                     batch=8, unit="frames/s"),
    "kaist32": dict(config=3, label="KAIST RGB+thermal dual detector (K = 1) -> binary ProbEn", models=[("rgb_only", 3, False), ("thermal_only", 3, False)],
                    K=1, scaling="strong", global_batch=32, unit="pairs/s"),
    "ensemble64": dict(config=4, label="FLIR 3-model ensemble (thermal_only + early_fusion + middle_fusion) -> ProbEn",
                       models=[("thermal_only", 3, False), ("early_fusion", 4, False), ("middle_fusion", 6, True)], K=3, scaling="strong",
                       global_batch=64, full_world=8, unit="pairs/s"),
}


def run_detector_workload(name, args, D, steps, warmup, target_s, instrument, sample_clocks):
    """Builds the M detectors + ProbEn pipeline of a workload, captures one batch as a CUDA graph and times it.
    Returns the record dict (rank 0) or None."""
    import torch
    from probenb200 import detector, pipeline, weights
    spec = WORKLOADS[name]
    K, depth = spec["K"], args.depth
    method = (args.score_fusion, args.box_fusion)
    M = len(spec["models"])
    if spec["scaling"] == "weak":
        B = args.batch if (name == "pairs" and args.batch) else spec["batch"]
        shard_note = "%d per GPU" % B
    else:
        world_cfg = spec.get("full_world", D.world) if D.world < spec.get("full_world", 1) else D.world
        if spec["global_batch"] % world_cfg:
            return {"skipped": "global batch %d does not divide over %d ranks" % (spec["global_batch"], world_cfg)} if D.rank == 0 else None
        B = spec["global_batch"] // world_cfg
        shard_note = "%d global = %d per GPU over %d GPUs" % (spec["global_batch"], B, world_cfg) + \
                     ("" if world_cfg == D.world else " (this run times %d such shard(s))" % D.world)
    nh, nw = detector.resize_shortest_edge_shape(512, 640)
    canvas = ((nh + 31) // 32 * 32, (nw + 31) // 32 * 32)
    dets = []
    for m, (fm, cin, mid) in enumerate(spec["models"]):
        cfg = detector.fusion_method_config(fm)
        sd = weights.random_state_dict(depth, 3 if mid else cin, K, seed=11 + m, middle_fusion=mid, head_gain=3.0 if K == 1 else 1.0)
        dets.append(detector.Detector(sd, depth=depth, num_classes=K, max_batch=B, canvas=canvas, device=D.dev, **cfg))
    pipe = pipeline.ProbEnPipeline(dets, method, frame_size=(512, 640))
    flat = pipe.out.flat
    frames = synth_frames(B, 777 + D.rank, channels=[c for _, c, _ in spec["models"]])
    host = [torch.from_numpy(f).pin_memory() for f in frames]
    slots = [[h.to(D.dev) for h in host] for _ in range(2)]      # two device-side frame buffers (upload of batch i+1 overlaps batch i)
    graphs = None
    if not args.no_graph:
        graphs = [pipe.capture(s, net_hw=(nh, nw)) for s in slots]
    gather = pipeline.AsyncGather(flat) if D.world > 1 else None
    host_out = torch.empty(flat.numel() * D.world, dtype=torch.int32).pin_memory()
    main = torch.cuda.current_stream(D.dev)

    def compute(slot):
        if graphs is not None:
            graphs[slot].replay()
        else:
            pipe.forward_device(slots[slot], net_hw=(nh, nw))

    def batch_device(i):
        compute(0)
        if gather is not None:
            gather.submit(flat)

    # end-to-end: every batch uploads its own frames from pinned host memory (copy stream, double buffered) and
    # downloads its results (gathered over the ranks at N > 1); all copies are inside the timed region
    copy_stream = torch.cuda.Stream(device=D.dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    state = {"i": 0, "primed": False}

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            for m in range(M):
                slots[slot][m].copy_(host[m], non_blocking=True)
            ready[slot].record(copy_stream)

    def batch_e2e(i):
        if not state["primed"]:
            for sl in range(2):
                consumed[sl].record(main)
            upload(0)
            state["primed"] = True
        slot = state["i"] & 1
        upload(slot ^ 1)
        main.wait_event(ready[slot])
        compute(slot)
        consumed[slot].record(main)
        if gather is not None:
            gather.submit(flat, host_out=host_out)
        else:
            host_out.copy_(flat, non_blocking=True)
        state["i"] += 1

    def timed(batch_fn, bps):
        for _ in range(warmup):
            batch_fn(0)
        if gather is not None:
            gather.finish()
        D.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(steps * bps):
            batch_fn(0)
        if gather is not None:
            gather.finish()
        e1.record()
        D.barrier()
        t1 = time.time()
        return D.max_over_ranks(e0.elapsed_time(e1)), t0, t1

    # batches per step: the timed region should last >= target_s
    for _ in range(max(3, warmup)):
        batch_device(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        batch_device(0)
    e1.record()
    torch.cuda.synchronize()
    est_ms = D.max_over_ranks(e0.elapsed_time(e1) / 3)
    bps = args.batches_per_step or max(1, int(np.ceil(target_s * 1e3 / (steps * est_ms))))
    sampler = ClockSampler(D.local_rank)
    if D.rank == 0 and sample_clocks:
        sampler.start()
        time.sleep(0.3)
    ms, t0, t1 = timed(batch_device, bps)
    clocks = sampler.stop(t0, t1) if (D.rank == 0 and sample_clocks) else None
    ms_e2e, _, _ = timed(batch_e2e, bps)
    launches = pipe.launches_per_step()
    n_batches = steps * bps
    rec = {"value": D.world * B * n_batches / (ms * 1e-3), "unit": spec["unit"], "ms_per_step": ms / steps, "ms_per_batch": ms / n_batches,
           "batch_per_gpu": B, "batches_per_step": bps, "timed_region_s": ms * 1e-3, "scaling": spec["scaling"], "shard": shard_note,
           "e2e": {"value": D.world * B * n_batches / (ms_e2e * 1e-3), "unit": spec["unit"],
                   "h2d_bytes_per_step": int(sum(h.numel() for h in host)) * bps, "d2h_bytes_per_step": int(host_out.numel() * 4) * bps},
           "kernel_launches_per_batch": launches, "graph_launches_per_batch": 0 if graphs is None else 1,
           "clocks": clocks, "models": [m[0] for m in spec["models"]], "K": K, "depth": depth,
           "canvas": list(canvas), "net_hw": [nh, nw]}
    if instrument:
        # instrumented passes: device time of the tensor-core GEMM launches, detectors back to back on one stream, on the now
        # warm (power-steady) GPU.  Mode 1 (the roofline number and the per-layer table): one event pair per launch = the launch
        # durations.  Mode 2 (reported next to it): one pair per RUN of back-to-back GEMM launches, i.e. durations + the gaps
        # between consecutive layers with PDL active - what the GEMM chain costs inside a real forward.
        saved = pipe.streams
        pipe.streams = None
        rec["_profile"], rec["_profile_runs"] = [], []
        for mode, key in ((2, "_profile_runs"), (1, "_profile")):
            for d in dets:
                d.set_profiling(mode)
            for i in range(args.profile_passes + 1):
                pipe.forward_device(slots[0], net_hw=(nh, nw))
                torch.cuda.synchronize()
                if i:
                    rec[key].append(([d.last_profile() for d in dets], [d.profile_launches() for d in dets]))
                    if mode == 1:
                        rec.setdefault("_profile_kernels", []).append([d.profile_kernels() for d in dets])
        for d in dets:
            d.set_profiling(False)
        pipe.streams = saved
        torch.cuda.synchronize()
    counts = [int(pipe.dets[m].counts[:B].sum().item()) for m in range(M)]
    rec["detections_per_image"] = [c / B for c in counts]
    rec["fused_per_item"] = int(pipe.out.counts[:B].sum().item()) / B
    del graphs
    return rec if D.rank == 0 else None


def conv_roofline(rec, args, n_models_flops):
    """Tensor roofline of the conv/GEMM kernel family from the instrumented passes of ``rec``."""
    passes = rec.pop("_profile")
    run_passes = rec.pop("_profile_runs")
    peaks, peak_kind = measured_peaks()
    peak_tf = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    gemm_ms = float(np.mean([sum(p[0] for p in prof) for prof, _ in passes]))            # sum of the launch durations
    gemm_ms_runs = float(np.mean([sum(p[0] for p in prof) for prof, _ in run_passes]))   # ... plus the gaps inside GEMM runs
    span_ms = float(np.mean([sum(p[1] for p in prof) for prof, _ in passes]))
    n_gemm = sum(p[3] for p in passes[0][0])
    l_ms = np.mean([np.concatenate([l[0] for l in layers]) for _, layers in passes], axis=0)
    l_fl = np.concatenate([l[1] for l in passes[0][1]])
    l_by = np.concatenate([l[2] for l in passes[0][1]])
    flop_batch = n_models_flops * 1e9
    achieved = flop_batch / (gemm_ms * 1e-3) / 1e12
    t_tensor = l_fl / (peak_tf * 1e12) * 1e3
    t_hbm = l_by / (peaks["hbm_gbs"] * 1e9) * 1e3
    bound_ms = np.maximum(t_tensor, t_hbm)
    if args.dump_launches:
        with open(args.dump_launches, "w") as f:
            f.write("idx,ms,gflop,mbytes,tensor_bound_ms,hbm_bound_ms,frac_of_bound,lost_ms\n")
            for i in range(len(l_ms)):
                f.write("%d,%.5f,%.3f,%.3f,%.5f,%.5f,%.3f,%.5f\n" % (i, l_ms[i], l_fl[i] / 1e9, l_by[i] / 1e6, t_tensor[i], t_hbm[i],
                                                                  bound_ms[i] / max(1e-9, l_ms[i]), l_ms[i] - bound_ms[i]))
    hb = t_hbm > t_tensor
    per_launch = {"frac": float(bound_ms.sum() / max(1e-9, l_ms.sum())), "bound_ms_per_batch": float(bound_ms.sum()),
                  "measured_ms_per_batch": float(l_ms.sum()), "hbm_bound_launches": int(hb.sum()), "tensor_bound_launches": int((~hb).sum()),
                  "frac_hbm_bound": float(t_hbm[hb].sum() / max(1e-9, l_ms[hb].sum())),
                  "frac_tensor_bound": float(t_tensor[~hb].sum() / max(1e-9, l_ms[~hb].sum())),
                  "algorithmic_gb_per_batch": float(l_by.sum() / 1e9)}
    # DRAM traffic + tensor-pipe activity of the same launches from the committed ncu --set full capture of one detector forward
    # at batch 16 (tools/ncu_conv_all.sh); null when the capture is not in the tree or the shape differs
    traffic, traffic_src, tensor_pipe = None, None, None
    cap = os.path.join(ROOT, "profiles", "r02_conv_all_layers_b16_final2_ncu_summary.csv")
    if os.path.isfile(cap) and rec["depth"] == 50 and rec["batch_per_gpu"] == 16 and len(rec["models"]) == 2:
        import csv
        rows = list(csv.DictReader(open(cap)))
        mb = sum(float(r["dram_read[Mbyte]"]) + float(r["dram_write[Mbyte]"]) for r in rows)
        us = np.array([float(r["us[us]"]) for r in rows])
        tp = np.array([float(r["tensor_pipe_active_pct[%]"]) for r in rows])
        traffic = mb * 1e6 * 2
        tensor_pipe = float((us * tp).sum() / us.sum())
        traffic_src = "profiles/r02_conv_all_layers_b16_final2_ncu_summary.csv (%d GEMM launches of one detector at batch 16) x 2 detectors" % len(rows)
    return {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
            "traffic": traffic, "traffic_unit": "DRAM bytes per batch over the GEMM launches", "traffic_source": traffic_src,
            "tensor_pipe_active_pct_time_weighted": tensor_pipe,
            "peak_kind": peak_kind + " (sustained cuBLAS bf16)", "kernel": "conv_gemm_kernel<*> (tcgen05)",
            "algorithmic_gflop_per_batch": flop_batch / 1e9, "gemm_ms_per_batch": gemm_ms,
            "gemm_ms_method": "sum of the per-launch durations (one CUDA-event pair per launch, serial instrumented passes)",
            "gemm_ms_per_batch_incl_gaps": gemm_ms_runs,
            "gemm_ms_incl_gaps_method": "one CUDA-event pair per RUN of back-to-back GEMM launches (%d runs per batch): launch gaps "
                                        "between consecutive layers included, PDL overlap active" % sum(len(l[0]) for l in run_passes[0][1]),
            "frac_incl_gaps": flop_batch / (gemm_ms_runs * 1e-3) / 1e12 / peak_tf,
            "gemm_launches_per_batch": n_gemm,
            "gemm_share_of_serial_batch": gemm_ms / max(1e-9, span_ms), "profile_passes": len(passes),
            "batch_frac_of_peak": flop_batch / (rec["ms_per_batch"] * 1e-3) / 1e12 / peak_tf,
            "per_launch_roofline": per_launch}


def secondary_kernels(rec):
    """Device time of the non-GEMM launch groups (event pair per group in the instrumented passes, both detectors summed) with the
    HBM roofline of the ones that stream: algorithmic bytes = every input and output once."""
    passes = rec.pop("_profile_kernels", None)
    if not passes:
        return None
    peaks, peak_kind = measured_peaks()
    B, (ch, cw), (nh, nw), n_models = rec["batch_per_gpu"], rec["canvas"], rec["net_hw"], len(rec["models"])
    tot = {}
    for p in passes:
        for det in p:
            for name, ms in det:
                tot[name] = tot.get(name, 0.0) + ms / len(passes)
    lv = [(ch >> s, cw >> s) for s in (2, 3, 4, 5)]
    feat_b = sum(h * w for h, w in lv) * 256 * 2
    algo = {  # bytes per batch and detector
        "launch_stem_im2col_u8": B * (512 * 640 * 3 + (ch + 6) * (cw + 8) * 8),
        "launch_maxpool": B * ((ch // 2) * (cw // 2) + (ch // 4) * (cw // 4)) * 64 * 2,
        "launch_roi_align": B * (1000 * 49 * 256 * 2 + feat_b),
        "launch_rpn_proposals": B * (sum(h * w for h, w in lv) + (ch >> 6) * (cw >> 6)) * 16 * 4,
    }
    label = {"launch_stem_im2col_u8": "canvas staging (Pillow-exact resize + normalise, fused)", "launch_maxpool": "stem max-pool 3x3/2",
             "launch_roi_align": "ROIAlign 7x7, 4 levels, 1000 ROIs per image", "launch_rpn_proposals": "RPN decode + top-k + NMS + merge (4 kernels)",
             "launch_head_post": "softmax + decode + per-class NMS + top-100", "launch_subsample2": "p6 = p5[::2, ::2]",
             "launch_concat_channels": "middle-fusion channel concat"}
    out = {}
    for name, ms in sorted(tot.items(), key=lambda kv: -kv[1]):
        e = {"what": label.get(name, name), "ms_per_batch": ms, "share_of_serial_batch": None}
        if name in algo:
            by = algo[name] * n_models
            e["algorithmic_mb_per_batch"] = by / 1e6
            e["hbm_frac"] = by / (ms * 1e-3) / (peaks["hbm_gbs"] * 1e9)
        else:
            e["bound"] = "latency (one block per image / level)"
        out[name.replace("launch_", "")] = e
    return out


def workload_flops(name, depth, canvas, B):
    spec = WORKLOADS[name]
    tot = 0.0
    for fm, cin, mid in spec["models"]:
        c, h = detector_gflop(depth, canvas[0], canvas[1], spec["K"], cin=3 if mid else cin, fc=512 if mid else 256,
                              backbone_passes=2 if mid else 1)
        tot += c + h
    return tot * B


def run_pairs(args, D):
    name = args.workload
    rec = run_detector_workload(name, args, D, args.steps, args.warmup, target_s=args.target_seconds, instrument=True, sample_clocks=True)
    extra = {}
    if name == "pairs" and not args.no_extras:
        small = dict(steps=max(5, args.steps // 2), warmup=3, target_s=1.0, instrument=False, sample_clocks=False)
        if D.world == 1:
            r = run_detector_workload("thermal8", args, D, **small)
            if D.rank == 0:
                extra["thermal8"] = r
        for wl in ("kaist32", "ensemble64"):
            r = run_detector_workload(wl, args, D, **small)
            if D.rank == 0:
                extra[wl] = r
        method = (args.score_fusion, args.box_fusion)
        fus = {}
        for tag, md, force in (("5_dets_per_model", 5.0, None), ("7.5_dets_per_model", 7.5, None), ("100_dets_per_model_mid_kernel", None, 100)):
            n_img = (1 << 18) if force is None else (1 << 13)
            r = fusion_regime(D, n_img, 2, md, method, 10, 3, e2e=False, force_count=force)
            r.pop("_window", None)
            fus[tag] = r
        if D.rank == 0:
            extra["fusion"] = fus
    if D.rank != 0:
        return None
    spec = WORKLOADS[name]
    B = rec["batch_per_gpu"]
    sec = secondary_kernels(rec)
    roof = conv_roofline(rec, args, workload_flops(name, rec["depth"], rec["canvas"], B))
    if sec:
        span = roof["gemm_ms_per_batch"] / max(1e-9, roof["gemm_share_of_serial_batch"])
        for e in sec.values():
            e["share_of_serial_batch"] = e["ms_per_batch"] / span
        extra["secondary_kernels"] = sec
    for k, r in extra.items():
        if isinstance(r, dict):
            r.pop("_profile", None)
    out = {
        "metric": "RGB+thermal image-pairs/sec end-to-end (dual Faster R-CNN R%d-FPN -> ProbEn)" % rec["depth"] if name == "pairs" else
                  "%s, R%d-FPN: %s end-to-end" % (spec["label"], rec["depth"], spec["unit"]),
        "value": rec["value"], "unit": rec["unit"], "n_gpus": D.world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": rec["scaling"], "vs_baseline": None,
        "dtype": "bf16 (tensor-core convs/FCs, fp32 accumulate; fp16 stem operands; fp32 box/NMS/fusion math)",
        "data": "synthetic",
        "config": {"workload": "%s (%s/%s), batch %s, 512x640 frames resized to %dx%d (canvas %dx%d), R%d-FPN x%d, K=%d, 1000 proposals, "
                               "seeded random weights (BASELINE.json configs[%d])" %
                               (spec["label"], args.score_fusion, args.box_fusion, rec["shard"], rec["net_hw"][0], rec["net_hw"][1],
                                rec["canvas"][0], rec["canvas"][1], rec["depth"], len(spec["models"]), rec["K"], spec["config"]),
                   "global_batch": B * D.world, "batches_per_step": rec["batches_per_step"], "timed_region_s": rec["timed_region_s"],
                   "submission": "one CUDA graph launch per batch (%d kernels inside)" % rec["kernel_launches_per_batch"]
                                 if rec["graph_launches_per_batch"] else "%d kernel launches per batch" % rec["kernel_launches_per_batch"],
                   "parallelism": "dp%d (pairs sharded, one NCCL all-gather of detections per batch on a side stream)" % D.world,
                   "l2": "per-batch activation working set (several GB) >> 126 MB L2; no explicit flush needed",
                   "detections_per_image": rec["detections_per_image"], "fused_per_pair": rec["fused_per_item"]},
        "e2e": rec["e2e"],
        "gpu_launches": rec["kernel_launches_per_batch"] * rec["batches_per_step"] * args.steps,
        "clocks": rec["clocks"],
        "roofline": roof,
        "extra": extra,
    }
    if D.world == 1 and args.cpu_seconds > 0 and name == "pairs":
        out["cpu_baseline"] = cpu_pairs_baseline(rec["depth"], (args.score_fusion, args.box_fusion), n_pairs=1)
    else:
        out["cpu_baseline"] = {"value": None, "unit": "pairs/s", "cores": 0, "kind": "port",
                               "sample": "timed on rank 0 at N = 1 only" if D.world > 1 else "skipped (--cpu-seconds 0)"}
    return out


# -------------------------------------------------------------------------------- CPU legs (oracle port)
def _cpu_pairs_worker(a):
    depth, method, n_pairs, threads, seed = a
    return cpu_pairs_run(depth, method, n_pairs, threads, seed)


def cpu_pairs_baseline(depth, method, n_pairs=1, threads_per_proc=16):
    """The oracle port of the whole reference path on ALL host cores: cores // 16 worker processes (the reference
    itself is a batch-1 PyTorch loop; 16 intra-op threads each is where torch CPU convs stop scaling), every
    worker runs n_pairs pairs; throughput = total pairs / wall time."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    threads = min(threads_per_proc, cores)
    procs = max(1, cores // threads)
    if procs == 1:
        cpu_pairs_run(depth, method, 0, threads, 0)
        dt = cpu_pairs_run(depth, method, n_pairs, threads, 4242)
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            pool.map(_cpu_pairs_worker, [(depth, method, 0, threads, 0)] * procs)  # import / warm-up
            t = time.perf_counter()
            pool.map(_cpu_pairs_worker, [(depth, method, n_pairs, threads, 4242 + i) for i in range(procs)])
            dt = time.perf_counter() - t
    total = n_pairs * procs
    return {"value": total / dt, "unit": "pairs/s", "cores": procs * threads, "kind": "port",
            "sample": "%d synthetic pair(s) over %d processes x %d threads: oracle/detector_oracle.py (torch CPU fp32, batch 1 per "
                      "modality, the reference's own Pillow BILINEAR resize) + oracle/proben_oracle.py; /root/reference is not present on "
                      "the GPU box, so the pinned port stands in for the reference's GeneralizedRCNN" % (total, procs, threads)}


def cpu_pairs_run(depth, method, n_pairs, threads, seed):
    """Runs n_pairs pairs through the oracle (GeneralizedRCNN x2 on torch CPU fp32 + numpy ProbEn); returns seconds."""
    import torch
    from oracle import detector_oracle as Dm
    from oracle import proben_oracle as O
    from probenb200 import detector, weights
    torch.set_num_threads(threads)
    nh, nw = detector.resize_shortest_edge_shape(512, 640)
    sds = [weights.random_state_dict(depth, 3, 3, seed=11 + m) for m in range(2)]
    cfg = Dm.DetCfg(depth=depth)
    rgb, th = synth_frames(max(n_pairs, 1), seed)
    try:  # DefaultPredictor's resize of a 3-channel uint8 frame is PIL.Image.resize(BILINEAR) (transform.py:92-96)
        from PIL import Image

        def resize(fr):
            return np.asarray(Image.fromarray(fr).resize((nw, nh), Image.BILINEAR))
    except ImportError:
        from oracle import resize_oracle as R

        def resize(fr):
            return R.pil_bilinear_resize_u8(fr, nh, nw)
    t = time.perf_counter()
    for i in range(n_pairs):
        infos = []
        for m, fr in enumerate((rgb, th)):
            x = torch.from_numpy(resize(fr[i]).astype(np.float32)).permute(2, 0, 1).contiguous()
            r = Dm.detector_forward([x], [(512, 640)], sds[m], cfg)[0]
            infos.append({"bbox": r["pred_boxes"].tolist(), "score": r["scores"].tolist(), "class": r["pred_classes"].tolist(),
                          "prob": r["prob_score"].tolist(), "vars": r["vars"].tolist()})
        O.late_fusion_dispatch(method, infos)
    return time.perf_counter() - t


def run_reference_pairs(args):
    rank, _, world = dist_env()
    if rank != 0:
        return None
    method = (args.score_fusion, args.box_fusion)
    vals = []
    for i in range(min(args.warmup, 1) + min(args.steps, 3)):
        r = cpu_pairs_baseline(args.depth, method, n_pairs=1)
        if i >= min(args.warmup, 1):
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    r = vals[-1]
    r["value"] = v
    return {
        "impl": "reference", "metric": "RGB+thermal image-pairs/sec end-to-end (dual Faster R-CNN R%d-FPN -> ProbEn)" % args.depth,
        "value": v, "unit": "pairs/s", "n_gpus": world, "steps": len(vals), "warmup": min(args.warmup, 1), "ms_per_step": 1e3 / v,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (CPU)", "data": "synthetic",
        "config": {"workload": "FLIR RGB+thermal dual detector -> ProbEn (%s/%s), 1 pair/step per worker process (bounded sample; the "
                               "reference runs batch 1), 512x640 frames resized to 800x1000, R%d-FPN x2, K=3 "
                               "(BASELINE.json configs[2])" % (method[0], method[1], args.depth)},
        "cpu_baseline": r, "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


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
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pairs", choices=["pairs", "fusion", "thermal8", "kaist32", "ensemble64"])
    ap.add_argument("--dump_launches", default="", help="write the per-GEMM-launch roofline table of the instrumented pass to this CSV")
    ap.add_argument("--batch", type=int, default=0, help="image pairs per GPU per batch (pairs workload; default 16)")
    ap.add_argument("--batches_per_step", type=int, default=0, help="back-to-back batches per step (default: auto, timed region >= --target-seconds)")
    ap.add_argument("--target-seconds", type=float, default=3.0, dest="target_seconds")
    ap.add_argument("--profile-passes", type=int, default=5, dest="profile_passes")
    ap.add_argument("--no-graph", action="store_true", dest="no_graph", help="launch the kernels of a batch one by one instead of one CUDA graph")
    ap.add_argument("--no-extras", action="store_true", dest="no_extras", help="skip the sub-records of the other BASELINE configs")
    ap.add_argument("--depth", type=int, default=50, choices=[50, 101])
    ap.add_argument("--images", type=int, default=1 << 20)
    ap.add_argument("--models", type=int, default=2)
    ap.add_argument("--mean-dets", type=float, default=7.5, dest="mean_dets")
    ap.add_argument("--score_fusion", default="probEn")
    ap.add_argument("--box_fusion", default="v-avg")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, dest="cpu_seconds")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        # the reference arm runs on rank 0's host cores only; the other ranks leave before any NCCL initialisation
        out = run_reference_fusion(args) if args.workload == "fusion" else run_reference_pairs(args)
        if out is not None:
            print(json.dumps(out))
        return
    D = Dist()
    out = run_fusion(args, D) if args.workload == "fusion" else run_pairs(args, D)
    if out is not None:
        print(json.dumps(out))
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
