"""RGB+thermal detection -> ProbEn pipeline: M per-modality detectors feed the late-fusion kernel without
leaving the device (the reference does this through JSON files: demo_FLIR_save_predictions.py writes them,
demo_probEn.py reads them back and fuses on the CPU).

One stream, no host synchronisation: detector m -> fixed-stride detection records -> ``pe_pack_detections``
(CSR layout, model order preserved = prepare_data's concatenation order) -> ``pe_fuse_batch``.  All results
of a step live in ONE flat device buffer so that the multi-GPU path needs a single NCCL all-gather.
"""
import ctypes
import os

import torch

from . import _lib
from .detector import MAX_DET, DetectionBuffers, Detector
from .fusion import _method_codes


class FusedOutput:
    """Views into one flat int32/float32 buffer: [counts B | offsets B*M+1 | classes N | scores N | boxes 4N]."""

    @staticmethod
    def words_for(B, M):
        N = B * M * MAX_DET
        return B + (B * M + 1) + 2 * N + 4 * N + ((-(B + B * M + 1 + 2 * N)) % 4)

    def __init__(self, B, M, device, storage=None):
        """``storage``: optional int32 view (``words_for(B, M)`` words, 16-byte aligned) inside a larger buffer, so that
        several sub-batch pipelines publish their results into ONE tensor (one all-gather)."""
        self.B, self.M = B, M
        self.N = N = B * M * MAX_DET
        self.words = B + (B * M + 1) + N + N + 4 * N
        pad = (-(B + B * M + 1 + 2 * N)) % 4  # keep boxes 16-byte aligned
        self.words += pad
        self.flat = torch.zeros(self.words, dtype=torch.int32, device=device) if storage is None else storage
        assert self.flat.numel() == self.words and self.flat.data_ptr() % 16 == 0
        o = 0
        self.counts = self.flat[o:o + B]; o += B
        self.offsets = self.flat[o:o + B * M + 1]; o += B * M + 1
        self.classes = self.flat[o:o + N]; o += N
        self.scores = self.flat[o:o + N].view(torch.float32); o += N
        o += pad
        self.boxes = self.flat[o:o + 4 * N].view(torch.float32).view(N, 4)

    @staticmethod
    def split(flat, B, M):
        """Host-side unpack of one rank's flat buffer -> list over images of (boxes, scores, classes) or None."""
        N = B * M * MAX_DET
        pad = (-(B + B * M + 1 + 2 * N)) % 4
        f = flat.cpu()
        o = 0
        counts = f[o:o + B].tolist(); o += B
        offsets = f[o:o + B * M + 1].tolist(); o += B * M + 1
        classes = f[o:o + N]; o += N
        scores = f[o:o + N].view(torch.float32); o += N + pad
        boxes = f[o:o + 4 * N].view(torch.float32).view(N, 4)
        out = []
        for b in range(B):
            n, lo = counts[b], offsets[b * M]
            out.append(None if n == 0 else (boxes[lo:lo + n].clone(), scores[lo:lo + n].clone(), classes[lo:lo + n].to(torch.float32)))
        return out


class ProbEnPipeline:
    """``detectors``: list of M ``Detector`` objects (same num_classes); model order = fusion order."""

    def __init__(self, detectors, method=("probEn", "v-avg"), iou_thr=0.5, frame_size=(512, 640), concurrent=True, out_storage=None,
                 stagger=None):
        self.lib = _lib.load()
        self.detectors = list(detectors)
        self.M = len(self.detectors)
        if not 1 <= self.M <= 4:
            raise ValueError("1..4 models supported")
        self.K = self.detectors[0].num_classes
        self.B = min(d.max_batch for d in self.detectors)
        self.device = self.detectors[0].device
        self.method = method
        self.codes = _method_codes(method)
        self.iou_thr = float(iou_thr)
        self.frame_h, self.frame_w = frame_size
        # every model runs on its own stream with its own scratch arena, so the latency-bound stages of one
        # detector (top-k, NMS, head post-processing: a handful of CTAs) overlap the other detector's GEMMs
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.M)] if concurrent and self.M > 1 else None
        if self.streams is None:
            big = max(self.detectors, key=lambda d: d.ws_bytes)
            for d in self.detectors:
                if d is not big:
                    d.share_workspace(big)
        self.ev_start = torch.cuda.Event()
        self.ev_done = [torch.cuda.Event() for _ in range(self.M)]
        # stagger between the model streams: model m + 1 starts once model m has launched ``stagger`` kernels (canvas staging, stem,
        # max-pool and the first res2 layers: ~0.5 ms), so the detectors never reach their latency-bound stages (RPN top-k / NMS /
        # merge: ~0.35 ms on a few CTAs) together and each of them runs under the other model's GEMMs.  0 = start together.
        # measured on B200 (profiles/r02_pipeline_stagger_sweep.txt): every offset is slower than starting together, hence 0
        if stagger is None:
            stagger = int(os.environ.get("PE_PIPE_STAGGER", "0"))
        self.stagger = int(stagger) if self.streams is not None else 0
        self.ev_stagger = [torch.cuda.Event() for _ in range(self.M - 1)] if self.stagger else []
        for e in self.ev_stagger:
            e.record()  # creates the handle the engine records on
        self.dets = [DetectionBuffers(self.B, self.K, self.device) for _ in range(self.M)]
        self.det_structs = (_lib.Detections * self.M)(*[d.struct() for d in self.dets])
        N = self.B * self.M * MAX_DET
        self.in_boxes = torch.zeros((N, 4), dtype=torch.float32, device=self.device)
        self.in_scores = torch.zeros(N, dtype=torch.float32, device=self.device)
        self.in_classes = torch.zeros(N, dtype=torch.int32, device=self.device)
        self.in_probs = torch.zeros((N, self.K), dtype=torch.float32, device=self.device)
        self.in_vars = torch.zeros(N, dtype=torch.float32, device=self.device)
        self.out = FusedOutput(self.B, self.M, self.device, storage=out_storage)
        self.ws_bytes = int(self.lib.pe_fuse_workspace_bytes(self.B))
        self.fuse_ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)

    def forward_device(self, images, net_hw=None):
        """images: list of M CUDA tensors, either float32 [B, C_m, h, w] (already resized) or - with ``net_hw`` -
        raw uint8 frames [B, H, W, C_m] that the engine resizes to ``net_hw`` on the fly.  Asynchronous.
        Returns the ``FusedOutput`` (boxes in the frame_size coordinate system)."""
        B = images[0].shape[0]
        nvtx = torch.cuda.nvtx

        def run(det, img, buf):
            if net_hw is not None:
                # 3-channel uint8 frames take Pillow's 8-bit resize in the reference, 4-/6-channel arrays cv2's float path
                # (data/transforms/transform.py:81-99)
                det.forward_frames_device(img, net_hw, out=buf, round_u8=img.shape[3] == 3)
            else:
                det.forward_device(img, (self.frame_h, self.frame_w), out=buf)
        if B != self.B:
            raise RuntimeError("pipeline was built for batch %d" % self.B)
        main = torch.cuda.current_stream(self.device)
        stream = _lib.current_stream_ptr(self.device)
        # NVTX ranges (nsys / ncu --nvtx): one per model forward, one for pack + fuse
        if self.streams is None:
            for m, (det, img, buf) in enumerate(zip(self.detectors, images, self.dets)):
                nvtx.range_push("probenb200.detector[%d]" % m)
                run(det, img, buf)
                nvtx.range_pop()
        else:
            self.ev_start.record(main)
            for m, (det, img, buf) in enumerate(zip(self.detectors, images, self.dets)):
                with torch.cuda.stream(self.streams[m]):
                    self.streams[m].wait_event(self.ev_start)
                    if self.stagger and m > 0:
                        self.streams[m].wait_event(self.ev_stagger[m - 1])  # recorded by model m - 1's forward, see __init__
                    if self.stagger:
                        det.set_stagger_event(self.ev_stagger[m] if m + 1 < self.M else None, self.stagger)
                    nvtx.range_push("probenb200.detector[%d]" % m)
                    run(det, img, buf)
                    nvtx.range_pop()
                    if self.stagger:
                        det.set_stagger_event(None, 0)
                    self.ev_done[m].record(self.streams[m])
            for m in range(self.M):
                main.wait_event(self.ev_done[m])
        o = self.out
        nvtx.range_push("probenb200.pack+fuse")
        with torch.cuda.device(self.device):
            st = self.lib.pe_pack_detections(self.det_structs, self.M, B, self.K, _lib.ptr(o.offsets), _lib.ptr(self.in_boxes),
                                             _lib.ptr(self.in_scores), _lib.ptr(self.in_classes), _lib.ptr(self.in_probs),
                                             _lib.ptr(self.in_vars), stream)
            _lib.check(st, "pe_pack_detections")
            st = self.lib.pe_fuse_batch(_lib.ptr(self.in_boxes), _lib.ptr(self.in_scores), _lib.ptr(self.in_classes), _lib.ptr(self.in_probs),
                                        _lib.ptr(self.in_vars), _lib.ptr(o.offsets), B, self.M, self.K, self.iou_thr, self.codes[0],
                                        self.codes[1], float(self.frame_w), float(self.frame_h), _lib.ptr(o.boxes), _lib.ptr(o.scores),
                                        _lib.ptr(o.classes), _lib.ptr(o.counts), _lib.ptr(self.fuse_ws), self.ws_bytes, stream)
            _lib.check(st, "pe_fuse_batch")
        nvtx.range_pop()
        return o

    def gather(self, out, group=None):
        return all_gather_flat(out.flat, group)

    def capture(self, images, net_hw=None):
        """Records one ``forward_device`` over ``images`` (device buffers whose ADDRESSES the graph keeps: refill them in
        place) into a CUDA graph - the ~200 kernel launches of a step (M detector plans on their streams, pack, fuse)
        become ONE submission per step.  Returns the ``torch.cuda.CUDAGraph``; ``graph.replay()`` on the current stream
        runs the step.  The reference needs several host syncs per image here (rpn.py:174-177, rpn_outputs.py:132,145,
        ROIAlign_cuda.cu:363); the engine has none, which is what makes the capture possible."""
        self.forward_device(images, net_hw)  # outside the capture: first-launch attribute setup, lazy module loading
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.forward_device(images, net_hw)
        return graph

    def launches_per_step(self):
        """Kernel launches of the last forward (detector plans + 2 pack + 2 fuse kernels)."""
        return sum(d.last_profile()[2] for d in self.detectors) + 4


def all_gather_flat(flat, group=None):
    """The one collective of the path: every rank's flat result buffer -> every rank, one all-gather (NCCL on
    GPUs; the reference gathers pickled predictions to rank 0 over gloo, evaluation/FLIR_evaluation.py:125-131,
    utils/comm.py:177-217)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    full = torch.empty(world * flat.numel(), dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(full, flat, group=group)
    return full.view(world, flat.numel())


class AsyncGather:
    """The path's one collective, taken off the compute stream: step i's flat result is snapshotted (double buffer) and
    all-gathered on a side stream while step i + 1 computes, so a rank never waits for the slowest rank of the SAME step
    (with the gather on the compute stream every step ends in a rank barrier and straggler skew accumulates).
    ``submit`` returns (gathered [world, n] tensor, event recorded on the gather stream when it is complete)."""

    def __init__(self, flat, group=None):
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.stream = torch.cuda.Stream(device=flat.device)
        self.snap = [torch.empty_like(flat) for _ in range(2)]
        self.full = [torch.empty(self.world * flat.numel(), dtype=flat.dtype, device=flat.device) for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]
        self.i = 0

    def submit(self, flat, host_out=None):
        import torch.distributed as dist
        s = self.i & 1
        main = torch.cuda.current_stream(flat.device)
        main.wait_event(self.done[s])             # the gather of step i - 2 has consumed this snapshot slot
        self.snap[s].copy_(flat, non_blocking=True)
        self.ready[s].record(main)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.ready[s])
            dist.all_gather_into_tensor(self.full[s], self.snap[s], group=self.group)
            if host_out is not None:
                host_out.copy_(self.full[s], non_blocking=True)
            self.done[s].record(self.stream)
        self.i += 1
        return self.full[s].view(self.world, flat.numel()), self.done[s]

    def finish(self):
        """Joins the gather stream into the current stream (end of a timed region / before reading results)."""
        torch.cuda.current_stream(self.snap[0].device).wait_stream(self.stream)


def all_gather_ragged(rows, group=None):
    """Gather of UNEQUAL shards (a 1013-image validation set over 8 ranks: InferenceSampler's last shard is short,
    data/samplers/distributed_sampler.py:190-193): ``rows`` [n_local, w] -> list over ranks of [n_r, w] tensors.  Two
    collectives: the shard lengths, then the rows padded to the longest shard."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    counts = torch.empty(world, dtype=torch.int64, device=rows.device)
    dist.all_gather_into_tensor(counts, n, group=group)
    counts = counts.tolist()
    cap = max(counts)
    padded = torch.zeros((cap,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    padded[: rows.shape[0]] = rows
    full = torch.empty((world * cap,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    dist.all_gather_into_tensor(full, padded, group=group)
    return [full[r * cap: r * cap + counts[r]] for r in range(world)]


def shard_range(num_items, rank, world):
    """Contiguous shards, InferenceSampler's rule (data/samplers/distributed_sampler.py:190-193)."""
    shard = (num_items - 1) // world + 1
    lo = min(shard * rank, num_items)
    return lo, min(lo + shard, num_items)
