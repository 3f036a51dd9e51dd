"""RGB+thermal detection -> ProbEn pipeline: M per-modality detectors feed the late-fusion kernel without
leaving the device (the reference does this through JSON files: demo_FLIR_save_predictions.py writes them,
demo_probEn.py reads them back and fuses on the CPU).

One stream, no host synchronisation: detector m -> fixed-stride detection records -> ``pe_pack_detections``
(CSR layout, model order preserved = prepare_data's concatenation order) -> ``pe_fuse_batch``.  All results
of a step live in ONE flat device buffer so that the multi-GPU path needs a single NCCL all-gather.
"""
import ctypes

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

    def __init__(self, detectors, method=("probEn", "v-avg"), iou_thr=0.5, frame_size=(512, 640), concurrent=True, out_storage=None):
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
        def run(det, img, buf):
            if net_hw is not None:
                det.forward_frames_device(img, net_hw, out=buf)
            else:
                det.forward_device(img, (self.frame_h, self.frame_w), out=buf)
        if B != self.B:
            raise RuntimeError("pipeline was built for batch %d" % self.B)
        main = torch.cuda.current_stream(self.device)
        stream = _lib.current_stream_ptr(self.device)
        if self.streams is None:
            for det, img, buf in zip(self.detectors, images, self.dets):
                run(det, img, buf)
        else:
            self.ev_start.record(main)
            for m, (det, img, buf) in enumerate(zip(self.detectors, images, self.dets)):
                with torch.cuda.stream(self.streams[m]):
                    self.streams[m].wait_event(self.ev_start)
                    run(det, img, buf)
                    self.ev_done[m].record(self.streams[m])
            for m in range(self.M):
                main.wait_event(self.ev_done[m])
        o = self.out
        st = self.lib.pe_pack_detections(self.det_structs, self.M, B, self.K, _lib.ptr(o.offsets), _lib.ptr(self.in_boxes),
                                         _lib.ptr(self.in_scores), _lib.ptr(self.in_classes), _lib.ptr(self.in_probs),
                                         _lib.ptr(self.in_vars), stream)
        _lib.check(st, "pe_pack_detections")
        st = self.lib.pe_fuse_batch(_lib.ptr(self.in_boxes), _lib.ptr(self.in_scores), _lib.ptr(self.in_classes), _lib.ptr(self.in_probs),
                                    _lib.ptr(self.in_vars), _lib.ptr(o.offsets), B, self.M, self.K, self.iou_thr, self.codes[0], self.codes[1],
                                    float(self.frame_w), float(self.frame_h), _lib.ptr(o.boxes), _lib.ptr(o.scores), _lib.ptr(o.classes),
                                    _lib.ptr(o.counts), _lib.ptr(self.fuse_ws), self.ws_bytes, stream)
        _lib.check(st, "pe_fuse_batch")
        return o

    def gather(self, out, group=None):
        return all_gather_flat(out.flat, group)


def all_gather_flat(flat, group=None):
    """The one collective of the path: every rank's flat result buffer -> every rank, one all-gather (NCCL on
    GPUs; the reference gathers pickled predictions to rank 0 over gloo, evaluation/FLIR_evaluation.py:125-131,
    utils/comm.py:177-217)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    full = torch.empty(world * flat.numel(), dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(full, flat, group=group)
    return full.view(world, flat.numel())


def shard_range(num_items, rank, world):
    """Contiguous shards, InferenceSampler's rule (data/samplers/distributed_sampler.py:190-193)."""
    shard = (num_items - 1) // world + 1
    lo = min(shard * rank, num_items)
    return lo, min(lo + shard, num_items)
