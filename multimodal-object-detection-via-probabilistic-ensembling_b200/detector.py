"""Host side of the native detector engine (``pe_detector_*`` in include/probenb200.h).

``Detector`` mirrors the reference's model interface: ``Detector.forward(batched_inputs)`` takes the same
``[{"image": CHW float tensor, "height": H, "width": W}, ...]`` list as ``GeneralizedRCNN.forward``
(detectron2/modeling/meta_arch/rcnn.py:146-170) and returns ``[{"instances": Instances}, ...]`` with the fork's
fields ``pred_boxes, scores, pred_classes, class_logits, prob_score, vars``.  ``DefaultPredictor`` mirrors
engine/defaults.py:161-198 (resize shortest edge to 800 / max 1333, HWC->CHW float32, batch of one).

All arithmetic on activations runs in the CUDA library; this module folds FrozenBatchNorm2d into the conv
weights once at load time (layers/batch_norm.py:45-64: y = (x - mean) * w / sqrt(var + 1e-5) + b) and
repacks them into the engine's blob layout.
"""
import ctypes

import torch

from . import _lib
from .structures import Boxes, Instances

FLIR_PIXEL_MEAN = (103.530, 116.280, 123.675)
THERMAL_MEAN = 135.438
MAX_DET = 100


def fusion_method_config(method):
    """Input format / pixel statistics per ``--fusion_method`` (demo_FLIR_save_predictions.py:58-73)."""
    if method in ("rgb_only", "thermal_only"):
        return dict(in_channels=3, middle_fusion=False, pixel_mean=FLIR_PIXEL_MEAN, pixel_std=(1.0,) * 3)
    if method == "early_fusion":
        return dict(in_channels=4, middle_fusion=False, pixel_mean=FLIR_PIXEL_MEAN + (THERMAL_MEAN,), pixel_std=(1.0,) * 4)
    if method == "middle_fusion":
        return dict(in_channels=6, middle_fusion=True, pixel_mean=FLIR_PIXEL_MEAN + (THERMAL_MEAN,) * 3, pixel_std=(1.0,) * 6)
    raise ValueError("The method is not supported: %r" % (method,))


def _fold_bn(sd, name):
    w = sd[name + ".weight"].float()
    if name + ".norm.weight" in sd:
        scale = sd[name + ".norm.weight"].float() / torch.sqrt(sd[name + ".norm.running_var"].float() + 1e-5)
        bias = sd[name + ".norm.bias"].float() - sd[name + ".norm.running_mean"].float() * scale
        w = w * scale.view(-1, 1, 1, 1)
    else:
        bias = sd[name + ".bias"].float() if name + ".bias" in sd else torch.zeros(w.shape[0])
    return w, bias


def pack_weights(sd, manifest, total_bytes, num_classes):
    """state dict (reference names) -> uint8 blob following the engine manifest."""
    blob = torch.zeros(total_bytes, dtype=torch.uint8)

    def put(off, t):
        raw = t.contiguous().view(torch.uint8).reshape(-1)
        blob[off:off + raw.numel()] = raw

    K = num_classes
    for p in manifest:
        name, kind = p["name"], p["kind"]
        if kind in (0, 1):
            w, b = _fold_bn(sd, name)
            w = w.permute(0, 2, 3, 1)
            assert tuple(w.shape) == (p["cout"], p["kh"], p["kw"], p["cin"]), (name, tuple(w.shape), p)
            put(p["weight_offset"], w.to(torch.bfloat16))
        elif kind == 7:
            # conv3 | projection shortcut of a stage's first block, concatenated along K (biases add): see engine.cu
            w3, b3 = _fold_bn(sd, name)
            wsc, bsc = _fold_bn(sd, name[: -len("conv3")] + "shortcut")
            w = torch.cat([w3.reshape(w3.shape[0], -1), wsc.reshape(wsc.shape[0], -1)], 1)
            b = b3 + bsc
            assert tuple(w.shape) == (p["cout"], p["cin"]), (name, tuple(w.shape), p)
            put(p["weight_offset"], w.to(torch.bfloat16))
        elif kind == 2:
            # stem: K index = kh*32 + kw*4 + c over a 7 x 8 x 4 window (kw = 7 and c >= C are zero)
            w, b = _fold_bn(sd, name)
            wp = torch.zeros(p["cout"], 7, 8, 4)
            wp[:, :, :7, : w.shape[1]] = w.permute(0, 2, 3, 1)
            wp = wp.reshape(p["cout"], -1)
            assert wp.shape[1] == p["cin"], (wp.shape, p["cin"])
            put(p["weight_offset"], wp.to(torch.float16))
        elif kind == 3:
            wo, bo = sd[name + ".objectness_logits.weight"].float(), sd[name + ".objectness_logits.bias"].float()
            wd, bd = sd[name + ".anchor_deltas.weight"].float(), sd[name + ".anchor_deltas.bias"].float()
            w = torch.zeros(p["cout"], p["cin"])
            b = torch.zeros(p["cout"])
            # rows: 3 objectness | 1 zero pad | 12 deltas (a*4+j) so that each anchor's deltas are one aligned float4
            w[: wo.shape[0]] = wo.reshape(wo.shape[0], -1)
            w[4: 4 + wd.shape[0]] = wd.reshape(wd.shape[0], -1)
            b[: wo.shape[0]] = bo
            b[4: 4 + wd.shape[0]] = bd
            put(p["weight_offset"], w.to(torch.bfloat16))
        elif kind == 4:
            w, b = sd[name + ".weight"].float(), sd[name + ".bias"].float()
            c = p["cin"] // 49
            w = w.view(w.shape[0], c, 7, 7).permute(0, 2, 3, 1).reshape(w.shape[0], -1)
            put(p["weight_offset"], w.to(torch.bfloat16))
        elif kind == 5:
            w, b = sd[name + ".weight"].float(), sd[name + ".bias"].float()
            put(p["weight_offset"], w.to(torch.bfloat16))
        elif kind == 6:
            parts = [(sd[name + ".cls_score.weight"], sd[name + ".cls_score.bias"]),
                     (sd[name + ".bbox_pred.weight"], sd[name + ".bbox_pred.bias"]),
                     (sd[name + ".var_pred.weight"], sd[name + ".var_pred.bias"])]
            w = torch.zeros(p["cout"], p["cin"])
            b = torch.zeros(p["cout"])
            r = 0
            for pw, pb in parts:
                w[r: r + pw.shape[0]] = pw.float()
                b[r: r + pw.shape[0]] = pb.float()
                r += pw.shape[0]
            assert r == (K + 1) + 4 * K + 1
            put(p["weight_offset"], w.to(torch.bfloat16))
        else:
            raise RuntimeError("unknown manifest kind %d" % kind)
        put(p["bias_offset"], b.float())
    return blob


class DetectionBuffers:
    """Device-side ``pe_detections`` record arrays for a batch."""

    def __init__(self, B, K, device):
        self.B, self.K = B, K
        self.boxes = torch.zeros((B, MAX_DET, 4), dtype=torch.float32, device=device)
        self.scores = torch.zeros((B, MAX_DET), dtype=torch.float32, device=device)
        self.classes = torch.zeros((B, MAX_DET), dtype=torch.int32, device=device)
        self.class_logits = torch.zeros((B, MAX_DET, K + 1), dtype=torch.float32, device=device)
        self.probs = torch.zeros((B, MAX_DET, K), dtype=torch.float32, device=device)
        self.vars = torch.zeros((B, MAX_DET), dtype=torch.float32, device=device)
        self.roi_index = torch.zeros((B, MAX_DET), dtype=torch.int32, device=device)
        self.counts = torch.zeros((B,), dtype=torch.int32, device=device)

    def struct(self):
        return _lib.Detections(*[t.data_ptr() for t in (self.boxes, self.scores, self.classes, self.class_logits,
                                                        self.probs, self.vars, self.roi_index, self.counts)])

    def to_instances(self, out_sizes):
        """D2H + split into per-image ``Instances`` (the schema of rcnn.py:288-302 / fast_rcnn.py:133-145)."""
        counts = self.counts.cpu().tolist()
        host = {k: getattr(self, k).cpu() for k in ("boxes", "scores", "classes", "class_logits", "probs", "vars")}
        res = []
        for b, n in enumerate(counts[: len(out_sizes)]):
            inst = Instances(out_sizes[b])
            inst.pred_boxes = Boxes(host["boxes"][b, :n].clone())
            inst.scores = host["scores"][b, :n].clone()
            inst.pred_classes = host["classes"][b, :n].to(torch.int64)
            inst.class_logits = host["class_logits"][b, :n].clone()
            inst.prob_score = host["probs"][b, :n].clone()
            inst.vars = host["vars"][b, :n].clone().view(-1, 1)
            res.append(inst)
        return res


class Detector:
    """Faster R-CNN R50/R101-FPN inference engine bound to one weight set."""

    def __init__(self, state_dict, depth=50, num_classes=3, in_channels=3, middle_fusion=False,
                 pixel_mean=FLIR_PIXEL_MEAN, pixel_std=(1.0, 1.0, 1.0), max_batch=1, canvas=(800, 1024),
                 score_thresh=0.5, nms_thresh=0.5, rpn_nms_thresh=0.7, pre_nms_topk=1000, post_nms_topk=1000,
                 detections_per_image=100, device="cuda"):
        lib = _lib.load()
        self.lib = lib
        self.device = torch.device(device)
        self._ctor = dict(depth=depth, num_classes=num_classes, in_channels=in_channels, middle_fusion=middle_fusion,
                          pixel_mean=tuple(pixel_mean), pixel_std=tuple(pixel_std), score_thresh=score_thresh, nms_thresh=nms_thresh,
                          rpn_nms_thresh=rpn_nms_thresh, pre_nms_topk=pre_nms_topk, post_nms_topk=post_nms_topk,
                          detections_per_image=detections_per_image)
        self.num_classes, self.in_channels, self.max_batch = num_classes, in_channels, max_batch
        self.canvas = (int(canvas[0]), int(canvas[1]))
        cfg = _lib.DetectorConfig()
        cfg.depth, cfg.in_channels, cfg.middle_fusion, cfg.num_classes = depth, in_channels, int(middle_fusion), num_classes
        cfg.max_batch, cfg.canvas_h, cfg.canvas_w = max_batch, self.canvas[0], self.canvas[1]
        for i in range(8):
            cfg.pixel_mean[i] = float(pixel_mean[i]) if i < len(pixel_mean) else 0.0
            cfg.pixel_std[i] = float(pixel_std[i]) if i < len(pixel_std) else 1.0
        cfg.score_thresh, cfg.nms_thresh, cfg.rpn_nms_thresh = score_thresh, nms_thresh, rpn_nms_thresh
        cfg.pre_nms_topk, cfg.post_nms_topk, cfg.detections_per_image = pre_nms_topk, post_nms_topk, detections_per_image
        handle = ctypes.c_void_p()
        _lib.check(lib.pe_detector_create(ctypes.byref(cfg), ctypes.byref(handle)), "pe_detector_create")
        self.handle = handle
        self.manifest = []
        info = _lib.ParamInfo()
        for i in range(lib.pe_detector_num_params(handle)):
            _lib.check(lib.pe_detector_param_info(handle, i, ctypes.byref(info)), "pe_detector_param_info")
            self.manifest.append({"name": info.name.decode(), "kind": info.kind, "cout": info.cout, "kh": info.kh, "kw": info.kw,
                                  "cin": info.cin, "weight_offset": info.weight_offset, "bias_offset": info.bias_offset})
        self.weight_bytes = int(lib.pe_detector_weight_bytes(handle))
        self.ws_bytes = int(lib.pe_detector_workspace_bytes(handle))
        self.weights = None
        if state_dict is not None:
            self.load_state_dict(state_dict)
        self.workspace = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        self.out = DetectionBuffers(max_batch, num_classes, self.device)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.pe_detector_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def load_state_dict(self, sd):
        blob = pack_weights(sd, self.manifest, self.weight_bytes, self.num_classes)
        self.weights = blob.to(self.device)

    def clone_shared_weights(self, max_batch=None):
        """A second engine over the SAME weight blob (own scratch + outputs): lets sub-batches of one model run on
        different streams."""
        c = self._ctor
        twin = Detector(None, depth=c["depth"], num_classes=c["num_classes"], in_channels=c["in_channels"], middle_fusion=c["middle_fusion"],
                        pixel_mean=c["pixel_mean"], pixel_std=c["pixel_std"], max_batch=max_batch or self.max_batch, canvas=self.canvas,
                        score_thresh=c["score_thresh"], nms_thresh=c["nms_thresh"], rpn_nms_thresh=c["rpn_nms_thresh"],
                        pre_nms_topk=c["pre_nms_topk"], post_nms_topk=c["post_nms_topk"],
                        detections_per_image=c["detections_per_image"], device=self.device)
        twin.weights = self.weights
        return twin

    def share_workspace(self, other):
        """Two detectors with the same plan (e.g. the RGB and the thermal model) can share scratch memory."""
        assert other.ws_bytes >= self.ws_bytes
        self.workspace = other.workspace

    def buffer(self, name):
        """Named intermediate of the last forward (stage-wise parity tests)."""
        off, dims, elem = ctypes.c_size_t(), (ctypes.c_int * 4)(), ctypes.c_int()
        _lib.check(self.lib.pe_detector_buffer_info(self.handle, name.encode(), ctypes.byref(off), ctypes.byref(dims), ctypes.byref(elem)),
                   "pe_detector_buffer_info(%s)" % name)
        n = dims[0] * dims[1] * dims[2] * dims[3]
        raw = self.workspace[off.value: off.value + n * elem.value]
        return raw, tuple(dims), elem.value

    def forward_device(self, images, out_hw, out=None):
        """images: [B, C, h, w] float32 CUDA tensor.  Asynchronous; returns the DetectionBuffers."""
        _lib.require_cuda(images)
        if images.dtype != torch.float32 or images.dim() != 4 or images.shape[1] != self.in_channels:
            raise RuntimeError("probenb200.Detector: images must be float32 [B,%d,h,w]" % self.in_channels)
        B, _, h, w = images.shape
        out = out or self.out
        det = out.struct()
        with torch.cuda.device(images.device):  # the engine launches on the CURRENT device: make it the tensors' device
            st = self.lib.pe_detector_forward(self.handle, _lib.ptr(self.weights), _lib.ptr(images), B, h, w, float(out_hw[0]), float(out_hw[1]),
                                              ctypes.byref(det), _lib.ptr(self.workspace), self.ws_bytes, _lib.current_stream_ptr(images.device))
        _lib.check(st, "pe_detector_forward")
        return out

    def forward_frames_device(self, frames_u8, net_hw, out=None, round_u8=True):
        """frames_u8: [B, H, W, C] uint8 CUDA tensor (raw frames); resized to ``net_hw`` inside the engine
        (DefaultPredictor semantics); detections are reported in the frame's own H x W coordinates."""
        _lib.require_cuda(frames_u8)
        if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[3] != self.in_channels:
            raise RuntimeError("probenb200.Detector: frames must be uint8 [B,H,W,%d]" % self.in_channels)
        B, H, W, _ = frames_u8.shape
        out = out or self.out
        det = out.struct()
        with torch.cuda.device(frames_u8.device):
            st = self.lib.pe_detector_forward_frames(self.handle, _lib.ptr(self.weights), _lib.ptr(frames_u8), B, H, W, int(net_hw[0]), int(net_hw[1]),
                                                     int(round_u8), float(H), float(W), ctypes.byref(det), _lib.ptr(self.workspace), self.ws_bytes,
                                                     _lib.current_stream_ptr(frames_u8.device))
        _lib.check(st, "pe_detector_forward_frames")
        return out

    def set_profiling(self, enabled):
        _lib.check(self.lib.pe_detector_set_profiling(self.handle, int(enabled)), "pe_detector_set_profiling")

    def set_stagger_event(self, event, after_launches):
        """``event``: a ``torch.cuda.Event`` (recorded at least once, so that its handle exists) or None; see
        ``pe_detector_set_stagger_event``."""
        handle = None if event is None else ctypes.c_void_p(event.cuda_event)
        _lib.check(self.lib.pe_detector_set_stagger_event(self.handle, handle, int(after_launches)), "pe_detector_set_stagger_event")

    def last_profile(self):
        """(gemm_ms, span_ms, launches, gemm_launches) of the last forward (syncs on its last GEMM event)."""
        g, s = ctypes.c_float(), ctypes.c_float()
        n, ng = ctypes.c_int(), ctypes.c_int()
        _lib.check(self.lib.pe_detector_last_profile(self.handle, ctypes.byref(g), ctypes.byref(s), ctypes.byref(n), ctypes.byref(ng)),
                   "pe_detector_last_profile")
        return g.value, s.value, n.value, ng.value

    def profile_launches(self, capacity=512):
        """Per GEMM launch of the last profiled forward: (ms, algorithmic flops, algorithmic bytes) numpy arrays."""
        import numpy as np
        ms = (ctypes.c_float * capacity)()
        fl = (ctypes.c_double * capacity)()
        by = (ctypes.c_double * capacity)()
        n = min(capacity, self.lib.pe_detector_profile_launches(self.handle, ms, fl, by, capacity))
        return np.array(ms[:n]), np.array(fl[:n]), np.array(by[:n])

    def profile_kernels(self, capacity=64):
        """Non-GEMM launch groups of the last forward profiled with mode 1: [(launcher name, device ms), ...]."""
        names = ctypes.create_string_buffer(32 * capacity)
        ms = (ctypes.c_float * capacity)()
        n = min(capacity, self.lib.pe_detector_profile_kernels(self.handle, names, ms, capacity))
        return [(names.raw[32 * i: 32 * i + 32].split(b"\0", 1)[0].decode(), float(ms[i])) for i in range(n)]

    def forward(self, batched_inputs):
        """GeneralizedRCNN.forward-compatible entry (inference only): all images must share one size."""
        imgs = torch.stack([x["image"].to(torch.float32) for x in batched_inputs]).to(self.device)
        h, w = imgs.shape[-2:]
        oh = batched_inputs[0].get("height", h)
        ow = batched_inputs[0].get("width", w)
        out = self.forward_device(imgs.contiguous(), (oh, ow))
        inst = out.to_instances([(oh, ow)] * len(batched_inputs))
        return [{"instances": i} for i in inst]

    __call__ = forward


class DefaultPredictor:
    """engine/defaults.py:161-198: one BGR uint8 HxWxC frame in, ``{"instances": Instances}`` out; the resize runs
    on the GPU (``pe_resize_frames``)."""

    def __init__(self, model, min_size=800, max_size=1333):
        self.model, self.min_size, self.max_size = model, min_size, max_size

    def __call__(self, original_image):
        from . import ops
        import numpy as np
        h, w = original_image.shape[:2]
        nh, nw = resize_shortest_edge_shape(h, w, self.min_size, self.max_size)
        u8 = torch.from_numpy(np.ascontiguousarray(original_image.astype(np.uint8)))[None].to(self.model.device)
        x = ops.resize_frames(u8, (nh, nw), round_u8=original_image.shape[2] == 3)
        out = self.model.forward_device(x, (h, w))
        return {"instances": out.to_instances([(h, w)])[0]}


def resize_shortest_edge_shape(h, w, short=800, max_size=1333):
    """ResizeShortestEdge.get_transform arithmetic (data/transforms/transform_gen.py:192-213)."""
    scale = short * 1.0 / min(h, w)
    newh, neww = (short, scale * w) if h < w else (scale * h, short)
    if max(newh, neww) > max_size:
        s = max_size * 1.0 / max(newh, neww)
        newh, neww = newh * s, neww * s
    return int(newh + 0.5), int(neww + 0.5)
