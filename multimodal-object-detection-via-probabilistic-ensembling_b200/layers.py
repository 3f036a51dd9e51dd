"""Drop-ins for the two native operators of ``detectron2.layers`` that sit on the hot path:

* ``batched_nms`` / ``nms``  <- detectron2/layers/nms.py:9-26 (torchvision.ops.boxes.batched_nms / ops.nms)
* ``ROIAlign`` / ``roi_align`` <- detectron2/layers/roi_align.py:10-105 (``detectron2._C.roi_align_forward``)

Same names, argument meaning, return types and error behaviour (``RuntimeError``) as the reference's operators;
CUDA tensors only - there is no CPU path behind them.  Inference only: ROIAlign has no backward here (the
reference's ``_ROIAlign.backward`` belongs to training, which is out of scope).
"""
import torch

from . import _lib


def _nms_call(boxes, scores, idxs, iou_threshold, mode):
    if boxes.dim() != 2 or boxes.shape[-1] != 4:
        raise RuntimeError("boxes should be a 2d tensor of shape [N, 4], got %s" % (tuple(boxes.shape),))
    if scores.dim() != 1 or scores.shape[0] != boxes.shape[0]:
        raise RuntimeError("boxes and scores should have the same number of elements in dimension 0")
    n = boxes.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    lib = _lib.load()
    boxes = boxes.contiguous().float()
    scores = scores.contiguous().float()
    idxs = None if idxs is None else idxs.contiguous().to(torch.int64)
    _lib.require_cuda(boxes, scores, idxs)
    if n > lib.pe_batched_nms_max_boxes():
        raise RuntimeError("probenb200.batched_nms: %d boxes exceed the supported %d" % (n, lib.pe_batched_nms_max_boxes()))
    with torch.cuda.device(boxes.device):
        ws = torch.empty(lib.pe_batched_nms_workspace_bytes(n), dtype=torch.uint8, device=boxes.device)
        keep = torch.empty(n, dtype=torch.int64, device=boxes.device)
        n_keep = torch.empty(1, dtype=torch.int32, device=boxes.device)
        _lib.check(lib.pe_batched_nms(_lib.ptr(boxes), _lib.ptr(scores), _lib.ptr(idxs), n, float(iou_threshold), mode,
                                      _lib.ptr(keep), _lib.ptr(n_keep), _lib.ptr(ws), ws.numel(),
                                      _lib.current_stream_ptr(boxes.device)), "batched_nms")
    return keep[:int(n_keep.item())]


def nms(boxes, scores, iou_threshold):
    """torchvision.ops.nms: indices of the kept boxes, sorted by decreasing score."""
    return _nms_call(boxes, scores, None, iou_threshold, 0)


def batched_nms(boxes, scores, idxs, iou_threshold):
    """detectron2/layers/nms.py:9-26.  Below 40000 boxes the reference forwards to torchvision's batched_nms, which on
    CUDA uses the coordinate-offset trick up to 20000 box elements and per-category NMS above; from 40000 boxes the
    reference loops over categories itself - the same result as the per-category mode."""
    assert boxes.shape[-1] == 4
    mode = 1 if boxes.numel() > 20000 else 0
    return _nms_call(boxes, scores, idxs, iou_threshold, mode)


def roi_align(input, rois, output_size, spatial_scale, sampling_ratio, aligned):
    """``_C.roi_align_forward(input, rois, spatial_scale, pooled_h, pooled_w, sampling_ratio, aligned)``."""
    if isinstance(output_size, int):
        output_size = (output_size, output_size)
    ph, pw = int(output_size[0]), int(output_size[1])
    if rois.dim() != 2 or rois.size(1) != 5:
        raise RuntimeError("rois must be [M, 5] = (batch index, x1, y1, x2, y2)")
    if input.dim() != 4:
        raise RuntimeError("input must be [N, C, H, W]")
    n, c, h, w = input.shape
    m = rois.shape[0]
    out = torch.empty((m, c, ph, pw), dtype=torch.float32, device=input.device)
    if m == 0 or c == 0:
        return out
    x = input.contiguous().float()
    r = rois.contiguous().float()
    _lib.require_cuda(x, r)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.pe_roi_align_forward(_lib.ptr(x), n, c, h, w, _lib.ptr(r), m, float(spatial_scale), ph, pw,
                                            int(sampling_ratio), int(bool(aligned)), _lib.ptr(out),
                                            _lib.current_stream_ptr(x.device)), "roi_align_forward")
    return out.to(input.dtype)


class ROIAlign(torch.nn.Module):
    """detectron2/layers/roi_align.py:47-105: ``ROIAlign(output_size, spatial_scale, sampling_ratio, aligned=True)``;
    ``forward(input NCHW, rois Bx5)``.  ``aligned=True`` shifts the box by -0.5 pixel (the reference's corrected
    variant), ``sampling_ratio=0`` samples ceil(roi_size / output_size) points per bin."""

    def __init__(self, output_size, spatial_scale, sampling_ratio, aligned=True):
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio
        self.aligned = aligned

    def forward(self, input, rois):
        assert rois.dim() == 2 and rois.size(1) == 5
        return roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio, self.aligned)

    def __repr__(self):
        return "%s(output_size=%s, spatial_scale=%s, sampling_ratio=%s, aligned=%s)" % (
            self.__class__.__name__, self.output_size, self.spatial_scale, self.sampling_ratio, self.aligned)
