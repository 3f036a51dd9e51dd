"""Operator seams of detectron2/layers: ROIAlign and batched_nms, written like the reference's own
tests (tests/test_roi_align.py:12-116; tests/test_rpn.py and tests/test_nms_rotated.py for the NMS semantics) and
checked against torchvision's CPU operators, the arithmetic the reference delegates to."""
import numpy as np
import pytest
import torch
import torchvision

from probenb200.layers import ROIAlign, batched_nms, nms

pytestmark = pytest.mark.gpu


def _simple_roialign(img, box, resolution, aligned=True):
    """tests/test_roi_align.py:62-77: scale 1.0, sampling ratio 0; CUDA result vs the CPU operator."""
    if isinstance(resolution, int):
        resolution = (resolution, resolution)
    op = ROIAlign(resolution, 1.0, 0, aligned=aligned)
    inp = torch.from_numpy(img[None, None, :, :].astype("float32"))
    rois = torch.from_numpy(np.asarray([0] + list(box))[None, :].astype("float32"))
    out = op.forward(inp.cuda(), rois.cuda()).cpu()
    want = torchvision.ops.roi_align(inp, rois, resolution, 1.0, 0, aligned)
    assert torch.allclose(out, want)
    return out[0, 0]


def test_roi_align_forward_output_known_answers():
    """The two golden tables of tests/test_roi_align.py:13-45."""
    inp = np.arange(25).reshape(5, 5).astype("float32")
    old = _simple_roialign(inp, [1, 1, 3, 3], (4, 4), aligned=False)
    new = _simple_roialign(inp, [1, 1, 3, 3], (4, 4), aligned=True)
    old_results = [[7.5, 8, 8.5, 9], [10, 10.5, 11, 11.5], [12.5, 13, 13.5, 14], [15, 15.5, 16, 16.5]]
    correct_results = [[4.5, 5.0, 5.5, 6.0], [7.0, 7.5, 8.0, 8.5], [9.5, 10.0, 10.5, 11.0], [12.0, 12.5, 13.0, 13.5]]
    assert np.allclose(old.numpy().flatten(), np.asarray(old_results).flatten())
    assert np.allclose(new.numpy().flatten(), np.asarray(correct_results).flatten())


def test_roi_align_resize_equivalence():
    """tests/test_roi_align.py:50-60: pooling a 2x-downsampled image with a 2x-smaller box gives the same output."""
    rng = np.random.default_rng(0)
    H, W = 30, 30
    inp = rng.random((H, W)).astype("float32") * 100
    box = [10, 10, 20, 20]
    out = _simple_roialign(inp, box, (5, 5), aligned=True)
    inp2 = torch.nn.functional.interpolate(torch.from_numpy(inp)[None, None], size=(H // 2, W // 2), mode="bilinear",
                                           align_corners=False)[0, 0].numpy()
    out2 = _simple_roialign(inp2, [x / 2 for x in box], (5, 5), aligned=True)
    assert np.abs(out2.numpy() - out.numpy()).max() < 1e-4


def test_roi_align_empty_box_and_empty_batch():
    """tests/test_roi_align.py:95-116."""
    img = np.random.default_rng(1).random((5, 5))
    o = _simple_roialign(img, [3, 4, 5, 4], 7)
    assert o.shape == (7, 7) and bool((o == 0).all())
    out = ROIAlign((7, 7), 1.0, 0, aligned=True).forward(torch.zeros(0, 3, 10, 10).cuda(), torch.zeros(0, 5).cuda())
    assert out.shape == (0, 3, 7, 7)


@pytest.mark.parametrize("sampling_ratio,aligned,scale", [(0, True, 0.25), (2, True, 0.125), (0, False, 1.0 / 16), (3, False, 0.5)])
def test_roi_align_random_rois_match_torchvision(sampling_ratio, aligned, scale):
    g = torch.Generator().manual_seed(3)
    N, C, H, W = 3, 19, 50, 64
    x = torch.randn(N, C, H, W, generator=g)
    M = 200
    xy = torch.rand(M, 2, generator=g) * torch.tensor([W / scale, H / scale]) * 0.9
    wh = torch.rand(M, 2, generator=g) * torch.tensor([W / scale, H / scale]) * 0.6 + 1
    bi = torch.randint(0, N, (M, 1), generator=g).float()
    rois = torch.cat([bi, xy, xy + wh], 1)
    rois[5, 3:] = rois[5, 1:3] - 4  # inverted box
    rois[6, 1:] = torch.tensor([-50., -60., -10., -20.]) / scale  # fully outside
    rois[7, 1:] = torch.tensor([0., 0., W / scale, H / scale])  # whole image: largest adaptive grid
    got = ROIAlign((7, 7), scale, sampling_ratio, aligned).forward(x.cuda(), rois.cuda()).cpu()
    want = torchvision.ops.roi_align(x, rois, (7, 7), scale, sampling_ratio, aligned)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=5e-5)  # float32 sums, FMA contraction differs


def test_roi_align_huge_adaptive_grid_takes_direct_path():
    g = torch.Generator().manual_seed(4)
    x = torch.randn(1, 2, 1300, 40, generator=g)
    rois = torch.tensor([[0, 2.0, 3.0, 30.0, 1290.0]])  # 1287 rows / 1 bin -> grid_h 1287 > tap table
    got = ROIAlign((1, 2), 1.0, 0, True).forward(x.cuda(), rois.cuda()).cpu()
    want = torchvision.ops.roi_align(x, rois, (1, 2), 1.0, 0, True)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=5e-5)  # float32 sums, FMA contraction differs


def test_roi_align_rejects_cpu_tensors():
    with pytest.raises(RuntimeError):
        ROIAlign((7, 7), 1.0, 0).forward(torch.zeros(1, 1, 8, 8), torch.zeros(1, 5))


def _random_boxes(n, g, size=800.0, cluster=True):
    ctr = torch.rand(max(1, n // 6), 2, generator=g) * size
    idx = torch.randint(0, ctr.shape[0], (n,), generator=g)
    xy = (ctr[idx] + torch.randn(n, 2, generator=g) * 12) if cluster else torch.rand(n, 2, generator=g) * size
    wh = torch.rand(n, 2, generator=g) * 120 + 4
    return torch.cat([xy, xy + wh], 1)


@pytest.mark.parametrize("n", [1, 33, 1000, 4624])
@pytest.mark.parametrize("thr", [0.5, 0.7])
def test_nms_matches_torchvision(n, thr):
    g = torch.Generator().manual_seed(n)
    boxes = _random_boxes(n, g)
    scores = torch.rand(n, generator=g)
    scores[::7] = scores[0]  # ties: the CPU operator sorts stably
    got = nms(boxes.cuda(), scores.cuda(), thr).cpu()
    want = torchvision.ops.nms(boxes, scores, thr)
    assert got.dtype == torch.int64
    assert torch.equal(got, want)


@pytest.mark.parametrize("n,ncls", [(50, 3), (1000, 5), (4624, 5), (6000, 80)])
def test_batched_nms_matches_torchvision(n, ncls):
    """4624 boxes x 5 levels is the RPN call (rpn_outputs.py:147); 6000 boxes exceed 20000 elements -> per-category
    mode, as torchvision's _batched_nms_vanilla."""
    g = torch.Generator().manual_seed(n + ncls)
    boxes = _random_boxes(n, g)
    scores = torch.rand(n, generator=g)
    idxs = torch.randint(0, ncls, (n,), generator=g)
    got = batched_nms(boxes.cuda(), scores.cuda(), idxs.cuda(), 0.7).cpu()
    if boxes.numel() > 20000:
        # the vanilla path re-sorts the kept indices with an unstable sort: equal scores may swap places
        want = torchvision.ops.boxes._batched_nms_vanilla(boxes, scores, idxs, 0.7)
        assert torch.equal(torch.sort(got).values, torch.sort(want).values)
        assert torch.equal(scores[got], scores[want])
    else:
        want = torchvision.ops.boxes._batched_nms_coordinate_trick(boxes, scores, idxs, 0.7)
        assert torch.equal(got, want)


def test_batched_nms_semantics_and_empty():
    """Greedy semantics documented by tests/test_nms_rotated.py:11-33: sort by score, keep iou <= thr."""
    boxes = torch.tensor([[0., 0., 10., 10.], [1., 1., 11., 11.], [0., 0., 10., 10.], [50., 50., 60., 60.]])
    scores = torch.tensor([0.9, 0.8, 0.7, 0.6])
    same = torch.zeros(4, dtype=torch.int64)
    assert batched_nms(boxes.cuda(), scores.cuda(), same.cuda(), 0.5).cpu().tolist() == [0, 3]
    diff = torch.tensor([0, 1, 2, 0])
    assert batched_nms(boxes.cuda(), scores.cuda(), diff.cuda(), 0.5).cpu().tolist() == [0, 1, 2, 3]
    e = batched_nms(torch.zeros(0, 4).cuda(), torch.zeros(0).cuda(), torch.zeros(0, dtype=torch.int64).cuda(), 0.5)
    assert e.shape == (0,) and e.dtype == torch.int64
