"""tcgen05 implicit-GEMM conv / linear (pe_conv2d_fwd) against a plain PyTorch fp32 reference of the same
op evaluated on the same bf16-rounded operands (floating-point kernel: tolerance = bf16 output rounding +
fp32 accumulation-order noise, stated per test)."""
import pytest
import torch
import torch.nn.functional as F

from probenb200 import ops

pytestmark = pytest.mark.gpu


def ref_conv(x, w, bias, residual, stride, relu, residual_mode):
    xf = x.float().permute(0, 3, 1, 2)
    wf = w.float().permute(0, 3, 1, 2)
    y = F.conv2d(xf, wf, bias, stride=stride, padding=w.shape[1] // 2)
    if residual_mode == 1:
        y = y + residual.float().permute(0, 3, 1, 2)
    elif residual_mode == 2:
        y = y + F.interpolate(residual.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")[:, :, :y.shape[2], :y.shape[3]]
    if relu:
        y = F.relu(y)
    return y.permute(0, 2, 3, 1).contiguous()


CASES = [
    # N, H, W, Cin, Cout, k, stride, relu, residual_mode, out_fp32
    (1, 8, 16, 64, 64, 1, 1, False, 0, False),      # single tile, single k-chunk
    (1, 8, 16, 128, 128, 1, 1, True, 0, False),     # two k-chunks
    (2, 20, 24, 64, 256, 1, 1, True, 1, False),     # ragged tiles, residual add
    (2, 13, 16, 256, 16, 1, 1, False, 0, True),     # RPN-like narrow fp32 output
    (1, 25, 32, 128, 64, 3, 1, True, 0, False),     # 3x3, borders
    (2, 50, 64, 64, 64, 3, 1, True, 0, False),
    (1, 26, 32, 256, 512, 1, 2, False, 0, False),   # stride-2 1x1 (shortcut)
    (1, 25, 31, 128, 128, 1, 2, True, 0, False),    # odd sizes, stride 2
    (2, 26, 32, 256, 256, 1, 1, False, 2, False),   # FPN lateral + upsampled top-down
    (1, 100, 128, 256, 256, 3, 1, False, 0, False), # many tiles per CTA (persistent loop, both acc stages)
    (4, 7, 9, 512, 1024, 1, 1, True, 1, False),     # tiny maps, 4 n-tiles
    (1, 1, 300, 1024, 32, 1, 1, False, 0, True),    # predictor-like linear, N tile 32
    (2, 200, 256, 64, 256, 1, 1, True, 1, False),   # res2.conv3-like: K=64, 5+ tiles per CTA, deep residual queue
    (1, 100, 128, 256, 256, 1, 1, False, 2, False), # FPN lateral + top-down through the coarse TMA residual
    (3, 50, 64, 256, 1024, 1, 1, True, 1, False),   # res4.conv3-like: 4 n-tiles, 4 k-chunks
    (1, 56, 64, 224, 64, 1, 1, True, 0, False),     # ragged K (224 = 3.5 chunks): TMA zero fill on both operands
]


@pytest.mark.parametrize("case", CASES)
def test_conv_matches_fp32_reference(case):
    N, H, W, Cin, Cout, k, stride, relu, rmode, out_fp32 = case
    g = torch.Generator(device="cuda").manual_seed(sum(int(v) * (i + 1) for i, v in enumerate(case)))
    x = torch.randn(N, H, W, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, k, k, Cin, device="cuda", generator=g) / (k * k * Cin) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device="cuda", generator=g)
    Ho, Wo = ((H - 1) // 2 + 1, (W - 1) // 2 + 1) if stride == 2 else (H, W)
    res = None
    if rmode == 1:
        res = torch.randn(N, Ho, Wo, Cout, device="cuda", generator=g).bfloat16()
    elif rmode == 2:
        res = torch.randn(N, (Ho + 1) // 2, (Wo + 1) // 2, Cout, device="cuda", generator=g).bfloat16()
    y = ops.conv2d_nhwc(x, w, bias, res, stride, relu, rmode, out_fp32)
    torch.cuda.synchronize()
    want = ref_conv(x, w, bias, res, stride, relu, rmode)
    assert y.shape == want.shape
    err = (y.float() - want).abs()
    # outputs are O(1..4); bf16 rounding of the result is <= 2^-8 relative, fp32 accumulation noise ~1e-4
    tol = 1e-3 + (0.0 if out_fp32 else 1.0) * want.abs() * 2 ** -8
    assert bool((err <= tol).all()), "max err %g" % float(err.max())


def test_linear_large_k():
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(1000, 12544, device="cuda", generator=g).bfloat16()
    w = (torch.randn(1024, 12544, device="cuda", generator=g) / 112).bfloat16()
    b = torch.randn(1024, device="cuda", generator=g)
    y = ops.linear(x, w, b, relu=True)
    want = F.relu(x.float() @ w.float().t() + b)
    err = (y.float() - want).abs()
    assert bool((err <= 2e-3 + want.abs() * 2 ** -8).all()), float(err.max())


@pytest.mark.parametrize("case", [
    # N, H, W, Cin (t2), Cin2 (block input), H2, W2, stride2, Cout
    (2, 50, 64, 64, 64, 50, 64, 1, 256),      # res2.0: conv3 64->256 + shortcut 64->256
    (2, 25, 32, 128, 256, 50, 64, 2, 512),    # res3.0: conv3 128->512 + stride-2 shortcut 256->512
    (1, 13, 16, 256, 512, 25, 31, 2, 1024),   # res4.0-like with odd input size, 4 n-tiles
])
def test_dual_input_1x1_equals_conv3_plus_projection_shortcut(case):
    """relu(conv3(t) + shortcut(x)) of a stage's first bottleneck block (resnet.py:205-221) as ONE GEMM over K = [t | x]."""
    N, H, W, Cin, Cin2, H2, W2, s2, Cout = case
    g = torch.Generator(device="cuda").manual_seed(sum(case))
    t = torch.randn(N, H, W, Cin, device="cuda", generator=g).bfloat16()
    x = torch.randn(N, H2, W2, Cin2, device="cuda", generator=g).bfloat16()
    w3 = (torch.randn(Cout, 1, 1, Cin, device="cuda", generator=g) / Cin ** 0.5).bfloat16()
    wsc = (torch.randn(Cout, 1, 1, Cin2, device="cuda", generator=g) / Cin2 ** 0.5).bfloat16()
    b3, bsc = torch.randn(Cout, device="cuda", generator=g), torch.randn(Cout, device="cuda", generator=g)
    w = torch.cat([w3.view(Cout, Cin), wsc.view(Cout, Cin2)], 1).contiguous()
    y = ops.conv1x1_dual_nhwc(t, x, w, b3 + bsc, stride2=s2, relu=True)
    torch.cuda.synchronize()
    want = F.relu(ref_conv(t, w3, b3, None, 1, False, 0) + ref_conv(x, wsc, bsc, None, s2, False, 0))
    assert y.shape == want.shape
    err = (y.float() - want).abs()
    assert bool((err <= 1e-3 + want.abs() * 2 ** -8).all()), "max err %g" % float(err.max())


@pytest.mark.parametrize("case", [
    # N, H, W, Cin, Cout, N_chain, residual, Cin2 (dual input), H2, W2, stride2
    (2, 40, 48, 64, 256, 64, True, 0, 0, 0, 1),       # res2.x conv3 (+residual) -> next conv1 256->64; one n tile
    (16, 50, 64, 64, 256, 64, True, 0, 0, 0, 1),      # many tiles per CTA: staging ring wraps, both chained accumulator stages
    (2, 25, 32, 128, 512, 128, True, 0, 0, 0, 1),     # res3: two n tiles per m tile, chained K accumulates over them
    (3, 13, 16, 256, 1024, 256, True, 0, 0, 0, 1),    # res4: four n tiles, single chained accumulator stage
    (2, 25, 32, 128, 512, 128, False, 256, 50, 64, 2),  # first block of a stage: [conv3 | stride-2 shortcut] dual input + chain
])
def test_chained_1x1_equals_conv3_then_next_conv1(case):
    """conv3 (+residual / + projection shortcut) followed by the NEXT bottleneck's conv1 in one kernel (the second GEMM reads the
    first one's bf16 output tiles from shared memory): both outputs against the fp32 reference of the two ops, the second one
    evaluated on the bf16-rounded first output exactly as the separate launches would."""
    N, H, W, Cin, Cout, Nc, use_res, Cin2, H2, W2, s2 = case
    g = torch.Generator(device="cuda").manual_seed(sum(case))
    x = torch.randn(N, H, W, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, 1, 1, Cin, device="cuda", generator=g) / Cin ** 0.5).bfloat16()
    b = torch.randn(Cout, device="cuda", generator=g)
    wc = (torch.randn(Nc, 1, 1, Cout, device="cuda", generator=g) / Cout ** 0.5).bfloat16()
    bc = torch.randn(Nc, device="cuda", generator=g)
    res = torch.randn(N, H, W, Cout, device="cuda", generator=g).bfloat16() if use_res else None
    x2 = wsc = None
    want = ref_conv(x, w, b, res, 1, False, 1 if use_res else 0)
    wfull = w.view(Cout, Cin)
    if Cin2:
        x2 = torch.randn(N, H2, W2, Cin2, device="cuda", generator=g).bfloat16()
        wsc = (torch.randn(Cout, 1, 1, Cin2, device="cuda", generator=g) / Cin2 ** 0.5).bfloat16()
        want = want + ref_conv(x2, wsc, None, None, s2, False, 0)
        wfull = torch.cat([wfull, wsc.view(Cout, Cin2)], 1).contiguous()
    want = F.relu(want)
    y, yc = ops.conv1x1_chain_nhwc(x, wfull.contiguous(), b, wc.view(Nc, Cout).contiguous(), bc, residual=res, x2=x2, stride2=s2)
    torch.cuda.synchronize()
    err = (y.float() - want).abs()
    assert bool((err <= 1e-3 + want.abs() * 2 ** -8).all()), "main output: max err %g" % float(err.max())
    want_c = ref_conv(y, wc, bc, None, 1, True, 0)   # the chained conv sees the bf16 output, like a separate launch would
    err_c = (yc.float() - want_c).abs()
    assert bool((err_c <= 2e-3 + want_c.abs() * 2 ** -8).all()), "chained output: max err %g" % float(err_c.max())


@pytest.mark.parametrize("case", [
    # N, H, W, Cin
    (2, 13, 16, 256),     # p6-like: one ragged tile per CTA (only the final drain of the chained GEMMs)
    (2, 50, 64, 256),     # one tile per CTA
    (4, 100, 128, 256),   # 400 tiles on 148 CTAs: chained GEMMs issued between the k-iterations of the next tile, both TMEM stages
    (16, 50, 64, 256),    # same with batch 16 (tile decomposition over images)
    (1, 200, 256, 64),    # short K (9 k-iterations per tile)
])
def test_rpn_head_single_kernel(case):
    """3x3 conv + ReLU -> objectness | deltas (16 fp32 outputs) in ONE kernel (the hidden tensor never leaves the SM) against
    (a) the two separate launches the engine used before - the chained GEMM consumes the same bf16-rounded hidden tile with the
    same K order, so the results must agree to fp32 accumulation noise - and (b) the fp32 reference of the two ops."""
    N, H, W, Cin = case
    g = torch.Generator(device="cuda").manual_seed(sum(case))
    x = torch.randn(N, H, W, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(256, 3, 3, Cin, device="cuda", generator=g) / (9 * Cin) ** 0.5).bfloat16()
    b = torch.randn(256, device="cuda", generator=g)
    wc = (torch.randn(16, 1, 1, 256, device="cuda", generator=g) / 16.0).bfloat16()
    bc = torch.randn(16, device="cuda", generator=g)
    t = ops.conv2d_nhwc(x, w, b, None, 1, True, 0, False)
    two_step = ops.conv2d_nhwc(t, wc, bc, None, 1, False, 0, True)
    for _ in range(2):  # twice: the second launch starts from warm TMEM / smem state
        got = ops.conv_rpn_head_nhwc(x, w, b, wc.view(16, 256).contiguous(), bc)
        torch.cuda.synchronize()
        assert got.shape == two_step.shape and got.dtype == torch.float32
        d = (got - two_step).abs()
        assert bool((d <= 1e-5 + two_step.abs() * 1e-5).all()), "vs separate launches: max diff %g" % float(d.max())
    want = ref_conv(t, wc, bc, None, 1, False, 0)
    err = (got - want).abs()
    assert bool((err <= 2e-3 + want.abs() * 2 ** -8).all()), "vs fp32 reference: max err %g" % float(err.max())
