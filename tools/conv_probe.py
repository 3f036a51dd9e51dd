"""Debug helper: run ONE conv test case (index argv[1]) in its own process and print error statistics.
Used on the GPU box so that a trapped/hung kernel cannot poison the other cases."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import test_conv_gpu as T  # noqa: E402
from probenb200 import ops  # noqa: E402

i = int(sys.argv[1])
case = T.CASES[i]
N, H, W, Cin, Cout, k, stride, relu, rmode, out_fp32 = case
g = torch.Generator(device="cuda").manual_seed(i)
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
want = T.ref_conv(x, w, bias, res, stride, relu, rmode)
err = (y.float() - want).abs()
bad = err > (1e-3 + want.abs() * 2 ** -7)
print("case %d %s: max_err %.4g mean_err %.4g bad %d/%d  |want| mean %.3g" % (
    i, case, float(err.max()), float(err.mean()), int(bad.sum()), bad.numel(), float(want.abs().mean())))
if bad.any():
    idx = bad.nonzero()[:5].tolist()
    print("  first bad idx:", idx, [float(y.float()[tuple(j)]) for j in idx], [float(want[tuple(j)]) for j in idx])
    # which channels / pixels are bad
    print("  bad per-channel count (first 16 ch):", bad.sum(dim=(0, 1, 2))[:16].tolist())
    print("  bad rows (h) sum:", bad.sum(dim=(0, 2, 3))[:16].tolist())
