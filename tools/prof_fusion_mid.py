"""Profiling target: pe_fuse_batch on 2048 synthetic pairs with exactly 100 detections per model (the detector pipeline's
regime -> fuse_mid_kernel), three launches.  Not a benchmark."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from probenb200 import fusion, synth  # noqa: E402

p = synth.synth_packed(2048, num_models=2, seed=3, force_count=100)
dev = fusion.to_device(p)
for _ in range(3):
    buf = fusion.fuse_packed(dev, ("probEn", "v-avg"))
    torch.cuda.synchronize()
print("fused per pair:", float(buf.out_counts[:2048].float().mean()))
