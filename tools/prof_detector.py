"""Profiling target: ONE detector (R50-FPN, K=3) forward over a batch of 512x640 uint8 frames on one stream, repeated
``--reps`` times.  Under ncu the launches of repetition r are [r * L, (r + 1) * L) in plan order, so
``-k regex:conv_gemm --launch-skip 74 --launch-count 74`` captures exactly the 74 tensor-core launches of one warm
forward (tools/ncu_conv_all.sh).  Not a benchmark: numbers printed under a profiler are never bench values."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from probenb200 import detector, weights  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--depth", type=int, default=50)
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
dev = torch.device("cuda:0")
nh, nw = detector.resize_shortest_edge_shape(512, 640)
canvas = ((nh + 31) // 32 * 32, (nw + 31) // 32 * 32)
det = detector.Detector(weights.random_state_dict(args.depth, 3, 3, seed=11), depth=args.depth, num_classes=3, max_batch=args.batch,
                        canvas=canvas, device=dev)
g = torch.Generator().manual_seed(777)
frames = torch.randint(0, 256, (args.batch, 512, 640, 3), dtype=torch.uint8, generator=g).to(dev)
for _ in range(args.reps):
    out = det.forward_frames_device(frames, (nh, nw))
    torch.cuda.synchronize()
print("detections per image:", out.counts.float().mean().item())
