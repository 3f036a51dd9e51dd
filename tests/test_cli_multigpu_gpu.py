"""Sharded save_predictions under torchrun (2 GPUs): every rank runs a contiguous shard of the validation pairs
(InferenceSampler's rule, data/samplers/distributed_sampler.py:190-193; 5 pairs -> 3 + 2, a ragged last shard), rank 0
receives the fixed-stride detection rows through the padded NCCL gather and writes the JSON - which must equal the JSON of
the single-process run.  Skipped unless two CUDA devices are visible (`gpurun --gpus 2`)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from probenb200 import weights
from probenb200.opt import config_parser

cv2 = pytest.importorskip("cv2")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_torchrun_sharded_save_predictions_equals_single_process(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import test_cli_gpu as T
    root = str(tmp_path / "val")
    T._make_dataset(root, n=5)
    ck = str(tmp_path / "thermal.pth")
    torch.save({"model": weights.random_state_dict(50, 3, 3, seed=20)}, ck)
    save = T._load_cli("demo_FLIR_save_predictions")
    out1 = str(tmp_path / "single") + "/"
    args = config_parser(["--dataset_path", root, "--fusion_method", "thermal_only", "--model_path", ck, "--outfolder", out1])
    single = json.load(open(save.save_predictions(args, batch=2, depth=50)))
    out2 = str(tmp_path / "sharded") + "/"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "demo", "FLIR", "demo_FLIR_save_predictions.py"), "--dataset_path", root,
           "--fusion_method", "thermal_only", "--model_path", ck, "--outfolder", out2, "--batch", "2", "--depth", "50"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    sharded = json.load(open(os.path.join(out2, "val_thermal_only_predictions.json")))
    assert sharded["image_id"] == single["image_id"]
    assert sum(len(b) for b in single["boxes"]) > 0
    for k in ("boxes", "scores", "classes", "class_logits", "probs", "vars"):
        assert sharded[k] == single[k], k
