#!/usr/bin/env python
"""Drop-in for the reference's demo/FLIR/demo_FLIR_save_predictions.py (same flags, same output file / schema):

    python demo/FLIR/demo_FLIR_save_predictions.py --dataset_path /path/to/FLIR/val \
        --fusion_method thermal_only --model_path out_model_thermal_only.pth

Runs the B200 detector engine over the validation pairs and writes ``<outfolder>/val_<method>_predictions.json``
with the keys of demo_FLIR_save_predictions.py:166-176 (image, boxes, scores, classes, image_id, class_logits,
probs, vars; detections with class > 2 dropped, :148-164).  Rows follow ``data['images']`` like the reference's
loop (:93-99); the ``image`` column repeats its quirk of listing ``os.listdir(RGB)`` names by position (:42,157).

The input side runs on the GPU: the JPEG files are read as bytes, decoded by nvJPEG into HBM,
the RGB frame is resized to the thermal frame's size with OpenCV's 8-bit bilinear arithmetic and the 3-/4-/6-channel
input is assembled on the device (``probenb200.io``); DefaultPredictor's ResizeShortestEdge is fused into the
detector's stem staging.  Extra flags (not in the reference): ``--batch N`` frames per launch, ``--depth {50,101}``,
``--cpu_decode`` to read and assemble the frames with cv2 exactly as the reference does - REQUIRED for bit-level decoder
parity with the reference: nvJPEG and libjpeg-turbo differ by 1-2 grey levels.  ``rgb_only`` runs the 80-class COCO
zoo model from ``trained_models/Detectron2_pretrained/model_final_f6e8b1.pkl`` like the reference (falling back to
``--model_path`` when that file is absent); ``.pkl`` (model-zoo format) and ``.pth`` checkpoints are both accepted.
"""
import json
import os
import sys
from os import listdir
from os.path import isfile, join

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from probenb200 import detector, weights  # noqa: E402
from probenb200.opt import config_parser  # noqa: E402


# rgb_only: the COCO model-zoo R101-FPN 3x checkpoint, 80 classes (demo_FLIR_save_predictions.py:58-60)
ZOO_RGB_MODEL = "trained_models/Detectron2_pretrained/model_final_f6e8b1.pkl"


def load_frame(t_path, rgb_path, name, method):
    """Input assembly of demo_FLIR_save_predictions.py:100-121 (needs OpenCV)."""
    import cv2
    stem = name.split(".")[0]
    thermal = cv2.imread(join(t_path, stem + ".jpeg"))
    if method == "thermal_only":
        return thermal.astype(np.float32)
    rgb = cv2.imread(join(rgb_path, name))
    rgb = cv2.resize(rgb, (thermal.shape[1], thermal.shape[0]))  # default interpolation = bilinear
    if method == "rgb_only":
        return rgb.astype(np.float32)
    if method == "early_fusion":
        return np.concatenate([rgb, thermal[:, :, :1]], axis=2).astype(np.float32)
    return np.concatenate([rgb, thermal], axis=2).astype(np.float32)  # middle_fusion: BGRTTT


def resize_like_predictor(img, short=800, max_size=1333):
    """DefaultPredictor's ResizeShortestEdge (engine/defaults.py:186-190; cv2 bilinear as transform.py:81-99 uses
    for float / >3-channel inputs)."""
    import cv2
    h, w = img.shape[:2]
    nh, nw = detector.resize_shortest_edge_shape(h, w, short, max_size)
    return cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR)


def extra_flags(argv):
    """Flags this CLI adds to the reference's (stripped before the reference parser sees the command line)."""
    import argparse
    ap = argparse.ArgumentParser(add_help=False)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--depth", type=int, default=101, choices=[50, 101])
    ap.add_argument("--cpu_decode", action="store_true")
    return ap.parse_known_args(argv)


def row_width(K):
    return 1 + detector.MAX_DET * (4 + 1 + 1 + (K + 1) + K + 1)


def detection_rows(buf, n):
    """The first n images of a ``DetectionBuffers`` as one float32 row each: count | boxes | scores | classes | logits | probs | vars
    (fixed stride, stays on the device: this is what the multi-GPU gather moves)."""
    return torch.cat([buf.counts[:n, None].float(), buf.boxes[:n].reshape(n, -1), buf.scores[:n], buf.classes[:n].float(),
                      buf.class_logits[:n].reshape(n, -1), buf.probs[:n].reshape(n, -1), buf.vars[:n]], 1)


def rows_to_instances(rows, K, hw):
    """Inverse of ``detection_rows`` on the host -> list of Instances with the fork's fields (fast_rcnn.py:133-145)."""
    from probenb200.structures import Boxes, Instances
    D = detector.MAX_DET
    out = []
    for r in rows:
        n = int(r[0])
        o = 1
        inst = Instances(hw)
        inst.pred_boxes = Boxes(r[o:o + 4 * D].view(D, 4)[:n].clone()); o += 4 * D
        inst.scores = r[o:o + D][:n].clone(); o += D
        inst.pred_classes = r[o:o + D][:n].to(torch.int64); o += D
        inst.class_logits = r[o:o + D * (K + 1)].view(D, K + 1)[:n].clone(); o += D * (K + 1)
        inst.prob_score = r[o:o + D * K].view(D, K)[:n].clone(); o += D * K
        inst.vars = r[o:o + D][:n].clone().view(-1, 1)
        out.append(inst)
    return out


def dist_setup():
    """(rank, world) from the torchrun environment; initialises NCCL and selects cuda:LOCAL_RANK when world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
    return dist.get_rank(), world


def save_predictions(args, batch=8, depth=101, cpu_decode=False):
    from probenb200 import io as pio
    from probenb200.pipeline import shard_range
    rank, world = dist_setup()
    val_folder = args.dataset_path
    val_json_path = val_folder + "/FLIR_thermal_RGBT_pairs_val.json"
    rgb_path, t_path = val_folder + "/RGB/", val_folder + "/thermal_8_bit/"
    method = args.fusion_method
    print("==========================")
    print("model:", method)
    print("==========================")
    data = json.load(open(val_json_path, "r"))
    stems = [im["file_name"].split("/")[1].split(".")[0] for im in data["images"]]
    name_to_id = {st: im["id"] for st, im in zip(stems, data["images"])}
    files_names = [f for f in listdir(rgb_path) if isfile(join(rgb_path, f))]
    if not os.path.exists(args.outfolder):
        os.mkdir(args.outfolder)
    mcfg = detector.fusion_method_config(method)
    K = 80 if method == "rgb_only" else 3
    model_path = args.model_path
    if method == "rgb_only" and os.path.isfile(ZOO_RGB_MODEL):
        model_path = ZOO_RGB_MODEL  # the reference ignores --model_path for rgb_only (:58-60)
    sd = weights.load_checkpoint(model_path, num_classes=K)
    print("model loaded:", model_path)
    decoder = None if cpu_decode else pio.JpegDecoder()
    shard = shard_range(len(stems), rank, world)

    def load_batch(names):
        """uint8 device frames [B, H, W, C] of the pairs `names` (file stems), C = 3 / 4 / 6."""
        if cpu_decode:
            return torch.from_numpy(np.stack([load_frame(t_path, rgb_path, n + ".jpg", method).astype(np.uint8) for n in names])).cuda()
        return pio.load_pair_batch(decoder, [join(rgb_path, n + ".jpg") for n in names], [join(t_path, n + ".jpeg") for n in names],
                                   method)

    first = load_batch(stems[:1])
    h0, w0 = int(first.shape[1]), int(first.shape[2])
    net_hw = detector.resize_shortest_edge_shape(h0, w0)
    canvas = ((net_hw[0] + 31) // 32 * 32, (net_hw[1] + 31) // 32 * 32)
    det = detector.Detector(sd, depth=depth, num_classes=K, max_batch=batch, canvas=canvas,
                            score_thresh=0.5, device="cuda:%d" % torch.cuda.current_device(), **mcfg)
    # multi-GPU: `torchrun --nproc-per-node N demo/FLIR/demo_FLIR_save_predictions.py ...` shards the pairs contiguously over
    # the ranks (InferenceSampler's rule, data/samplers/distributed_sampler.py:190-193); every rank keeps its detections as
    # fixed-stride device rows and rank 0 receives them through one padded NCCL all-gather (the reference pickles them through
    # gloo, evaluation/FLIR_evaluation.py:125-131) and writes the JSON.  Single process: the shard is the whole set.
    lo, hi = shard
    rows = []
    for i0 in range(lo, hi, batch):
        names = stems[i0:min(i0 + batch, hi)]
        frames = load_batch(names)
        # 3-channel uint8 frames take Pillow's resize in the reference, 4-/6-channel arrays cv2's float path
        buf = det.forward_frames_device(frames, net_hw, round_u8=frames.shape[3] == 3)
        rows.append(detection_rows(buf, len(names)))
    rows = torch.cat(rows) if rows else torch.zeros((0, row_width(K)), device="cuda")
    if world > 1:
        from probenb200 import pipeline
        rows = torch.cat(pipeline.all_gather_ragged(rows))
        if rank != 0:
            return None
    out = {k: [] for k in ("image", "boxes", "scores", "classes", "image_id", "class_logits", "probs", "vars")}
    for i, (n, inst) in enumerate(zip(stems, rows_to_instances(rows.cpu(), K, (h0, w0)))):
        inst = inst[inst.pred_classes <= 2]
        out["image"].append(files_names[i] if i < len(files_names) else n + ".jpg")
        out["image_id"].append(name_to_id[n])
        out["boxes"].append(inst.pred_boxes.tensor.tolist())
        out["scores"].append(inst.scores.tolist())
        out["classes"].append(inst.pred_classes.tolist())
        out["class_logits"].append(inst.class_logits.tolist())
        out["probs"].append(inst.prob_score.tolist())
        out["vars"].append(inst.vars.tolist())
    path = join(args.outfolder, "val_" + method + "_predictions.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=2)
    print("saved", path)
    # the same columns as a binary columnar file; demo_probEn.py prefers it when present
    from probenb200 import detfile
    detfile.DetFile.from_json_dict(out, K=K).save(path[:-5] + ".pedet")
    return path


if __name__ == "__main__":
    extra, rest = extra_flags(sys.argv[1:])
    save_predictions(config_parser(rest), batch=extra.batch, depth=extra.depth, cpu_decode=extra.cpu_decode)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
