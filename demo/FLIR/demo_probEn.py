#!/usr/bin/env python
"""Drop-in for the reference's demo/FLIR/demo_probEn.py (same flags, same input files, same printed lines):

    python demo/FLIR/demo_probEn.py --dataset_path /path/to/FLIR/val --prediction_path out/ \
        --score_fusion probEn --box_fusion v-avg

Reads ``<prediction_path>val_{thermal_only,early_fusion,middle_fusion}_predictions.json`` (schema of
demo_FLIR_save_predictions.py:166-176), fuses ALL images in one ``pe_fuse_batch`` launch on the GPU
(the reference loops image by image in numpy, demo_probEn.py:204-292), and evaluates COCO bbox mAP against
``<dataset_path>/FLIR_thermal_RGBT_pairs_val.json`` when that file exists.  Extra: ``--save_fused FILE`` writes
the fused detections as JSON.  The class-offset tile of the clustering is 640 x 512 whatever the frames' size, as in
the reference (demo_probEn.py:100-103 hard-codes it); ``--tiles_from_annotations`` takes it from the first image of
the annotation file instead (needed for frames larger than 640 x 512, where the reference's tiles overlap).  Binary
``.pedet`` sidecars are read instead of the JSON only when they are at least as new as their JSON file (or with
``--binary``); the source used is printed.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from probenb200 import evaluation, fusion  # noqa: E402
from probenb200.opt import config_parser  # noqa: E402
from probenb200.structures import Boxes, Instances  # noqa: E402
from probenb200.synth import image_info  # noqa: E402


def apply_late_fusion(det_list, method, img_w=640, img_h=512):
    """Whole-set version of apply_late_fusion_and_evaluate's loop body; returns list over images of
    ``Instances`` (or None where no model detected anything, which the reference skips)."""
    n_img = len(det_list[1]["image"]) if len(det_list) > 1 else len(det_list[0]["image"])
    images = [[image_info(d, i) for d in det_list] for i in range(n_img)]
    fused = fusion.late_fusion_batch(method, images, img_w=img_w, img_h=img_h)
    out = []
    for r in fused:
        if r is None:
            out.append(None)
            continue
        inst = Instances([img_h, img_w])
        inst.pred_boxes = Boxes(r[0])
        import torch
        inst.scores = torch.from_numpy(r[1])
        inst.pred_classes = torch.from_numpy(r[2])
        out.append(inst)
    return out


def apply_late_fusion_columnar(cols, method, img_w=640, img_h=512):
    """Same as ``apply_late_fusion`` from ``detfile.DetFile`` columns: one vectorised pack, one launch."""
    import torch
    from probenb200 import detfile
    packed = detfile.pack_models(cols)
    buf = fusion.fuse_packed(fusion.to_device(packed), method, img_w=float(img_w), img_h=float(img_h))
    out = []
    for r in fusion.unpack_results(packed, buf):
        if r is None:
            out.append(None)
            continue
        inst = Instances([img_h, img_w])
        inst.pred_boxes = Boxes(r[0])
        inst.scores = torch.from_numpy(r[1])
        inst.pred_classes = torch.from_numpy(r[2])
        out.append(inst)
    return out


def extra_flags(argv):
    """Flags this CLI adds to the reference's (stripped before the reference parser sees the command line)."""
    import argparse
    ap = argparse.ArgumentParser(add_help=False)
    ap.add_argument("--tiles_from_annotations", action="store_true",
                    help="class-offset tile = size of the first annotated image (reference: always 640 x 512)")
    ap.add_argument("--binary", action="store_true", help="read the .pedet sidecars even when they are older than the JSON files")
    ap.add_argument("--no_binary", action="store_true", help="always parse the JSON prediction files")
    return ap.parse_known_args(argv)


def _fresh_sidecars(files, binaries, force):
    """A .pedet is used only if it exists and is not older than its JSON (the JSON is the reference's contract: a
    regenerated JSON must never be shadowed by a stale sidecar)."""
    for f, b in zip(files, binaries):
        if not os.path.isfile(b):
            return False
        if not force and os.path.isfile(f) and os.path.getmtime(b) < os.path.getmtime(f):
            print("stale sidecar %s (older than %s): reading the JSON files" % (b, f))
            return False
    return True


def main(argv=None):
    extra, rest = extra_flags(sys.argv[1:] if argv is None else argv)
    args = config_parser(rest)
    pred = args.prediction_path
    files = [pred + "val_thermal_only_predictions.json", pred + "val_early_fusion_predictions.json",
             pred + "val_middle_fusion_predictions.json"]
    for i, f in enumerate(files):
        print("detection file %d:" % (i + 1), f)
    if not os.path.exists(args.outfolder):
        os.mkdir(args.outfolder)
    method = [args.score_fusion, args.box_fusion]
    # frame size for the class-offset tiles (the reference imreads every thermal JPEG for it, demo_probEn.py:269-271)
    val_json = os.path.join(args.dataset_path or "", "FLIR_thermal_RGBT_pairs_val.json")
    gt = json.load(open(val_json)) if args.dataset_path and os.path.isfile(val_json) else None
    img_w, img_h = 640, 512  # demo_probEn.py:100-103
    if extra.tiles_from_annotations and gt and gt.get("images") and "width" in gt["images"][0]:
        img_w, img_h = int(gt["images"][0]["width"]), int(gt["images"][0]["height"])
    binaries = [f[:-5] + ".pedet" for f in files]
    if not extra.no_binary and _fresh_sidecars(files, binaries, extra.binary):
        print("reading binary columnar detections:", ", ".join(os.path.basename(b) for b in binaries))
        # binary columnar files written next to the JSON by demo_FLIR_save_predictions.py: disk -> HBM without a
        # per-detection Python step
        from probenb200 import detfile
        cols = [detfile.DetFile.load(f) for f in binaries]
        print("Method: ", method)
        start = time.time()
        results = apply_late_fusion_columnar(cols, method, img_w, img_h)
        total = time.time() - start
        image_ids = [int(i) for i in cols[1].image_id]
    else:
        print("reading JSON detections")
        dets = [json.load(open(f, "r")) for f in files if os.path.isfile(f)]
        if len(dets) < 2:
            raise FileNotFoundError("need at least two prediction files under %r" % pred)
        print("Method: ", method)
        start = time.time()
        results = apply_late_fusion(dets, method, img_w, img_h)
        total = time.time() - start
        image_ids = dets[1]["image_id"] if len(dets) > 1 else dets[0]["image_id"]
    print("Average time:", total / max(1, len(results)))
    coco_dets = []
    for inst, iid in zip(results, image_ids):
        if inst is not None:
            coco_dets += evaluation.instances_to_coco_json(inst.pred_boxes.tensor.numpy(), inst.scores.numpy(),
                                                           inst.pred_classes.numpy(), iid)
    if gt is not None:
        # FLIREvaluator un-maps the contiguous class index to the dataset's category id first (FLIR_evaluation.py:163-175)
        evaluation.unmap_category_ids(coco_dets, gt.get("categories"))
    out_json = os.path.join(args.outfolder, "probEn_%s_%s_fused.json" % tuple(method))
    json.dump(coco_dets, open(out_json, "w"))
    if gt is not None:
        ev = evaluation.COCOBBoxEval(gt["annotations"], coco_dets, image_ids=[im["id"] for im in gt["images"]])
        res = ev.evaluate(device="cuda")  # matching on the GPU (pe_coco_match); decisions identical to the host evaluator
        print("Evaluation results for bbox:")
        print(" | ".join("%s %.3f" % (k, v) for k, v in res.items()))
        return res
    print("no annotation file at %r: fused detections written to %s" % (val_json, out_json))
    return None


if __name__ == "__main__":
    main()
