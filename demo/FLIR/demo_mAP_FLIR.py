#!/usr/bin/env python
"""Drop-in for the reference's demo/FLIR/demo_mAP_FLIR.py: single-model COCO bbox mAP on the FLIR validation pairs
(``inference_on_dataset(predictor.model, val_loader, FLIREvaluator(...))``, demo_mAP_FLIR.py:11-16,67).

The reference hard-codes the dataset paths, the method and the checkpoint in the file (:25-65); here they are the
flags of its sibling demos:

    python demo/FLIR/demo_mAP_FLIR.py --dataset_path /path/to/FLIR/val --fusion_method thermal_only \
        --model_path out_model_thermal_only.pth [--outfolder out]

Runs the detector over every pair (GPU decode + assembly + inference, same code as demo_FLIR_save_predictions.py),
keeps the classes FLIREvaluator keeps (FLIR_evaluation.py:313-382) and evaluates with the COCOeval protocol
(``probenb200.evaluation.COCOBBoxEval``, pinned to the reference's vendored cocoeval.py).  Prints the
``Evaluation results for bbox`` numbers (AP, AP50, AP75, APs, APm, APl, per-category AP) and stores them in
``<outfolder>/FLIR_<method>_mAP.json`` (the reference pickles the COCOeval object to ``FLIR_<method>_mAP.out``).
"""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from probenb200 import evaluation  # noqa: E402
from probenb200.opt import config_parser  # noqa: E402


def _save_cli():
    spec = importlib.util.spec_from_file_location("cli_save_predictions", os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                                                        "demo_FLIR_save_predictions.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main(argv=None):
    save = _save_cli()
    extra, rest = save.extra_flags(sys.argv[1:] if argv is None else argv)
    args = config_parser(rest)
    pred = json.load(open(save.save_predictions(args, batch=extra.batch, depth=extra.depth, cpu_decode=extra.cpu_decode)))
    gt = json.load(open(os.path.join(args.dataset_path, "FLIR_thermal_RGBT_pairs_val.json")))
    dets = []
    for iid, boxes, scores, classes in zip(pred["image_id"], pred["boxes"], pred["scores"], pred["classes"]):
        dets += evaluation.instances_to_coco_json(boxes, scores, classes, iid)
    evaluation.unmap_category_ids(dets, gt.get("categories"))  # FLIR_evaluation.py:163-175
    ev = evaluation.COCOBBoxEval(gt["annotations"], dets, image_ids=[im["id"] for im in gt["images"]])
    res = ev.evaluate(device="cuda")  # matching on the GPU (pe_coco_match)
    print("Evaluation results for bbox:")
    print(" | ".join("%s %.3f" % (k, v) for k, v in res.items()))
    out = os.path.join(args.outfolder, "FLIR_%s_mAP.json" % args.fusion_method)
    json.dump(res, open(out, "w"))
    return res


if __name__ == "__main__":
    main()
