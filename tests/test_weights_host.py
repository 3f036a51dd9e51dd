"""Checkpoint readers (detectron2/checkpoint/detection_checkpoint.py:26-45): ``.pth`` state dicts and Detectron2
model-zoo ``.pkl`` files (the reference's rgb_only branch, demo_FLIR_save_predictions.py:58-60)."""
import os
import pickle

import pytest
import torch

from probenb200 import weights


def test_model_zoo_pkl_round_trip(tmp_path):
    sd = weights.random_state_dict(50, 3, 80, seed=1)
    zoo = {k: v.numpy() for k, v in sd.items() if "var_pred" not in k}  # zoo models predate the fork's variance head
    path = str(tmp_path / "model_final.pkl")
    pickle.dump({"model": zoo, "__author__": "Detectron2 Model Zoo"}, open(path, "wb"))
    got = weights.load_checkpoint(path, num_classes=80)
    assert all(torch.equal(got[k], sd[k]) for k in zoo)
    assert got["roi_heads.box_predictor.var_pred.weight"].shape == (1, 1024) and float(got["roi_heads.box_predictor.var_pred.weight"].abs().max()) == 0
    with pytest.raises(RuntimeError, match="80 classes"):
        weights.load_checkpoint(path, num_classes=3)


def test_pth_both_layouts_and_no_code_execution(tmp_path):
    sd = weights.random_state_dict(50, 3, 3, seed=2)
    for name, obj in (("bare.pth", sd), ("wrapped.pth", {"model": sd, "iteration": 7})):
        p = str(tmp_path / name)
        torch.save(obj, p)
        got = weights.load_checkpoint(p)
        assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)

    class Evil:
        def __reduce__(self):
            return (os.system, ("true",))

    p = str(tmp_path / "evil.pkl")
    pickle.dump({"model": {"a": Evil()}, "__author__": "x"}, open(p, "wb"))
    with pytest.raises(pickle.UnpicklingError):
        weights.load_checkpoint(p)
    p = str(tmp_path / "caffe2.pkl")
    pickle.dump({"blobs": {}}, open(p, "wb"))
    with pytest.raises(RuntimeError, match="model-zoo"):
        weights.load_checkpoint(p)
