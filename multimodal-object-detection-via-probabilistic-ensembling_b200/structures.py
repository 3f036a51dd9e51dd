"""Output schema of the reference: ``Boxes`` and ``Instances`` (detectron2/structures/boxes.py:125-297,
instances.py:9-187) - the subset of behaviour the demo scripts and the evaluator use.  Pure containers
around torch tensors; no compute kernels live here."""
import itertools

import numpy as np
import torch


class Boxes:
    """N x 4 float32 xyxy absolute boxes (boxes.py:125-160: anything array-like is cast to float32,
    an empty input becomes shape (0, 4))."""

    def __init__(self, tensor):
        if isinstance(tensor, (list, tuple)) and len(tensor) and isinstance(tensor[0], np.ndarray):
            tensor = np.stack(tensor)
        device = tensor.device if isinstance(tensor, torch.Tensor) else torch.device("cpu")
        tensor = torch.as_tensor(tensor, dtype=torch.float32, device=device)
        if tensor.numel() == 0:
            tensor = tensor.reshape((0, 4)).to(dtype=torch.float32)
        assert tensor.dim() == 2 and tensor.size(-1) == 4, tensor.size()
        self.tensor = tensor

    def clone(self):
        return Boxes(self.tensor.clone())

    def to(self, device):
        return Boxes(self.tensor.to(device))

    def area(self):
        b = self.tensor
        return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])

    def clip(self, box_size):
        h, w = box_size
        self.tensor[:, 0].clamp_(min=0, max=w)
        self.tensor[:, 1].clamp_(min=0, max=h)
        self.tensor[:, 2].clamp_(min=0, max=w)
        self.tensor[:, 3].clamp_(min=0, max=h)

    def nonempty(self, threshold=0):
        b = self.tensor
        return ((b[:, 2] - b[:, 0]) > threshold) & ((b[:, 3] - b[:, 1]) > threshold)

    def scale(self, scale_x, scale_y):
        self.tensor[:, 0::2] *= scale_x
        self.tensor[:, 1::2] *= scale_y

    def __getitem__(self, item):
        if isinstance(item, int):
            return Boxes(self.tensor[item].view(1, -1))
        return Boxes(self.tensor[item])

    def __len__(self):
        return self.tensor.shape[0]

    def __iter__(self):
        yield from self.tensor

    def __repr__(self):
        return "Boxes(" + str(self.tensor) + ")"

    @property
    def device(self):
        return self.tensor.device

    @staticmethod
    def cat(boxes_list):
        if len(boxes_list) == 0:
            return Boxes(torch.empty(0))
        return Boxes(torch.cat([b.tensor for b in boxes_list], dim=0))


class Instances:
    """Per-image dict of equal-length fields plus ``image_size`` (instances.py:9-187)."""

    def __init__(self, image_size, **kwargs):
        self._image_size = tuple(image_size)
        self._fields = {}
        for k, v in kwargs.items():
            self.set(k, v)

    @property
    def image_size(self):
        return self._image_size

    def __setattr__(self, name, val):
        if name.startswith("_"):
            super().__setattr__(name, val)
        else:
            self.set(name, val)

    def __getattr__(self, name):
        if name == "_fields" or name not in self._fields:
            raise AttributeError("Cannot find field '{}' in the given Instances!".format(name))
        return self._fields[name]

    def set(self, name, value):
        data_len = len(value)
        if len(self._fields):
            assert len(self) == data_len, "Adding a field of length {} to a Instances of length {}".format(data_len, len(self))
        self._fields[name] = value

    def has(self, name):
        return name in self._fields

    def remove(self, name):
        del self._fields[name]

    def get(self, name):
        return self._fields[name]

    def get_fields(self):
        return self._fields

    def to(self, device):
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            ret.set(k, v.to(device) if hasattr(v, "to") else v)
        return ret

    def __getitem__(self, item):
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            ret.set(k, v[item])
        return ret

    def __len__(self):
        for v in self._fields.values():
            return len(v)
        raise NotImplementedError("Empty Instances does not support __len__!")

    @staticmethod
    def cat(instance_lists):
        assert len(instance_lists) > 0
        if len(instance_lists) == 1:
            return instance_lists[0]
        ret = Instances(instance_lists[0].image_size)
        for k in instance_lists[0]._fields.keys():
            values = [i.get(k) for i in instance_lists]
            v0 = values[0]
            if isinstance(v0, torch.Tensor):
                values = torch.cat(values, dim=0)
            elif isinstance(v0, list):
                values = list(itertools.chain(*values))
            elif hasattr(type(v0), "cat"):
                values = type(v0).cat(values)
            ret.set(k, values)
        return ret

    def __repr__(self):
        s = self.__class__.__name__ + "(num_instances={}, image_height={}, image_width={}, fields=[{}])".format(
            len(self) if self._fields else 0, self._image_size[0], self._image_size[1], ", ".join(self._fields.keys()))
        return s
