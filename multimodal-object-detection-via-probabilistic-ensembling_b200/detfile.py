"""Binary columnar detections file.

The reference exchanges detections between its two demos as indented JSON (demo_FLIR_save_predictions.py:166-176
writes, demo_probEn.py:335-337 reads): ~14 decimal floats per detection, parsed into Python lists and converted row
by row (prepare_data, demo_probEn.py:79-90).  This container stores the same columns as flat little-endian arrays in
exactly the layout ``pe_fuse_batch`` consumes (SoA rows + CSR offsets per image), so a validation set goes from disk
to HBM without a per-detection Python step; ``to_json_dict`` / ``from_json_dict`` convert to and from the reference
schema losslessly (float32 values print and parse exactly).

Layout: 8-byte magic ``PEDET\\x01\\x00\\x00`` | int64 header_len | UTF-8 JSON header {"K", "n_images", "n_rows",
"columns": [[name, dtype, shape, offset, nbytes], ...]} | 64-byte aligned column blobs:
``image_id`` int64[B], ``offsets`` int64[B+1], ``boxes`` f32[N,4], ``scores`` f32[N], ``classes`` i32[N],
``class_logits`` f32[N,K+1], ``probs`` f32[N,K], ``vars`` f32[N], ``names`` (newline-joined UTF-8).
"""
import json

import numpy as np

MAGIC = b"PEDET\x01\x00\x00"
COLUMNS = ("image_id", "offsets", "boxes", "scores", "classes", "class_logits", "probs", "vars")


class DetFile:
    def __init__(self, K, image_id, offsets, boxes, scores, classes, class_logits, probs, vars, names=None):
        self.K = int(K)
        self.image_id = np.ascontiguousarray(image_id, np.int64)
        self.offsets = np.ascontiguousarray(offsets, np.int64)
        n = int(self.offsets[-1]) if len(self.offsets) else 0
        self.boxes = np.ascontiguousarray(boxes, np.float32).reshape(n, 4)
        self.scores = np.ascontiguousarray(scores, np.float32).reshape(n)
        self.classes = np.ascontiguousarray(classes, np.int32).reshape(n)
        self.class_logits = np.ascontiguousarray(class_logits, np.float32).reshape(n, self.K + 1)
        self.probs = np.ascontiguousarray(probs, np.float32).reshape(n, self.K)
        self.vars = np.ascontiguousarray(vars, np.float32).reshape(n)
        self.names = list(names) if names is not None else ["" for _ in range(len(self.image_id))]
        if len(self.offsets) != len(self.image_id) + 1 or len(self.names) != len(self.image_id):
            raise ValueError("detfile: offsets / image_id / names lengths disagree")

    @property
    def n_images(self):
        return len(self.image_id)

    # ---- reference JSON schema <-> columns
    @classmethod
    def from_json_dict(cls, d, K=None):
        """``d``: the dict of demo_FLIR_save_predictions.py:166-176."""
        counts = np.array([len(b) for b in d["boxes"]], np.int64)
        offsets = np.zeros(len(counts) + 1, np.int64)
        np.cumsum(counts, out=offsets[1:])
        flat = lambda key, width: np.array([r for rows in d[key] for r in rows], np.float64).reshape(-1, width) if offsets[-1] else np.zeros((0, width))
        if K is None:
            K = next((len(rows[0]) for rows in d["probs"] if len(rows)), 3)
        vars_ = np.array([np.ravel(v)[0] for rows in d["vars"] for v in rows], np.float64)
        return cls(K, d["image_id"], offsets, flat("boxes", 4), np.array([s for rows in d["scores"] for s in rows], np.float64),
                   np.array([c for rows in d["classes"] for c in rows], np.int64), flat("class_logits", K + 1), flat("probs", K),
                   vars_, d.get("image"))

    def to_json_dict(self):
        out = {k: [] for k in ("image", "boxes", "scores", "classes", "image_id", "class_logits", "probs", "vars")}
        for i in range(self.n_images):
            lo, hi = int(self.offsets[i]), int(self.offsets[i + 1])
            out["image"].append(self.names[i])
            out["image_id"].append(int(self.image_id[i]))
            out["boxes"].append(self.boxes[lo:hi].astype(np.float64).tolist())
            out["scores"].append(self.scores[lo:hi].astype(np.float64).tolist())
            out["classes"].append(self.classes[lo:hi].tolist())
            out["class_logits"].append(self.class_logits[lo:hi].astype(np.float64).tolist())
            out["probs"].append(self.probs[lo:hi].astype(np.float64).tolist())
            out["vars"].append([[float(v)] for v in self.vars[lo:hi]])
        return out

    # ---- file format
    def save(self, path):
        blobs, cols, pos = [], [], 0
        arrays = [(c, getattr(self, c)) for c in COLUMNS] + [("names", np.frombuffer("\n".join(self.names).encode("utf-8"), np.uint8))]
        for name, arr in arrays:
            arr = np.ascontiguousarray(arr)
            pad = (-pos) % 64
            pos += pad
            blobs.append((pad, arr))
            cols.append([name, arr.dtype.str, list(arr.shape), pos, int(arr.nbytes)])
            pos += arr.nbytes
        header = json.dumps({"K": self.K, "n_images": self.n_images, "n_rows": int(self.offsets[-1]) if len(self.offsets) else 0,
                             "columns": cols}).encode("utf-8")
        with open(path, "wb") as f:
            f.write(MAGIC)
            f.write(np.int64(len(header)).tobytes())
            f.write(header)
            f.write(b"\0" * ((-(16 + len(header))) % 64))
            for pad, arr in blobs:
                f.write(b"\0" * pad)
                f.write(arr.tobytes())
        return path

    @classmethod
    def load(cls, path, mmap=True):
        with open(path, "rb") as f:
            if f.read(8) != MAGIC:
                raise ValueError("%s is not a PEDET file" % path)
            hlen = int(np.frombuffer(f.read(8), np.int64)[0])
            header = json.loads(f.read(hlen).decode("utf-8"))
        base = 16 + hlen + ((-(16 + hlen)) % 64)
        data = np.memmap(path, np.uint8, "r") if mmap else np.fromfile(path, np.uint8)
        got = {}
        for name, dtype, shape, off, nbytes in header["columns"]:
            got[name] = data[base + off: base + off + nbytes].view(np.dtype(dtype)).reshape(shape)
        names = bytes(got.pop("names")).decode("utf-8").split("\n") if header["n_images"] else []
        return cls(header["K"], names=names, **got)


def pack_models(files):
    """M ``DetFile``s over the same images -> the packed dict ``fusion.to_device`` / ``fuse_packed`` take (rows of one
    image = its models' detections in model order, offsets [B*M+1]); fully vectorised."""
    M = len(files)
    B = files[0].n_images
    K = files[0].K
    for f in files:
        if f.n_images != B or f.K != K:
            raise ValueError("detfile: all models must cover the same images with the same K")
    counts = np.stack([np.diff(f.offsets) for f in files], axis=1).reshape(-1)  # (b, m) order
    offsets = np.zeros(B * M + 1, np.int64)
    np.cumsum(counts, out=offsets[1:])
    N = int(offsets[-1])
    out = {"boxes": np.zeros((N, 4), np.float32), "scores": np.zeros(N, np.float32), "classes": np.zeros(N, np.int32),
           "probs": np.zeros((N, K), np.float32), "vars": np.ones(N, np.float32)}
    for m, f in enumerate(files):
        n_m = np.diff(f.offsets)
        dest0 = offsets[np.arange(B) * M + m]                      # first packed row of (image b, model m)
        dest = np.repeat(dest0 - f.offsets[:-1], n_m) + np.arange(int(f.offsets[-1]))
        out["boxes"][dest] = f.boxes
        out["scores"][dest] = f.scores
        out["classes"][dest] = f.classes
        out["probs"][dest] = f.probs
        out["vars"][dest] = f.vars
    out.update({"offsets": offsets.astype(np.int32), "B": B, "M": M, "K": K})
    return out
