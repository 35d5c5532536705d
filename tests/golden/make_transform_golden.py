"""Golden vectors of the reference's input transforms, from the REAL reference in the build container.

    python tests/golden/make_transform_golden.py      # writes tests/golden/transform_golden.npz

/root/reference/utils/data_utils.py imports skimage / cv2-style modules that are not installed here and are not
needed by the classes exercised (Normalization :94-105, RandomFlip :107-126, ToTensor :159-168 and the
``/255`` + dtype conversions of PatchDataset.__getitem__ :216-219).  The harness registers empty stand-in modules
for the missing imports before importing the reference file; the reference source is untouched.
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference_data_utils():
    for name in ("skimage", "skimage.color", "skimage.io", "skimage.transform", "cv2", "openslide", "matplotlib",
                 "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.__dict__.setdefault("__path__", [])
                for attr in ("rgb2hed", "hed2rgb", "rgb2gray", "imread", "resize"):
                    setattr(m, attr, lambda *a, **k: (_ for _ in ()).throw(RuntimeError("stub")))
                sys.modules[name] = m
    sys.path.insert(0, REF)
    import utils.data_utils as du                         # /root/reference/utils/data_utils.py
    return du


def main():
    du = _import_reference_data_utils()
    rng = np.random.default_rng(7)
    out = {}
    n, size = 6, 16
    imgs = rng.integers(0, 256, size=(n, size, size, 3), dtype=np.uint8)
    imgs[0, 0, :8, 0] = np.arange(8) * 36 + 3              # a few hand-picked byte values incl. 0 / 255
    imgs[0, 1, :4, 1] = [0, 255, 127, 128]
    labs = rng.choice(np.array([0, 255, 254, 1, 128], dtype=np.uint8), size=(n, size, size), p=[0.5, 0.4, 0.04, 0.03, 0.03])
    out["img_u8"], out["label_u8"] = imgs, labs
    xs, ys, flips = [], [], []
    norm, flip, tot = du.Normalization(mean=0.5, std=0.5), du.RandomFlip(), du.ToTensor()
    for i in range(n):
        # PatchDataset.__getitem__ :216-219
        inp, lab = imgs[i] / 255.0, labs[i] / 255.0
        inp, lab = inp.astype(np.float32), lab.astype(np.uint8)
        np.random.seed(100 + i)
        r = np.random.rand(2)                              # the two draws RandomFlip makes, in order
        np.random.seed(100 + i)
        d = tot(flip(norm({"input": inp, "label": lab})))
        xs.append(d["input"].numpy())
        ys.append(d["label"].numpy())
        flips.append(int(r[0] > 0.5) | (int(r[1] > 0.5) << 1))
    out["x"] = np.stack(xs)                                # float32 [n,3,size,size]
    out["label"] = np.stack(ys)                            # int64   [n,size,size]
    out["flip"] = np.array(flips, dtype=np.uint8)          # bit 0: left-right, bit 1: up-down
    assert out["x"].dtype == np.float32 and out["label"].dtype == np.int64
    np.savez_compressed(os.path.join(HERE, "transform_golden.npz"), **out)
    print("flips", flips, "x range", out["x"].min(), out["x"].max(), "label values", np.unique(out["label"]))


if __name__ == "__main__":
    main()
