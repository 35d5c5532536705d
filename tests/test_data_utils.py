"""The reader of the reference's on-disk layout (utils/data_utils.py): fold lists of (input_file, label_file)
pairs, the seeded 80/20 train/valid split, drop_last=False batching — on a tiny fake data directory, no GPU."""
import os

import numpy as np
import pytest


def _make_fake(tmp_path, n_per_fold=10, size=16):
    from PIL import Image
    root = tmp_path / "data"
    pdir = root / f"200x_{size}"
    os.makedirs(pdir)
    rng = np.random.RandomState(0)
    for k in range(1, 6):
        for kind in ("tumorable", "non_tumorable"):
            rows = []
            for i in range(n_per_fold):
                sid = f"s{k}{kind[0]}_{i}_{i * 3}"
                img = rng.randint(0, 256, (size, size, 3)).astype(np.uint8)
                lab = (rng.rand(size, size) < 0.4).astype(np.uint8) * 255
                Image.fromarray(img).save(pdir / f"{sid}_input.png")      # png: lossless, so values can be checked
                Image.fromarray(lab).save(pdir / f"{sid}_label.png")
                rows.append((f"{sid}_input.png", f"{sid}_label.png"))
            np.save(root / f"{k}-fold_{kind}_data.npy", np.array(rows))
    return str(root), size


def test_train_valid_split_is_the_references(tmp_path):
    from selectivenet_for_semantic_segmentation_binary_b200.utils import data_utils as D
    root, size = _make_fake(tmp_path)
    train, valid = D.construct_train_valid(root, test_fold=2)
    # restatement of /root/reference/utils/data_utils.py:46-76 with the module-level np.random.seed(42)
    np.random.seed(42)
    t = np.concatenate([np.load(f"{root}/{i}-fold_tumorable_data.npy") for i in (1, 3, 4, 5)])
    n = np.concatenate([np.load(f"{root}/{i}-fold_non_tumorable_data.npy") for i in (1, 3, 4, 5)])

    def split(lst):
        vi = np.random.choice(len(lst), size=int(len(lst) * 0.2), replace=False)
        ti = np.setdiff1d(list(range(len(lst))), vi)
        return lst[ti], lst[vi]
    tt, tv = split(t)
    nt, nv = split(n)
    assert np.array_equal(train, np.vstack([tt, nt])) and np.array_equal(valid, np.vstack([tv, nv]))
    assert len(train) == 64 and len(valid) == 16
    assert not set(map(tuple, train)) & set(map(tuple, valid))                  # no leakage
    assert all("s2" not in r[0] for r in np.vstack([train, valid]))             # the test fold is held out
    test = D.construct_test(root, test_fold=2)
    assert len(test) == 20 and all(r[0].startswith("s2") for r in test)


def test_patch_arrays_reads_pairs_and_keeps_the_tail(tmp_path):
    from PIL import Image
    from selectivenet_for_semantic_segmentation_binary_b200.utils import data_utils as D
    root, size = _make_fake(tmp_path)
    test = D.construct_test(root, test_fold=1)
    ds = D.PatchArrays(root, test, 200, size, "RGB", train=False)
    assert len(ds) == 20 and ds.n_batches(8) == 3
    batches = list(ds.batches(8))
    assert [b[0].shape[0] for b in batches] == [8, 8, 4]                        # drop_last=False (eval.py:92)
    x0, y0 = batches[0]
    assert x0.shape == (8, 3, size, size) and x0.dtype.is_floating_point and y0.shape == (8, size, size)
    # first sample == PatchDataset.__getitem__ + Normalization(0.5, 0.5) + ToTensor, no flip in eval mode
    img = np.array(Image.open(os.path.join(root, f"200x_{size}", test[0][0])))
    lab = np.array(Image.open(os.path.join(root, f"200x_{size}", test[0][1])).convert("L"))
    ex = (((img / 255.0).astype(np.float32)) - 0.5) / 0.5
    assert np.array_equal(x0[0].numpy(), ex.transpose(2, 0, 1).astype(np.float32))
    assert np.array_equal(y0[0].numpy(), (lab / 255.0).astype(np.uint8).astype(np.float32))
    assert set(np.unique(y0.numpy())) <= {0.0, 1.0}
    # training mode: every sample appears once per epoch, possibly flipped
    tr = D.PatchArrays(root, test, 200, size, "RGB", train=True, seed=3)
    xs = np.concatenate([b[0].numpy() for b in tr.batches(8)])
    assert xs.shape[0] == 20
    ref_sums = sorted(round(float(b.sum()), 3) for bb in batches for b in bb[0].numpy())
    assert sorted(round(float(v.sum()), 3) for v in xs) == ref_sums             # flips permute pixels only
    with pytest.raises(AssertionError):
        D.PatchArrays(root, np.array([("a_input.png", "b_label.png")]), 200, size)
    with pytest.raises(SystemExit):
        D.PatchArrays(root, test, 200, size, "GH")
