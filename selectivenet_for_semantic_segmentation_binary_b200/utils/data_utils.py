"""Reader for the reference's on-disk patch layout and its train/valid/test lists, so train.py / eval.py work on
real data too.

Mirrors the data-format side of /root/reference/utils/data_utils.py: ``split_train_valid`` / ``construct_train_valid``
/ ``construct_test`` (:46-90 — every row of ``{data_dir}/{k}-fold_{tumorable,non_tumorable}_data.npy`` is a pair
``(input_file, label_file)``; the training folds are split 80/20 into train/valid with ``np.random.seed(42)``) and
``PatchDataset`` (:170-236 — ``{data_dir}/{mag}x_{size}/{input_file}`` read with PIL, ``/255`` -> float32, label
``convert('L')/255`` -> uint8).  JPEG decode and the 16-worker DataLoader are OUT OF SCOPE of the hot path (SURVEY.md
§8(f) "next"); the benchmark and the tests use synthetic tensors, and the fast path for decoded patches is
``SUNetTrainer.step_u8`` (normalisation, flips, layout and im2col in one kernel).  RGB input only (the stain
transforms ``RGB2GH`` / ``H_RGB`` need skimage and cv2, which this image does not have).
"""
import os

import numpy as np
import torch


def _fold_rows(data_dir, folds, kind):
    """All (input_file, label_file) rows of `kind` ('tumorable' | 'non_tumorable') over the given folds, in fold order."""
    return np.concatenate([np.load(os.path.join(data_dir, f'{k}-fold_{kind}_data.npy'), allow_pickle=True)
                           for k in folds])


def split_train_valid(rows, valid_ratio=0.2, rng=None):
    """Hold out int(n * valid_ratio) rows drawn without replacement (data_utils.py:49-53); the kept rows stay in their
    original order.  `rng` is the stream that np.random.seed(42) starts at import time in the reference."""
    rng = np.random if rng is None else rng
    n = len(rows)
    held_out = rng.choice(n, size=int(n * valid_ratio), replace=False)
    keep = np.ones(n, dtype=bool)
    keep[held_out] = False
    return rows[keep], rows[held_out]


def construct_train_valid(data_dir, test_fold=5):
    """Train / valid lists of the four folds that are not `test_fold` (data_utils.py:55-76): the tumorable and the
    non-tumorable rows are split 80/20 separately — tumorable first, which fixes the random stream — and stacked."""
    folds = [k for k in (1, 2, 3, 4, 5) if k != test_fold]
    rng = np.random.RandomState(42)          # == the global stream right after the reference's np.random.seed(42)
    parts = [split_train_valid(_fold_rows(data_dir, folds, kind), 0.2, rng) for kind in ('tumorable', 'non_tumorable')]
    return np.vstack([tr for tr, _ in parts]), np.vstack([va for _, va in parts])


def construct_test(data_dir, test_fold=1):
    """Test list = both row kinds of `test_fold`, tumorable first (data_utils.py:78-90)."""
    return np.vstack([np.asarray(_fold_rows(data_dir, [test_fold], kind)) for kind in ('tumorable', 'non_tumorable')])


class PatchArrays:
    """PatchDataset + the reference's transforms + DataLoader(drop_last=False) as a plain batch iterator.

    data_list rows are ``(input_file, label_file)`` (data_utils.py:173-199).  ``train=True``: shuffled every epoch,
    Normalization(0.5, 0.5) + RandomFlip + ToTensor (train.py:367); ``train=False``: in order, no flips (:368)."""

    def __init__(self, data_dir, data_list, patch_mag=200, patch_size=256, input_type='RGB', train=True, seed=0):
        if input_type != 'RGB':
            raise SystemExit("only --input_type RGB is read from disk (GH / H_RGB need skimage + cv2; SURVEY.md §2)")
        self.dir = os.path.join(data_dir, f'{patch_mag}x_{patch_size}')
        self.inputs, self.labels = [], []
        for f in data_list:
            assert str(f[0]).split('_input')[0] == str(f[1]).split('_label')[0], \
                f'check the pairness btw input {f[0]} and label {f[1]}'
            self.inputs.append(str(f[0]))
            self.labels.append(str(f[1]))
        self.train = train
        self.rng = np.random.RandomState(seed)

    def __len__(self):
        return len(self.inputs)

    def ids(self):
        return [f.split('_input')[0] for f in self.inputs]

    def _load(self, i):
        from PIL import Image
        x = np.array(Image.open(os.path.join(self.dir, self.inputs[i])))
        y = np.array(Image.open(os.path.join(self.dir, self.labels[i])).convert("L"))
        x, y = (x / 255.0).astype(np.float32), (y / 255.0).astype(np.uint8)        # data_utils.py:216-217
        x = (x - 0.5) / 0.5                                                        # Normalization(mean=0.5, std=0.5)
        if self.train:                                                             # RandomFlip: two draws per sample
            if self.rng.rand() > 0.5:
                y, x = np.fliplr(y), np.fliplr(x)
            if self.rng.rand() > 0.5:
                y, x = np.flipud(y), np.flipud(x)
        return np.ascontiguousarray(x.transpose(2, 0, 1).astype(np.float32)), np.ascontiguousarray(y)

    def n_batches(self, batch_size):
        return -(-len(self) // batch_size)

    def batches(self, batch_size):
        """float32 NCHW inputs and float32 {0,1} labels; the last batch may be short (drop_last=False)."""
        order = self.rng.permutation(len(self)) if self.train else np.arange(len(self))
        for i in range(0, len(order), batch_size):
            xs, ys = zip(*(self._load(j) for j in order[i:i + batch_size]))
            yield torch.from_numpy(np.stack(xs)), torch.from_numpy(np.stack(ys).astype(np.float32))


# ------------------------------------------------------------------ the reference's transform objects
# Same names / call convention as /root/reference/utils/data_utils.py:94-126,159-168 (dict in, dict out) for scripts
# that build ``transforms.Compose([Normalization(0.5, 0.5), RandomFlip(), ToTensor()])``.  The fast path does not use
# them: ``SUNetTrainer.step_u8`` takes the decoded uint8 patches plus the two coin flips per image and performs
# normalisation, flips, layout and the first layer's im2col in one kernel (kernels.pack_input_u8_im2col32).
class Normalization:
    def __init__(self, mean=0.5, std=0.5):
        self.mean, self.std = mean, std

    def __call__(self, data):
        data['input'] = (data['input'] - self.mean) / self.std
        return data


class RandomFlip:
    """Left-right, then up-down, each with probability 1/2 (two ``np.random.rand()`` draws per sample)."""

    @staticmethod
    def draw() -> int:
        """The two coin flips as the bit mask step_u8 expects (bit 0 left-right, bit 1 up-down)."""
        lr = np.random.rand() > 0.5
        ud = np.random.rand() > 0.5
        return int(lr) | (int(ud) << 1)

    def __call__(self, data):
        bits = self.draw()
        label, image = data['label'], data['input']
        if bits & 1:
            label, image = np.fliplr(label).copy(), np.fliplr(image)
        if bits & 2:
            label, image = np.flipud(label).copy(), np.flipud(image)
        data['input'], data['label'] = image, label
        return data


class ToTensor:
    def __call__(self, data):
        data['input'] = torch.from_numpy(np.ascontiguousarray(data['input'].transpose((2, 0, 1)).astype(np.float32)))
        data['label'] = torch.from_numpy(np.ascontiguousarray(data['label'])).type(torch.LongTensor)
        return data
