"""Thin reader for the reference's on-disk patch layout, so train.py works on real data too.

The reference's input pipeline (/root/reference/utils/data_utils.py:94-236: PatchDataset + Normalization
+ RandomFlip + ToTensor behind a 16-worker DataLoader) is OUT OF SCOPE of the hot path (SURVEY.md §8(f)
"next"); the benchmark and the tests use synthetic tensors.  This module only keeps the CLI usable:
it reads ``{data_dir}/{mag}x_{size}/{id}_input.jpg`` / ``{id}_label.png`` for the ids listed in
``{data_dir}/{k}-fold_{tumorable,non_tumorable}_data.npy`` (all folds except ``fold`` = training set,
data_utils.py:56-74), applies ``/255`` and ``Normalization(0.5, 0.5)`` and a random horizontal/vertical
flip, and yields float32 NCHW batches.  RGB input only (the stain transforms need skimage/cv2).
"""
import os

import numpy as np
import torch


class PatchArrays:
    def __init__(self, data_dir, fold, patch_mag, patch_size, input_type='RGB', seed=42):
        if 'RGB' not in input_type or input_type != 'RGB':
            raise SystemExit("only --input_type RGB is read from disk (GH / H_RGB need skimage; SURVEY.md §2)")
        self.dir = os.path.join(data_dir, f'{patch_mag}x_{patch_size}')
        ids = []
        for k in range(1, 6):
            if k == fold:
                continue
            for kind in ('tumorable', 'non_tumorable'):
                p = os.path.join(data_dir, f'{k}-fold_{kind}_data.npy')
                if os.path.exists(p):
                    ids += [str(i) for i in np.load(p, allow_pickle=True).tolist()]
        if not ids:
            raise SystemExit(f'no fold lists under {data_dir}: pass --synthetic N to train on synthetic patches')
        self.ids = ids
        self.rng = np.random.default_rng(seed)

    def _load(self, pid):
        from PIL import Image
        x = np.asarray(Image.open(os.path.join(self.dir, f'{pid}_input.jpg')).convert('RGB'), dtype=np.float32) / 255.0
        y = np.asarray(Image.open(os.path.join(self.dir, f'{pid}_label.png')), dtype=np.float32)
        if y.ndim == 3:
            y = y[..., 0]
        y = (y / 255.0 if y.max() > 1 else y).astype(np.uint8).astype(np.float32)
        x = (x - 0.5) / 0.5                                   # Normalization(mean=0.5, std=0.5)
        if self.rng.random() > 0.5:                           # RandomFlip
            x, y = x[:, ::-1], y[:, ::-1]
        if self.rng.random() > 0.5:
            x, y = x[::-1], y[::-1]
        return np.ascontiguousarray(x.transpose(2, 0, 1)), np.ascontiguousarray(y)

    def batches(self, batch_size):
        order = self.rng.permutation(len(self.ids))
        for i in range(0, len(order) - batch_size + 1, batch_size):
            xs, ys = zip(*(self._load(self.ids[j]) for j in order[i:i + batch_size]))
            yield torch.from_numpy(np.stack(xs)), torch.from_numpy(np.stack(ys))


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
