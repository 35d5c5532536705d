"""CPU oracle for the SelectiveUNet hot path  —  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain fp32 (optionally fp64) CPU restatement of the reference algorithm, function by
function, each citing the reference file:line it follows.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module, and only as the checker / the CPU baseline — the product path
(``selectivenet_for_semantic_segmentation_binary_b200``) never does, and has no CPU fallback.

Pinning: ``tests/golden/make_golden.py`` (committed) imports the real reference from
/root/reference in the build container, runs it on seeded inputs and stores the outputs in
``tests/golden/*.npz``; ``tests/test_oracle.py`` checks this restatement against those vectors
and against the reference's own notebook known-answers.  Parity is therefore
pinned for: forward logits, both losses, coverage, all 68 parameter gradients, BN running
statistics, thresholded masks, reject counts, confusion matrices and derived metrics.

The arithmetic lives in third-party PyTorch / numpy in the reference too (torch 2.11.0+cu128,
numpy 2.3.5 here; the reference pins no versions); this file calls the same primitives on CPU.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# (name, cin, cout) in registration order — /root/reference/model.py:29-61
CBR_BLOCKS = [
    ("encoder_layer_1_1", None, 64), ("encoder_layer_1_2", 64, 64),
    ("encoder_layer_2_1", 64, 128), ("encoder_layer_2_2", 128, 128),
    ("encoder_layer_3_1", 128, 256), ("encoder_layer_3_2", 256, 256),
    ("decoder_layer_4_2", 256, 512), ("decoder_layer_4_1", 512, 512),
    ("unpool3", 512, 256),
    ("decoder_layer_3_2", 512, 256), ("decoder_layer_3_1", 256, 256),
    ("unpool2", 256, 128),
    ("decoder_layer_2_2", 256, 128), ("decoder_layer_2_1", 128, 128),
    ("unpool1", 128, 64),
    ("decoder_layer_1_2", 128, 64), ("decoder_layer_1_1", 64, 64),
]
BN_EPS = 1e-5        # nn.BatchNorm2d default, model.py:12
BN_MOMENTUM = 0.1


def input_channels(input_type: str) -> int:
    """model.py:24-27"""
    if "RGB" in input_type:
        return 3
    if input_type == "GH":
        return 2
    raise ValueError(input_type)


def init_state_dict(seed: int, input_type: str = "RGB", selective: bool = True,
                    n_cls: int = 1) -> "OrderedDict[str, torch.Tensor]":
    """Default-initialised parameters/buffers in the reference's key order (Appendix B of SURVEY.md).

    Restates the constructor model.py:19-66: layers are created in the same order with the same
    torch.nn classes, so ``torch.manual_seed(seed)`` yields the same tensors as
    ``torch.manual_seed(seed); UNet_B(input_type, selective)`` (checked by make_golden.py).
    """
    torch.manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    cin0 = input_channels(input_type)
    for name, cin, cout in CBR_BLOCKS:
        if name.startswith("unpool"):
            m = torch.nn.ConvTranspose2d(cin, cout, kernel_size=2, stride=2, padding=0, bias=True)
            sd[f"{name}.weight"] = m.weight.detach().clone()
            sd[f"{name}.bias"] = m.bias.detach().clone()
        else:
            conv = torch.nn.Conv2d(cin if cin is not None else cin0, cout, kernel_size=3, stride=1, padding=1, bias=True)
            bn = torch.nn.BatchNorm2d(cout)
            sd[f"{name}.0.weight"] = conv.weight.detach().clone()
            sd[f"{name}.0.bias"] = conv.bias.detach().clone()
            sd[f"{name}.1.weight"] = bn.weight.detach().clone()
            sd[f"{name}.1.bias"] = bn.bias.detach().clone()
            sd[f"{name}.1.running_mean"] = bn.running_mean.detach().clone()
            sd[f"{name}.1.running_var"] = bn.running_var.detach().clone()
            sd[f"{name}.1.num_batches_tracked"] = bn.num_batches_tracked.detach().clone()
    heads = ["conv1x1"] + (["conv_select", "conv_aux"] if selective else [])
    for h in heads:
        # n_cls = 1: UNet_B (model.py:62-66); n_cls = 2: UNet (model.py:151-155, conv_select always has 2 channels)
        m = torch.nn.Conv2d(64, n_cls if h != "conv_select" else (1 if n_cls == 1 else 2), kernel_size=1)
        sd[f"{h}.weight"] = m.weight.detach().clone()
        sd[f"{h}.bias"] = m.bias.detach().clone()
    return sd


def _cbr(sd, name, x, training, update_running=True):
    """CBR_2D: Conv2d(k3,s1,p1,bias) -> BatchNorm2d -> ReLU  (model.py:9-15)"""
    y = F.conv2d(x, sd[f"{name}.0.weight"], sd[f"{name}.0.bias"], stride=1, padding=1)
    rm, rv = sd[f"{name}.1.running_mean"], sd[f"{name}.1.running_var"]
    if training and not update_running:
        rm, rv = rm.clone(), rv.clone()
    y = F.batch_norm(y, rm, rv, sd[f"{name}.1.weight"], sd[f"{name}.1.bias"], training=training,
                     momentum=BN_MOMENTUM, eps=BN_EPS)
    if training and update_running:
        sd[f"{name}.1.num_batches_tracked"] += 1
    return F.relu(y)


def unet_b_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, selective: bool = True, training: bool = True,
                   update_running: bool = True):
    """UNet_B.forward, model.py:68-103.  x: [N,C,H,W]; returns [N,H,W] or (out, select, aux)."""
    def cbr(name, t):
        return _cbr(sd, name, t, training, update_running)

    enc1_1 = cbr("encoder_layer_1_1", x)                       # :69
    enc1_2 = cbr("encoder_layer_1_2", enc1_1)                  # :70
    pool1 = F.max_pool2d(enc1_2, kernel_size=2)                # :71
    enc2_1 = cbr("encoder_layer_2_1", pool1)
    enc2_2 = cbr("encoder_layer_2_2", enc2_1)
    pool2 = F.max_pool2d(enc2_2, kernel_size=2)                # :75
    enc3_1 = cbr("encoder_layer_3_1", pool2)
    enc3_2 = cbr("encoder_layer_3_2", enc3_1)
    pool3 = F.max_pool2d(enc3_2, kernel_size=2)                # :79
    bottom = cbr("decoder_layer_4_2", pool3)                   # :81
    bottom = cbr("decoder_layer_4_1", bottom)                  # :82

    def up(name, t):                                           # ConvTranspose2d k2 s2, model.py:44-45
        return F.conv_transpose2d(t, sd[f"{name}.weight"], sd[f"{name}.bias"], stride=2, padding=0)

    unpool3 = torch.cat((up("unpool3", bottom), enc3_2), dim=1)   # :83  [up | skip]
    dec3_2 = cbr("decoder_layer_3_2", unpool3)
    dec3_1 = cbr("decoder_layer_3_1", dec3_2)
    unpool2 = torch.cat((up("unpool2", dec3_1), enc2_2), dim=1)   # :87
    dec2_2 = cbr("decoder_layer_2_2", unpool2)
    dec2_1 = cbr("decoder_layer_2_1", dec2_2)
    unpool1 = torch.cat((up("unpool1", dec2_1), enc1_2), dim=1)   # :91
    dec1_2 = cbr("decoder_layer_1_2", unpool1)
    dec1_1 = cbr("decoder_layer_1_1", dec1_2)
    output = F.conv2d(dec1_1, sd["conv1x1.weight"], sd["conv1x1.bias"])            # :96
    if selective:
        select = F.conv2d(dec1_1, sd["conv_select.weight"], sd["conv_select.bias"])  # :99
        aux = F.conv2d(dec1_1, sd["conv_aux.weight"], sd["conv_aux.bias"])           # :100
        if output.shape[1] != 1:            # UNet (CE variant), model.py:185-191: (N, C, H, W) outputs, no squeeze
            return output, select, aux
        return torch.squeeze(output, 1), torch.squeeze(select, 1), torch.squeeze(aux, 1)  # :101
    if output.shape[1] != 1:
        return output
    return torch.squeeze(output, 1)                                                # :103


def bce_with_logits_mean(x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """torch.nn.BCEWithLogitsLoss() as used at train.py:78,195 (mean reduction)."""
    return F.binary_cross_entropy_with_logits(x, t)


def selective_risk_b(output, selection, target, target_coverage=0.8, lamb=8) -> Tuple[torch.Tensor, torch.Tensor]:
    """calc_selective_risk_image_b, selective_loss.py:58-85 (hard_selection=False path), including the
    reference's naive log(sigmoid) form (:79-80) — minus the hard-coded .cuda() of :73."""
    selection = torch.sigmoid(selection)                                      # :71
    coverage = torch.mean(selection)                                          # :72
    zero = torch.zeros(coverage.shape, dtype=coverage.dtype)                  # :73
    prob = torch.sigmoid(output)                                              # :79
    loss_risk = -torch.mean((target * torch.log(prob) + (1 - target) * torch.log(1 - prob)) * selection) / coverage
    diff, _ = torch.max(torch.stack([target_coverage - coverage, zero], dim=-1), dim=0)   # :81
    loss_constraint = torch.square(diff)                                      # :82
    loss = loss_risk + lamb * loss_constraint                                 # :84
    return loss, coverage


def train_losses(sd, x, label, s_lamb=2, selective=True, update_running=True):
    """The loss part of the training step body, train.py:193-204."""
    if selective:
        output, selection, aux = unet_b_forward(sd, x, True, True, update_running)
        aux_loss = bce_with_logits_mean(aux, label)                            # :195
        select_loss, coverage = selective_risk_b(output, selection, target=label, lamb=s_lamb)   # :196
        loss = aux_loss + select_loss                                          # :201
        return loss, dict(output=output, selection=selection, aux=aux, aux_loss=aux_loss, select_loss=select_loss,
                          coverage=coverage)
    output = unet_b_forward(sd, x, False, True, update_running)
    loss = bce_with_logits_mean(output, label)                                 # :204
    return loss, dict(output=output)


# ----------------------------------------------------------------------------- host post-processing
def sigmoid_np(x: np.ndarray, dtype) -> np.ndarray:
    """train.py:150 (float64) / eval.py:175 (float32, input dtype kept)."""
    if dtype == np.float64:
        return 1 / (1 + np.exp(-x.astype("float64")))
    return 1 / (1 + np.exp(-x))


def postprocess(output: np.ndarray, selection: Optional[np.ndarray], path: str = "train", cut_off: float = 0.5,
                s_cut_off: float = 0.5, scale: str = "sigmoid"):
    """Thresholding of train.py:216-231 (path='train', float64 sigmoid, cut 0.5) or
    eval.py:231-243 (path='eval', float32 sigmoid, --cut_off / --s_cut_off).
    Returns (pred uint8, selection float {0.,1.} or None)."""
    dt = np.float64 if path == "train" else np.float32
    o = sigmoid_np(output, dt) if scale == "sigmoid" else output
    pred = (1.0 * (o > cut_off)).astype("uint8")
    sel = None
    if selection is not None:
        s = sigmoid_np(selection, dt) if scale == "sigmoid" else selection
        sel = 1.0 * (s > s_cut_off)
    return pred, sel


def _f32_to_key(x) -> int:
    """Monotone map float32 -> uint32 (total order of finite floats)."""
    b = int(np.array([x], dtype=np.float32).view(np.uint32)[0])
    return (~b & 0xFFFFFFFF) if (b & 0x80000000) else (b | 0x80000000)


def _key_to_f32(k: int) -> np.float32:
    b = (k ^ 0x80000000) if (k & 0x80000000) else (~k & 0xFFFFFFFF)
    return np.array([b], dtype=np.uint32).view(np.float32)[0]


def logit_threshold(cut_off: float, path: str) -> np.float32:
    """Smallest float32 logit x with sigmoid_np(x) > cut_off under the numpy arithmetic of `path`
    ('train': float64 exp, train.py:150; 'eval': float32 exp, eval.py:175), found by bisection
    over float32 bit patterns.  pred == (x >= threshold)."""
    dt = np.float64 if path == "train" else np.float32

    def is_one(k: int) -> bool:
        x = np.array([_key_to_f32(k)], dtype=np.float32)
        with np.errstate(over="ignore"):
            return bool(sigmoid_np(x, dt)[0] > cut_off)

    lo, hi = _f32_to_key(np.float32(-200.0)), _f32_to_key(np.float32(200.0))
    if is_one(lo):
        return np.float32(-np.inf)
    if not is_one(hi):
        return np.float32(np.inf)
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if is_one(mid):
            hi = mid
        else:
            lo = mid
    return _key_to_f32(hi)


class Evaluator:
    """utils/compute_metric.py:4-84, restated (float64 confusion matrix, rows = label, cols = pred)."""

    def __init__(self, num_class, selective):
        self.num_class = num_class
        self.confusion_matrix = np.zeros((self.num_class,) * 2)
        self.selective = selective

    def _generate_matrix(self, label, pred, selection=None):            # :10-22
        mask = (label >= 0) & (label < self.num_class)
        if self.selective:
            mask = mask & (selection == 1)
        idx = self.num_class * label[mask].astype("int") + pred[mask]
        count = np.bincount(idx, minlength=self.num_class * 2)
        return count.reshape(self.num_class, self.num_class)

    def add_batch(self, label, pred, selection=None):                   # :24-26
        assert label.shape == pred.shape
        self.confusion_matrix += self._generate_matrix(label, pred, selection=selection)

    def reset(self):
        self.confusion_matrix = np.zeros((self.num_class,) * 2)

    def get_Pixel_Accuracy(self):                                       # :35-37
        return np.diag(self.confusion_matrix).sum() / self.confusion_matrix.sum()

    def get_Pixel_Accuracy_Class(self):                                 # :39-42
        return np.nanmean(np.diag(self.confusion_matrix) / self.confusion_matrix.sum(axis=1))

    def get_Precision(self):                                            # :48-50
        return np.diag(self.confusion_matrix) / self.confusion_matrix.sum(axis=0)

    def get_Recall(self):                                               # :52-54
        return np.diag(self.confusion_matrix) / self.confusion_matrix.sum(axis=1)

    def get_F1_Score(self, Prec, Recall):                               # :56-58
        return 2 * (Prec * Recall) / (Prec + Recall)

    def get_IoU_Class(self):                                            # :67-71
        cm = self.confusion_matrix
        return np.diag(cm) / (np.sum(cm, axis=1) + np.sum(cm, axis=0) - np.diag(cm))

    def get_mIoU(self):                                                 # :60-65
        return np.nanmean(self.get_IoU_Class())

    def get_FWIoU(self):                                                # :73-80
        cm = self.confusion_matrix
        freq = np.sum(cm, axis=1) / np.sum(cm)
        iu = self.get_IoU_Class()
        return (freq[freq > 0] * iu[freq > 0]).sum()

    def get_Dice_Score(self):                                           # :82-84
        cm = self.confusion_matrix
        return 2 * np.diag(cm) / (np.sum(cm, axis=1) + np.sum(cm, axis=0))


def synthetic_batch(batch: int, size: int, seed: int = 0, in_ch: int = 3, p_label: float = 0.4):
    """Synthetic inputs of the 200x_256 shape family: input U(-1,1) = Normalization(0.5,0.5)
    of uniform [0,1] images (train.py:367), label Bernoulli(p) float32."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, in_ch, size, size, generator=g) * 2 - 1
    g2 = torch.Generator().manual_seed(seed + 1)
    label = (torch.rand(batch, size, size, generator=g2) < p_label).float()
    return x, label


# ------------------------------------------------------------------ input transforms (SURVEY.md §8(f) "next" #2)
def input_lut(mean: float = 0.5, std: float = 0.5) -> np.ndarray:
    """float32 value of every input byte after PatchDataset.__getitem__'s ``input/255.0`` -> float32
    (utils/data_utils.py:216-217) and Normalization (:100), with numpy's own arithmetic."""
    x = (np.arange(256, dtype=np.uint8) / 255.0).astype(np.float32)
    return ((x - mean) / std).astype(np.float32)


def input_transform(img_u8: np.ndarray, label_u8: np.ndarray, flip: int = 0, mean: float = 0.5, std: float = 0.5):
    """One sample through PatchDataset.__getitem__ (:216-219), Normalization (:94-105), RandomFlip (:107-126, with the
    two coin flips decided by the caller: bit 0 = left-right, bit 1 = up-down) and ToTensor (:159-168).
    img_u8 [H,W,3], label_u8 [H,W] -> (float32 [3,H,W], int64 [H,W])."""
    x, lab = img_u8 / 255.0, label_u8 / 255.0
    x, lab = x.astype(np.float32), lab.astype(np.uint8)
    x = (x - mean) / std
    if flip & 1:
        lab, x = np.fliplr(lab).copy(), np.fliplr(x)
    if flip & 2:
        lab, x = np.flipud(lab).copy(), np.flipud(x)
    return np.ascontiguousarray(x.transpose((2, 0, 1)).astype(np.float32)), lab.astype(np.int64)


# ------------------------------------------------------------------ cross-entropy variant (SURVEY.md §8(f) "next" #3)
def selective_risk_ce(output, selection, target, target_coverage=0.8, lamb=8) -> Tuple[torch.Tensor, torch.Tensor]:
    """calc_selective_risk_image, selective_loss.py:24-56 (hard_selection=False): output/selection (N,C,H,W),
    target (N,H,W) int64 class indices."""
    onehot = torch.zeros(target.size(0), output.size(1), target.size(1), target.size(2)).scatter_(
        1, target.view(target.size(0), 1, target.size(1), target.size(2)), 1)               # :40-41
    sel = F.softmax(selection, dim=1)[:, 1, :, :]                                           # :43
    coverage = torch.mean(sel)                                                              # :44
    risk = -torch.mean(torch.sum(F.log_softmax(output, dim=1) * onehot, dim=1) * sel) / coverage   # :52
    diff = torch.clamp(target_coverage - coverage, min=0)                                   # :53
    return risk + lamb * diff * diff, coverage                                              # :54-56


def train_losses_ce(sd, x, label, s_lamb=2, update_running=True):
    """train.py:193-201 with --model_arch UNet --loss CE --selective 1: label is int64 (N,H,W)."""
    out, sel, aux = unet_b_forward(sd, x, True, True, update_running)
    aux_loss = F.cross_entropy(aux, label)                                                  # train.py:80,195
    s_loss, cov = selective_risk_ce(out, sel, label, lamb=s_lamb)
    return aux_loss + s_loss, dict(output=out, selection=sel, aux=aux, aux_loss=aux_loss, select_loss=s_loss,
                                   coverage=cov)


def postprocess_ce(output: np.ndarray, selection: Optional[np.ndarray]):
    """train.py:216-226 for (N,C,H,W) outputs: argmax over the class axis (first maximum wins ties)."""
    pred = np.argmax(output.transpose(0, 2, 3, 1), axis=-1).astype("uint8")
    sel = None if selection is None else np.argmax(selection.transpose(0, 2, 3, 1), -1).astype("uint8")
    return pred, sel
