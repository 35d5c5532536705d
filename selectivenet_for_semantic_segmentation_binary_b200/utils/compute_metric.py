"""``Evaluator`` with the reference's interface, backed by exact integer counters on the GPU.

Mirrors /root/reference/utils/compute_metric.py:4-84.  The reference copies logits to the host,
applies a numpy sigmoid + threshold (train.py:211-235, eval.py:228-247) and counts with
``np.bincount`` on one core (~0.4-0.5 s per 128-patch batch).  Here thresholding and counting are
one coalesced integer-histogram kernel over the logits (``add_batch_from_logits``); the derived
metrics keep the reference's float64 numpy formulas verbatim, so printed values are
bit-identical given identical logits.

``add_batch(label, pred, selection)`` keeps the reference signature; numpy inputs are uploaded
and counted by the same kernel (there is no host counting path).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import kernels as K


def _sigmoid_np(x: np.ndarray, dtype) -> np.ndarray:
    if dtype == np.float64:
        return 1 / (1 + np.exp(-x.astype('float64')))      # train.py:150
    return 1 / (1 + np.exp(-x))                            # eval.py:175 (float32 in, float32 out)


_THR_CACHE = {}


def logit_threshold(cut_off: float, path: str = 'train', scale: str = 'sigmoid') -> float:
    """Smallest float32 logit x that the reference's host code classifies as 1.

    The reference decides ``sigmoid_numpy(x) > cut_off`` in float64 (train.py:150,155) or
    float32 (eval.py:175,179); that is *not* ``x > 0`` (e.g. 1.5612511e-16 vs 1.2318974e-07 for
    cut 0.5).  numpy's sigmoid is monotone in x, so the decision is ``x >= x*`` with x* found by
    bisection over float32 bit patterns against numpy itself.  With ``scale != 'sigmoid'`` the
    reference compares the raw logit: ``x > cut_off`` == ``x >= nextafter(cut_off)``.
    """
    key = (float(cut_off), path, scale)
    if key in _THR_CACHE:
        return _THR_CACHE[key]
    if scale != 'sigmoid':
        # numpy >= 2 compares a float32 array with a python float in float32 (NEP 50):
        # x > float32(cut_off)  ==  x >= the next float32 above it
        c = np.float32(cut_off)
        thr = np.nextafter(c, np.float32(np.inf))
        _THR_CACHE[key] = float(thr)
        return float(thr)
    dt = np.float64 if path == 'train' else np.float32

    def to_key(x) -> int:
        b = int(np.array([x], dtype=np.float32).view(np.uint32)[0])
        return (~b & 0xFFFFFFFF) if (b & 0x80000000) else (b | 0x80000000)

    def from_key(k: int) -> np.float32:
        b = (k ^ 0x80000000) if (k & 0x80000000) else (~k & 0xFFFFFFFF)
        return np.array([b], dtype=np.uint32).view(np.float32)[0]

    def is_one(k: int) -> bool:
        with np.errstate(over='ignore'):
            return bool(_sigmoid_np(np.array([from_key(k)], dtype=np.float32), dt)[0] > cut_off)

    lo, hi = to_key(np.float32(-200.0)), to_key(np.float32(200.0))
    if is_one(lo):
        thr = float('-inf')
    elif not is_one(hi):
        thr = float('inf')
    else:
        while hi - lo > 1:
            mid = (lo + hi) // 2
            if is_one(mid):
                hi = mid
            else:
                lo = mid
        thr = float(from_key(hi))
    _THR_CACHE[key] = thr
    return thr


class Evaluator(object):
    def __init__(self, num_class, selective, device=None):
        if num_class != 2:
            raise ValueError("B200-native Evaluator covers the binary path (num_class == 2) of UNet_B/BCElogit")
        self.num_class = num_class
        self.selective = selective  # (N, H, W)
        self._device = torch.device(device) if device is not None else None
        self._counts = None          # int64[6] on the device: cm00 cm01 cm10 cm11 selected total
        self._host_cm = np.zeros((self.num_class,) * 2)

    # ------------------------------------------------------------------ device counters
    def _ensure(self, device):
        if self._counts is None or self._counts.device != device:
            old = None if self._counts is None else self._counts.cpu()
            self._counts = torch.zeros(6, dtype=torch.int64, device=device)
            if old is not None:
                self._counts += old.to(device)
            self._device = device
        return self._counts

    def _as_cuda(self, a, dtype=None):
        if isinstance(a, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(a))
        else:
            t = a
        dev = self._device
        if dev is None:
            dev = t.device if t.is_cuda else torch.device('cuda', torch.cuda.current_device())
        if dtype is not None:
            t = t.to(dtype)
        return t.to(dev).contiguous()

    def add_batch_from_logits(self, label, output, selection=None, cut_off=0.5, s_cut_off=0.5, path='train',
                              scale='sigmoid'):
        """Fused thresholding + counting on device tensors (the fast path used by train.py / eval.py).

        label: uint8 / float32 {0,1} / int64, output/selection: fp32 logits, all ``[N,H,W]`` on the GPU.
        Also accumulates the reference's ``total`` / ``reject`` bookkeeping (train.py:233-235)."""
        assert label.shape == output.shape
        out = output.detach().to(torch.float32).contiguous()
        sel = None if selection is None else selection.detach().to(torch.float32).contiguous()
        lab = label.detach()
        if lab.dtype not in (torch.uint8, torch.float32, torch.int64):
            lab = lab.to(torch.int64)
        counts = self._ensure(out.device)
        masked = bool(self.selective)
        if masked and sel is None:
            raise ValueError("Evaluator(selective=True).add_batch needs a selection map")
        K.metric_hist(out, sel, lab.contiguous(), logit_threshold(cut_off, path, scale),
                      logit_threshold(s_cut_off, path, scale), masked, counts)

    def add_batch(self, label, pred, selection=None):
        """Reference signature (compute_metric.py:24-26): label/pred uint8 masks, selection {0.,1.}."""
        assert label.shape == pred.shape  # (N, H, W)
        lab = self._as_cuda(label)
        if lab.dtype not in (torch.uint8, torch.float32, torch.int64):
            lab = lab.to(torch.int64)
        prd = self._as_cuda(pred, torch.float32)
        sel = None if selection is None else self._as_cuda(selection, torch.float32)
        counts = self._ensure(prd.device)
        masked = bool(self.selective)
        if masked and sel is None:
            raise TypeError("selection is required when Evaluator(selective=True)")
        # masks are already thresholded: pred in {0,1} -> (pred >= 0.5); selection == 1 -> (sel >= 1)
        K.metric_hist(prd, sel, lab, 0.5, 1.0, masked, counts)

    # ------------------------------------------------------------------ host view
    @property
    def confusion_matrix(self):
        cm = self._host_cm.copy()
        if self._counts is not None:
            c = self._counts.cpu().numpy()
            cm = cm + c[:4].reshape(2, 2).astype(np.float64)
        return cm

    @confusion_matrix.setter
    def confusion_matrix(self, value):
        self._host_cm = np.array(value, dtype=np.float64)
        if self._counts is not None:
            self._counts[:4].zero_()

    @property
    def total(self) -> int:
        return 0 if self._counts is None else int(self._counts[5].item())

    @property
    def total_reject(self) -> int:
        """pixels seen - pixels selected (train.py:233-235, eval.py:245-247)"""
        if self._counts is None:
            return 0
        c = self._counts.cpu()
        return int(c[5] - c[4])

    def counts_tensor(self) -> Optional[torch.Tensor]:
        """int64[6] device counters (cm00, cm01, cm10, cm11, selected, total) for all-reduce across ranks."""
        return self._counts

    def reset(self):
        self._host_cm = np.zeros((self.num_class,) * 2)
        if self._counts is not None:
            self._counts.zero_()

    # ------------------------------------------------------------------ metrics (formulas verbatim, float64)
    def Confusion_Matrix(self):
        print(self.confusion_matrix)
        return self.confusion_matrix

    def get_Pixel_Accuracy(self):
        cm = self.confusion_matrix
        Acc = np.diag(cm).sum() / cm.sum()
        return Acc

    def get_Pixel_Accuracy_Class(self):
        cm = self.confusion_matrix
        Acc = np.diag(cm) / cm.sum(axis=1)
        Acc = np.nanmean(Acc)
        return Acc

    def get_Pixel_Accuracy_Class_S(self):
        cm = self.confusion_matrix
        Acc = np.diag(cm) / cm.sum(axis=1)
        return Acc

    def get_Precision(self):
        cm = self.confusion_matrix
        Prec = np.diag(cm) / cm.sum(axis=0)
        return Prec

    def get_Recall(self):
        cm = self.confusion_matrix
        Recall = np.diag(cm) / cm.sum(axis=1)
        return Recall

    def get_F1_Score(self, Prec, Recall):
        F1_score = 2 * (Prec * Recall) / (Prec + Recall)
        return F1_score

    def get_mIoU(self):
        cm = self.confusion_matrix
        MIoU = np.diag(cm) / (np.sum(cm, axis=1) + np.sum(cm, axis=0) - np.diag(cm))
        MIoU = np.nanmean(MIoU)
        return MIoU

    def get_IoU_Class(self):
        cm = self.confusion_matrix
        MIoU = np.diag(cm) / (np.sum(cm, axis=1) + np.sum(cm, axis=0) - np.diag(cm))
        return MIoU

    def get_FWIoU(self):
        cm = self.confusion_matrix
        freq = np.sum(cm, axis=1) / np.sum(cm)
        iu = np.diag(cm) / (np.sum(cm, axis=1) + np.sum(cm, axis=0) - np.diag(cm))
        FWIoU = (freq[freq > 0] * iu[freq > 0]).sum()
        return FWIoU

    def get_Dice_Score(self):
        cm = self.confusion_matrix
        dice_score = 2 * np.diag(cm) / (np.sum(cm, axis=1) + np.sum(cm, axis=0))
        return dice_score
