"""SelectiveNet losses with the reference's signatures, computed by fused CUDA reductions.

Mirrors /root/reference/selective_loss.py:58-85 (``calc_selective_risk_image_b``) and the
``torch.nn.BCEWithLogitsLoss()`` instance of /root/reference/train.py:78.

Both losses are two-phase: phase 1 reduces (sum sigmoid(sel), sum bce*sigmoid(sel), sum bce(aux))
in one pass over the logits; phase 2 (backward) writes per-pixel gradients from those sums.
Under batch-sharded data parallelism the sums — and the pixel count — are all-reduced between
the phases (``set_data_parallel_group``), because the reference computes its losses on the
gathered *global* batch (train.py:194-201): coverage and the risk ratio are global quantities.

Numerics: the BCE term uses the stable softplus form; the reference's naive
``log(sigmoid(x))`` (selective_loss.py:79-80) agrees with it to fp32 rounding wherever the
naive form is finite and overflows to inf/NaN for |x| >~ 17..104.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import kernels as K

_DP_GROUP = None          # torch.distributed process group, or None
_DP_ENABLED = False


def set_data_parallel_group(group=None, enabled: bool = True) -> None:
    """Make the losses global-batch losses across `group` (default group if None)."""
    global _DP_GROUP, _DP_ENABLED
    _DP_GROUP, _DP_ENABLED = group, enabled


_WS = {}


def _workspace(device) -> torch.Tensor:
    ws = _WS.get(device)
    if ws is None:
        ws = K.new_workspace(device)
        _WS[device] = ws
    return ws


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("B200-native losses have no CPU path: tensors must be on a CUDA device")


def _global_sums(sums: torch.Tensor, local_pixels: int) -> int:
    """All-reduce [S, R, A, pixels] (one fp64 buffer) when data parallel; returns the global pixel count.
    `sums` is the fp64 [4] buffer sunet_loss_sums filled (sums[3] = this shard's pixel count)."""
    if not _DP_ENABLED:
        return local_pixels
    from .trainer import exchange_loss_sums
    exchange_loss_sums(sums, _DP_GROUP)
    return int(round(sums[3].item()))


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


class _SelectiveRisk(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, selection, target, target_coverage, lamb):
        _require_cuda(output, selection, target)
        out, sel, tgt = _f32c(output), _f32c(selection), _f32c(target)
        dev = out.device
        sums = torch.zeros(4, dtype=torch.float64, device=dev)
        K.loss_sums(out, sel, None, tgt, sums, _workspace(dev), pixels_out=sums[3:4])
        P = _global_sums(sums, tgt.numel())
        res = torch.empty(4, device=dev)
        K.loss_finalize(sums, P, lamb, target_coverage, res)
        ctx.save_for_backward(out, sel, tgt, sums)
        ctx.meta = (P, float(lamb), float(target_coverage), output.shape)
        loss, coverage = res[0].clone(), res[1].clone()
        ctx.mark_non_differentiable(coverage)
        return loss, coverage

    @staticmethod
    def backward(ctx, g_loss, _g_cov):
        out, sel, tgt, sums = ctx.saved_tensors
        P, lamb, tc, shape = ctx.meta
        d_out = torch.empty_like(out)
        d_sel = torch.empty_like(sel)
        g = g_loss.detach().to(torch.float32).reshape(1).contiguous()
        K.loss_bwd(out, sel, None, tgt, sums, P, lamb, tc, g, None, d_out, d_sel, None)
        return d_out.view(shape), d_sel.view(shape), None, None, None


def calc_selective_risk_image_b(output, selection, target, target_coverage=0.8, lamb=8, hard_selection=False):
    """
    the modificated selective risk for image segmentation with BCEwithLogitLoss (Binary Class)

    Args
        output: (N, H, W)
        selection: (N, H, W)
        target: (N, H, W)
    Return
        selective loss, coverage          (selective_loss.py:58-85)
    """
    if hard_selection:
        # never used by train.py; the reference branch prints and re-wraps tensors
        raise NotImplementedError("hard_selection=True is dead code in the reference and is not implemented")
    return _SelectiveRisk.apply(output, selection, target, float(target_coverage), float(lamb))


class _BCEMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target):
        _require_cuda(logits, target)
        x, tgt = _f32c(logits), _f32c(target)
        dev = x.device
        sums = torch.zeros(4, dtype=torch.float64, device=dev)
        K.loss_sums(None, None, x, tgt, sums, _workspace(dev), pixels_out=sums[3:4])
        P = _global_sums(sums, tgt.numel())
        res = torch.empty(4, device=dev)
        K.loss_finalize(sums, P, 0.0, 0.0, res)
        ctx.save_for_backward(x, tgt, sums)
        ctx.meta = (P, logits.shape)
        return res[2].clone()

    @staticmethod
    def backward(ctx, g):
        x, tgt, sums = ctx.saved_tensors
        P, shape = ctx.meta
        d = torch.empty_like(x)
        gg = g.detach().to(torch.float32).reshape(1).contiguous()
        K.loss_bwd(None, None, x, tgt, sums, P, 0.0, 0.0, None, gg, None, None, d)
        return d.view(shape), None


class BCEWithLogitsLoss(torch.nn.Module):
    """Drop-in for the ``torch.nn.BCEWithLogitsLoss()`` of train.py:78 (mean reduction, no weights)."""

    def forward(self, input, target):
        return _BCEMean.apply(input, target)


# ------------------------------------------------------------------ cross-entropy variant (`UNet`, n_cls = 2)
# For two classes, log_softmax(z)[t] = -bce_with_logits(z1 - z0, t) and softmax(s)[1] = sigmoid(s1 - s0), so the
# reference's CE losses are the binary losses above applied to the channel differences; autograd carries the
# gradient back through the subtraction (d/dz1 = g, d/dz0 = -g).
def _logit_difference(x: torch.Tensor) -> torch.Tensor:
    if x.dim() != 4 or x.shape[1] != 2:
        raise NotImplementedError("only two-class (N, 2, H, W) logits are implemented")
    return x[:, 1] - x[:, 0]


def calc_selective_risk_image(output, selection, target, target_coverage=0.8, lamb=8, hard_selection=False):
    """
    the modificated selective risk for image segmentation with Cross Entropy Loss (2 classes)

    Args
        output: (N, 2, H, W)
        selection: (N, 2, H, W)
        target: (N, H, W) class indices, or (N, 2, H, W) one-hot
    Return
        selective loss, coverage          (selective_loss.py:24-56)
    """
    if hard_selection:
        raise NotImplementedError("hard_selection=True is dead code in the reference and is not implemented")
    if target.dim() == 4:
        target = target[:, 1]
    return _SelectiveRisk.apply(_logit_difference(output), _logit_difference(selection), target.to(torch.float32),
                                float(target_coverage), float(lamb))


class CrossEntropyLoss(torch.nn.Module):
    """Drop-in for the ``torch.nn.CrossEntropyLoss()`` of train.py:80 on (N, 2, H, W) logits and (N, H, W) class
    indices (mean reduction, no weights)."""

    def forward(self, input, target):
        return _BCEMean.apply(_logit_difference(input), target.to(torch.float32))
