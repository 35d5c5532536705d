"""``UNet_B`` with the reference's constructor, forward signature and state_dict, executed by
hand-written sm_100a kernels.

Mirrors /root/reference/model.py:9-15 (``CBR_2D``) and :18-103 (``UNet_B``): same attribute
names, same registration order (so ``torch.manual_seed(s); UNet_B(...)`` draws identical
initial weights and ``state_dict()`` has the reference's 110 keys / 106 without the selective
heads), same outputs — ``[N,H,W]`` logits, or ``(output, select, aux)`` when ``selective``.

The ``nn.Conv2d`` / ``nn.BatchNorm2d`` / ``nn.ConvTranspose2d`` children only *hold* parameters
and buffers (fp32, reference layouts — that is the checkpoint contract); their ``forward`` is
never called.  ``UNet_B.forward`` runs a :class:`~.engine.SUNetPlan`.  There is no CPU path:
inputs must live on a CUDA device.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn as nn

import os

from .engine import FlatGrads, SUNetPlan, param_order


def _plan_class():
    """SUNET_CHECK_FP32=1: the slow fp32 check mode (engine_fp32.SUNetPlanF32) instead of the bf16 tensor-core plan."""
    if os.environ.get("SUNET_CHECK_FP32", "0") != "0":
        from .engine_fp32 import SUNetPlanF32
        return SUNetPlanF32
    return SUNetPlan


def CBR_2D(in_ch, out_ch, k_size=3, stride=1, padding=1, bias=True):
    """Conv2d -> BatchNorm2d -> ReLU container (model.py:9-15).  Parameter holder only."""
    layers = []
    layers += [nn.Conv2d(in_channels=in_ch, out_channels=out_ch, kernel_size=k_size, stride=stride, padding=padding,
                         bias=bias)]
    layers += [nn.BatchNorm2d(num_features=out_ch)]
    layers += [nn.ReLU()]
    return nn.Sequential(*layers)


class _UNetBFunction(torch.autograd.Function):
    """Autograd bridge: forward/backward are sequences of C-ABI kernel launches."""

    @staticmethod
    def forward(ctx, net: "UNet_B", x: torch.Tensor, training: bool, *param_tensors):
        plan = net._plan_for(x)
        params, buffers = net._param_dict(), net._buffer_dict()
        logits = plan.forward(x.contiguous(), params, buffers, training)
        ctx.net, ctx.plan, ctx.generation = net, plan, plan.generation
        B, H, W = plan.B, plan.H, plan.W
        outs = tuple(logits[h].view(B, H, W).clone() for h in range(plan.n_cls * plan.nheads))
        return outs

    @staticmethod
    def backward(ctx, *douts):
        net, plan = ctx.net, ctx.plan
        if plan.generation != ctx.generation:
            raise RuntimeError("UNet_B.backward: the activation buffers of this forward pass were overwritten by a "
                               "later forward pass (one in-flight training step per model/shape)")
        P = plan.P
        dl = torch.empty(plan.n_cls * plan.nheads, P, device=plan.device)
        for h, d in enumerate(douts):
            if d is None:
                dl[h].zero_()
            else:
                dl[h].copy_(d.reshape(P))
        params = net._param_dict()
        fg = net._flat_grads()
        plan.backward(dl, params, fg.views)
        # hand autograd a private copy so .grad accumulation can never alias the plan's buffer
        flat = fg.flat.clone()
        grads = tuple(flat[o:o + k].view_as(params[n]) for n, (o, k) in ((n, fg.offsets[n]) for n in fg.order))
        return (None, None, None) + grads


class UNet_B(nn.Module):
    _N_CLS = 1          # output channels per head; the `UNet` subclass (CE variant) has 2

    def __init__(self, input_type='RGB', selective=False):
        super(UNet_B, self).__init__()
        self.selective = selective
        if 'RGB' in input_type:
            input_ch = 3
        elif input_type == 'GH':
            input_ch = 2
        else:
            raise ValueError(f"unknown input_type {input_type!r}")
        self.input_ch = input_ch

        self.encoder_layer_1_1 = CBR_2D(in_ch=input_ch, out_ch=64)
        self.encoder_layer_1_2 = CBR_2D(in_ch=64, out_ch=64)
        self.pool1 = nn.MaxPool2d(kernel_size=2)
        self.encoder_layer_2_1 = CBR_2D(in_ch=64, out_ch=128)
        self.encoder_layer_2_2 = CBR_2D(in_ch=128, out_ch=128)
        self.pool2 = nn.MaxPool2d(kernel_size=2)
        self.encoder_layer_3_1 = CBR_2D(in_ch=128, out_ch=256)
        self.encoder_layer_3_2 = CBR_2D(in_ch=256, out_ch=256)
        self.pool3 = nn.MaxPool2d(kernel_size=2)
        self.decoder_layer_4_2 = CBR_2D(in_ch=256, out_ch=512)
        self.decoder_layer_4_1 = CBR_2D(in_ch=512, out_ch=512)
        self.unpool3 = nn.ConvTranspose2d(in_channels=512, out_channels=256, kernel_size=2, stride=2, padding=0,
                                          bias=True)
        self.decoder_layer_3_2 = CBR_2D(in_ch=512, out_ch=256)
        self.decoder_layer_3_1 = CBR_2D(in_ch=256, out_ch=256)
        self.unpool2 = nn.ConvTranspose2d(in_channels=256, out_channels=128, kernel_size=2, stride=2, padding=0,
                                          bias=True)
        self.decoder_layer_2_2 = CBR_2D(in_ch=256, out_ch=128)
        self.decoder_layer_2_1 = CBR_2D(in_ch=128, out_ch=128)
        self.unpool1 = nn.ConvTranspose2d(in_channels=128, out_channels=64, kernel_size=2, stride=2, padding=0,
                                          bias=True)
        self.decoder_layer_1_2 = CBR_2D(in_ch=128, out_ch=64)
        self.decoder_layer_1_1 = CBR_2D(in_ch=64, out_ch=64)
        self._build_heads()

        self._plans: Dict[Tuple, SUNetPlan] = {}
        self._fg = None

    def _build_heads(self):
        self.conv1x1 = nn.Conv2d(in_channels=64, out_channels=1, kernel_size=1)
        if self.selective:
            self.conv_select = nn.Conv2d(64, 1, 1)
            self.conv_aux = nn.Conv2d(64, 1, 1)

    # ------------------------------------------------------------------ plumbing
    def _param_dict(self) -> Dict[str, torch.Tensor]:
        return dict(self.named_parameters())

    def _buffer_dict(self) -> Dict[str, torch.Tensor]:
        return dict(self.named_buffers())

    def _plan_for_shape(self, batch: int, height: int, width: int, device) -> SUNetPlan:
        """Plan lookup by shape (the uint8 input pipeline has no float32 NCHW tensor to derive it from)."""
        device = torch.device(device)
        key = (batch, height, width, device.index)
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) >= 4:
                self._plans.pop(next(iter(self._plans)))
            plan = _plan_class()(batch, height, width, self.input_ch, self.selective, device, n_cls=self._N_CLS)
            self._plans[key] = plan
        return plan

    def _plan_for(self, x: torch.Tensor) -> SUNetPlan:
        if not x.is_cuda:
            raise RuntimeError("UNet_B (B200-native) has no CPU path: move the input and the model to a CUDA device")
        if x.dim() != 4 or x.shape[1] != self.input_ch:
            raise RuntimeError(f"expected input [N,{self.input_ch},H,W], got {tuple(x.shape)}")
        if x.dtype != torch.float32:
            raise RuntimeError(f"expected float32 input, got {x.dtype}")
        p0 = next(self.parameters())
        if p0.device != x.device:
            raise RuntimeError(f"model on {p0.device} but input on {x.device}")
        key = (x.shape[0], x.shape[2], x.shape[3], x.device.index)
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) >= 4:        # plans own GBs of activations: keep a few shapes only
                self._plans.pop(next(iter(self._plans)))
            plan = _plan_class()(x.shape[0], x.shape[2], x.shape[3], self.input_ch, self.selective, x.device,
                                 n_cls=self._N_CLS)
            self._plans[key] = plan
        return plan

    def _flat_grads(self) -> FlatGrads:
        p0 = next(self.parameters())
        if self._fg is None or self._fg.flat.device != p0.device:
            params = self._param_dict()
            order = param_order(self.selective)
            self._fg = FlatGrads({n: tuple(params[n].shape) for n in order}, order, p0.device)
        return self._fg

    def _apply(self, fn, *args, **kwargs):
        # .to()/.cuda() move parameters: drop device-bound caches
        self._plans = {}
        self._fg = None
        return super()._apply(fn, *args, **kwargs)

    # ------------------------------------------------------------------ forward
    def forward(self, x):
        params = self._param_dict()
        order = param_order(self.selective)
        outs = _UNetBFunction.apply(self, x, self.training, *[params[n] for n in order])
        if self.selective:
            return outs[0], outs[1], outs[2]
        return outs[0]


class UNet(UNet_B):
    """The reference's cross-entropy variant (/root/reference/model.py:106-191): same trunk, heads with ``n_cls``
    output channels — ``conv1x1`` 64 -> n_cls, ``conv_select`` 64 -> 2, ``conv_aux`` 64 -> n_cls — and outputs of shape
    ``(N, C, H, W)``.  Only ``n_cls == 2`` is built (the task and the Evaluator are binary).  Same
    kernels as UNet_B: the fused BN+ReLU+heads pass and the heads backward run once per output channel, the second
    backward call adding the first call's activation gradient before the BatchNorm reduction."""
    _N_CLS = 2

    def __init__(self, input_type='RGB', n_cls=2, selective=False):
        if n_cls != 2:
            raise NotImplementedError("UNet: only n_cls = 2 is implemented (binary segmentation)")
        self.n_cls = n_cls
        super(UNet, self).__init__(input_type, selective)

    def _build_heads(self):
        self.conv1x1 = nn.Conv2d(in_channels=64, out_channels=self.n_cls, kernel_size=1, stride=1)
        if self.selective:
            self.conv_select = nn.Conv2d(64, 2, 1, 1)
            self.conv_aux = nn.Conv2d(64, self.n_cls, 1, 1)

    def forward(self, x):
        params = self._param_dict()
        order = param_order(self.selective)
        rows = _UNetBFunction.apply(self, x, self.training, *[params[n] for n in order])
        nh = 3 if self.selective else 1
        heads = [torch.stack([rows[c * nh + h] for c in range(2)], dim=1) for h in range(nh)]    # (N, 2, H, W)
        if self.selective:
            return heads[0], heads[1], heads[2]
        return heads[0]
