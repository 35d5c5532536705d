"""fp32 CHECK MODE of the execution plan (``SUNET_CHECK_FP32=1``).

BASELINE.json's tolerance clause: logits and loss within 2e-2 relative in bf16, *or 1e-4 in an fp32 check mode*.
``SUNetPlanF32`` has the interface of :class:`~.engine.SUNetPlan` (``forward`` / ``backward`` / ``logits`` / ``P``), so
``UNet_B``, the loss functions and ``SUNetTrainer`` run unchanged on top of it, but every op is a slow SIMT fp32 kernel
(``csrc/check_fp32.cu``): NHWC fp32 activations, parameters read in the reference's own layouts, fp64 accumulation
for the reductions over pixels.  With fp32 numerics the whole step agrees with the oracle to ~1e-6, which is what lets
``tests/test_gpu_check_fp32.py`` pin the ORCHESTRATION of the step (layer order, concat order ``[up | skip]``,
BatchNorm bookkeeping, max-pool gradient routing, loss plumbing, gradient layout) with a gradient cosine bound of
0.99999 instead of the 0.90 that bf16 allows.  Layer graph: /root/reference/model.py:68-103.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib
from . import kernels as K

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
_CH = {1: 64, 2: 128, 3: 256, 4: 512}


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous() and t.dtype in (torch.float32, torch.int64), (t.dtype, t.shape)
    return t.data_ptr()


def _call(name, *args):
    _lib.check(getattr(_lib.load(), name)(*args, torch.cuda.current_stream().cuda_stream), name)


class SUNetPlanF32:
    def __init__(self, batch: int, height: int, width: int, in_ch: int, selective: bool, device, n_cls: int = 1):
        if height % 8 or width % 8:
            raise ValueError("UNet_B needs H and W divisible by 8 (three 2x2 pools)")
        self.B, self.H, self.W, self.in_ch, self.selective = batch, height, width, in_ch, selective
        self.device = torch.device(device)
        self.n_cls = n_cls
        self.nheads = 3 if selective else 1
        self.P = batch * height * width
        self.hw = {L: (height >> (L - 1), width >> (L - 1)) for L in (1, 2, 3, 4)}
        self.logits = torch.empty(self.n_cls * self.nheads, self.P, device=self.device)
        self.ws = K.new_workspace(self.device)
        self.generation = 0
        self.t: Dict[str, torch.Tensor] = {}      # every intermediate of the last forward, by name

    def _act(self, L, ch):
        h, w = self.hw[L]
        return torch.empty(self.B, h, w, ch, device=self.device)

    # ------------------------------------------------------------------ forward
    def _cbr(self, name, L, src0, src1, params, buffers, training, pool=False):
        B = self.B
        h, w = self.hw[L]
        wt = params[f"{name}.0.weight"]
        cout = wt.shape[0]
        c0, c1 = src0.shape[3], (0 if src1 is None else src1.shape[3])
        y = self._act(L, cout)
        _call("sunet_f32_conv3x3_fwd", _p(src0), c0, _p(src1), c1, _p(wt), None, _p(y), B, h, w, cout)
        scale, shift, mean, invstd = (torch.empty(cout, device=self.device) for _ in range(4))
        if training:
            _call("sunet_f32_bn_stats", _p(y), B * h * w, cout, _p(params[f"{name}.1.weight"]),
                  _p(params[f"{name}.1.bias"]), _p(params[f"{name}.0.bias"]), _p(buffers[f"{name}.1.running_mean"]),
                  _p(buffers[f"{name}.1.running_var"]), _p(buffers[f"{name}.1.num_batches_tracked"]), BN_MOMENTUM,
                  BN_EPS, _p(scale), _p(shift), _p(mean), _p(invstd))
        else:
            K.bn_eval_affine(params[f"{name}.1.weight"], params[f"{name}.1.bias"], params[f"{name}.0.bias"],
                             buffers[f"{name}.1.running_mean"], buffers[f"{name}.1.running_var"], BN_EPS, scale, shift)
        a = self._act(L, cout)
        pooled = self._act(L + 1, cout) if pool else None
        _call("sunet_f32_bn_relu_pool", _p(y), _p(scale), _p(shift), _p(a), _p(pooled), B, h, w, cout)
        self.t[name] = dict(y=y, a=a, scale=scale, shift=shift, mean=mean, invstd=invstd, src0=src0, src1=src1,
                            pooled=pooled, L=L)
        return a, pooled

    def forward(self, x: torch.Tensor, params, buffers, training: bool) -> torch.Tensor:
        if x is None:
            raise RuntimeError("fp32 check mode takes float32 NCHW inputs (no uint8 pipeline)")
        assert x.shape == (self.B, self.in_ch, self.H, self.W)
        self.t = {}
        x0 = x.permute(0, 2, 3, 1).contiguous()
        a, _ = self._cbr("encoder_layer_1_1", 1, x0, None, params, buffers, training)
        e1, p1 = self._cbr("encoder_layer_1_2", 1, a, None, params, buffers, training, pool=True)
        a, _ = self._cbr("encoder_layer_2_1", 2, p1, None, params, buffers, training)
        e2, p2 = self._cbr("encoder_layer_2_2", 2, a, None, params, buffers, training, pool=True)
        a, _ = self._cbr("encoder_layer_3_1", 3, p2, None, params, buffers, training)
        e3, p3 = self._cbr("encoder_layer_3_2", 3, a, None, params, buffers, training, pool=True)
        a, _ = self._cbr("decoder_layer_4_2", 4, p3, None, params, buffers, training)
        a, _ = self._cbr("decoder_layer_4_1", 4, a, None, params, buffers, training)
        for L, skip in ((3, e3), (2, e2), (1, e1)):
            hh, ww = self.hw[L + 1]
            up = self._act(L, _CH[L])
            _call("sunet_f32_convT_fwd", _p(a), _p(params[f"unpool{L}.weight"]), _p(params[f"unpool{L}.bias"]), _p(up),
                  self.B, hh, ww, _CH[L + 1], _CH[L])
            self.t[f"unpool{L}"] = dict(x=a, up=up)
            a, _ = self._cbr(f"decoder_layer_{L}_2", L, up, skip, params, buffers, training)     # concat = [up | skip]
            a, _ = self._cbr(f"decoder_layer_{L}_1", L, a, None, params, buffers, training)
        heads = ["conv1x1"] + (["conv_select", "conv_aux"] if self.selective else [])
        for c in range(self.n_cls):
            wh = torch.stack([params[f"{h}.weight"][c].reshape(-1) for h in heads]).contiguous()
            bh = torch.stack([params[f"{h}.bias"][c] for h in heads]).contiguous()
            _call("sunet_f32_heads_fwd", _p(a), _p(wh), _p(bh), self.nheads,
                  self.logits[c * self.nheads:(c + 1) * self.nheads].data_ptr(), self.P, 64)
        if training:
            self.generation += 1
        return self.logits

    # ------------------------------------------------------------------ backward
    def _cbr_bwd(self, name, dA, dPool, params, grads, need_dx=True):
        """(dA, dPool) -> dgamma, dbeta, dW and the gradient(s) w.r.t. the block's input(s)."""
        t = self.t[name]
        B = self.B
        h, w = self.hw[t["L"]]
        y = t["y"]
        cout = y.shape[3]
        dy = torch.empty_like(y)
        _call("sunet_f32_bn_relu_pool_bwd", _p(dA), _p(dPool), _p(y), _p(t["a"]), _p(t["scale"]), _p(t["mean"]),
              _p(t["invstd"]), _p(grads[f"{name}.1.weight"]), _p(grads[f"{name}.1.bias"]), _p(dy), B, h, w, cout,
              self.ws.data_ptr(), self.ws.numel())
        src0, src1 = t["src0"], t["src1"]
        c0, c1 = src0.shape[3], (0 if src1 is None else src1.shape[3])
        _call("sunet_f32_conv3x3_wgrad", _p(dy), _p(src0), c0, _p(src1), c1, _p(grads[f"{name}.0.weight"]), B, h, w, cout)
        # the conv bias feeds BatchNorm: its gradient (sum of dy over pixels) is identically zero — left untouched
        if not need_dx:
            return None, None
        dx0 = torch.empty_like(src0)
        dx1 = None if src1 is None else torch.empty_like(src1)
        _call("sunet_f32_conv3x3_dgrad", _p(dy), _p(params[f"{name}.0.weight"]), _p(dx0), c0, _p(dx1), c1, B, h, w, cout)
        return dx0, dx1

    def backward(self, dlogits: torch.Tensor, params, grads, on_group_done=None) -> None:
        heads = ["conv1x1"] + (["conv_select", "conv_aux"] if self.selective else [])
        nh = self.nheads
        a_last = self.t["decoder_layer_1_1"]["a"]
        dA = torch.empty_like(a_last)
        for c in range(self.n_cls):
            wh = torch.stack([params[f"{h}.weight"][c].reshape(-1) for h in heads]).contiguous()
            dw = torch.empty(nh, 64, device=self.device)
            db = torch.empty(nh, device=self.device)
            _call("sunet_f32_heads_bwd", dlogits[c * nh:(c + 1) * nh].data_ptr(), _p(a_last), _p(wh), nh, _p(dA),
                  1 if c > 0 else 0, _p(dw), _p(db), self.P, 64)
            for i, h in enumerate(heads):
                grads[f"{h}.weight"][c].reshape(-1).copy_(dw[i])
                grads[f"{h}.bias"][c:c + 1].copy_(db[i:i + 1])
        d_skip = {}
        for L in (1, 2, 3):
            dA, _ = self._cbr_bwd(f"decoder_layer_{L}_1", dA, None, params, grads)
            d_up, d_skip[L] = self._cbr_bwd(f"decoder_layer_{L}_2", dA, None, params, grads)
            u = self.t[f"unpool{L}"]
            hh, ww = self.hw[L + 1]
            _call("sunet_f32_convT_wgrad", _p(d_up), _p(u["x"]), _p(grads[f"unpool{L}.weight"]),
                  _p(grads[f"unpool{L}.bias"]), self.B, hh, ww, _CH[L + 1], _CH[L])
            dA = torch.empty_like(u["x"])
            _call("sunet_f32_convT_dgrad", _p(d_up), _p(params[f"unpool{L}.weight"]), _p(dA), self.B, hh, ww, _CH[L + 1],
                  _CH[L])
            if on_group_done is not None:
                on_group_done(f"dec{L}")
        dA, _ = self._cbr_bwd("decoder_layer_4_1", dA, None, params, grads)
        d_pool, _ = self._cbr_bwd("decoder_layer_4_2", dA, None, params, grads)
        if on_group_done is not None:
            on_group_done("dec4")
        for L in (3, 2, 1):
            dA, _ = self._cbr_bwd(f"encoder_layer_{L}_2", d_skip[L], d_pool, params, grads)
            d_pool, _ = self._cbr_bwd(f"encoder_layer_{L}_1", dA, None, params, grads, need_dx=(L > 1))
            if on_group_done is not None:
                on_group_done(f"enc{L}")
