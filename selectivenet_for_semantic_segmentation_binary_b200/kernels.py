"""Typed torch-tensor wrappers over the C ABI (one Python function per entry point).

PyTorch is plumbing here: it owns device memory and streams; every function below only
extracts ``data_ptr()`` / strides and enqueues hand-written sm_100a kernels on the current
stream.  Activations are NHWC bf16 tensors ``[B, H, W, C]`` and may be channel slices of a
wider buffer (``stride(2)`` is the pixel stride).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import A_CONV3X3, A_GATHER2X2, A_PLAIN, D_NHWC, D_SCATTER2X2, WORKSPACE_BYTES  # noqa: F401


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _act(t: torch.Tensor):
    """(ptr, channels, pix_stride) of an NHWC bf16 activation view."""
    assert t.dtype == torch.bfloat16 and t.dim() == 4 and t.is_cuda, (t.dtype, t.shape)
    B, H, W, Cc = t.shape
    assert t.stride(3) == 1
    ps = t.stride(2)
    assert t.stride(1) == W * ps and t.stride(0) == H * W * ps, "activation view must be pixel-contiguous"
    return t.data_ptr(), Cc, ps


def _f32(t: Optional[torch.Tensor]):
    if t is None:
        return None
    assert t.dtype == torch.float32 and t.is_cuda and t.is_contiguous(), (t.dtype, t.shape)
    return t.data_ptr()


def new_workspace(device) -> torch.Tensor:
    return torch.empty(WORKSPACE_BYTES, dtype=torch.uint8, device=device)


# ----------------------------------------------------------------------------- G1
def conv_gemm_stat_rows(batch: int, height: int, width: int, n_total: int, a_mode: int = A_CONV3X3,
                        d_mode: int = D_NHWC, bias: bool = False) -> int:
    """Partial-statistics rows a conv_gemm call of this shape writes (depends on the kernel variant chosen)."""
    a = _lib.ConvGemmArgs()
    a.batch, a.height, a.width, a.n_total, a.a_mode, a.d_mode = batch, height, width, n_total, a_mode, d_mode
    a.bias = 1 if bias else None          # only null / non-null matters here
    r = _lib.load().sunet_conv_gemm_stat_rows(C.byref(a))
    if r <= 0:
        raise _lib.SunetError("conv_gemm_stat_rows: bad shape")
    return r


def _conv_gemm_args(a_mode: int, grid, src0: torch.Tensor, weights: torch.Tensor, dst: torch.Tensor,
                    src1: Optional[torch.Tensor], bias: Optional[torch.Tensor], d_mode: int) -> _lib.ConvGemmArgs:
    a = _lib.ConvGemmArgs()
    a.batch, a.height, a.width = grid
    a.a_mode = a_mode
    a.src0, a.src0_channels, a.src0_pix_stride = _act(src0)
    if src1 is not None:
        a.src1, a.src1_channels, a.src1_pix_stride = _act(src1)
    assert weights.dtype == torch.bfloat16 and weights.dim() == 2 and weights.is_contiguous()
    a.weights = weights.data_ptr()
    a.n_total, a.k_total = weights.shape
    a.bias = _f32(bias)
    a.dst, _, a.dst_pix_stride = _act(dst)
    a.d_mode = d_mode
    return a


def conv_gemm_bnb_supported(a_mode: int, grid, src0, weights, dst, *, src1=None, bias=None,
                            d_mode: int = D_NHWC) -> bool:
    """True if conv_gemm(..., bnb=...) is available for this shape (the CTA-pair halo kernel serves it)."""
    a = _conv_gemm_args(a_mode, grid, src0, weights, dst, src1, bias, d_mode)
    return bool(_lib.load().sunet_conv_gemm_bnb_supported(C.byref(a)))


def conv_gemm_pro_supported(a_mode: int, grid, src0, weights, dst, *, src1=None, bias=None,
                            d_mode: int = D_NHWC) -> bool:
    """True if conv_gemm(..., pro=...) is available for this shape (the CTA-pair halo kernel serves it)."""
    a = _conv_gemm_args(a_mode, grid, src0, weights, dst, src1, bias, d_mode)
    return bool(_lib.load().sunet_conv_gemm_pro_supported(C.byref(a)))


def conv_gemm(a_mode: int, grid, src0: torch.Tensor, weights: torch.Tensor, dst: torch.Tensor, *,
              src1: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
              stats: Optional[torch.Tensor] = None, d_mode: int = D_NHWC, bnb=None, ep=None, pro=None) -> None:
    """grid = (batch, height, width) of the GEMM-M pixel grid.
    bnb = (y, scale, shift, mean, invstd): fuse the BatchNorm-backward reduction of the block whose input
    gradient this launch produces into the epilogue; `stats` then receives (sum g, sum g*xhat) rows.
    pro = (scale, shift[, mask]): the sources in `mask` (bit 0 src0, bit 1 src1; default src0) hold the RAW conv output
    of the previous block, scale/shift are indexed by concatenated input channel; the kernel applies
    relu(scale*y + shift) (BatchNorm(train) + ReLU) to each staged tile in shared memory before the MMAs read it,
    so the activation is never written to HBM."""
    lib = _lib.load()
    a = _conv_gemm_args(a_mode, grid, src0, weights, dst, src1, bias, d_mode)
    if stats is not None:
        rows = _lib.load().sunet_conv_gemm_stat_rows(C.byref(a))
        assert stats.dtype == torch.float32 and stats.is_contiguous() and stats.numel() >= rows * a.n_total * 2
        a.stats = stats.data_ptr()
    if bnb is not None:
        y, scale, shift, mean, invstd = bnb[:5]
        col0 = bnb[5] if len(bnb) > 5 else 0       # the reduction covers output columns [col0, n_total)
        assert stats is not None and y.shape[:3] == dst.shape[:3] and y.shape[3] == dst.shape[3] - col0
        a.bnb_col0 = col0
        a.bnb_y, _, a.bnb_y_pix_stride = _act(y)
        a.bnb_scale, a.bnb_shift, a.bnb_mean, a.bnb_invstd = _f32(scale), _f32(shift), _f32(mean), _f32(invstd)
    if ep is not None:               # inference: dst = relu(acc * scale + shift), BatchNorm(eval) + ReLU folded in
        a.ep_scale, a.ep_shift = _f32(ep[0]), _f32(ep[1])
    if pro is not None:
        a.pro_scale, a.pro_shift = _f32(pro[0]), _f32(pro[1])
        a.pro_mask = int(pro[2]) if len(pro) > 2 else 1       # bit s: source s holds a raw conv output
    _lib.check(lib.sunet_conv_gemm(C.byref(a), _stream()), "sunet_conv_gemm")


# ----------------------------------------------------------------------------- G2
def _wgrad_args(grid, a: torch.Tensor, b_mode: int, b0: torch.Tensor, b1: Optional[torch.Tensor],
                partials: Optional[torch.Tensor]) -> _lib.WgradGemmArgs:
    w = _lib.WgradGemmArgs()
    w.batch, w.height, w.width = grid
    w.a, w.a_channels, w.a_pix_stride = _act(a)
    w.b_mode = b_mode
    w.b0, w.b0_channels, w.b0_pix_stride = _act(b0)
    if b1 is not None:
        w.b1, w.b1_channels, w.b1_pix_stride = _act(b1)
    if partials is not None:
        w.partials = partials.data_ptr()
        w.partials_bytes = partials.numel() * partials.element_size()
    return w


def wgrad_splits(grid, a, b_mode, b0, b1=None) -> int:
    w = _wgrad_args(grid, a, b_mode, b0, b1, None)
    r = _lib.load().sunet_wgrad_gemm_splits(C.byref(w))
    if r <= 0:
        raise _lib.SunetError("wgrad_gemm_splits: bad shape")
    return r


def wgrad_pro_supported(grid, a, b_mode, b0, b1=None) -> bool:
    w = _wgrad_args(grid, a, b_mode, b0, b1, None)
    return bool(_lib.load().sunet_wgrad_gemm_pro_supported(C.byref(w)))


def wgrad_gemm(grid, a, b_mode, b0, partials, b1=None, b_pro=None) -> int:
    """Returns the number of split-K partial slabs written.
    b_pro = (scale, shift): b0 holds a RAW conv output; relu(scale*y + shift) is applied to each staged B tile."""
    w = _wgrad_args(grid, a, b_mode, b0, b1, partials)
    if b_pro is not None:
        w.b_pro_scale, w.b_pro_shift = _f32(b_pro[0]), _f32(b_pro[1])
    lib = _lib.load()
    splits = lib.sunet_wgrad_gemm_splits(C.byref(w))
    _lib.check(lib.sunet_wgrad_gemm(C.byref(w), _stream()), "sunet_wgrad_gemm")
    return splits


def wgrad_reduce(partials, splits, taps, a_channels, b_channels, layout, grad, real_cin=0) -> None:
    _lib.check(_lib.load().sunet_wgrad_reduce(partials.data_ptr(), splits, taps, a_channels, b_channels, layout,
                                              real_cin, _f32(grad), _stream()), "sunet_wgrad_reduce")


# ----------------------------------------------------------------------------- packing
def pack_input_im2col(x: torch.Tensor, out: torch.Tensor) -> None:
    B, Cin, H, W = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and out.dtype == torch.bfloat16
    assert out.shape == (B, H, W, 64) and out.is_contiguous()
    _lib.check(_lib.load().sunet_pack_input_im2col(x.data_ptr(), out.data_ptr(), B, Cin, H, W, _stream()),
               "sunet_pack_input_im2col")


def pack_conv3x3_weights(w, wf, wd=None) -> None:
    co, ci = w.shape[0], w.shape[1]
    _lib.check(_lib.load().sunet_pack_conv3x3_weights(_f32(w), wf.data_ptr(), _ptr(wd), co, ci, _stream()),
               "sunet_pack_conv3x3_weights")


def pack_input_im2col32(x: torch.Tensor, out: torch.Tensor) -> None:
    """fp32 NCHW -> bf16 [B,H,W,32] (paired-pixel first layer)."""
    B, Cin, H, W = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and x.is_cuda
    assert out.dtype == torch.bfloat16 and out.is_contiguous() and tuple(out.shape) == (B, H, W, 32)
    _lib.check(_lib.load().sunet_pack_input_im2col32(x.data_ptr(), out.data_ptr(), B, Cin, H, W, _stream()),
               "sunet_pack_input_im2col32")


def pack_input_u8_im2col32(img: torch.Tensor, lut: torch.Tensor, flip: Optional[torch.Tensor],
                           out: torch.Tensor) -> None:
    """uint8 [B,H,W,3] patches (+ per-image flip bits) -> the first layer's [B,H,W,32] bf16 operand."""
    B, H, W, Cc = img.shape
    assert img.dtype == torch.uint8 and img.is_contiguous() and img.is_cuda and Cc == 3
    assert lut.dtype == torch.float32 and lut.numel() == 256 and lut.is_cuda and lut.is_contiguous()
    assert flip is None or (flip.dtype == torch.uint8 and flip.numel() == B and flip.is_cuda)
    assert out.dtype == torch.bfloat16 and out.is_contiguous() and tuple(out.shape) == (B, H, W, 32)
    _lib.check(_lib.load().sunet_pack_input_u8_im2col32(img.data_ptr(), lut.data_ptr(), _ptr(flip), out.data_ptr(), B, H,
                                                        W, _stream()), "sunet_pack_input_u8_im2col32")


def pack_label_u8(label: torch.Tensor, flip: Optional[torch.Tensor], out: torch.Tensor) -> None:
    B, H, W = label.shape
    assert label.dtype == torch.uint8 and label.is_contiguous() and label.is_cuda
    assert flip is None or (flip.dtype == torch.uint8 and flip.numel() == B and flip.is_cuda)
    assert out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (B, H, W)
    _lib.check(_lib.load().sunet_pack_label_u8(label.data_ptr(), _ptr(flip), out.data_ptr(), B, H, W, _stream()),
               "sunet_pack_label_u8")


def pack_conv1_pair_weights(w, wf) -> None:
    co, ci = w.shape[0], w.shape[1]
    assert wf.dtype == torch.bfloat16 and wf.is_contiguous() and tuple(wf.shape) == (128, 64)
    _lib.check(_lib.load().sunet_pack_conv1_pair_weights(_f32(w), wf.data_ptr(), co, ci, _stream()),
               "sunet_pack_conv1_pair_weights")


def pack_conv1_weights(w, wf) -> None:
    co, ci = w.shape[0], w.shape[1]
    _lib.check(_lib.load().sunet_pack_conv1_weights(_f32(w), wf.data_ptr(), co, ci, _stream()),
               "sunet_pack_conv1_weights")


def pack_convT_weights(w, bias, wf, wd, bias4) -> None:
    ci, co = w.shape[0], w.shape[1]
    _lib.check(_lib.load().sunet_pack_convT_weights(_f32(w), _f32(bias), wf.data_ptr(), _ptr(wd), _f32(bias4), ci, co,
                                                    _stream()), "sunet_pack_convT_weights")


def pack_weights_table(table_dev: torch.Tensor, n_jobs: int, total_tiles: int) -> None:
    _lib.check(_lib.load().sunet_pack_weights_table(table_dev.data_ptr(), n_jobs, total_tiles, _stream()),
               "sunet_pack_weights_table")


# ----------------------------------------------------------------------------- BN / pool
def bn_finalize(stats, rows, channels, count, gamma, beta, conv_bias, running_mean, running_var, nbt, momentum, eps,
                scale, shift, mean, invstd) -> None:
    _lib.check(_lib.load().sunet_bn_finalize(_f32(stats), rows, channels, count, _f32(gamma), _f32(beta),
                                             _f32(conv_bias), _f32(running_mean), _f32(running_var), _ptr(nbt),
                                             momentum, eps, _f32(scale), _f32(shift), _f32(mean), _f32(invstd),
                                             _stream()), "sunet_bn_finalize")


def bn_eval_affine(gamma, beta, conv_bias, running_mean, running_var, eps, scale, shift) -> None:
    _lib.check(_lib.load().sunet_bn_eval_affine(_f32(gamma), _f32(beta), _f32(conv_bias), _f32(running_mean),
                                                _f32(running_var), eps, _f32(scale), _f32(shift), gamma.numel(),
                                                _stream()), "sunet_bn_eval_affine")


def bn_relu_pool(y, scale, shift, a, pooled=None, ywin=None) -> None:
    """ywin (with pooled): also store the conv output y of each window's first maximum (the element the pool
    gradient is routed to), which lets the producer of the pooled gradient reduce it for BatchNorm backward."""
    B, H, W, Cc = y.shape
    yp, _, ys = _act(y)
    ap, _, as_ = _act(a)
    pp, ps = (None, 0)
    if pooled is not None:
        pp, _, ps = _act(pooled)
    if ywin is not None:
        assert pooled is not None and ywin.shape == pooled.shape
        wp, _, wps = _act(ywin)
        _lib.check(_lib.load().sunet_bn_relu_pool_ywin(yp, ys, _f32(scale), _f32(shift), ap, as_, pp, ps, wp, wps, B, H,
                                                       W, Cc, _stream()), "sunet_bn_relu_pool_ywin")
        return
    _lib.check(_lib.load().sunet_bn_relu_pool(yp, ys, _f32(scale), _f32(shift), ap, as_, pp, ps, B, H, W, Cc,
                                              _stream()), "sunet_bn_relu_pool")


def maxpool2x2(a, pooled) -> None:
    B, H, W, Cc = a.shape
    ap, _, as_ = _act(a)
    pp, _, ps = _act(pooled)
    _lib.check(_lib.load().sunet_maxpool2x2(ap, as_, pp, ps, B, H, W, Cc, _stream()), "sunet_maxpool2x2")


def bn_pool_bwd_apply(dA, dPool, y, scale, shift, mean, invstd, src0, src1, dgamma, dbeta, dy, workspace) -> None:
    """Pooled block, reduction rows already written by the two producers of its gradient.
    src = (partials fp32 tensor, rows, row stride in channels, first column)."""
    B, H, W, Cc = y.shape
    yp, _, ys = _act(y)
    dyp, _, dys = _act(dy)
    dAp, _, das = _act(dA)
    dPp, _, dps = _act(dPool)
    (p0, r0, s0, c0), (p1, r1, s1, c1) = src0, src1
    for pt, r, st in ((p0, r0, s0), (p1, r1, s1)):
        assert pt.dtype == torch.float32 and pt.is_contiguous() and pt.numel() >= r * st * 2
    _lib.check(_lib.load().sunet_bn_pool_bwd_apply(dAp, das, dPp, dps, yp, ys, _f32(scale), _f32(shift), _f32(mean),
                                                   _f32(invstd), p0.data_ptr(), r0, s0, c0, p1.data_ptr(), r1, s1, c1,
                                                   _f32(dgamma), _f32(dbeta), dyp, dys, B, H, W, Cc,
                                                   workspace.data_ptr(), workspace.numel(), _stream()),
               "sunet_bn_pool_bwd_apply")


def bn_relu_pool_bwd(dA, dPool, y, scale, shift, mean, invstd, gamma, dgamma, dbeta, dy, workspace) -> None:
    B, H, W, Cc = y.shape
    yp, _, ys = _act(y)
    dyp, _, dys = _act(dy)
    dAp, das = (None, 0) if dA is None else _act(dA)[::2]
    dPp, dps = (None, 0) if dPool is None else _act(dPool)[::2]
    _lib.check(_lib.load().sunet_bn_relu_pool_bwd(dAp, das, dPp, dps, yp, ys, _f32(scale), _f32(shift), _f32(mean),
                                                  _f32(invstd), _f32(gamma), _f32(dgamma), _f32(dbeta), dyp, dys, B,
                                                  H, W, Cc, workspace.data_ptr(), workspace.numel(), _stream()),
               "sunet_bn_relu_pool_bwd")


def bn_bwd_apply(dA, y, scale, shift, mean, invstd, partials, partial_rows, dgamma, dbeta, dy, workspace) -> None:
    """Finalize + apply half of the BN/ReLU backward when the reduction rows come from the producer of dA."""
    B, H, W, Cc = y.shape
    yp, _, ys = _act(y)
    dyp, _, dys = _act(dy)
    dAp, _, das = _act(dA)
    assert partials.dtype == torch.float32 and partials.is_contiguous() and partials.numel() >= partial_rows * Cc * 2
    _lib.check(_lib.load().sunet_bn_bwd_apply(dAp, das, yp, ys, _f32(scale), _f32(shift), _f32(mean), _f32(invstd),
                                              partials.data_ptr(), partial_rows, _f32(dgamma), _f32(dbeta), dyp, dys,
                                              B, H, W, Cc, workspace.data_ptr(), workspace.numel(), _stream()),
               "sunet_bn_bwd_apply")


def colsum_finalize(stats, rows, n_total, col_offset, channels, out) -> None:
    _lib.check(_lib.load().sunet_colsum_finalize(_f32(stats), rows, n_total, col_offset, channels, _f32(out),
                                                 _stream()), "sunet_colsum_finalize")


# ----------------------------------------------------------------------------- heads / loss / metric / adam
def heads_fwd(a, weights, biases, logits) -> None:
    """weights/biases: lists (1 or 3) of fp32 tensors; logits: fp32 [nheads, P]."""
    ap, _, as_ = _act(a)
    n = len(weights)
    w = [_f32(t.reshape(-1)) for t in weights] + [None] * (3 - n)
    b = [_f32(t.reshape(-1)) for t in biases] + [None] * (3 - n)
    P = a.shape[0] * a.shape[1] * a.shape[2]
    assert logits.dtype == torch.float32 and logits.is_contiguous() and logits.numel() == n * P
    _lib.check(_lib.load().sunet_heads_fwd(ap, as_, w[0], b[0], w[1], b[1], w[2], b[2], n, logits.data_ptr(), P,
                                           _stream()), "sunet_heads_fwd")


def bn_relu_heads(y, scale, shift, a, weights, biases, logits) -> None:
    """a may be None: the activation is then not stored (heads_bwd_bn recomputes it from y)."""
    yp, _, ys = _act(y)
    ap, as_ = (None, 0) if a is None else _act(a)[::2]
    n = len(weights)
    w = [_f32(t.reshape(-1)) for t in weights] + [None] * (3 - n)
    b = [_f32(t.reshape(-1)) for t in biases] + [None] * (3 - n)
    P = y.shape[0] * y.shape[1] * y.shape[2]
    assert y.shape[3] == 64 and logits.dtype == torch.float32 and logits.is_contiguous() and logits.numel() == n * P
    _lib.check(_lib.load().sunet_bn_relu_heads(yp, ys, _f32(scale), _f32(shift), ap, as_, w[0], b[0], w[1], b[1],
                                               w[2], b[2], n, logits.data_ptr(), P, _stream()), "sunet_bn_relu_heads")


def heads_bwd(dlogits, a, weights, dA, dws, dbs, workspace) -> None:
    ap, _, as_ = _act(a)
    dp, _, ds = _act(dA)
    n = len(weights)
    w = [_f32(t.reshape(-1)) for t in weights] + [None] * (3 - n)
    dw = [_f32(t.reshape(-1)) for t in dws] + [None] * (3 - n)
    db = [_f32(t.reshape(-1)) for t in dbs] + [None] * (3 - n)
    P = a.shape[0] * a.shape[1] * a.shape[2]
    assert dlogits.dtype == torch.float32 and dlogits.is_contiguous() and dlogits.numel() == n * P
    _lib.check(_lib.load().sunet_heads_bwd(dlogits.data_ptr(), ap, as_, w[0], w[1], w[2], n, dp, ds, dw[0], db[0],
                                           dw[1], db[1], dw[2], db[2], P, workspace.data_ptr(), workspace.numel(),
                                           _stream()), "sunet_heads_bwd")


def heads_bwd_bn_rows(pixels: int) -> int:
    return int(_lib.load().sunet_heads_bwd_bn_rows(pixels))


def heads_bwd_bn(dlogits, y, scale, shift, mean, invstd, weights, dA, dws, dbs, bn_partials, workspace,
                 addend=None) -> None:
    """heads_bwd on y (activation recomputed) + the last block's BN-backward reduction rows into bn_partials.
    addend: gradient of the same activation from heads handled by an earlier call (may be dA itself)."""
    yp, _, ys = _act(y)
    dp, _, ds = _act(dA)
    n = len(weights)
    w = [_f32(t.reshape(-1)) for t in weights] + [None] * (3 - n)
    dw = [_f32(t.reshape(-1)) for t in dws] + [None] * (3 - n)
    db = [_f32(t.reshape(-1)) for t in dbs] + [None] * (3 - n)
    P = y.shape[0] * y.shape[1] * y.shape[2]
    assert y.shape[3] == 64 and dlogits.dtype == torch.float32 and dlogits.is_contiguous() and dlogits.numel() == n * P
    assert bn_partials.dtype == torch.float32 and bn_partials.is_contiguous()
    assert bn_partials.numel() >= heads_bwd_bn_rows(P) * 64 * 2
    _lib.check(_lib.load().sunet_heads_bwd_bn(dlogits.data_ptr(), yp, ys, _f32(scale), _f32(shift), _f32(mean),
                                              _f32(invstd), w[0], w[1], w[2], n, dp, ds, dw[0], db[0], dw[1], db[1],
                                              dw[2], db[2], bn_partials.data_ptr(),
                                              None if addend is None else _act(addend)[0],
                                              0 if addend is None else _act(addend)[2], P, workspace.data_ptr(),
                                              workspace.numel(), _stream()), "sunet_heads_bwd_bn")


def loss_sums(out, sel, aux, target, sums, workspace, pixels_out=None) -> None:
    """sums: fp64 [>=3] device tensor; pixels_out: optional fp64 [1] device tensor that receives float(pixels)."""
    P = target.numel()
    _lib.check(_lib.load().sunet_loss_sums(_f32(out), _f32(sel), _f32(aux), _f32(target), P, sums.data_ptr(),
                                           _ptr(pixels_out), workspace.data_ptr(), workspace.numel(), _stream()),
               "sunet_loss_sums")


def loss_finalize(sums, global_pixels, lamb, target_coverage, results, pixels_dev=None) -> None:
    """pixels_dev: fp64 [1] device tensor holding the global pixel count (overrides global_pixels)."""
    _lib.check(_lib.load().sunet_loss_finalize(sums.data_ptr(), int(global_pixels), _ptr(pixels_dev), float(lamb),
                                               float(target_coverage), _f32(results), _stream()),
               "sunet_loss_finalize")


def loss_bwd(out, sel, aux, target, sums, global_pixels, lamb, target_coverage, g_sel, g_aux, d_out, d_sel,
             d_aux, pixels_dev=None) -> None:
    P = target.numel()
    _lib.check(_lib.load().sunet_loss_bwd(_f32(out), _f32(sel), _f32(aux), _f32(target), P, sums.data_ptr(),
                                          int(global_pixels), _ptr(pixels_dev), float(lamb), float(target_coverage),
                                          _f32(g_sel), _f32(g_aux), _f32(d_out), _f32(d_sel), _f32(d_aux), _stream()),
               "sunet_loss_bwd")


_LABEL_DTYPES = {torch.uint8: 0, torch.float32: 1, torch.int64: 2}


def metric_hist(out, sel, label, thr_out, thr_sel, masked, counts) -> None:
    assert label.is_contiguous() and label.dtype in _LABEL_DTYPES, label.dtype
    assert counts.dtype == torch.int64 and counts.numel() >= 6
    P = out.numel()
    assert label.numel() == P
    _lib.check(_lib.load().sunet_metric_hist(_f32(out), _f32(sel), label.data_ptr(), _LABEL_DTYPES[label.dtype], P,
                                             float(thr_out), float(thr_sel), int(bool(masked)), counts.data_ptr(),
                                             _stream()), "sunet_metric_hist")


def minmax_f32(x: torch.Tensor, out: torch.Tensor, workspace: torch.Tensor) -> None:
    """out[0], out[1] = min, max of a contiguous fp32 tensor (the batch-wide extrema eval.py's fn_scale_minmax uses)."""
    assert x.dtype == torch.float32 and x.is_contiguous() and x.is_cuda and out.dtype == torch.float32 and out.numel() >= 2
    _lib.check(_lib.load().sunet_minmax_f32(x.data_ptr(), x.numel(), out.data_ptr(), workspace.data_ptr(),
                                            workspace.numel(), _stream()), "sunet_minmax_f32")


def ensemble_mean(maps, scale: str, mean: torch.Tensor, minmax=None) -> None:
    """mean = np.mean([scale(m) for m in maps], axis=0) with numpy's float32 order of operations (eval.py:209-222).
    maps: list of contiguous fp32 CUDA tensors of equal numel; minmax: list of device (min, max) pairs (scale='minmax')."""
    n = len(maps)
    P = maps[0].numel()
    for m in maps:
        assert m.dtype == torch.float32 and m.is_contiguous() and m.is_cuda and m.numel() == P
    assert mean.dtype == torch.float32 and mean.is_contiguous() and mean.numel() == P
    arr = (C.c_void_p * n)(*[m.data_ptr() for m in maps])
    mm = None
    if minmax is not None:
        assert len(minmax) == n
        mm = (C.c_void_p * n)(*[t.data_ptr() for t in minmax])
    _lib.check(_lib.load().sunet_ensemble_mean(arr, mm, n, P, _lib.SCALE_MODES[scale], mean.data_ptr(), _stream()),
               "sunet_ensemble_mean")


def adam_step(table_dev, n_tensors, max_numel, lr, beta1, beta2, eps, weight_decay, step, lr_dev=None,
              step_dev=None) -> None:
    _lib.check(_lib.load().sunet_adam_step(table_dev.data_ptr(), n_tensors, max_numel, lr, beta1, beta2, eps,
                                           weight_decay, step, _ptr(lr_dev), _ptr(step_dev), _stream()),
               "sunet_adam_step")
