"""Execution plan of one UNet_B forward/backward on one GPU.

A ``SUNetPlan`` is built for a fixed (batch, height, width, input channels, selective) and owns
every device buffer the step needs — NHWC bf16 activations, packed bf16 weights, BatchNorm
statistics, gradient scratch and one flat fp32 gradient buffer laid out in parameter order — so
a steady-state step performs no allocation and is CUDA-graph capturable.  It drives the
C-ABI kernels (``kernels.py``); it does no arithmetic of its own.

Layer graph: /root/reference/model.py:68-103.  HBM layout (per level L = 1..4, resolution
H/2^(L-1), channels 64*2^(L-1)):
    y[layer]   raw conv output (bias-free), bf16 NHWC      kept for BN/ReLU backward
    a[layer]   relu(bn(y)), bf16 NHWC                      the next conv's TMA source
    pool[L]    2x2 max of the encoder output of level L    (L = 1..3)
    up[L]      ConvTranspose output feeding level L        (L = 1..3); the decoder conv reads
               (up[L], skip a[enc_L_2]) through two tensor maps — the concat is never built
    dcat[L]    gradient w.r.t. that virtual concat, [.., 2C]: first C -> ConvT backward,
               last C -> encoder skip
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from . import kernels as K

BN_EPS = 1e-5
BN_MOMENTUM = 0.1

# (block name, level, cin from previous act, cout, kind)
#   kind: 'first' (im2col'ed input), 'conv' (single source), 'cat' (up + skip), plus pool flag
_ENC = [("encoder_layer_1_1", 1, "first"), ("encoder_layer_1_2", 1, "conv"),
        ("encoder_layer_2_1", 2, "conv"), ("encoder_layer_2_2", 2, "conv"),
        ("encoder_layer_3_1", 3, "conv"), ("encoder_layer_3_2", 3, "conv"),
        ("decoder_layer_4_2", 4, "conv"), ("decoder_layer_4_1", 4, "conv")]
_CH = {1: 64, 2: 128, 3: 256, 4: 512}


_FUSE_PROLOGUE_DEFAULT = "1"


def kind_first(ly) -> bool:
    return ly.kind == "first"


class _Layer:
    """One CBR block: buffers + parameter handles."""

    def __init__(self, name: str, level: int, kind: str, cin: int, cout: int):
        self.name, self.level, self.kind, self.cin, self.cout = name, level, kind, cin, cout
        self.pool = False          # followed by MaxPool2d(2)
        # filled by the plan
        self.y = self.a = None
        self.wf = self.wd = None
        self.stats = None
        self.stat_rows = 0
        self.scale = self.shift = self.mean = self.invstd = None
        self.bnb_stats = None      # rows of (sum g, sum g*xhat) written by the dgrad that produces this block's dA
        self.bnb_rows = 0


class SUNetPlan:
    def __init__(self, batch: int, height: int, width: int, in_ch: int, selective: bool, device, n_cls: int = 1):
        if height % 8 or width % 8:
            raise ValueError("UNet_B needs H and W divisible by 8 (three 2x2 pools)")
        self.B, self.H, self.W, self.in_ch, self.selective = batch, height, width, in_ch, selective
        self.device = torch.device(device)
        # n_cls = 1: UNet_B (one logit per head); n_cls = 2: the reference's `UNet` (model.py:106-191), whose heads
        # have two output channels each.  Logit row c * nheads + h = channel c of head h.
        if n_cls not in (1, 2):
            raise NotImplementedError("n_cls must be 1 (UNet_B) or 2 (UNet): the metrics are binary (SURVEY.md §2)")
        self.n_cls = n_cls
        self.nheads = 3 if selective else 1
        dev = self.device
        bf = torch.bfloat16
        B = batch
        self.hw = {L: (height >> (L - 1), width >> (L - 1)) for L in (1, 2, 3, 4)}

        def act(L, ch):
            h, w = self.hw[L]
            return torch.empty(B, h, w, ch, dtype=bf, device=dev)

        # ---- layers in forward order
        self.layers: Dict[str, _Layer] = {}
        order: List[_Layer] = []

        def add(name, L, kind, cin, cout):
            ly = _Layer(name, L, kind, cin, cout)
            self.layers[name] = ly
            order.append(ly)
            return ly

        add("encoder_layer_1_1", 1, "first", in_ch, 64)
        add("encoder_layer_1_2", 1, "conv", 64, 64).pool = True
        add("encoder_layer_2_1", 2, "conv", 64, 128)
        add("encoder_layer_2_2", 2, "conv", 128, 128).pool = True
        add("encoder_layer_3_1", 3, "conv", 128, 256)
        add("encoder_layer_3_2", 3, "conv", 256, 256).pool = True
        add("decoder_layer_4_2", 4, "conv", 256, 512)
        add("decoder_layer_4_1", 4, "conv", 512, 512)
        add("decoder_layer_3_2", 3, "cat", 512, 256)
        add("decoder_layer_3_1", 3, "conv", 256, 256)
        add("decoder_layer_2_2", 2, "cat", 256, 128)
        add("decoder_layer_2_1", 2, "conv", 128, 128)
        add("decoder_layer_1_2", 1, "cat", 128, 64)
        add("decoder_layer_1_1", 1, "conv", 64, 64)
        self.order = order

        # First layer (Cin = 2 or 3), paired-pixel form: im2col to 32 channels per pixel, so a 128-byte row holds two
        # adjacent pixels; a PLAIN GEMM over the (B, H, W/2) pair grid with block-diagonal weights [128][64] then
        # writes both pixels' 64 outputs side by side = the NHWC output itself.  SUNET_FIRST_PAIR=0: 64-wide form.
        self.first_pair = os.environ.get("SUNET_FIRST_PAIR", "1") != "0" and 9 * in_ch <= 32 and width % 2 == 0
        if self.first_pair:
            self.col = torch.empty(B, height, width, 32, dtype=bf, device=dev)
            self.col2 = self.col.view(B, height, width // 2, 64)
        else:
            self.col = act(1, 64)                               # im2col'ed input (9*in_ch real channels)
        for ly in order:
            h, w = self.hw[ly.level]
            ly.y = act(ly.level, ly.cout)
            ly.a = act(ly.level, ly.cout)
            kdim = 64 if ly.kind == "first" else 9 * ly.cin
            ly.wf = torch.empty(ly.cout, kdim, dtype=bf, device=dev)
            ly.wd = None if ly.kind == "first" else torch.empty(ly.cin, 9 * ly.cout, dtype=bf, device=dev)
            if kind_first(ly) and self.first_pair:
                ly.wf = torch.empty(128, 64, dtype=bf, device=dev)
                # [rows][128][2] partials == [2*rows][64][2]: the finalize kernel just sees twice the rows
                ly.stat_rows = 2 * K.conv_gemm_stat_rows(B, h, w // 2, 128, K.A_PLAIN)
            else:
                ly.stat_rows = K.conv_gemm_stat_rows(B, h, w, ly.cout, K.A_PLAIN if kind_first(ly) else K.A_CONV3X3)
            ly.stats = torch.zeros(ly.stat_rows, ly.cout, 2, device=dev)
            ly.scale, ly.shift, ly.mean, ly.invstd = (torch.empty(ly.cout, device=dev) for _ in range(4))
        self.pool = {L: act(L + 1, _CH[L]) for L in (1, 2, 3)}   # pooled encoder output of level L
        self.up = {L: act(L, _CH[L]) for L in (1, 2, 3)}         # ConvT output feeding level L
        # ConvT packs: unpool{L}: C_{L+1} -> C_L
        self.upw = {}
        for L in (1, 2, 3):
            ci, co = _CH[L + 1], _CH[L]
            self.upw[L] = dict(wf=torch.empty(4 * co, ci, dtype=bf, device=dev),
                               wd=torch.empty(ci, 4 * co, dtype=bf, device=dev),
                               b4=torch.empty(4 * co, device=dev))
        P = B * height * width
        self.P = P
        self.logits = torch.empty(self.n_cls * self.nheads, P, device=dev)

        # ---- backward scratch
        # Backward scratch of levels 2-4 lives in the storage of level-1 DECODER forward tensors that are dead by the
        # time it is first written (SUNET_ALIAS_SCRATCH=0: separate allocations).  Backward runs level 1 first, so
        #   decoder_layer_1_1.y (last read: its own BN backward)          -> dcat[2]
        #   decoder_layer_1_2.y (last read: its own BN backward)          -> gB[2] (both buffers)
        #   decoder_layer_1_2.a (last read: wgrad of decoder_layer_1_1)   -> gA[2], dcat[3]
        #   up[1]               (last read: wgrad of decoder_layer_1_2)   -> gB[3] (both), gA[3], gB[4] (both)
        #   decoder_layer_2_1.y (last read: its own BN backward)          -> dpool[1], dpool[2], gA[4]
        #   decoder_layer_2_2.y (last read: its own BN backward)          -> dpool[3]
        # The two wgrads run on the side stream: backward() makes the main stream wait for their events before the
        # first aliased write (the y tensors are only ever read on the main stream).  Sizes match exactly (channels
        # double and pixels quarter per level): 4.56 of the 26 level-1-sized units of a plan, 38 MB per 256^2 patch.
        # The next forward rewrites these tensors only after backward has joined the side stream.
        self.alias_scratch = os.environ.get("SUNET_ALIAS_SCRATCH", "1") != "0"
        # One dY buffer at level 1 for large batches (SUNET_LOW_MEM=0: always two, =1: always one).  The second buffer lets
        # the BN backward of the next layer start while the weight-gradient GEMM of this one still reads dY on the side
        # stream; at batch 128 that GEMM always finishes before the dgrad conv it runs beside (0.55 vs 0.76 ms, 1.04 vs
        # 1.20 ms), so the buffer buys nothing (32.33-32.49 vs 32.42-32.49 ms per step) and costs 8.4 MB per 256^2
        # patch; at batch 16 it is worth 0.03 ms of a 4.3 ms step and 134 MB are no concern.
        lm = os.environ.get("SUNET_LOW_MEM", "auto")
        self.low_mem = (batch * height * width >= 32 * 256 * 256) if lm == "auto" else lm != "0"

        def carve(buf, shapes):
            flat, off, out = buf.view(-1), 0, []
            for shp in shapes:
                n = 1
                for d in shp:
                    n *= d
                out.append(flat[off:off + n].view(*shp))
                off += n
            assert off <= flat.numel(), (off, flat.numel())
            return out

        def shp(L, ch):
            return (B,) + self.hw[L] + (ch,)

        alias = {}
        if self.alias_scratch:
            d11, d12 = self.layers["decoder_layer_1_1"], self.layers["decoder_layer_1_2"]
            alias["dcat2"], = carve(d11.y, [shp(2, 2 * _CH[2])])
            alias["gB2a"], alias["gB2b"] = carve(d12.y, [shp(2, _CH[2])] * 2)
            alias["gA2"], alias["dcat3"] = carve(d12.a, [shp(2, _CH[2]), shp(3, 2 * _CH[3])])
            (alias["gB3a"], alias["gB3b"], alias["gA3"], alias["gB4a"], alias["gB4b"]) = carve(
                self.up[1], [shp(3, _CH[3])] * 3 + [shp(4, _CH[4])] * 2)
            d21, d22 = self.layers["decoder_layer_2_1"], self.layers["decoder_layer_2_2"]
            alias["dpool1"], alias["dpool2"], alias["gA4"] = carve(d21.y, [shp(2, _CH[1]), shp(3, _CH[2]), shp(4, _CH[4])])
            alias["dpool3"], = carve(d22.y, [shp(4, _CH[3])])
        self._wgrad_ev: Dict[str, Optional[torch.cuda.Event]] = {}
        self.gA = {L: alias[f"gA{L}"] if f"gA{L}" in alias else act(L, _CH[L]) for L in (1, 2, 3, 4)}
        # dY buffers, two per level: the weight-gradient GEMM of layer L runs on a side stream while the main
        # stream already produces dY of layer L-1 into the other buffer
        self.gB = {}
        for L in (1, 2, 3, 4):
            if f"gB{L}a" in alias:
                self.gB[L] = (alias[f"gB{L}a"], alias[f"gB{L}b"])
            elif L == 1 and self.low_mem:
                one = act(L, _CH[L])
                self.gB[L] = (one, one)
            else:
                self.gB[L] = (act(L, _CH[L]), act(L, _CH[L]))
        self._gB_next = {L: 0 for L in (1, 2, 3, 4)}
        self._gB_busy = {L: [None, None] for L in (1, 2, 3, 4)}      # event: last wgrad that read the buffer
        # SUNET_WGRAD_AFTER_DGRAD=1: the side-stream wgrad of layer l starts when dgrad(l) has finished, so it runs
        # beside the HBM-bound BN backward of layer l-1 instead of competing with dgrad(l) for the SMs
        self.wgrad_after_dgrad = os.environ.get("SUNET_WGRAD_AFTER_DGRAD", "0") != "0"
        self.side = torch.cuda.Stream(device=dev, priority=int(os.environ.get("SUNET_SIDE_PRIO", "0")))
        self.overlap_wgrad = os.environ.get("SUNET_OVERLAP_WGRAD", "1") != "0"
        self.dcat = {L: alias[f"dcat{L}"] if f"dcat{L}" in alias else act(L, 2 * _CH[L]) for L in (1, 2, 3)}
        self.dpool = {L: alias[f"dpool{L}"] if f"dpool{L}" in alias else act(L + 1, _CH[L]) for L in (1, 2, 3)}
        self.dcat_stats = {}
        for L in (1, 2, 3):
            h, w = self.hw[L]
            rows = K.conv_gemm_stat_rows(B, h, w, 2 * _CH[L])
            self.dcat_stats[L] = (torch.zeros(rows, 2 * _CH[L], 2, device=dev), rows)
        # Fused BN-backward reduction: the dgrad conv of layer p that produces dA of a non-pooled block c also
        # emits c's (sum g, sum g*xhat) rows from its epilogue (producer name -> consumer layer)
        self.bnb_of: Dict[str, _Layer] = {}
        if os.environ.get("SUNET_FUSE_BNB", "1") != "0":
            pairs = [("decoder_layer_1_1", "decoder_layer_1_2"), ("decoder_layer_2_1", "decoder_layer_2_2"),
                     ("decoder_layer_3_1", "decoder_layer_3_2"), ("decoder_layer_4_1", "decoder_layer_4_2"),
                     ("encoder_layer_3_2", "encoder_layer_3_1"), ("encoder_layer_2_2", "encoder_layer_2_1"),
                     ("encoder_layer_1_2", "encoder_layer_1_1")]
            for pn, cn in pairs:
                pl, cl = self.layers[pn], self.layers[cn]
                h, w = self.hw[pl.level]
                dy, out = self.gB[pl.level][0], self.gA[pl.level]
                if K.conv_gemm_bnb_supported(K.A_CONV3X3, (B, h, w), dy, pl.wd, out):
                    rows = K.conv_gemm_stat_rows(B, h, w, cl.cout)
                    cl.bnb_rows = rows
                    cl.bnb_stats = torch.zeros(rows, cl.cout, 2, device=dev)
                    self.bnb_of[pn] = cl
        # ... and the ConvTranspose backward-data GEMM of level L does the same for the block that feeds the ConvT
        self.bnb_convT: Dict[int, _Layer] = {}
        if os.environ.get("SUNET_FUSE_BNB", "1") != "0" and os.environ.get("SUNET_FUSE_BNB_CONVT", "1") != "0":
            for L in (1, 2, 3):
                cl = self.layers[self._convT_input(L)]
                hh, ww = self.hw[L + 1]
                dup = self.dcat[L][..., :_CH[L]]
                if K.conv_gemm_bnb_supported(K.A_GATHER2X2, (B, hh, ww), dup, self.upw[L]["wd"], self.gA[L + 1]):
                    cl.bnb_rows = K.conv_gemm_stat_rows(B, hh, ww, cl.cout, K.A_GATHER2X2)
                    cl.bnb_stats = torch.zeros(cl.bnb_rows, cl.cout, 2, device=dev)
                    self.bnb_convT[L] = cl
        # Pooled encoder blocks (level L output feeds a skip AND a max-pool): their reduction is split between the
        # two producers of their gradient — the decoder dgrad writing [d_up | d_skip] reduces the skip half
        # (bnb_col0 = C), the deeper dgrad writing d_pool reduces the pool-routed half against ywin (the conv output
        # of each window's winner, stored by the forward pool kernel).  pool_fuse[L] = (skip producer, pool producer)
        self.pool_fuse: Dict[int, tuple] = {}
        self.ywin: Dict[int, torch.Tensor] = {}
        self.pool_stats: Dict[int, tuple] = {}
        if os.environ.get("SUNET_FUSE_BNB", "1") != "0" and os.environ.get("SUNET_FUSE_BNB_POOL", "1") != "0":
            for L in (1, 2, 3):
                c = _CH[L]
                skip_p = self.layers[f"decoder_layer_{L}_2"]
                pool_p = self.layers["decoder_layer_4_2" if L == 3 else f"encoder_layer_{L + 1}_1"]
                h, w = self.hw[L]
                h2, w2 = self.hw[L + 1]
                ok1 = K.conv_gemm_bnb_supported(K.A_CONV3X3, (B, h, w), self.gB[L][0], skip_p.wd, self.dcat[L]) and \
                    h % 16 == 0 and w % 16 == 0          # the column-offset form lives in the halo kernel only
                ok2 = K.conv_gemm_bnb_supported(K.A_CONV3X3, (B, h2, w2), self.gB[L + 1][0], pool_p.wd, self.dpool[L])
                if ok1 and ok2:
                    rows = K.conv_gemm_stat_rows(B, h2, w2, c)
                    self.pool_stats[L] = (torch.zeros(rows, c, 2, device=dev), rows)
                    self.ywin[L] = act(L + 1, c)
                    self.pool_fuse[L] = (skip_p.name, pool_p.name)
        # the last block: heads_bwd recomputes relu(bn(y)) from y and emits the BN-backward reduction rows, so the
        # activation of decoder_layer_1_1 is never stored (SUNET_FUSE_HEADS_BN=0 restores the two-pass form)
        self.fuse_heads_bn = os.environ.get("SUNET_FUSE_HEADS_BN", "1") != "0"
        if self.fuse_heads_bn:
            last = self.layers["decoder_layer_1_1"]
            last.bnb_rows = K.heads_bwd_bn_rows(B * height * width)
            last.bnb_stats = torch.zeros(last.bnb_rows, last.cout, 2, device=dev)
            last.a = None
        self.eval_fused = os.environ.get("SUNET_EVAL_FUSED", "1") != "0"
        # Training PROLOGUE fusion (SUNET_FUSE_PROLOGUE=1): for conv -> BN -> ReLU -> conv chains with no pool, skip or
        # ConvT in between, the consumer's forward conv and its weight-gradient GEMM read the producer's RAW conv output
        # y and apply relu(scale * y + shift) to each staged tile in shared memory, so the producer's activation is
        # never written to or read from HBM.  Only where both kernels have the variant (CTA-pair halo conv and CTA-pair
        # shifted-window wgrad: the level-3 blocks at 256^2).  pro_of: consumer name -> producer layer.
        self.pro_of: Dict[str, _Layer] = {}
        if os.environ.get("SUNET_FUSE_PROLOGUE", _FUSE_PROLOGUE_DEFAULT) != "0":
            for pn, cn in (("encoder_layer_1_1", "encoder_layer_1_2"), ("encoder_layer_2_1", "encoder_layer_2_2"),
                           ("encoder_layer_3_1", "encoder_layer_3_2"), ("decoder_layer_4_2", "decoder_layer_4_1"),
                           ("decoder_layer_3_2", "decoder_layer_3_1"), ("decoder_layer_2_2", "decoder_layer_2_1")):
                pl, cl = self.layers[pn], self.layers[cn]
                h, w = self.hw[cl.level]
                dy = self.gB[cl.level][0]
                if K.conv_gemm_pro_supported(K.A_CONV3X3, (B, h, w), pl.y, cl.wf, cl.y) and \
                        K.wgrad_pro_supported((B, h, w), dy, K.A_CONV3X3, pl.y):
                    self.pro_of[cn] = pl
        self._pro_producers = {pl.name for pl in self.pro_of.values()}
        self.ws = K.new_workspace(dev)
        self.partials = None      # sized lazily from the wgrad split plan
        self._partials_bytes = 0
        self._size_partials()
        self.generation = 0       # bumped by every training forward; backward checks it

    # ------------------------------------------------------------------ helpers
    def _size_partials(self):
        need = 0
        B = self.B
        for ly in self.order:
            h, w = self.hw[ly.level]
            dy = self.gB[ly.level][0]
            if ly.kind == "first" and self.first_pair:
                s = K.wgrad_splits((B, h, w // 2), dy.view(B, h, w // 2, 128), K.A_PLAIN, self.col2)
                need = max(need, s * 128 * 64)
            elif ly.kind == "first":
                s = K.wgrad_splits((B, h, w), dy, K.A_PLAIN, self.col)
                need = max(need, s * 1 * ly.cout * 64)
            elif ly.kind == "cat":
                s = K.wgrad_splits((B, h, w), dy, K.A_CONV3X3, self.up[ly.level], self._skip(ly.level))
                need = max(need, s * 9 * ly.cout * ly.cin)
            else:
                src = self._conv_src(ly)
                s = K.wgrad_splits((B, h, w), dy, K.A_CONV3X3, src)
                need = max(need, s * 9 * ly.cout * ly.cin)
        for L in (1, 2, 3):
            h, w = self.hw[L + 1]
            xin = self.layers[self._convT_input(L)].a
            s = K.wgrad_splits((B, h, w), xin, K.A_GATHER2X2, self.dcat[L][..., :_CH[L]])
            need = max(need, s * 4 * _CH[L + 1] * _CH[L])
        self.partials = torch.empty(need, device=self.device)

    def _skip(self, L) -> torch.Tensor:
        return self.layers[f"encoder_layer_{L}_2"].a

    @staticmethod
    def _convT_input(L) -> str:
        return {3: "decoder_layer_4_1", 2: "decoder_layer_3_1", 1: "decoder_layer_2_1"}[L]

    def _conv_src(self, ly: _Layer) -> torch.Tensor:
        """Activation tensor a single-source conv layer reads."""
        n = ly.name
        prev = {
            "encoder_layer_1_2": self.layers["encoder_layer_1_1"].a,
            "encoder_layer_2_1": self.pool[1], "encoder_layer_2_2": self.layers["encoder_layer_2_1"].a,
            "encoder_layer_3_1": self.pool[2], "encoder_layer_3_2": self.layers["encoder_layer_3_1"].a,
            "decoder_layer_4_2": self.pool[3], "decoder_layer_4_1": self.layers["decoder_layer_4_2"].a,
            "decoder_layer_3_1": self.layers["decoder_layer_3_2"].a,
            "decoder_layer_2_1": self.layers["decoder_layer_2_2"].a,
            "decoder_layer_1_1": self.layers["decoder_layer_1_2"].a,
        }
        return prev[n]

    # ------------------------------------------------------------------ weights
    def pack_weights(self, params: Dict[str, torch.Tensor]) -> None:
        """fp32 reference-layout parameters -> bf16 kernel operands (every forward; 7.7 M values), one launch."""
        key = tuple(params[f"{ly.name}.0.weight"].data_ptr() for ly in self.order) + tuple(
            params[f"unpool{L}.weight"].data_ptr() for L in (1, 2, 3))
        cache = self.__dict__.setdefault("_pack_cache", {})      # one job table per parameter set (ensembles share a plan)
        if key not in cache:
            jobs = (_lib.PackJob * (len(self.order) + 3))()
            for i, ly in enumerate(self.order):
                w = params[f"{ly.name}.0.weight"]
                assert w.dtype == torch.float32 and w.is_contiguous()
                jobs[i].kind = (3 if self.first_pair else 1) if ly.kind == "first" else 0
                jobs[i].a, jobs[i].b = ly.cout, ly.cin
                jobs[i].w, jobs[i].wf = w.data_ptr(), ly.wf.data_ptr()
                jobs[i].wd = None if ly.wd is None else ly.wd.data_ptr()
            for k, L in enumerate((1, 2, 3)):
                j = jobs[len(self.order) + k]
                u = self.upw[L]
                w, b = params[f"unpool{L}.weight"], params[f"unpool{L}.bias"]
                j.kind, j.a, j.b = 2, w.shape[0], w.shape[1]
                j.w, j.bias, j.wf, j.wd, j.bias4 = (w.data_ptr(), b.data_ptr(), u["wf"].data_ptr(), u["wd"].data_ptr(),
                                                   u["b4"].data_ptr())
            start = 0                                     # flat grid: job j owns blocks [tile_start_j, tile_start_{j+1})
            for j in jobs:
                j.tile_start = start
                start += (j.a // 32) * (j.b // 32) if j.kind in (0, 2) else 1
            if len(cache) >= 32:
                cache.pop(next(iter(cache)))
            cache[key] = (torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8).to(self.device), start)
        self._pack_jobs, self._pack_tiles = cache[key]
        self._pack_n = len(self.order) + 3
        K.pack_weights_table(self._pack_jobs, self._pack_n, self._pack_tiles)

    # ------------------------------------------------------------------ forward
    def _cbr_fwd_eval_fused(self, ly: _Layer, params, buffers) -> bool:
        """Inference: BatchNorm (running statistics) + ReLU folded into the conv's epilogue, so y is never written
        and there is no BN/ReLU pass (model.py:11-13 under net.eval()).  All blocks except the first (pixel-pair
        GEMM) and the last (its BN+ReLU is fused with the heads)."""
        n = ly.name
        if ly.kind == "first" or n == "decoder_layer_1_1" or not self.eval_fused:
            return False
        B = self.B
        h, w = self.hw[ly.level]
        K.bn_eval_affine(params[f"{n}.1.weight"], params[f"{n}.1.bias"], params[f"{n}.0.bias"],
                         buffers[f"{n}.1.running_mean"], buffers[f"{n}.1.running_var"], BN_EPS, ly.scale, ly.shift)
        ep = (ly.scale, ly.shift)
        if ly.kind == "cat":
            K.conv_gemm(K.A_CONV3X3, (B, h, w), self.up[ly.level], ly.wf, ly.a, src1=self._skip(ly.level), ep=ep)
        else:
            K.conv_gemm(K.A_CONV3X3, (B, h, w), self._conv_src(ly), ly.wf, ly.a, ep=ep)
        if ly.pool:
            K.maxpool2x2(ly.a, self.pool[ly.level])
        return True

    def _cbr_fwd(self, ly: _Layer, params, buffers, training: bool):
        if not training and self._cbr_fwd_eval_fused(ly, params, buffers):
            return
        B = self.B
        h, w = self.hw[ly.level]
        grid = (B, h, w)
        stats = ly.stats if training else None
        if ly.kind == "first" and self.first_pair:
            K.conv_gemm(K.A_PLAIN, (B, h, w // 2), self.col2, ly.wf, ly.y.view(B, h, w // 2, 128), stats=stats)
        elif ly.kind == "first":
            K.conv_gemm(K.A_PLAIN, grid, self.col, ly.wf, ly.y, stats=stats)
        elif ly.kind == "cat":
            K.conv_gemm(K.A_CONV3X3, grid, self.up[ly.level], ly.wf, ly.y, src1=self._skip(ly.level), stats=stats)
        elif training and ly.name in self.pro_of:
            pl = self.pro_of[ly.name]
            K.conv_gemm(K.A_CONV3X3, grid, pl.y, ly.wf, ly.y, stats=stats, pro=(pl.scale, pl.shift))
        else:
            K.conv_gemm(K.A_CONV3X3, grid, self._conv_src(ly), ly.wf, ly.y, stats=stats)
        n = ly.name
        if training:
            K.bn_finalize(ly.stats, ly.stat_rows, ly.cout, B * h * w, params[f"{n}.1.weight"], params[f"{n}.1.bias"],
                          params[f"{n}.0.bias"], buffers[f"{n}.1.running_mean"], buffers[f"{n}.1.running_var"],
                          buffers[f"{n}.1.num_batches_tracked"], BN_MOMENTUM, BN_EPS, ly.scale, ly.shift, ly.mean,
                          ly.invstd)
        else:
            K.bn_eval_affine(params[f"{n}.1.weight"], params[f"{n}.1.bias"], params[f"{n}.0.bias"],
                             buffers[f"{n}.1.running_mean"], buffers[f"{n}.1.running_var"], BN_EPS, ly.scale,
                             ly.shift)
        if n == "decoder_layer_1_1":
            # last block: BN + ReLU fused with the three 1x1 heads
            heads = ["conv1x1"] + (["conv_select", "conv_aux"] if self.selective else [])
            for c in range(self.n_cls):          # one pass over y per output channel of the heads
                K.bn_relu_heads(ly.y, ly.scale, ly.shift, ly.a if c == 0 else None,
                                [params[f"{h}.weight"][c] for h in heads], [params[f"{h}.bias"][c:c + 1] for h in heads],
                                self.logits[c * self.nheads:(c + 1) * self.nheads])
        elif training and n in self._pro_producers:
            pass        # relu(bn(y)) is applied by the consumer's prologue; the activation is never materialised
        else:
            K.bn_relu_pool(ly.y, ly.scale, ly.shift, ly.a, self.pool[ly.level] if ly.pool else None,
                           ywin=self.ywin.get(ly.level) if (ly.pool and training) else None)

    def forward(self, x: torch.Tensor, params: Dict[str, torch.Tensor], buffers: Dict[str, torch.Tensor],
                training: bool) -> torch.Tensor:
        """x: fp32 NCHW on this plan's device, or None when the caller has already filled ``self.col`` (the uint8
        input pipeline, kernels.pack_input_u8_im2col32).  Returns the plan-owned logits buffer [nheads, P]."""
        # the weight repack (one launch, ~70 us) does not depend on the input: it runs on the side stream beside the
        # input im2col and is joined before the first conv (inside a captured graph this is a fork / join)
        cur = torch.cuda.current_stream()
        overlap_pack = self.overlap_wgrad and x is not None
        if overlap_pack:
            fork = torch.cuda.Event()
            fork.record(cur)
            self.side.wait_event(fork)
            with torch.cuda.stream(self.side):
                self.pack_weights(params)
                packed = torch.cuda.Event()
                packed.record(self.side)
        else:
            self.pack_weights(params)
        if x is None:
            if not self.first_pair:
                raise RuntimeError("the uint8 input pipeline needs the paired-pixel first layer (SUNET_FIRST_PAIR=1)")
        else:
            assert x.shape == (self.B, self.in_ch, self.H, self.W), (x.shape, (self.B, self.in_ch, self.H, self.W))
            if self.first_pair:
                K.pack_input_im2col32(x, self.col)
            else:
                K.pack_input_im2col(x, self.col)
        if overlap_pack:
            cur.wait_event(packed)
        L = self.layers
        for name in ("encoder_layer_1_1", "encoder_layer_1_2", "encoder_layer_2_1", "encoder_layer_2_2",
                     "encoder_layer_3_1", "encoder_layer_3_2", "decoder_layer_4_2", "decoder_layer_4_1"):
            self._cbr_fwd(L[name], params, buffers, training)
        for lvl, (n2, n1) in ((3, ("decoder_layer_3_2", "decoder_layer_3_1")),
                              (2, ("decoder_layer_2_2", "decoder_layer_2_1")),
                              (1, ("decoder_layer_1_2", "decoder_layer_1_1"))):
            u = self.upw[lvl]
            hh, ww = self.hw[lvl + 1]
            K.conv_gemm(K.A_PLAIN, (self.B, hh, ww), L[self._convT_input(lvl)].a, u["wf"], self.up[lvl], bias=u["b4"],
                        d_mode=K.D_SCATTER2X2)
            self._cbr_fwd(L[n2], params, buffers, training)
            self._cbr_fwd(L[n1], params, buffers, training)
        if training:
            self.generation += 1
        return self.logits

    # ------------------------------------------------------------------ backward
    def _on_side(self, fn):
        """Run fn() (weight-gradient launches) on the side stream, ordered after everything enqueued so far on
        the current stream.  Returns the completion event (None when overlap is off)."""
        if not self.overlap_wgrad:
            fn()
            return None
        cur = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(cur)
        self.side.wait_event(ready)
        with torch.cuda.stream(self.side):
            fn()
            done = torch.cuda.Event()
            done.record(self.side)
        return done

    def _cbr_bwd(self, ly: _Layer, dA: Optional[torch.Tensor], dPool: Optional[torch.Tensor], params, grads,
                 dgrad_out: Optional[torch.Tensor], dgrad_stats: Optional[torch.Tensor] = None,
                 fused_reduce: bool = False) -> bool:
        """dA/dPool -> (dgamma, dbeta, dy) -> weight grad (side stream), and the input gradient into dgrad_out.
        fused_reduce: this block's BN-backward reduction rows were already written by the producer of dA.
        Returns True if the dgrad launched here wrote the reduction rows of the block consuming dgrad_out."""
        B = self.B
        h, w = self.hw[ly.level]
        grid = (B, h, w)
        n = ly.name
        lvl = ly.level
        idx = self._gB_next[lvl]
        if self.gB[lvl][0] is not self.gB[lvl][1]:      # (low-memory mode: one buffer, its busy event is always waited on)
            self._gB_next[lvl] = idx ^ 1
        dy = self.gB[lvl][idx]
        busy = self._gB_busy[lvl][idx]
        if busy is not None:                       # the wgrad that last read this buffer must be finished
            torch.cuda.current_stream().wait_event(busy)
        if fused_reduce and dPool is not None:
            c = ly.cout
            st0, rows0 = self.dcat_stats[lvl]
            st1, rows1 = self.pool_stats[lvl]
            K.bn_pool_bwd_apply(dA, dPool, ly.y, ly.scale, ly.shift, ly.mean, ly.invstd, (st0, rows0, 2 * c, c),
                                (st1, rows1, c, 0), grads[f"{n}.1.weight"], grads[f"{n}.1.bias"], dy, self.ws)
        elif fused_reduce:
            assert ly.bnb_stats is not None
            K.bn_bwd_apply(dA, ly.y, ly.scale, ly.shift, ly.mean, ly.invstd, ly.bnb_stats, ly.bnb_rows,
                           grads[f"{n}.1.weight"], grads[f"{n}.1.bias"], dy, self.ws)
        else:
            K.bn_relu_pool_bwd(dA, dPool, ly.y, ly.scale, ly.shift, ly.mean, ly.invstd, params[f"{n}.1.weight"],
                               grads[f"{n}.1.weight"], grads[f"{n}.1.bias"], dy, self.ws)
        gw = grads[f"{n}.0.weight"]

        def wgrad():
            if ly.kind == "first" and self.first_pair:
                s = K.wgrad_gemm((B, h, w // 2), dy.view(B, h, w // 2, 128), K.A_PLAIN, self.col2, self.partials)
                K.wgrad_reduce(self.partials, s, 1, 128, 64, 3, gw, real_cin=ly.cin)
            elif ly.kind == "first":
                s = K.wgrad_gemm(grid, dy, K.A_PLAIN, self.col, self.partials)
                K.wgrad_reduce(self.partials, s, 1, ly.cout, 64, 2, gw, real_cin=ly.cin)
            elif ly.kind == "cat":
                s = K.wgrad_gemm(grid, dy, K.A_CONV3X3, self.up[ly.level], self.partials, self._skip(ly.level))
                K.wgrad_reduce(self.partials, s, 9, ly.cout, ly.cin, 0, gw)
            elif n in self.pro_of:
                pl = self.pro_of[n]
                s = K.wgrad_gemm(grid, dy, K.A_CONV3X3, pl.y, self.partials, b_pro=(pl.scale, pl.shift))
                K.wgrad_reduce(self.partials, s, 9, ly.cout, ly.cin, 0, gw)
            else:
                s = K.wgrad_gemm(grid, dy, K.A_CONV3X3, self._conv_src(ly), self.partials)
                K.wgrad_reduce(self.partials, s, 9, ly.cout, ly.cin, 0, gw)

        nxt = self.bnb_of.get(n) if dgrad_stats is None else None
        bnb = None if nxt is None else (nxt.y, nxt.scale, nxt.shift, nxt.mean, nxt.invstd)
        if nxt is not None:
            dgrad_stats = nxt.bnb_stats
        for Lp, (skip_name, pool_name) in self.pool_fuse.items():
            pl = self.layers[f"encoder_layer_{Lp}_2"]
            if n == skip_name:           # [d_up | d_skip]: plain column sums for d_up, the reduction for d_skip
                bnb = (pl.y, pl.scale, pl.shift, pl.mean, pl.invstd, _CH[Lp])
            elif n == pool_name:         # d_pool against the winners' conv outputs
                bnb = (self.ywin[Lp], pl.scale, pl.shift, pl.mean, pl.invstd)
                dgrad_stats = self.pool_stats[Lp][0]

        def dgrad():
            if dgrad_out is not None:
                K.conv_gemm(K.A_CONV3X3, grid, dy, ly.wd, dgrad_out, stats=dgrad_stats, bnb=bnb)

        if self.wgrad_after_dgrad:
            dgrad()
            self._gB_busy[lvl][idx] = self._on_side(wgrad)
        else:
            self._gB_busy[lvl][idx] = self._on_side(wgrad)
            dgrad()
        self._wgrad_ev[n] = self._gB_busy[lvl][idx]
        return nxt is not None

    def backward(self, dlogits: torch.Tensor, params: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor],
                 on_group_done=None) -> None:
        """dlogits: fp32 [nheads, P].  Fills every tensor of `grads` (reference layouts, fp32);
        pre-BN conv bias gradients are identically zero and are left untouched (kept zero).

        on_group_done(tag) is called (host side, after the launches are enqueued) each time a group of
        parameters has all its gradients enqueued — 'dec1', 'dec2', 'dec3', 'dec4', 'enc3', 'enc2',
        'enc1' in that order; each group is a contiguous suffix slice of the flat gradient buffer,
        which is what the data-parallel trainer all-reduces while the rest of backward runs."""
        def done(tag):
            if on_group_done is not None:
                # the group's weight gradients are produced on the side stream
                torch.cuda.current_stream().wait_stream(self.side)
                on_group_done(tag)

        for lv in self._gB_busy:
            self._gB_busy[lv] = [None, None]
        L = self.layers
        B = self.B
        heads = ["conv1x1"] + (["conv_select", "conv_aux"] if self.selective else [])
        last = L["decoder_layer_1_1"]
        nh = self.nheads
        if self.fuse_heads_bn:
            # channel c of every head per call; the second call (UNet, n_cls = 2) adds the first call's gradient
            # and writes the reduction rows of the total
            for c in range(self.n_cls):
                K.heads_bwd_bn(dlogits[c * nh:(c + 1) * nh], last.y, last.scale, last.shift, last.mean, last.invstd,
                               [params[f"{h}.weight"][c] for h in heads], self.gA[1],
                               [grads[f"{h}.weight"][c] for h in heads], [grads[f"{h}.bias"][c:c + 1] for h in heads],
                               last.bnb_stats, self.ws, addend=self.gA[1] if c > 0 else None)
        else:
            if self.n_cls != 1:
                raise RuntimeError("UNet (n_cls = 2) needs SUNET_FUSE_HEADS_BN=1")
            K.heads_bwd(dlogits, last.a, [params[f"{h}.weight"] for h in heads], self.gA[1],
                        [grads[f"{h}.weight"] for h in heads], [grads[f"{h}.bias"] for h in heads], self.ws)
        dA = self.gA[1]
        up_fused = self.fuse_heads_bn
        for lvl, (n2, n1) in ((1, ("decoder_layer_1_2", "decoder_layer_1_1")),
                              (2, ("decoder_layer_2_2", "decoder_layer_2_1")),
                              (3, ("decoder_layer_3_2", "decoder_layer_3_1"))):
            c = _CH[lvl]
            f = self._cbr_bwd(L[n1], dA, None, params, grads, self.gA[lvl], fused_reduce=up_fused)
            st, rows = self.dcat_stats[lvl]
            self._cbr_bwd(L[n2], self.gA[lvl], None, params, grads, self.dcat[lvl], st, fused_reduce=f)
            # ConvTranspose backward: bias (column sums of d_up), weight, input
            dup = self.dcat[lvl][..., :c]
            K.colsum_finalize(st, rows, 2 * c, 0, c, grads[f"unpool{lvl}.bias"])
            hh, ww = self.hw[lvl + 1]
            xin = L[self._convT_input(lvl)].a
            def wgrad_t(lvl=lvl, c=c, hh=hh, ww=ww, xin=xin, dup=dup):
                s = K.wgrad_gemm((B, hh, ww), xin, K.A_GATHER2X2, dup, self.partials)
                K.wgrad_reduce(self.partials, s, 4, _CH[lvl + 1], c, 1, grads[f"unpool{lvl}.weight"])

            cl = self.bnb_convT.get(lvl)
            kw = {} if cl is None else dict(stats=cl.bnb_stats, bnb=(cl.y, cl.scale, cl.shift, cl.mean, cl.invstd))
            if self.alias_scratch and lvl in (1, 2):
                # gA[lvl + 1] (and the scratch written after it) shares storage with a tensor that a level-1 weight-
                # gradient GEMM on the side stream is still allowed to be reading: see __init__
                ev = self._wgrad_ev.get("decoder_layer_1_1" if lvl == 1 else "decoder_layer_1_2")
                if ev is not None:
                    torch.cuda.current_stream().wait_event(ev)
            if self.wgrad_after_dgrad:
                K.conv_gemm(K.A_GATHER2X2, (B, hh, ww), dup, self.upw[lvl]["wd"], self.gA[lvl + 1], **kw)
                self._on_side(wgrad_t)
            else:
                self._on_side(wgrad_t)
                K.conv_gemm(K.A_GATHER2X2, (B, hh, ww), dup, self.upw[lvl]["wd"], self.gA[lvl + 1], **kw)
            dA = self.gA[lvl + 1]
            up_fused = cl is not None        # the next block's reduction rows are already written
            done(f"dec{lvl}")
        # bottleneck
        f = self._cbr_bwd(L["decoder_layer_4_1"], self.gA[4], None, params, grads, self.gA[4], fused_reduce=up_fused)
        self._cbr_bwd(L["decoder_layer_4_2"], self.gA[4], None, params, grads, self.dpool[3], fused_reduce=f)
        done("dec4")
        # encoder, deepest first; skip gradient = second half of dcat, pooled gradient = dpool
        for lvl in (3, 2, 1):
            c = _CH[lvl]
            n2, n1 = f"encoder_layer_{lvl}_2", f"encoder_layer_{lvl}_1"
            f = self._cbr_bwd(L[n2], self.dcat[lvl][..., c:], self.dpool[lvl], params, grads, self.gA[lvl],
                              fused_reduce=lvl in self.pool_fuse)
            self._cbr_bwd(L[n1], self.gA[lvl], None, params, grads, self.dpool[lvl - 1] if lvl > 1 else None,
                          fused_reduce=f)
            done(f"enc{lvl}")
        # every weight gradient is complete before anything after backward (Adam, all-reduce) runs
        torch.cuda.current_stream().wait_stream(self.side)


def param_order(selective: bool) -> List[str]:
    """Trainable tensors in nn.Module registration order (model.py:29-66): 68 (or 64) names."""
    names = []
    seq = ["encoder_layer_1_1", "encoder_layer_1_2", "encoder_layer_2_1", "encoder_layer_2_2", "encoder_layer_3_1",
           "encoder_layer_3_2", "decoder_layer_4_2", "decoder_layer_4_1", "unpool3", "decoder_layer_3_2",
           "decoder_layer_3_1", "unpool2", "decoder_layer_2_2", "decoder_layer_2_1", "unpool1", "decoder_layer_1_2",
           "decoder_layer_1_1"]
    for n in seq:
        if n.startswith("unpool"):
            names += [f"{n}.weight", f"{n}.bias"]
        else:
            names += [f"{n}.0.weight", f"{n}.0.bias", f"{n}.1.weight", f"{n}.1.bias"]
    for h in ["conv1x1"] + (["conv_select", "conv_aux"] if selective else []):
        names += [f"{h}.weight", f"{h}.bias"]
    return names


class FlatGrads:
    """One flat fp32 buffer holding every parameter gradient, with per-parameter views."""

    def __init__(self, shapes: Dict[str, Sequence[int]], order: List[str], device):
        total = 0
        self.offsets = {}
        for n in order:
            numel = 1
            for s in shapes[n]:
                numel *= s
            self.offsets[n] = (total, numel)
            total += (numel + 3) // 4 * 4      # keep every view 16-byte aligned
        self.flat = torch.zeros(total, device=device)
        self.views = {n: self.flat[o:o + k].view(*shapes[n]) for n, (o, k) in self.offsets.items()}
        self.order = order
        self.total = total

    def group_ranges(self):
        """Flat [lo, hi) ranges of the gradient groups, keyed by the tags SUNetPlan.backward reports."""
        def off(name):
            return self.offsets[name][0]
        cuts = [("enc1", 0), ("enc2", off("encoder_layer_2_1.0.weight")), ("enc3", off("encoder_layer_3_1.0.weight")),
                ("dec4", off("decoder_layer_4_2.0.weight")), ("dec3", off("unpool3.weight")),
                ("dec2", off("unpool2.weight")), ("dec1", off("unpool1.weight"))]
        out = {}
        for i, (tag, lo) in enumerate(cuts):
            hi = cuts[i + 1][1] if i + 1 < len(cuts) else self.total
            out[tag] = (lo, hi)
        return out


def build_adam_table(params: List[torch.Tensor], grads: List[torch.Tensor], m: List[torch.Tensor],
                     v: List[torch.Tensor], device) -> torch.Tensor:
    tab = (_lib.AdamTensor * len(params))()
    for i, p in enumerate(params):
        tab[i].param, tab[i].grad = p.data_ptr(), grads[i].data_ptr()
        tab[i].exp_avg, tab[i].exp_avg_sq = m[i].data_ptr(), v[i].data_ptr()
        tab[i].numel = p.numel()
    return torch.frombuffer(bytearray(bytes(tab)), dtype=torch.uint8).to(device)
