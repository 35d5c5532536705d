"""Fused SUNet_B training step and batch-sharded data parallelism.

``SUNetTrainer.step(x, label)`` is the body of the reference's hot loop
(/root/reference/train.py:183-241) — forward, aux BCE + selective risk, backward, Adam, and the
thresholded confusion-matrix update — as one sequence of kernel launches with **no host
synchronisation**: loss values stay on the device (the reference's four ``.item()`` calls per step
become one optional read at logging time), the optimizer is one multi-tensor kernel, and on a
single GPU the whole sequence is captured once into a CUDA graph and replayed.

Data parallel (replaces ``torch.nn.DataParallel``, train.py:132-134): one process per GPU, weights
resident on every rank, the batch sharded on dim 0.  Two exchanges per step over NCCL/NVLink:
  1. all-reduce(sum) of the three loss sums between the loss phases, so coverage, risk and the
     per-pixel gradients are *global-batch* quantities (the reference computes the loss on the
     gathered batch; averaging per-shard losses is a different function);
  2. all-reduce(sum) of the flat gradient buffer in seven groups, each launched on a side stream
     as soon as backward has produced that group, overlapping the rest of backward.
BatchNorm statistics stay per-shard, exactly like DataParallel's replicas.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import kernels as K
from .engine import build_adam_table, param_order
from .model import UNet_B


class SUNetTrainer:
    def __init__(self, net: UNet_B, lr: float = 1e-3, s_lamb: float = 2, target_coverage: float = 0.8,
                 betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, process_group=None,
                 world_size: int = 1, evaluator=None, use_cuda_graph: bool = True):
        self.net = net
        self.selective = bool(net.selective)
        self.lamb, self.tc = float(s_lamb), float(target_coverage)
        self.betas, self.eps, self.wd = betas, eps, weight_decay
        self.group, self.world = process_group, int(world_size)
        self.evaluator = evaluator
        p0 = next(net.parameters())
        if not p0.is_cuda:
            raise RuntimeError("SUNetTrainer needs the model on a CUDA device (no CPU path)")
        self.device = p0.device
        self.order = param_order(self.selective)
        self.params = dict(net.named_parameters())
        self.buffers = dict(net.named_buffers())
        self.fg = net._flat_grads()
        plist = [self.params[n] for n in self.order]
        self.exp_avg = [torch.zeros_like(p) for p in plist]
        self.exp_avg_sq = [torch.zeros_like(p) for p in plist]
        self.table = build_adam_table(plist, [self.fg.views[n] for n in self.order], self.exp_avg, self.exp_avg_sq,
                                      self.device)
        self.max_numel = max(p.numel() for p in plist)
        self.lr_dev = torch.tensor([lr], dtype=torch.float32, device=self.device)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.sums = torch.zeros(3, dtype=torch.float64, device=self.device)
        self.results = torch.zeros(4, device=self.device)
        self.ws = K.new_workspace(self.device)
        # multi-GPU too: NCCL all-reduces issued through torch.distributed are graph-capturable, and at 16
        # patches per GPU the ~170 launches of a step would otherwise be host-bound
        self.use_graph = bool(use_cuda_graph)
        self._graph = None
        self._static = None           # (x, label, dlogits) static buffers of the captured shape
        self._warm = 0
        self._comm = torch.cuda.Stream(device=self.device) if self.world > 1 else None
        self._ranges = self.fg.group_ranges()
        import os
        self._buckets = bucket_plan(self._ranges, self.fg.total, os.environ.get("SUNET_DP_BUCKETS", DEFAULT_BUCKETS))
        self.launches_per_step = 0
        self.input_lut = None          # set_input_normalization(): byte -> float32 table of the uint8 pipeline
        self._u8 = None                # static (img, label_u8, flip, label_f32, dlogits) buffers + graph of step_u8
        self._u8_graph = None
        self._u8_warm = 0
        self.global_pixels_override = 0

    def set_lr(self, lr: float) -> None:
        self.lr_dev.fill_(lr)

    # ------------------------------------------------------------------ one step, eager
    def _step_impl(self, x: torch.Tensor, label: torch.Tensor, dl: torch.Tensor, u8=None) -> None:
        """u8 = (img uint8 [B,H,W,3], label uint8 [B,H,W], flip uint8 [B] or None): the device-side input pipeline;
        x is then None and `label` is the float32 buffer the label kernel fills."""
        net = self.net
        if u8 is not None:
            img, lab_u8, flip = u8
            plan = net._plan_for_shape(img.shape[0], img.shape[1], img.shape[2], img.device)
            K.pack_input_u8_im2col32(img, self.input_lut, flip, plan.col)
            K.pack_label_u8(lab_u8, flip, label)
        else:
            plan = net._plan_for(x)
        logits = plan.forward(x, self.params, self.buffers, True)
        lab = label.reshape(-1)
        P = plan.P
        if self.selective:
            K.loss_sums(logits[0], logits[1], logits[2], lab, self.sums, self.ws)
        else:
            K.loss_sums(None, None, logits[0], lab, self.sums, self.ws)
        # global pixel count: equal shards unless the caller states the true global count (uneven tail shard)
        Pg = self.global_pixels_override if self.global_pixels_override else P * self.world
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.sums, op=dist.ReduceOp.SUM, group=self.group)
        K.loss_finalize(self.sums, Pg, self.lamb, self.tc, self.results)
        if self.selective:
            K.loss_bwd(logits[0], logits[1], logits[2], lab, self.sums, Pg, self.lamb, self.tc, None, None, dl[0],
                       dl[1], dl[2])
        else:
            K.loss_bwd(None, None, logits[0], lab, self.sums, Pg, 0.0, 0.0, None, None, None, None, dl[0])
        if self.world > 1:
            import torch.distributed as dist
            comm, cur = self._comm, torch.cuda.current_stream()

            def reduce_group(tag):
                if tag not in self._buckets:
                    return
                lo, hi = self._buckets[tag]
                ev = torch.cuda.Event()
                ev.record(cur)
                comm.wait_event(ev)
                with torch.cuda.stream(comm):
                    dist.all_reduce(self.fg.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group)

            plan.backward(dl, self.params, self.fg.views, on_group_done=reduce_group)
            cur.wait_stream(comm)
        else:
            plan.backward(dl, self.params, self.fg.views)
        self.step_dev += 1
        K.adam_step(self.table, len(self.order), self.max_numel, 0.0, self.betas[0], self.betas[1], self.eps, self.wd,
                    1, lr_dev=self.lr_dev, step_dev=self.step_dev)
        if self.evaluator is not None:
            B, H, W = plan.B, plan.H, plan.W
            self.evaluator.add_batch_from_logits(label.reshape(B, H, W), logits[0].view(B, H, W),
                                                 logits[1].view(B, H, W) if self.selective else None, path='train')

    # ------------------------------------------------------------------ uint8 input pipeline
    def set_input_normalization(self, mean: float = 0.5, std: float = 0.5) -> None:
        """Byte -> float32 table of ``(b/255 - mean)/std`` built with numpy's arithmetic, i.e. the values the
        reference's PatchDataset + Normalization produce (utils/data_utils.py:100,216-217)."""
        import numpy as np
        x = (np.arange(256, dtype=np.uint8) / 255.0).astype(np.float32)
        self.input_lut = torch.from_numpy(((x - mean) / std).astype(np.float32)).to(self.device)

    def step_u8(self, img: torch.Tensor, label: torch.Tensor, flip: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One training step straight from decoded patches: img uint8 [N,H,W,3], label uint8 [N,H,W] (PNG values,
        255 = tumour), flip uint8 [N] (bit 0 left-right, bit 1 up-down — RandomFlip's two coin flips, drawn by the
        caller).  Normalisation, flips, NHWC/bf16 conversion and the first layer's im2col are one kernel; 4x less
        host->device traffic than float32 tensors.  Same return value as step()."""
        if not img.is_cuda or not label.is_cuda or img.dtype != torch.uint8 or label.dtype != torch.uint8:
            raise RuntimeError("SUNetTrainer.step_u8: uint8 CUDA tensors expected")
        if self.input_lut is None:
            self.set_input_normalization()
        N, H, W, _ = img.shape
        if flip is None:
            flip = torch.zeros(N, dtype=torch.uint8, device=self.device)
        plan = self.net._plan_for_shape(N, H, W, self.device)
        if self._u8 is None or tuple(self._u8[0].shape) != tuple(img.shape):
            self._u8 = (torch.empty_like(img), torch.empty_like(label), torch.empty_like(flip),
                        torch.empty(N, H, W, device=self.device), torch.empty(plan.nheads, plan.P, device=self.device))
            self._u8_graph, self._u8_warm = None, 0
        si, sl, sf, lab32, dl = self._u8
        si.copy_(img, non_blocking=True)
        sl.copy_(label, non_blocking=True)
        sf.copy_(flip, non_blocking=True)
        if not self.use_graph or self._u8_warm < 2:
            self._step_impl(None, lab32, dl, u8=(si, sl, sf))
            self._u8_warm += 1
            return self.results
        if self._u8_graph is None:
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(self.device)
            with torch.cuda.graph(g):
                self._step_impl(None, lab32, dl, u8=(si, sl, sf))
            self._u8_graph = g
        self._u8_graph.replay()
        return self.results

    # ------------------------------------------------------------------ public
    def step(self, x: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
        """x: fp32 [N,C,H,W], label: fp32 {0,1} [N,H,W], both on this trainer's device.
        Returns a device tensor [select_loss, coverage, aux_loss, total_loss] (no sync)."""
        if not x.is_cuda or not label.is_cuda:
            raise RuntimeError("SUNetTrainer.step: inputs must be CUDA tensors")
        if label.dtype != torch.float32:
            label = label.to(torch.float32)
        plan = self.net._plan_for(x)
        if not self.use_graph:
            if self._static is None or self._static[2].shape != (plan.nheads, plan.P):
                self._static = (None, None, torch.empty(plan.nheads, plan.P, device=self.device))
            self._step_impl(x.contiguous(), label.contiguous(), self._static[2])
            return self.results
        key = tuple(x.shape)
        if self._static is None or tuple(self._static[0].shape) != key:
            self._static = (torch.empty_like(x), torch.empty_like(label),
                            torch.empty(plan.nheads, plan.P, device=self.device))
            self._graph, self._warm = None, 0
        sx, sl, dl = self._static
        sx.copy_(x, non_blocking=True)
        sl.copy_(label, non_blocking=True)
        if self._graph is None:
            if self._warm < 2:
                # two eager steps first: lazy one-time work (function attributes, plan buffers) must not
                # happen inside capture
                self._step_impl(sx, sl, dl)
                self._warm += 1
                return self.results
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(self.device)
            try:
                with torch.cuda.graph(g):
                    self._step_impl(sx, sl, dl)
            except Exception as e:  # noqa: BLE001 - capture is an optimisation; eager is always valid
                import warnings
                warnings.warn(f"CUDA-graph capture of the training step failed ({e!r}); running eagerly")
                self.use_graph = False
                torch.cuda.synchronize(self.device)
                self._step_impl(sx, sl, dl)
                return self.results
            self._graph = g
            # capture does not execute: replay once so this call is a real step
        self._graph.replay()
        return self.results


GROUP_ORDER = ("dec1", "dec2", "dec3", "dec4", "enc3", "enc2", "enc1")     # order SUNetPlan.backward reports them
DEFAULT_BUCKETS = "dec1,dec2,dec3,dec4,enc3,enc2,enc1"


def bucket_plan(ranges, total: int, spec: str):
    """Gradient buckets of the data-parallel exchange.  `ranges` = FlatGrads.group_ranges(): the 7 groups are
    contiguous slices whose union, taken in GROUP_ORDER, is a suffix of the flat buffer growing towards 0.
    `spec` = comma list of group tags after which a bucket closes ('enc1' always closes the last one).
    Returns {tag: (lo, hi)}: when backward reports `tag`, all-reduce flat[lo:hi]."""
    close = set(t.strip() for t in spec.split(",") if t.strip()) | {"enc1"}
    unknown = close - set(GROUP_ORDER)
    if unknown:
        raise ValueError(f"SUNET_DP_BUCKETS: unknown group tags {sorted(unknown)}")
    out, hi = {}, total
    for tag in GROUP_ORDER:
        lo, ghi = ranges[tag]
        assert ghi <= hi, "groups must arrive from the end of the flat buffer towards its start"
        if tag in close:
            out[tag] = (lo, hi)
            hi = lo
    assert hi == 0
    return out




def chunk_bounds(total: int, world: int, rank: int):
    """[lo, hi) of this rank's shard, with torch.chunk's sizes (DataParallel's scatter,
    torch/nn/parallel/scatter_gather.py): ceil(total/world) per rank, remainder on the last."""
    size = -(-total // world)
    lo = min(total, rank * size)
    hi = min(total, lo + size)
    return lo, hi
