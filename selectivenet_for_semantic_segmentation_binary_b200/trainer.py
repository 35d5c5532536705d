"""Fused SUNet_B training step and batch-sharded data parallelism.

``SUNetTrainer.step(x, label)`` is the body of the reference's hot loop
(/root/reference/train.py:183-241) — forward, aux BCE + selective risk, backward, Adam, and the
thresholded confusion-matrix update — as one sequence of kernel launches with **no host
synchronisation**: loss values stay on the device (the reference's four ``.item()`` calls per step
become one optional read at logging time), the optimizer is one multi-tensor kernel, and the whole
sequence is captured once per input shape into a CUDA graph and replayed.
``SUNetTrainer.validate(x, label)`` is the body of the validation loop (train.py:275-331): eval-mode
forward (BatchNorm folded from running statistics into the conv epilogues), the same two losses without
gradients, and the Evaluator update.

Data parallel (replaces ``torch.nn.DataParallel``, train.py:132-134): one process per GPU, weights
resident on every rank, the batch sharded on dim 0.  Two exchanges per step over NCCL/NVLink:
  1. all-reduce(sum) of the three loss sums AND the pixel count between the loss phases, so coverage,
     risk and the per-pixel gradients are *global-batch* quantities (the reference computes the loss on
     the gathered batch; averaging per-shard losses is a different function) — uneven shards included;
  2. all-reduce(sum) of the flat gradient buffer in seven groups, each launched on a side stream
     as soon as backward has produced that group, overlapping the rest of backward.
BatchNorm statistics stay per-shard, exactly like DataParallel's replicas.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from . import kernels as K
from .engine import param_order
from .model import UNet_B
from .optim import Adam


def exchange_loss_sums(sums4: torch.Tensor, group=None) -> None:
    """Data-parallel exchange #1: sums4 = fp64 [sum sigmoid(sel), sum bce*sigmoid(sel), sum bce(aux), pixels] of this
    rank's shard -> the same four numbers of the GLOBAL batch, in place (train.py:194-201 computes the losses on the
    gathered batch).  Works on any backend (NCCL in the product, gloo in the CPU tests)."""
    import torch.distributed as dist
    dist.all_reduce(sums4, op=dist.ReduceOp.SUM, group=group)


class SUNetTrainer:
    def __init__(self, net: UNet_B, lr: float = 1e-3, s_lamb: float = 2, target_coverage: float = 0.8,
                 betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, process_group=None,
                 world_size: int = 1, evaluator=None, use_cuda_graph: bool = True, val_evaluator=None):
        self.net = net
        self.selective = bool(net.selective)
        self.lamb, self.tc = float(s_lamb), float(target_coverage)
        self.group, self.world = process_group, int(world_size)
        self.evaluator = evaluator
        self.val_evaluator = val_evaluator
        p0 = next(net.parameters())
        if not p0.is_cuda:
            raise RuntimeError("SUNetTrainer needs the model on a CUDA device (no CPU path)")
        self.device = p0.device
        self.order = param_order(self.selective)
        self.params = dict(net.named_parameters())
        self.buffers = dict(net.named_buffers())
        self.fg = net._flat_grads()
        plist = [self.params[n] for n in self.order]
        # torch.optim.Adam drop-in over the flat gradient buffer (train.py:88-92); `trainer.optimizer` is what
        # net_save / net_train_load (utils/net_utils.py:5-40) take as `optim`
        self.optimizer = Adam(plist, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                              grads=[self.fg.views[n] for n in self.order])
        # [S, R, A, pixels] of this shard; all-reduced as ONE buffer when data parallel
        self.sums = torch.zeros(4, dtype=torch.float64, device=self.device)
        self.results = torch.zeros(4, device=self.device)
        self.val_results = torch.zeros(4, device=self.device)
        self.ws = K.new_workspace(self.device)
        # multi-GPU too: NCCL all-reduces issued through torch.distributed are graph-capturable, and at 16
        # patches per GPU the ~170 launches of a step would otherwise be host-bound
        self.use_graph = bool(use_cuda_graph)
        self.strict_graph = os.environ.get("SUNET_STRICT_GRAPH", "0") != "0"      # capture failure = error, not eager
        self._graphs: Dict[tuple, dict] = {}      # (kind, shape) -> {static buffers, graph, warm}
        self._comm = torch.cuda.Stream(device=self.device) if self.world > 1 else None
        self._ranges = self.fg.group_ranges()
        self._buckets = bucket_plan(self._ranges, self.fg.total, os.environ.get("SUNET_DP_BUCKETS", DEFAULT_BUCKETS))
        self.input_lut = None          # set_input_normalization(): byte -> float32 table of the uint8 pipeline

    # the device scalars the captured graphs read
    @property
    def lr_dev(self):
        return self.optimizer.lr_dev

    @property
    def step_dev(self):
        return self.optimizer.step_dev

    @property
    def exp_avg(self):
        return self.optimizer.exp_avg

    @property
    def exp_avg_sq(self):
        return self.optimizer.exp_avg_sq

    def set_lr(self, lr: float) -> None:
        self.optimizer.set_lr(lr)

    # ------------------------------------------------------------------ losses (shared by step / validate)
    def _loss_phase1(self, logits, lab, P: int):
        """Shard sums -> global sums + global pixel count (device), and the four loss values."""
        if self.selective:
            K.loss_sums(logits[0], logits[1], logits[2], lab, self.sums, self.ws, pixels_out=self.sums[3:4])
        else:
            K.loss_sums(None, None, logits[0], lab, self.sums, self.ws, pixels_out=self.sums[3:4])
        if self.world > 1:
            exchange_loss_sums(self.sums, self.group)

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
        self._loss_phase1(logits, lab, P)
        # the global pixel count is read from device memory (sums[3]): uneven shards need no host knowledge
        K.loss_finalize(self.sums, 0, self.lamb, self.tc, self.results, pixels_dev=self.sums[3:4])
        if self.selective:
            K.loss_bwd(logits[0], logits[1], logits[2], lab, self.sums, 0, self.lamb, self.tc, None, None, dl[0],
                       dl[1], dl[2], pixels_dev=self.sums[3:4])
        else:
            K.loss_bwd(None, None, logits[0], lab, self.sums, 0, 0.0, 0.0, None, None, None, None, dl[0],
                       pixels_dev=self.sums[3:4])
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
        self.optimizer.enqueue_step()
        if self.evaluator is not None:
            B, H, W = plan.B, plan.H, plan.W
            self.evaluator.add_batch_from_logits(label.reshape(B, H, W), logits[0].view(B, H, W),
                                                 logits[1].view(B, H, W) if self.selective else None, path='train')

    def _val_impl(self, x: torch.Tensor, label: torch.Tensor) -> None:
        """train.py:275-331: net.eval() forward under no_grad, aux BCE + selective risk, Evaluator update."""
        plan = self.net._plan_for(x)
        logits = plan.forward(x, self.params, self.buffers, False)
        lab = label.reshape(-1)
        self._loss_phase1(logits, lab, plan.P)
        K.loss_finalize(self.sums, 0, self.lamb if self.selective else 0.0, self.tc if self.selective else 0.0,
                        self.val_results, pixels_dev=self.sums[3:4])
        ev = self.val_evaluator if self.val_evaluator is not None else self.evaluator
        if ev is not None:
            B, H, W = plan.B, plan.H, plan.W
            ev.add_batch_from_logits(label.reshape(B, H, W), logits[0].view(B, H, W),
                                     logits[1].view(B, H, W) if self.selective else None, path='train')

    # ------------------------------------------------------------------ uint8 input pipeline
    def set_input_normalization(self, mean: float = 0.5, std: float = 0.5) -> None:
        """Byte -> float32 table of ``(b/255 - mean)/std`` built with numpy's arithmetic, i.e. the values the
        reference's PatchDataset + Normalization produce (utils/data_utils.py:100,216-217)."""
        import numpy as np
        x = (np.arange(256, dtype=np.uint8) / 255.0).astype(np.float32)
        self.input_lut = torch.from_numpy(((x - mean) / std).astype(np.float32)).to(self.device)

    # ------------------------------------------------------------------ graph cache
    def _run(self, key: tuple, make_static, copy_in, body) -> None:
        """Run `body(static)` for the input shape `key`: two eager calls (lazy one-time work must not happen inside
        capture), then capture once and replay.  One graph per (kind, shape): a ragged last batch (the reference's
        DataLoader has drop_last=False) does not throw away the full-size graph."""
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= 6:
                self._graphs.pop(next(iter(self._graphs)))
            ent = dict(static=make_static(), graph=None, warm=0)
            self._graphs[key] = ent
        copy_in(ent["static"])
        if not self.use_graph or ent["warm"] < 2:
            body(ent["static"])
            ent["warm"] += 1
            return
        if ent["graph"] is None:
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(self.device)
            try:
                with torch.cuda.graph(g):
                    body(ent["static"])
            except Exception as e:  # noqa: BLE001
                if self.strict_graph:
                    raise
                import warnings
                warnings.warn(f"CUDA-graph capture of the {key[0]} step failed ({e!r}); running eagerly "
                              f"(SUNET_STRICT_GRAPH=1 turns this into an error)")
                self.use_graph = False
                torch.cuda.synchronize(self.device)
                body(ent["static"])
                return
            ent["graph"] = g
            # capture does not execute: replay once so this call is a real step
        ent["graph"].replay()

    def graph_active(self, kind: str = "train") -> bool:
        """True if some `kind` step has been captured and is being replayed (bench.py reports it)."""
        return any(k[0] == kind and e["graph"] is not None for k, e in self._graphs.items())

    # ------------------------------------------------------------------ public
    def step(self, x: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
        """x: fp32 [N,C,H,W], label: fp32 {0,1} [N,H,W], both on this trainer's device.
        Returns a device tensor [select_loss, coverage, aux_loss, total_loss] (no sync)."""
        if not x.is_cuda or not label.is_cuda:
            raise RuntimeError("SUNetTrainer.step: inputs must be CUDA tensors")
        if label.dtype != torch.float32:
            label = label.to(torch.float32)
        plan = self.net._plan_for(x)

        def make_static():
            return (torch.empty_like(x), torch.empty_like(label), torch.empty(plan.nheads, plan.P, device=self.device))

        def copy_in(st):
            st[0].copy_(x, non_blocking=True)
            st[1].copy_(label, non_blocking=True)

        self._run(("train",) + tuple(x.shape), make_static, copy_in, lambda st: self._step_impl(st[0], st[1], st[2]))
        return self.results

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

        def make_static():
            return (torch.empty_like(img), torch.empty_like(label), torch.empty_like(flip),
                    torch.empty(N, H, W, device=self.device), torch.empty(plan.nheads, plan.P, device=self.device))

        def copy_in(st):
            st[0].copy_(img, non_blocking=True)
            st[1].copy_(label, non_blocking=True)
            st[2].copy_(flip, non_blocking=True)

        self._run(("train_u8",) + tuple(img.shape), make_static, copy_in,
                  lambda st: self._step_impl(None, st[3], st[4], u8=(st[0], st[1], st[2])))
        return self.results

    def validate(self, x: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
        """One validation batch (train.py:275-331).  Returns a device tensor [select_loss, coverage, aux_loss,
        total_loss] computed in eval mode (running BatchNorm statistics), no gradients, no parameter update; updates
        ``val_evaluator`` (or ``evaluator``) with the float64-sigmoid thresholding of train.py:150,155."""
        if not x.is_cuda or not label.is_cuda:
            raise RuntimeError("SUNetTrainer.validate: inputs must be CUDA tensors")
        if label.dtype != torch.float32:
            label = label.to(torch.float32)

        def make_static():
            return (torch.empty_like(x), torch.empty_like(label))

        def copy_in(st):
            st[0].copy_(x, non_blocking=True)
            st[1].copy_(label, non_blocking=True)

        self._run(("val",) + tuple(x.shape), make_static, copy_in, lambda st: self._val_impl(st[0], st[1]))
        return self.val_results


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


def shard_bounds(total: int, world: int, rank: int):
    """[lo, hi) of this rank's shard of a batch of `total` samples.  torch.chunk sizes (== DataParallel's scatter,
    so the per-replica BatchNorm statistics see the same shard sizes) whenever that leaves no rank empty; otherwise
    (e.g. 9 samples on 8 ranks: chunk gives 2,2,2,2,1,0,0,0 and DataParallel would simply use five replicas, which
    a fixed one-process-per-GPU job cannot) a balanced split whose sizes differ by at most one.  total < world has no
    valid split: ValueError (callers skip such a ragged tail batch)."""
    if total < world:
        raise ValueError(f"a batch of {total} samples cannot be sharded over {world} ranks")
    lo, hi = chunk_bounds(total, world, world - 1)
    if hi - lo > 0:
        return chunk_bounds(total, world, rank)
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
