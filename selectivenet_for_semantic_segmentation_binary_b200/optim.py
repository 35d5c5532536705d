"""``Adam`` with ``torch.optim.Adam``'s interface, executed by one multi-tensor kernel launch.

Replaces ``optim = torch.optim.Adam(net.parameters(), lr=lr, weight_decay=args.w_decay)`` and
``optim.zero_grad(); loss.backward(); optim.step()`` of /root/reference/train.py:88-92,207-209.  The reference
checkpoints ``optim.state_dict()`` (utils/net_utils.py:9) and ``net_train_load`` restores it (utils/net_utils.py:37):
``state_dict()`` / ``load_state_dict()`` here produce and accept exactly torch's layout
(``{'state': {i: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [...]}``), so a checkpoint written by this class
loads into a stock ``torch.optim.Adam`` over the same parameters and vice versa.

Learning rate and step count live in device memory (``lr_dev`` / ``step_dev``): a CUDA graph that captured
``step()`` stays valid when a scheduler changes ``param_groups[0]['lr']`` (call ``sync_lr()`` or ``set_lr``).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch

from . import kernels as K
from .engine import build_adam_table


class Adam:
    def __init__(self, params: Iterable[torch.Tensor], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, grads: Optional[List[torch.Tensor]] = None):
        self.params: List[torch.Tensor] = [p for p in params]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        p0 = self.params[0]
        if not p0.is_cuda:
            raise RuntimeError("sunet Adam has no CPU path: parameters must live on a CUDA device")
        for p in self.params:
            if p.dtype != torch.float32 or not p.is_contiguous() or p.device != p0.device:
                raise RuntimeError("sunet Adam: contiguous float32 parameters on one device expected")
        self.device = p0.device
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False,
                             maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                             decoupled_weight_decay=False)
        self.param_groups = [dict(self.defaults, params=self.params)]
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self.lr_dev = torch.tensor([lr], dtype=torch.float32, device=self.device)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.max_numel = max(p.numel() for p in self.params)
        self._static_grads = grads          # fixed gradient tensors (the trainer's flat buffer views)
        self._table = None
        self._table_key = None
        self._lr_cached = float(lr)

    # ------------------------------------------------------------------ torch.optim.Optimizer surface
    def zero_grad(self, set_to_none: bool = True) -> None:
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.detach_()
                    p.grad.zero_()

    def set_lr(self, lr: float) -> None:
        self.param_groups[0]['lr'] = float(lr)
        self.sync_lr()

    def sync_lr(self) -> None:
        """Push param_groups[0]['lr'] (what torch's lr_scheduler objects write) to the device scalar."""
        lr = float(self.param_groups[0]['lr'])
        if lr != self._lr_cached:
            self.lr_dev.fill_(lr)
            self._lr_cached = lr

    def _grads(self) -> List[torch.Tensor]:
        if self._static_grads is not None:
            return self._static_grads
        gs = []
        for p in self.params:
            if p.grad is None:
                raise RuntimeError("sunet Adam.step(): every parameter needs a gradient (call backward() first)")
            g = p.grad
            if g.dtype != torch.float32 or not g.is_contiguous():
                g = g.to(torch.float32).contiguous()
                p.grad = g
            gs.append(g)
        return gs

    def step(self, closure=None):
        loss = closure() if closure is not None else None
        grads = self._grads()
        key = tuple(g.data_ptr() for g in grads)
        if key != self._table_key:        # autograd may hand out new .grad tensors after zero_grad(set_to_none=True)
            self._table = build_adam_table(self.params, grads, self.exp_avg, self.exp_avg_sq, self.device)
            self._table_key = key
        self.sync_lr()
        g = self.param_groups[0]
        self.step_dev += 1
        K.adam_step(self._table, len(self.params), self.max_numel, 0.0, g['betas'][0], g['betas'][1], g['eps'],
                    g['weight_decay'], 1, lr_dev=self.lr_dev, step_dev=self.step_dev)
        return loss

    def enqueue_step(self) -> None:
        """step() without any host-side decision: what a captured CUDA graph replays (fixed gradient tensors)."""
        assert self._static_grads is not None
        if self._table is None:
            self._table = build_adam_table(self.params, self._static_grads, self.exp_avg, self.exp_avg_sq, self.device)
            self._table_key = tuple(g.data_ptr() for g in self._static_grads)
        g = self.param_groups[0]
        self.step_dev += 1
        K.adam_step(self._table, len(self.params), self.max_numel, 0.0, g['betas'][0], g['betas'][1], g['eps'],
                    g['weight_decay'], 1, lr_dev=self.lr_dev, step_dev=self.step_dev)

    # ------------------------------------------------------------------ checkpoint (torch.optim.Adam layout)
    def state_dict(self) -> Dict:
        step = int(self.step_dev.item())
        state = {}
        if step > 0:
            for i in range(len(self.params)):
                state[i] = {'step': torch.tensor(float(step)), 'exp_avg': self.exp_avg[i].detach().clone(),
                            'exp_avg_sq': self.exp_avg_sq[i].detach().clone()}
        group = {k: v for k, v in self.param_groups[0].items() if k != 'params'}
        group['params'] = list(range(len(self.params)))
        return {'state': state, 'param_groups': [group]}

    def load_state_dict(self, sd: Dict) -> None:
        if 'param_groups' not in sd or 'state' not in sd:
            raise ValueError("not a torch.optim.Adam state_dict (expected keys 'state' and 'param_groups'); checkpoints "
                             "written by the round-1 build carried no optimizer state and cannot resume Adam")
        groups = sd['param_groups']
        if len(groups) != 1 or len(groups[0]['params']) != len(self.params):
            raise ValueError("loaded state dict has a different number of parameter groups / parameters")
        g = groups[0]
        for k in ('lr', 'betas', 'eps', 'weight_decay'):
            if k in g:
                self.param_groups[0][k] = tuple(g[k]) if k == 'betas' else g[k]
        if g.get('amsgrad', False) or g.get('maximize', False):
            raise NotImplementedError("amsgrad / maximize are not implemented")
        self._lr_cached = None
        self.sync_lr()
        steps = set()
        for j, idx in enumerate(g['params']):
            st = sd['state'].get(idx)
            if st is None:
                self.exp_avg[j].zero_()
                self.exp_avg_sq[j].zero_()
                continue
            self.exp_avg[j].copy_(st['exp_avg'].to(self.device, torch.float32))
            self.exp_avg_sq[j].copy_(st['exp_avg_sq'].to(self.device, torch.float32))
            steps.add(int(float(st['step'])))
        if len(steps) > 1:
            raise ValueError(f"per-parameter step counts differ ({sorted(steps)}): one shared step is kept on the device")
        self.step_dev.fill_(steps.pop() if steps else 0)
