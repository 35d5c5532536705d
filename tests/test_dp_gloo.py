"""world_size-2 gloo test of the data-parallel host logic (no GPU): batch sharding with torch.chunk
sizes, the loss-sum exchange that makes coverage / selective risk GLOBAL-batch quantities, and the
SUM-reduction of gradients that carry 1/P_global — checked against the oracle on the full batch."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    from oracle import sunet_oracle as O
    from selectivenet_for_semantic_segmentation_binary_b200.trainer import chunk_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    total = 5                                            # uneven on purpose: shards of 3 and 2
    g = torch.Generator().manual_seed(0)
    out = torch.randn(total, 8, 8, generator=g)
    sel = torch.randn(total, 8, 8, generator=g) + torch.arange(total).view(-1, 1, 1) * 0.7 - 1.0   # skewed shards
    aux = torch.randn(total, 8, 8, generator=g)
    tgt = (torch.rand(total, 8, 8, generator=g) < 0.4).float()
    lo, hi = chunk_bounds(total, world, rank)
    o, s, a, t = (v[lo:hi].clone().requires_grad_(v is not tgt) for v in (out, sel, aux, tgt))
    # phase 1 on the shard: the three sums + pixel count (what sunet_loss_sums produces)
    sg = torch.sigmoid(s)
    bce = torch.nn.functional.binary_cross_entropy_with_logits(o, t, reduction="none")
    bce_a = torch.nn.functional.binary_cross_entropy_with_logits(a, t, reduction="none")
    sums = torch.stack([sg.sum(), (bce * sg).sum(), bce_a.sum(), torch.tensor(float(t.numel()))]).double()
    red = sums.detach().clone()
    dist.all_reduce(red, op=dist.ReduceOp.SUM)            # the exchange between the loss phases
    S, R, A, P = red.tolist()
    cov = S / P
    lamb = 2.0
    loss_global = R / S + lamb * max(0.0, 0.8 - cov) ** 2 + A / P
    # phase 2: per-pixel gradients from the GLOBAL sums (formulas of SURVEY.md A.6 / sunet_loss_bwd)
    d = max(0.0, 0.8 - cov)
    d_out = sg * (torch.sigmoid(o) - t) / S
    d_sel = sg * (1 - sg) * (bce / S - R / S ** 2 - 2 * lamb * d / P)
    d_aux = (torch.sigmoid(a) - t) / P
    # a toy "network": logits = w * features, gradient of w is SUM-reduced across ranks
    w = torch.ones(3, requires_grad=True)
    (w[0] * o.detach() * 0 + 0).sum()
    gw = torch.stack([(d_out.detach() * o.detach()).sum(), (d_sel.detach() * s.detach()).sum(),
                      (d_aux.detach() * a.detach()).sum()]).double()
    dist.all_reduce(gw, op=dist.ReduceOp.SUM)
    # oracle on the full batch
    wf = torch.ones(3, requires_grad=True)
    l_sel, c_ref = O.selective_risk_b(wf[0] * out, wf[1] * sel, tgt, lamb=lamb)
    l_ref = l_sel + O.bce_with_logits_mean(wf[2] * aux, tgt)
    l_ref.backward()
    ok = (abs(loss_global - l_ref.item()) < 1e-5 and abs(cov - c_ref.item()) < 1e-6 and
          torch.allclose(gw.float(), wf.grad, rtol=1e-4, atol=1e-6))
    # averaging per-shard losses is NOT the same function (why the exchange is mandatory)
    l_local, _ = O.selective_risk_b(o.detach(), s.detach(), t, lamb=lamb)
    mean_of_shards = torch.tensor([l_local.item() * (hi - lo) / total])
    dist.all_reduce(mean_of_shards)
    differs = abs(mean_of_shards.item() + A / P - l_ref.item()) > 1e-3
    ret[rank] = (bool(ok), bool(differs), (lo, hi))
    dist.destroy_process_group()


def test_two_rank_sharded_loss_and_grad_sum():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][2] == (0, 3) and ret[1][2] == (3, 5)
    for r in range(world):
        ok, differs, _ = ret[r]
        assert ok, "global loss / coverage / SUM-reduced gradients must equal the full-batch oracle"
        assert differs, "mean of per-shard losses should differ from the global loss on skewed shards"
