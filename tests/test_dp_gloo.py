"""world_size-2 gloo tests of the data-parallel HOST logic of the product (no GPU): the functions
``SUNetTrainer`` itself calls — ``shard_bounds`` (batch sharding), ``exchange_loss_sums`` (the exchange between the
loss phases that makes coverage / selective risk GLOBAL-batch quantities, through both the trainer path and the
``selective_loss.set_data_parallel_group`` path), ``bucket_plan`` over ``FlatGrads`` (which gradient slices are
all-reduced when) and train.py's scheduler replication (the learning rate must stay identical on every rank).
The per-pixel arithmetic between the exchanges runs on the GPU in the product (``sunet_loss_sums`` /
``sunet_loss_bwd``); here it is taken from the oracle so that the exchanged quantities can be checked against the
oracle's full-batch loss and gradients.  The GPU end of the same path is checked by bench.py's ``dp_parity`` block
on real NCCL ranks."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import train as train_cli
    from oracle import sunet_oracle as O
    from selectivenet_for_semantic_segmentation_binary_b200 import selective_loss as SL
    from selectivenet_for_semantic_segmentation_binary_b200.engine import FlatGrads, param_order
    from selectivenet_for_semantic_segmentation_binary_b200.trainer import (DEFAULT_BUCKETS, GROUP_ORDER, bucket_plan,
                                                                            exchange_loss_sums, shard_bounds)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 5                                            # uneven on purpose: shards of 3 and 2
    g = torch.Generator().manual_seed(0)
    out = torch.randn(total, 8, 8, generator=g)
    sel = torch.randn(total, 8, 8, generator=g) + torch.arange(total).view(-1, 1, 1) * 0.7 - 1.0   # skewed shards
    aux = torch.randn(total, 8, 8, generator=g)
    tgt = (torch.rand(total, 8, 8, generator=g) < 0.4).float()
    lo, hi = shard_bounds(total, world, rank)
    o, s, a, t = (v[lo:hi].clone() for v in (out, sel, aux, tgt))
    # phase 1 on the shard: [S, R, A, pixels] (what sunet_loss_sums writes), then the PRODUCT exchange
    sg = torch.sigmoid(s)
    bce = torch.nn.functional.binary_cross_entropy_with_logits(o, t, reduction="none")
    bce_a = torch.nn.functional.binary_cross_entropy_with_logits(a, t, reduction="none")
    sums = torch.stack([sg.sum(), (bce * sg).sum(), bce_a.sum(), torch.tensor(float(t.numel()))]).double()
    local = sums.clone()
    exchange_loss_sums(sums, None)
    S, R, A, P = sums.tolist()
    # ... and the module-API path (selective_loss._global_sums with a data-parallel group set)
    SL.set_data_parallel_group(None, True)
    buf = local.clone()
    P2 = SL._global_sums(buf, int(local[3].item()))
    SL.set_data_parallel_group(None, False)
    same_paths = torch.equal(buf, sums) and P2 == int(P) == total * 64
    cov, lamb = S / P, 2.0
    loss_global = R / S + lamb * max(0.0, 0.8 - cov) ** 2 + A / P
    # phase 2: per-pixel gradients from the GLOBAL sums (SURVEY.md A.6 / sunet_loss_bwd); "gradient of a weight" =
    # SUM over ranks of shard contributions, all-reduced through the product's bucket plan on a flat buffer
    d = max(0.0, 0.8 - cov)
    d_out = sg * (torch.sigmoid(o) - t) / S
    d_sel = sg * (1 - sg) * (bce / S - R / S ** 2 - 2 * lamb * d / P)
    d_aux = (torch.sigmoid(a) - t) / P
    sd = O.init_state_dict(0, "RGB", True)
    order = param_order(True)
    fg = FlatGrads({n: tuple(sd[n].shape) for n in order}, order, "cpu")
    # toy gradients: three scalars of the full-batch loss w.r.t. scales of the three logit maps, written into three
    # different gradient GROUPS of the real flat buffer; everything else gets a rank-dependent fill
    fg.flat.fill_(float(rank + 1))
    fg.views["conv1x1.bias"][0] = (d_out * o).sum()                       # group dec1
    fg.views["decoder_layer_4_1.1.bias"][0] = (d_sel * s).sum()           # group dec4
    fg.views["encoder_layer_1_1.1.bias"][0] = (d_aux * a).sum()           # group enc1
    buckets = bucket_plan(fg.group_ranges(), fg.total, DEFAULT_BUCKETS)
    covered = torch.zeros(fg.total, dtype=torch.bool)
    for tag in GROUP_ORDER:                               # the order SUNetPlan.backward reports the groups
        if tag in buckets:
            blo, bhi = buckets[tag]
            assert not covered[blo:bhi].any()
            dist.all_reduce(fg.flat[blo:bhi], op=dist.ReduceOp.SUM)
            covered[blo:bhi] = True
    all_covered = bool(covered.all())
    gw = torch.stack([fg.views["conv1x1.bias"][0], fg.views["decoder_layer_4_1.1.bias"][0],
                      fg.views["encoder_layer_1_1.1.bias"][0]]).double()
    fill_ok = float(fg.views["unpool2.weight"].flatten()[0]) == float(sum(range(1, world + 1)))
    # oracle on the full batch
    wf = torch.ones(3, requires_grad=True)
    l_sel, c_ref = O.selective_risk_b(wf[0] * out, wf[1] * sel, tgt, lamb=lamb)
    l_ref = l_sel + O.bce_with_logits_mean(wf[2] * aux, tgt)
    l_ref.backward()
    ok = (abs(loss_global - l_ref.item()) < 1e-5 and abs(cov - c_ref.item()) < 1e-6 and
          torch.allclose(gw.float(), wf.grad, rtol=1e-4, atol=1e-6))
    # averaging per-shard losses is NOT the same function (why the exchange is mandatory)
    l_local, _ = O.selective_risk_b(o, s, t, lamb=lamb)
    mean_of_shards = torch.tensor([l_local.item() * (hi - lo) / total])
    dist.all_reduce(mean_of_shards)
    differs = abs(mean_of_shards.item() + A / P - l_ref.item()) > 1e-3
    # learning-rate schedule replication (ReduceLR must not diverge across ranks): every rank steps its own
    # scheduler copy with the all-reduced epoch loss; the resulting learning rates must be bit-identical
    args = train_cli.parse_arguments(["--lr_sche", "ReduceLR", "--patience", "1", "--factor", "0.5", "--lr", "0.01"])
    holder = torch.optim.SGD([torch.zeros(1, requires_grad=True)], lr=args.lr)
    sched = train_cli.make_scheduler(args, holder)
    lrs = []
    for epoch_loss_local in (1.0, 0.9, 0.95, 0.97, 0.99, 1.2):
        e = torch.tensor([epoch_loss_local + 0.01 * rank], dtype=torch.float64)   # local values differ ...
        dist.all_reduce(e)                                                          # ... the reduced one does not
        holder.step()
        sched.step(float(e) / world)
        lrs.append(holder.param_groups[0]["lr"])
    lr_t = torch.tensor(lrs, dtype=torch.float64)
    gathered = [torch.zeros_like(lr_t) for _ in range(world)]
    dist.all_gather(gathered, lr_t)
    lr_same = all(torch.equal(gathered[0], t_) for t_ in gathered) and lrs[-1] < args.lr
    ret[rank] = dict(ok=bool(ok), differs=bool(differs), bounds=(lo, hi), same_paths=bool(same_paths),
                     all_covered=all_covered, fill_ok=bool(fill_ok), lr_same=bool(lr_same), lrs=lrs)
    dist.destroy_process_group()


def test_two_rank_product_dp_host_path():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0]["bounds"] == (0, 3) and ret[1]["bounds"] == (3, 5)
    for r in range(world):
        v = ret[r]
        assert v["ok"], "global loss / coverage / SUM-reduced gradients must equal the full-batch oracle"
        assert v["differs"], "mean of per-shard losses should differ from the global loss on skewed shards"
        assert v["same_paths"], "trainer exchange and selective_loss exchange must produce the same global sums"
        assert v["all_covered"] and v["fill_ok"], "the bucket plan must all-reduce every gradient exactly once"
        assert v["lr_same"], f"learning rates diverged across ranks: {v['lrs']}"
