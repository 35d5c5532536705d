"""Plan footprint options: backward scratch of levels 2-4 aliased onto dead level-1 decoder tensors (default) and the
single level-1 dY buffer (default; SUNET_LOW_MEM=0 restores two) must not change a single bit of a training step, in eager mode and under
CUDA-graph replay (SUNetTrainer), and must shrink the plan by what DESIGN.md says."""
import pytest
import torch

from oracle import sunet_oracle as O

pytestmark = pytest.mark.gpu


def _module_step(monkeypatch, env, batch, size):
    from selectivenet_for_semantic_segmentation_binary_b200.model import UNet_B
    from selectivenet_for_semantic_segmentation_binary_b200.selective_loss import (BCEWithLogitsLoss,
                                                                                   calc_selective_risk_image_b)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    torch.manual_seed(0)
    net = UNet_B("RGB", selective=True).cuda()
    x, label = O.synthetic_batch(batch, size, seed=3)
    net.train()
    torch.cuda.synchronize()
    before = torch.cuda.memory_allocated()
    out, sel, aux = net(x.cuda())
    loss = BCEWithLogitsLoss()(aux, label.cuda()) + calc_selective_risk_image_b(out, sel, target=label.cuda(), lamb=2)[0]
    loss.backward()
    torch.cuda.synchronize()
    plan_bytes = torch.cuda.memory_allocated() - before
    res = dict(out=out.detach().clone(), sel=sel.detach().clone(), aux=aux.detach().clone(), loss=loss.detach().clone())
    res.update({"grad:" + n: p.grad.detach().clone() for n, p in net.named_parameters()})
    del net
    return res, plan_bytes


def test_aliased_scratch_is_bit_identical_and_smaller(monkeypatch):
    ref, b0 = _module_step(monkeypatch, {"SUNET_ALIAS_SCRATCH": "0", "SUNET_LOW_MEM": "0"}, 3, 128)
    got, b1 = _module_step(monkeypatch, {"SUNET_ALIAS_SCRATCH": "1", "SUNET_LOW_MEM": "0"}, 3, 128)
    low, b2 = _module_step(monkeypatch, {"SUNET_ALIAS_SCRATCH": "1", "SUNET_LOW_MEM": "1"}, 3, 128)      # the default
    for k in ref:
        assert torch.equal(ref[k], got[k]), (k, float((ref[k].float() - got[k].float()).abs().max()))
        assert torch.equal(ref[k], low[k]), (k, float((ref[k].float() - low[k].float()).abs().max()))
    unit = 3 * 128 * 128 * 64 * 2            # one level-1 activation tensor
    assert b0 - b1 >= 4.5 * unit - (1 << 22), (b0, b1)        # 4.56 units gone (allocator rounding: 2 MB blocks)
    assert b1 - b2 >= unit - (1 << 22), (b1, b2)


def test_aliased_scratch_under_graph_replay(monkeypatch):
    """Several trainer steps (eager warm-up, capture, replays): same parameters after 5 steps with and without aliasing."""
    from selectivenet_for_semantic_segmentation_binary_b200.model import UNet_B
    from selectivenet_for_semantic_segmentation_binary_b200.trainer import SUNetTrainer

    def run(alias):
        monkeypatch.setenv("SUNET_ALIAS_SCRATCH", alias)
        torch.manual_seed(0)
        net = UNet_B("RGB", selective=True).cuda()
        tr = SUNetTrainer(net, lr=1e-3)
        for i in range(5):
            x, label = O.synthetic_batch(2, 64, seed=10 + i)
            tr.step(x.cuda(), label.cuda())
        torch.cuda.synchronize()
        return {n: p.detach().clone() for n, p in net.named_parameters()}, tr.graph_active("train")

    a, ga = run("0")
    b, gb = run("1")
    assert ga and gb
    for n in a:
        assert torch.equal(a[n], b[n]), n
