"""Training-mode prologue fusion (``SUNET_FUSE_PROLOGUE=1``): the level-3 conv -> BN -> ReLU -> conv chains read the
producer's raw conv output and apply relu(scale * y + shift) to the staged tiles in shared memory, in the forward conv
and in the weight-gradient GEMM.  The transform uses the same fmaf / max / round-to-nearest-even as the stand-alone
BN+ReLU pass, so the whole step must be BIT-IDENTICAL to the unfused plan — logits, loss and every gradient.
Kernel-level bit-identity is in tests/test_gpu_kernels.py (g1_prologue, g2_wgrad_prologue)."""
import pytest
import torch

from oracle import sunet_oracle as O

pytestmark = pytest.mark.gpu


def _step(monkeypatch, fuse, batch, size):
    from selectivenet_for_semantic_segmentation_binary_b200.model import UNet_B
    from selectivenet_for_semantic_segmentation_binary_b200.selective_loss import (BCEWithLogitsLoss,
                                                                                   calc_selective_risk_image_b)
    monkeypatch.setenv("SUNET_FUSE_PROLOGUE", "1" if fuse else "0")
    torch.manual_seed(0)
    net = UNet_B("RGB", selective=True).cuda()
    x, label = O.synthetic_batch(batch, size, seed=3)
    net.train()
    out, sel, aux = net(x.cuda())
    loss = BCEWithLogitsLoss()(aux, label.cuda()) + calc_selective_risk_image_b(out, sel, target=label.cuda(), lamb=2)[0]
    loss.backward()
    torch.cuda.synchronize()
    plan = next(iter(net._plans.values()))
    fused = sorted(plan.pro_of)
    res = dict(out=out.detach().clone(), sel=sel.detach().clone(), aux=aux.detach().clone(), loss=loss.detach().clone())
    res.update({"grad:" + n: p.grad.detach().clone() for n, p in net.named_parameters()})
    res.update({"buf:" + n: b.detach().clone() for n, b in net.named_buffers()})
    return res, fused


@pytest.mark.parametrize("batch,size", [(2, 256), (1, 256)])
def test_fused_prologue_step_is_bit_identical(monkeypatch, batch, size):
    ref, none = _step(monkeypatch, False, batch, size)
    got, fused = _step(monkeypatch, True, batch, size)
    assert none == []
    assert fused == ["decoder_layer_3_1", "encoder_layer_3_2"], fused      # level 3 = 64 x 64: both kernels have the variant
    for k in ref:
        assert torch.equal(ref[k], got[k]), (k, float((ref[k].float() - got[k].float()).abs().max()))
