"""Properties at BASELINE.json's full patch size (256x256, one 8-GPU shard of 16 patches), where the CPU oracle is
too slow to be the checker: the fused and the un-fused forms of the step must agree, gradients must obey the
identities BatchNorm imposes, and the counting kernel must conserve pixels."""
import numpy as np
import pytest
import torch

from oracle import sunet_oracle as O

pytestmark = pytest.mark.gpu

B, S = 16, 256

FUSED_OFF = {"SUNET_FUSE_BNB": "0", "SUNET_FUSE_BNB_POOL": "0", "SUNET_FUSE_HEADS_BN": "0", "SUNET_FIRST_PAIR": "0", "SUNET_OVERLAP_WGRAD": "0"}


def _step(monkeypatch, env, batch=B):
    """One SUNet_B forward/backward on seeded inputs with the given environment; returns logits + gradients."""
    from selectivenet_for_semantic_segmentation_binary_b200.model import UNet_B
    from selectivenet_for_semantic_segmentation_binary_b200.selective_loss import (BCEWithLogitsLoss,
                                                                                   calc_selective_risk_image_b)
    for k in FUSED_OFF:
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    torch.manual_seed(0)
    net = UNet_B("RGB", selective=True).cuda()
    net.train()
    x, label = O.synthetic_batch(batch, S, seed=3)
    x, label = x.cuda(), label.cuda()
    out, sel, aux = net(x)
    loss = BCEWithLogitsLoss()(aux, label)
    s_loss, cov = calc_selective_risk_image_b(out, sel, target=label, lamb=2)
    (loss + s_loss).backward()
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().clone() for n, p in net.named_parameters()}
    bufs = {n: b.detach().clone() for n, b in net.named_buffers()}
    net._plans.clear()               # the plan owns ~230 MB of buffers per patch: free it before the next model
    del net
    torch.cuda.empty_cache()
    return dict(out=out.detach(), sel=sel.detach(), aux=aux.detach(), loss=float((loss + s_loss).detach()), cov=float(cov.detach()),
                grads=grads, bufs=bufs, label=label)


def _rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("batch", [16, 128])
def test_fused_and_unfused_steps_agree_at_full_size(monkeypatch, batch):
    """Fused BN-backward reductions (dgrad / ConvT-dgrad / heads epilogues), the paired-pixel first layer and the
    side-stream weight gradients are re-associations of the same arithmetic: logits must agree to fp32
    round-off, gradients to a small fraction of the bf16 noise floor.  batch 16 = one 8-GPU shard, batch 128 = the
    headline single-GPU configuration of BASELINE.json (8.4 M pixels per level-1 tensor)."""
    a = _step(monkeypatch, {}, batch)
    b = _step(monkeypatch, FUSED_OFF, batch)
    for k in ("out", "sel", "aux"):
        assert _rel_l2(a[k], b[k]) < 1e-5, k             # forward differs only in the first layer's K order
    assert abs(a["loss"] - b["loss"]) < 1e-5 * abs(b["loss"]) and abs(a["cov"] - b["cov"]) < 1e-6
    for n, v in b["bufs"].items():                       # running statistics / num_batches_tracked
        assert torch.allclose(a["bufs"][n].double(), v.double(), rtol=1e-5, atol=1e-7), n
    diffs = {}
    for n, g in b["grads"].items():
        if n.endswith(".0.bias") and "layer" in n:
            continue                                      # identically zero, checked below
        diffs[n] = _rel_l2(a["grads"][n], g)
    worst = max(diffs, key=diffs.get)
    print("fused vs unfused, gradient relative L2 difference per tensor (largest first):")
    for n in sorted(diffs, key=diffs.get, reverse=True)[:8]:
        print(f"   {n:32s} {diffs[n]:.2e}")
    # measured on B200: weights <= 4e-3 (deepest tensor, encoder_layer_1_1), BN affine gradients of the first blocks
    # up to 2e-2; the bf16 noise floor against the fp32 oracle is 1e-1 .. 3e-1 for the same tensors
    # (tests/test_gpu_model.py), so a bound of 5e-2 still separates "re-association" from "different function"
    assert diffs[worst] < 5e-2, (worst, diffs[worst])


def test_batchnorm_gradient_identities_at_full_size(monkeypatch):
    r = _step(monkeypatch, {})
    # a conv bias that feeds a BatchNorm has an identically zero gradient (sum of dy over pixels is 0)
    for n, g in r["grads"].items():
        if n.endswith(".0.bias") and "layer" in n:
            assert float(g.abs().max()) == 0.0, n
        else:
            assert bool(torch.isfinite(g).all()) and float(g.abs().max()) > 0.0, n
    # head bias gradients are plain sums of the per-pixel loss gradients: d aux bias = sum (sigmoid(aux) - t) / P
    P = B * S * S
    exp = float((torch.sigmoid(r["aux"].double()) - r["label"].double()).sum() / P)
    got = float(r["grads"]["conv_aux.bias"])
    assert abs(got - exp) < 1e-5 * max(1.0, abs(exp)) + 1e-7, (got, exp)


def test_histogram_conserves_pixels_at_full_size(monkeypatch):
    from selectivenet_for_semantic_segmentation_binary_b200.utils.compute_metric import Evaluator
    r = _step(monkeypatch, {})
    ev = Evaluator(2, True)
    ev.add_batch_from_logits(r["label"], r["out"], r["sel"], path="train")
    cm = ev.confusion_matrix
    out, sel, lab = r["out"].cpu().numpy(), r["sel"].cpu().numpy(), r["label"].cpu().numpy()
    pred, selm = O.postprocess(out, sel, path="train")          # numpy float64 sigmoid, the reference's rule
    oe = O.Evaluator(2, True)
    oe.add_batch(lab.astype(np.uint8), pred, selection=selm)
    assert (cm == oe.confusion_matrix).all()
    assert cm.sum() == selm.sum() and ev.total == B * S * S and ev.total_reject == B * S * S - selm.sum()
    # un-masked counting covers every pixel exactly once
    ev2 = Evaluator(2, False)
    ev2.add_batch_from_logits(r["label"], r["out"], None, path="train")
    assert ev2.confusion_matrix.sum() == B * S * S
