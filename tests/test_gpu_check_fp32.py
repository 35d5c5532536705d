"""fp32 CHECK MODE (SUNET_CHECK_FP32=1; BASELINE.json: "logits and loss ... within 1e-4 in an fp32 check mode"):
the same UNet_B / loss / trainer code on slow fp32 kernels must reproduce the CPU oracle to 1e-4 on logits and
losses and to a per-tensor gradient cosine of 0.99999 — a bound that catches a wrong tap, a swapped concat half, a
mis-routed pool gradient or a missing BatchNorm term, which the 0.90 bound of the bf16 path cannot."""
import numpy as np
import pytest
import torch

from oracle import sunet_oracle as O

pytestmark = pytest.mark.gpu


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-300)).item()


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("batch,h,w,selective", [(2, 64, 64, True), (3, 32, 64, True), (2, 32, 32, False)])
def test_fp32_check_mode_matches_oracle(monkeypatch, batch, h, w, selective):
    monkeypatch.setenv("SUNET_CHECK_FP32", "1")
    from selectivenet_for_semantic_segmentation_binary_b200.engine_fp32 import SUNetPlanF32
    from selectivenet_for_semantic_segmentation_binary_b200.model import UNet_B
    from selectivenet_for_semantic_segmentation_binary_b200.selective_loss import (BCEWithLogitsLoss,
                                                                                   calc_selective_risk_image_b)
    torch.manual_seed(0)
    net = UNet_B("RGB", selective=selective).cuda()
    sd = O.init_state_dict(0, "RGB", selective)
    names = [n for n, _ in net.named_parameters()]
    for n in names:
        sd[n].requires_grad_(True)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(batch, 3, h, w, generator=g) * 2 - 1
    label = (torch.rand(batch, h, w, generator=g) < 0.4).float()
    ref_loss, ref = O.train_losses(sd, x, label, s_lamb=2, selective=selective)
    ref_loss.backward()
    # the same oracle evaluated in float64: the yardstick for the gradient bound (two fp32 executions of this graph —
    # oneDNN on the CPU and any GPU path — already differ by ~3e-3 on the deepest tensor's gradient)
    sd64 = {k: (v.detach().double() if v.is_floating_point() else v.detach().clone()) for k, v in
            O.init_state_dict(0, "RGB", selective).items()}
    for n in names:
        sd64[n].requires_grad_(True)
    ref64_loss, ref64 = O.train_losses(sd64, x.double(), label.double(), s_lamb=2, selective=selective)
    ref64_loss.backward()
    net.train()
    lab = label.cuda()
    if selective:
        out, sel, aux = net(x.cuda())
        loss = BCEWithLogitsLoss()(aux, lab)
        s_loss, cov = calc_selective_risk_image_b(out, sel, target=lab, lamb=2)
        total = loss + s_loss
    else:
        out = net(x.cuda())
        total = BCEWithLogitsLoss()(out, lab)
    total.backward()
    torch.cuda.synchronize()
    assert isinstance(next(iter(net._plans.values())), SUNetPlanF32)
    assert _rel(out.detach().cpu(), ref["output"].detach()) < 1e-4
    if selective:
        assert _rel(sel.detach().cpu(), ref["selection"].detach()) < 1e-4
        assert _rel(aux.detach().cpu(), ref["aux"].detach()) < 1e-4
        assert abs(cov.item() - ref["coverage"].item()) < 1e-4 * ref["coverage"].item()
    assert abs(total.item() - ref_loss.item()) < 1e-4 * abs(ref_loss.item())
    params = dict(net.named_parameters())
    worst = 1.0
    for n in names:
        gg, gr = params[n].grad.cpu(), sd[n].grad
        if n.endswith(".0.bias") and "layer" in n:       # conv bias feeding BatchNorm: true gradient is exactly 0
            assert gg.abs().max().item() <= 1e-6, n
            continue
        c, c32 = _cos(gg, sd64[n].grad), _cos(gg, gr)
        worst = min(worst, c)
        assert c > 0.99999, (n, c)                       # vs the float64 evaluation of the reference graph
        assert c32 > 0.9999, (n, c32)                    # vs its fp32 CPU execution (which carries its own rounding)
        assert abs(gg.norm().item() / sd64[n].grad.norm().item() - 1) < 1e-3, n
    assert _rel(out.detach().cpu(), ref64["output"].detach()) < 1e-4
    print("fp32 check mode: worst gradient cosine vs the float64 oracle", worst)
    for k, v in net.named_buffers():                     # running statistics after the step
        if "num_batches" in k:
            assert int(v.item()) == 1
        else:
            assert _rel(v.cpu(), sd[k].detach()) < 1e-5, k


def test_fp32_check_mode_trainer_step_and_eval(monkeypatch):
    """The fused trainer (losses, Adam, Evaluator) and the eval-mode forward run unchanged on the check-mode plan."""
    monkeypatch.setenv("SUNET_CHECK_FP32", "1")
    from selectivenet_for_semantic_segmentation_binary_b200.model import UNet_B
    from selectivenet_for_semantic_segmentation_binary_b200.trainer import SUNetTrainer
    torch.manual_seed(2)
    net = UNet_B("RGB", selective=True).cuda()
    sd = O.init_state_dict(2, "RGB", True)
    names = [n for n, _ in net.named_parameters()]
    opt = torch.optim.Adam([sd[n].requires_grad_(True) for n in names], lr=1e-3)
    tr = SUNetTrainer(net, lr=1e-3, s_lamb=2, use_cuda_graph=False)
    x, label = O.synthetic_batch(2, 32, seed=7)
    for _ in range(3):
        opt.zero_grad()
        ref_loss, _ = O.train_losses(sd, x, label, s_lamb=2, selective=True)
        ref_loss.backward()
        opt.step()
        got = tr.step(x.cuda(), label.cuda())[3].item()
        assert abs(got - ref_loss.item()) < 2e-4 * abs(ref_loss.item()), (got, ref_loss.item())
    params = dict(net.named_parameters())
    # three Adam steps: early Adam moves every weight by ~lr * sign(g), so an element whose gradient is at rounding level
    # may step the other way; compare the update DIRECTION of whole tensors
    sd0 = O.init_state_dict(2, "RGB", True)
    for n in ("decoder_layer_4_1.0.weight", "encoder_layer_1_1.0.weight", "unpool2.weight", "conv_select.weight"):
        assert _cos(params[n].detach().cpu() - sd0[n], sd[n].detach() - sd0[n]) > 0.999, n
    net.eval()
    with torch.no_grad():
        out, sel, aux = net(x.cuda())
        r_out, r_sel, r_aux = O.unet_b_forward({k: v.detach().cpu() for k, v in net.state_dict().items()}, x, True,
                                               False)
    assert _rel(out.cpu(), r_out) < 1e-4 and _rel(sel.cpu(), r_sel) < 1e-4
