"""The reference's cross-entropy variant (SURVEY.md §8(f) "next" #3): ``UNet(input_type, n_cls=2, selective)`` with
``CrossEntropyLoss`` and ``calc_selective_risk_image``.  CPU: the oracle restatement against golden vectors produced
by the reference itself (tests/golden/make_unet_golden.py).  GPU: the CUDA path against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import sunet_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ug():
    return np.load(os.path.join(ROOT, "tests", "golden", "unet_ce_golden.npz"))


def _oracle_step(seed=0, batch=2, size=32):
    sd = O.init_state_dict(seed, "RGB", True, n_cls=2)
    names = [k for k in sd if not any(t in k for t in ("running_", "num_batches"))]
    for n in names:
        sd[n].requires_grad_(True)
    x, label = O.synthetic_batch(batch, size, seed=0)
    loss, ref = O.train_losses_ce(sd, x, label.long(), s_lamb=2)
    loss.backward()
    return sd, names, x, label, loss, ref


def test_oracle_ce_variant_matches_reference_golden(ug):
    sd, names, x, label, loss, ref = _oracle_step()
    assert names == [str(s) for s in ug["param_names"]]
    assert sum(sd[n].numel() for n in names) == int(ug["n_params"]) == 7703302
    for k in ("output", "selection", "aux"):
        assert ref[k].shape == ug[k].shape and np.allclose(ref[k].detach().numpy(), ug[k], rtol=1e-5, atol=1e-6), k
    assert abs(float(ref["aux_loss"]) - float(ug["aux_loss"])) < 1e-6
    assert abs(float(ref["select_loss"]) - float(ug["select_loss"])) < 1e-6
    assert abs(float(ref["coverage"]) - float(ug["coverage"])) < 1e-7
    for n in names:
        g = sd[n].grad
        assert abs(float(g.norm()) - float(ug[f"g_norm/{n}"])) <= 1e-4 * float(ug[f"g_norm/{n}"]) + 1e-7, n
    pred, selm = O.postprocess_ce(ref["output"].detach().numpy(), ref["selection"].detach().numpy())
    assert np.array_equal(pred, ug["pred"]) and np.array_equal(selm, ug["selm"])
    # two-class identity the CUDA path relies on: CE / softmax selection == BCE / sigmoid of the channel difference
    d = ref["aux"][:, 1] - ref["aux"][:, 0]
    assert abs(float(O.bce_with_logits_mean(d, label)) - float(ref["aux_loss"])) < 1e-6
    sl, cov = O.selective_risk_b(ref["output"][:, 1] - ref["output"][:, 0], ref["selection"][:, 1] - ref["selection"][:, 0],
                                 label, lamb=2)
    assert abs(float(sl) - float(ref["select_loss"])) < 1e-6 and abs(float(cov) - float(ref["coverage"])) < 1e-7


def test_unet_module_layout_matches_reference(ug):
    from selectivenet_for_semantic_segmentation_binary_b200.model import UNet
    torch.manual_seed(0)
    net = UNet("RGB", 2, selective=True)
    sd = O.init_state_dict(0, "RGB", True, n_cls=2)
    assert list(net.state_dict().keys()) == list(sd.keys())
    for k, v in net.state_dict().items():
        assert torch.equal(v, sd[k]), k
    assert [n for n, _ in net.named_parameters()] == [str(s) for s in ug["param_names"]]
    with pytest.raises(NotImplementedError):
        UNet("RGB", 3)


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-30)).item()


@pytest.mark.gpu
@pytest.mark.parametrize("batch,size", [(2, 32), (3, 64)])
def test_unet_ce_forward_backward_parity(batch, size):
    from selectivenet_for_semantic_segmentation_binary_b200.model import UNet
    from selectivenet_for_semantic_segmentation_binary_b200.selective_loss import (CrossEntropyLoss,
                                                                                   calc_selective_risk_image)
    sd, names, x, label, ref_loss, ref = _oracle_step(batch=batch, size=size)
    torch.manual_seed(0)
    net = UNet("RGB", 2, selective=True).cuda()
    net.train()
    out, sel, aux = net(x.cuda())
    lab = label.long().cuda()
    aux_loss = CrossEntropyLoss()(aux, lab)
    s_loss, cov = calc_selective_risk_image(out, sel, target=lab, lamb=2)
    (aux_loss + s_loss).backward()
    torch.cuda.synchronize()

    def rel(a, b):
        return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()

    for got, want in ((out, ref["output"]), (sel, ref["selection"]), (aux, ref["aux"])):
        assert got.shape == want.shape
        assert rel(got.detach().cpu(), want.detach()) < 2e-2              # north_star: 2e-2 relative in bf16
    assert abs(float(aux_loss) - float(ref["aux_loss"])) < 2e-2 * abs(float(ref["aux_loss"]))
    assert abs(float(s_loss) - float(ref["select_loss"])) < 2e-2 * abs(float(ref["select_loss"]))
    assert abs(float(cov) - float(ref["coverage"])) < 2e-2
    params = dict(net.named_parameters())
    for n in names:
        g = params[n].grad.cpu()
        if n.endswith(".0.bias") and "layer" in n:
            assert float(g.abs().max()) <= 1e-6                            # zero by construction (BatchNorm follows)
            continue
        assert _cos(g, sd[n].grad) >= 0.90, (n, _cos(g, sd[n].grad))
    for h in ("conv1x1", "conv_select", "conv_aux"):                        # head gradients: tight
        assert _cos(params[f"{h}.weight"].grad.cpu(), sd[f"{h}.weight"].grad) > 0.999, h
        assert torch.allclose(params[f"{h}.bias"].grad.cpu(), sd[f"{h}.bias"].grad, rtol=2e-2, atol=1e-4), h
    # argmax masks / counts exact given our logits
    pred, selm = O.postprocess_ce(out.detach().cpu().numpy(), sel.detach().cpu().numpy())
    from selectivenet_for_semantic_segmentation_binary_b200.utils.compute_metric import Evaluator
    ev = Evaluator(2, True)
    ev.add_batch(label.numpy().astype("uint8"), pred, selection=selm.astype(np.float64))
    oe = O.Evaluator(2, True)
    oe.add_batch(label.numpy().astype("uint8"), pred, selection=selm.astype(np.float64))
    assert (ev.confusion_matrix == oe.confusion_matrix).all()
