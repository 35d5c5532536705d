"""Generate golden vectors by running the REAL reference (/root/reference) in the build container.

    python tests/golden/make_golden.py        # writes tests/golden/sunet_b_golden.npz

The reference is imported read-only.  Its loss hard-codes ``.cuda()`` (selective_loss.py:73); on this
CPU-only container the harness shims ``torch.Tensor.cuda`` to the identity — the reference source is
untouched.  /root/reference does not exist on the GPU box, so the vectors are committed and the
tests never import the reference.

Contents (all from the reference's own code paths):
  * UNet_B('RGB', selective=True), torch.manual_seed(0) default init, one training step on a
    seeded 2x3x32x32 batch: three logit maps, aux / selective / total loss, coverage, every
    parameter gradient (norm, sum, first 8 values), BN running statistics after the step;
  * the same model in eval mode on a second batch: logits;
  * train-path and eval-path post-processing (numpy sigmoid + threshold) and Evaluator results;
  * calc_selective_risk_image_b and Evaluator on the notebook examples.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def main():
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self      # harness-side shim, CPU container only
    import model as ref_model                            # /root/reference/model.py
    import selective_loss as ref_loss                    # /root/reference/selective_loss.py
    from utils.compute_metric import Evaluator as RefEvaluator

    from oracle import sunet_oracle as O

    out = {}
    torch.manual_seed(0)
    net = ref_model.UNet_B("RGB", selective=True)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    # the oracle's constructor restatement must draw the same initial weights
    osd = O.init_state_dict(0, "RGB", True)
    assert list(osd.keys()) == list(sd0.keys()), "state_dict key order differs"
    for k in sd0:
        assert torch.equal(osd[k], sd0[k]), k
    out["n_keys"] = np.array(len(sd0))
    out["n_params"] = np.array(sum(p.numel() for p in net.parameters()))

    x, label = O.synthetic_batch(2, 32, seed=0)
    net.train()
    output, selection, aux = net(x)
    aux_loss = torch.nn.BCEWithLogitsLoss()(aux, label)
    select_loss, coverage = ref_loss.calc_selective_risk_image_b(output, selection, target=label, lamb=2)
    loss = aux_loss + select_loss
    loss.backward()
    out["train_output"] = output.detach().numpy()
    out["train_selection"] = selection.detach().numpy()
    out["train_aux"] = aux.detach().numpy()
    out["aux_loss"] = aux_loss.detach().numpy()
    out["select_loss"] = select_loss.detach().numpy()
    out["coverage"] = coverage.detach().numpy()
    out["loss"] = loss.detach().numpy()
    names = [n for n, _ in net.named_parameters()]
    out["param_names"] = np.array(names)
    out["grad_norm"] = np.array([p.grad.norm().item() for _, p in net.named_parameters()])
    out["grad_sum"] = np.array([p.grad.double().sum().item() for _, p in net.named_parameters()])
    out["grad_head8"] = np.stack([np.resize(p.grad.reshape(-1)[:8].numpy(), 8) for _, p in net.named_parameters()])
    sd1 = net.state_dict()
    bn_keys = [k for k in sd1 if "running" in k or "num_batches" in k]
    out["bn_keys"] = np.array(bn_keys)
    out["bn_sum"] = np.array([sd1[k].double().sum().item() for k in bn_keys])
    out["bn_abs_sum"] = np.array([sd1[k].double().abs().sum().item() for k in bn_keys])

    # train-path post-processing (train.py:211-237)
    o_np, s_np = output.detach().numpy(), selection.detach().numpy()
    lab_u8 = label.numpy().astype("uint8")
    fn_sigmoid64 = lambda a: 1 / (1 + np.exp(-a.astype("float64")))
    pred = (1.0 * (fn_sigmoid64(o_np) > 0.5)).astype("uint8")
    sel = 1.0 * (fn_sigmoid64(s_np) > 0.5)
    ev = RefEvaluator(2, True)
    ev.add_batch(lab_u8, pred, selection=sel)
    out["train_cm"] = ev.confusion_matrix.copy()
    out["train_reject"] = np.array(lab_u8.size - sel.sum())
    out["train_acc"] = np.array(ev.get_Pixel_Accuracy())
    out["train_miou"] = np.array(ev.get_mIoU())

    # eval mode (running stats after one step) on a second batch; eval-path post-processing (eval.py:228-251)
    net.train(False)
    x2, label2 = O.synthetic_batch(2, 32, seed=10)
    with torch.no_grad():
        o2, s2, _ = net(x2)
    out["eval_output"] = o2.numpy()
    out["eval_selection"] = s2.numpy()
    fn_sigmoid32 = lambda a: 1 / (1 + np.exp(-a))
    for cut, scut, tag in ((0.5, 0.5, "a"), (0.3, 0.6, "b")):
        pred2 = (1.0 * (fn_sigmoid32(o2.numpy()) > cut)).astype("uint8")
        sel2 = 1.0 * (fn_sigmoid32(s2.numpy()) > scut)
        for selective, t2 in ((True, "sel"), (False, "all")):
            ev2 = RefEvaluator(2, selective)
            if selective:
                ev2.add_batch(label2.numpy().astype("uint8"), pred2, selection=sel2)
            else:
                ev2.add_batch(label2.numpy().astype("uint8"), pred2)
            out[f"eval_cm_{tag}_{t2}"] = ev2.confusion_matrix.copy()
        out[f"eval_reject_{tag}"] = np.array(o2.numel() - sel2.sum())

    # notebook examples
    target = torch.tensor([[[1., 0., 1.], [1., 1., 1.], [0., 0., 1.]]])
    logit1 = torch.tensor([[[1., 0., 1.], [1., 1., 0.], [0., 0., 0.]]])
    selx = torch.tensor([[[2., -1., .5], [3., 0., -2.], [1., 1., -.5]]])
    l0, c0 = ref_loss.calc_selective_risk_image_b(logit1, torch.zeros_like(logit1), target, lamb=2)
    l1, c1 = ref_loss.calc_selective_risk_image_b(logit1, selx, target, lamb=2)
    out["nb_sel_loss"] = np.array([l0.item(), c0.item(), l1.item(), c1.item()])
    out["nb_bce"] = np.array(torch.nn.BCEWithLogitsLoss()(logit1, target).item())

    path = os.path.join(HERE, "sunet_b_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
