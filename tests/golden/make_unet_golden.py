"""Golden vectors of the reference's cross-entropy variant (UNet + CrossEntropyLoss + calc_selective_risk_image),
from the REAL reference in the build container (same harness shim as make_golden.py for the loss's ``.cuda()``).

    python tests/golden/make_unet_golden.py        # writes tests/golden/unet_ce_golden.npz
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
    from oracle import sunet_oracle as O

    out = {}
    torch.manual_seed(0)
    net = ref_model.UNet("RGB", 2, selective=True)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    osd = O.init_state_dict(0, "RGB", True, n_cls=2)
    assert list(osd.keys()) == list(sd0.keys())
    for k in sd0:
        assert torch.equal(osd[k], sd0[k]), k
    out["n_params"] = np.array(sum(p.numel() for p in net.parameters()))
    x, label = O.synthetic_batch(2, 32, seed=0)
    label = label.long()
    net.train()
    output, selection, aux = net(x)
    aux_loss = torch.nn.CrossEntropyLoss()(aux, label)
    select_loss, coverage = ref_loss.calc_selective_risk_image(output, selection, target=label, lamb=2)
    (aux_loss + select_loss).backward()
    out["output"], out["selection"], out["aux"] = (t.detach().numpy() for t in (output, selection, aux))
    out["aux_loss"], out["select_loss"], out["coverage"] = (t.detach().numpy() for t in (aux_loss, select_loss, coverage))
    names = [n for n, _ in net.named_parameters()]
    out["param_names"] = np.array(names)
    for n, p in net.named_parameters():
        g = p.grad.detach()
        out[f"g_norm/{n}"] = np.array(float(g.norm()))
        out[f"g_head/{n}"] = g.flatten()[:8].numpy()
    for h in ("conv1x1", "conv_select", "conv_aux"):
        out[f"g_full/{h}.weight"] = dict(net.named_parameters())[f"{h}.weight"].grad.numpy()
        out[f"g_full/{h}.bias"] = dict(net.named_parameters())[f"{h}.bias"].grad.numpy()
    pred = np.argmax(output.detach().numpy().transpose(0, 2, 3, 1), axis=-1).astype("uint8")   # train.py:216-217
    selm = np.argmax(selection.detach().numpy().transpose(0, 2, 3, 1), -1).astype("uint8")      # train.py:224-226
    out["pred"], out["selm"] = pred, selm
    np.savez_compressed(os.path.join(HERE, "unet_ce_golden.npz"), **out)
    print("aux", float(aux_loss), "select", float(select_loss), "coverage", float(coverage), "params", int(out["n_params"]))


if __name__ == "__main__":
    main()
