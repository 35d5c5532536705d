"""Numerics calibration (development aid, GPU box): how far is the CUDA bf16 path from the fp32 CPU oracle,
and how far is STOCK PyTorch (cuDNN, torch.autocast(bf16)) from the same oracle on the same inputs?
Prints logit errors and per-tensor gradient cosines for both, so the parity bounds written in
tests/test_gpu_model.py are grounded in what bf16 compute of this network can deliver at all.

    python scripts/parity_report.py [batch size]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn as nn

from oracle import sunet_oracle as O


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-30)).item()


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12)).item()


class StockUNetB(nn.Module):
    """Plain nn.Module with the reference's graph (stock PyTorch ops), used only as the "library bf16" yardstick."""

    def __init__(self, sd):
        super().__init__()
        self.sd = {k: v.clone().cuda() for k, v in sd.items()}
        self.params = nn.ParameterDict({k.replace(".", "__"): nn.Parameter(v) for k, v in self.sd.items()
                                        if "running" not in k and "num_batches" not in k})

    def forward(self, x):
        sd = dict(self.sd)
        for k, p in self.params.items():
            sd[k.replace("__", ".")] = p
        return O.unet_b_forward(sd, x, True, True, update_running=False)


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    from selectivenet_for_semantic_segmentation_binary_b200.model import UNet_B
    from selectivenet_for_semantic_segmentation_binary_b200.selective_loss import (BCEWithLogitsLoss,
                                                                                   calc_selective_risk_image_b)
    torch.manual_seed(0)
    net = UNet_B("RGB", selective=True).cuda()
    sd = O.init_state_dict(0, "RGB", True)
    names = [n for n, _ in net.named_parameters()]
    stock = StockUNetB(sd)
    for n in names:
        sd[n].requires_grad_(True)
    x, label = O.synthetic_batch(batch, size, seed=3)
    ref_loss, ref = O.train_losses(sd, x, label, s_lamb=2, selective=True)
    ref_loss.backward()

    net.train()
    out, sel, aux = net(x.cuda())
    loss = BCEWithLogitsLoss()(aux, label.cuda()) + calc_selective_risk_image_b(out, sel, target=label.cuda(), lamb=2)[0]
    loss.backward()

    with torch.autocast("cuda", dtype=torch.bfloat16):
        o2, s2, a2 = stock(x.cuda())
    o2, s2, a2 = o2.float(), s2.float(), a2.float()
    l2 = O.bce_with_logits_mean(a2, label.cuda())
    sg = torch.sigmoid(s2)
    cov = sg.mean()
    risk = (torch.nn.functional.binary_cross_entropy_with_logits(o2, label.cuda(), reduction="none") * sg).mean() / cov
    l2 = l2 + risk + 2 * torch.clamp(0.8 - cov, min=0) ** 2
    l2.backward()
    torch.cuda.synchronize()

    print(f"batch {batch} size {size}")
    print(f"loss: oracle {ref_loss.item():.6f}  ours {loss.item():.6f}  stock-bf16 {l2.item():.6f}")
    for nm, g, s_, r in (("output", out, o2, ref["output"]), ("selection", sel, s2, ref["selection"]),
                         ("aux", aux, a2, ref["aux"])):
        print(f"  logits {nm:10s} rel-err ours {rel(g.detach().cpu(), r.detach()):.4f}   stock-bf16 "
              f"{rel(s_.detach().cpu(), r.detach()):.4f}")
    params = dict(net.named_parameters())
    print(f"  {'tensor':34s} {'cos ours':>9s} {'cos stock':>9s} {'norm ratio ours':>15s}")
    for n in names:
        gr = sd[n].grad
        g = params[n].grad.cpu()
        gs = stock.params[n.replace(".", "__")].grad.cpu()
        if n.endswith(".0.bias"):
            print(f"  {n:34s} |g|max ours {g.abs().max().item():.2e} stock {gs.abs().max().item():.2e} "
                  f"oracle {gr.abs().max().item():.2e}")
            continue
        print(f"  {n:34s} {cos(g, gr):9.5f} {cos(gs, gr):9.5f} {g.norm().item() / gr.norm().item():15.4f}")


if __name__ == "__main__":
    main()
