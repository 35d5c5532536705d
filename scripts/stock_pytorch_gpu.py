"""The practical "kernel to beat" on the same B200: the reference's layer graph executed
by STOCK PyTorch (cuDNN / cuBLAS / ATen) — fp32 as the reference is written, and under
torch.autocast(bfloat16) with channels_last.  Development aid; prints patches/s.  Not part of the product.

    python scripts/stock_pytorch_gpu.py [batch] [steps]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import sunet_oracle as O


class StockUNetB(nn.Module):
    def __init__(self, sd):
        super().__init__()
        self.bufs = {k: v.clone().cuda() for k, v in sd.items() if "running" in k or "num_batches" in k}
        self.params = nn.ParameterDict({k.replace(".", "__"): nn.Parameter(v.clone().cuda()) for k, v in sd.items()
                                        if "running" not in k and "num_batches" not in k})

    def forward(self, x):
        sd = dict(self.bufs)
        for k, p in self.params.items():
            sd[k.replace("__", ".")] = p
        return O.unet_b_forward(sd, x, True, True, update_running=True)


def step(net, opt, x, label, autocast):
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        o, s, a = net(x)
    o, s, a = o.float(), s.float(), a.float()
    sg = torch.sigmoid(s)
    cov = sg.mean()
    loss = (F.binary_cross_entropy_with_logits(o, label, reduction="none") * sg).mean() / cov
    loss = loss + 2 * torch.clamp(0.8 - cov, min=0) ** 2 + F.binary_cross_entropy_with_logits(a, label)
    loss.backward()
    opt.step()
    return loss


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    torch.backends.cudnn.benchmark = True
    sd = O.init_state_dict(0, "RGB", True)
    x, label = O.synthetic_batch(batch, 256, seed=0)
    label = label.cuda()
    for autocast, cl in ((True, True), (False, False)):
        try:
            net = StockUNetB(sd)
            opt = torch.optim.Adam(net.parameters(), lr=1e-3)
            xx = x.cuda()
            if cl:
                xx = xx.contiguous(memory_format=torch.channels_last)
            for _ in range(2):
                step(net, opt, xx, label, autocast)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                loss = step(net, opt, xx, label, autocast)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / steps
            print(f"stock PyTorch {'autocast-bf16 channels_last' if autocast else 'fp32 (TF32 off)'}: batch {batch} "
                  f"{dt * 1e3:.1f} ms/step = {batch / dt:.1f} patches/s (loss {loss.item():.4f})", flush=True)
            del net, opt
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            print("stock run failed:", repr(e)[:300], flush=True)


if __name__ == "__main__":
    main()
