"""Timing of the Cout = 64 weight-gradient GEMM at the level-1 shapes of the headline step (128 x 256^2):
run once with SUNET_WGRAD64_WIDE=0 (tall form, wgrad64_kernel) and once with =1 (N = 192 form, wgrad64n_kernel)."""
import os
import sys

import torch

sys.path.insert(0, ".")
from selectivenet_for_semantic_segmentation_binary_b200 import kernels as K  # noqa: E402


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = "cuda"
    B, H, W = 128, 256, 256
    g = torch.Generator().manual_seed(1)
    dy = torch.randn(B, H, W, 64, generator=g).to(torch.bfloat16).to(dev)
    x0 = torch.randn(B, H, W, 64, generator=g).to(torch.bfloat16).to(dev)
    x1 = torch.randn(B, H, W, 64, generator=g).to(torch.bfloat16).to(dev)
    for name, b1 in (("64->64", None), ("(64+64)->64", x1)):
        splits = K.wgrad_splits((B, H, W), dy, K.A_CONV3X3, x0, b1)
        cin = 64 if b1 is None else 128
        part = torch.empty(splits, 9, 64, cin, device=dev)
        t = timed(lambda: K.wgrad_gemm((B, H, W), dy, K.A_CONV3X3, x0, part, b1))
        fl = 2.0 * B * H * W * 64 * cin * 9
        print(f"WIDE={os.environ.get('SUNET_WGRAD64_WIDE', '1')} wgrad {name}: {t:.3f} ms  {fl / t / 1e9:.0f} TFLOP/s  splits {splits}",
              flush=True)


if __name__ == "__main__":
    main()
