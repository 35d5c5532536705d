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
    tag = f"WIDE={os.environ.get('SUNET_WGRAD64_WIDE', '1')} MAXC={os.environ.get('SUNET_WGRAD64_MAXC', '128')}"
    for (B, H, W, cout, c0, c1) in [(128, 256, 256, 64, 64, 0), (128, 256, 256, 64, 64, 64), (128, 128, 128, 128, 128, 0),
                                    (128, 128, 128, 128, 128, 128)]:
        g = torch.Generator().manual_seed(1)
        dy = torch.randn(B, H, W, cout, generator=g).to(torch.bfloat16).to(dev)
        x0 = torch.randn(B, H, W, c0, generator=g).to(torch.bfloat16).to(dev)
        b1 = torch.randn(B, H, W, c1, generator=g).to(torch.bfloat16).to(dev) if c1 else None
        splits = K.wgrad_splits((B, H, W), dy, K.A_CONV3X3, x0, b1)
        cin = c0 + c1
        part = torch.empty(splits, 9, cout, cin, device=dev)
        t = timed(lambda: K.wgrad_gemm((B, H, W), dy, K.A_CONV3X3, x0, part, b1))
        fl = 2.0 * B * H * W * cout * cin * 9
        print(f"{tag} wgrad {H}x{W} {c0}{'+' + str(c1) if c1 else ''}->{cout}: {t:.3f} ms  {fl / t / 1e9:.0f} TFLOP/s  splits {splits}",
              flush=True)
        del dy, x0, b1, part


if __name__ == "__main__":
    main()
