"""Per-kernel A/B of the training prologue fusion at the level-3 shape of the headline step (128 x 64 x 64, 256 -> 256):
   unfused = bn_relu pass (y -> a) + conv(a) [+ wgrad(dy, a)]    fused = conv(y, pro) [+ wgrad(dy, y, b_pro)]
Prints ms per launch (CUDA events, 20 launches after 3 warm-ups; operands 268 MB each, larger than L2)."""
import sys

import torch

sys.path.insert(0, ".")
from selectivenet_for_semantic_segmentation_binary_b200 import kernels as K  # noqa: E402


def timed(fn, n=20):
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
    for (B, H, W, Cin, Cout) in [(128, 64, 64, 256, 256), (128, 32, 32, 512, 512)]:
        g = torch.Generator().manual_seed(1)
        y = torch.randn(B, H, W, Cin, generator=g).to(dev).to(torch.bfloat16)
        dy = torch.randn(B, H, W, Cout, generator=g).to(dev).to(torch.bfloat16)
        a = torch.empty_like(y)
        out = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device=dev)
        wf = (torch.randn(Cout, 9 * Cin, generator=g) / 50).to(dev).to(torch.bfloat16)
        scale = (torch.rand(Cin, generator=g) + 0.5).to(dev)
        shift = (torch.randn(Cin, generator=g) * 0.3).to(dev)
        rows = K.conv_gemm_stat_rows(B, H, W, Cout)
        st = torch.zeros(rows, Cout, 2, device=dev)
        grid = (B, H, W)
        t_bn = timed(lambda: K.bn_relu_pool(y, scale, shift, a))
        t_conv = timed(lambda: K.conv_gemm(K.A_CONV3X3, grid, a, wf, out, stats=st))
        line = f"B{B} {H}x{W} {Cin}->{Cout}: bn_relu {t_bn:.3f}  conv {t_conv:.3f}"
        if K.conv_gemm_pro_supported(K.A_CONV3X3, grid, y, wf, out):
            t_pro = timed(lambda: K.conv_gemm(K.A_CONV3X3, grid, y, wf, out, stats=st, pro=(scale, shift)))
            line += f"  conv+pro {t_pro:.3f} (unfused sum {t_bn + t_conv:.3f})"
        print(line, flush=True)
        splits = K.wgrad_splits(grid, dy, K.A_CONV3X3, a)
        part = torch.empty(splits, 9, Cout, Cin, device=dev)
        t_w = timed(lambda: K.wgrad_gemm(grid, dy, K.A_CONV3X3, a, part))
        line = f"    wgrad {t_w:.3f} (splits {splits})"
        if K.wgrad_pro_supported(grid, dy, K.A_CONV3X3, y):
            t_wp = timed(lambda: K.wgrad_gemm(grid, dy, K.A_CONV3X3, y, part, b_pro=(scale, shift)))
            line += f"  wgrad+pro {t_wp:.3f}"
        print(line, flush=True)


if __name__ == "__main__":
    main()
