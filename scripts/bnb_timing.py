"""Per-launch time of the dgrad conv with and without the fused BN-backward epilogue (step-sized shapes)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from selectivenet_for_semantic_segmentation_binary_b200 import kernels as K  # noqa: E402


def timeit(fn, iters=8):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    dev = "cuda"
    bf = torch.bfloat16
    for (H, Cd, Cn) in [(256, 64, 64), (128, 128, 128), (64, 256, 256), (32, 512, 512), (32, 512, 256)]:
        dy = torch.randn(B, H, H, Cd, device=dev).to(bf)
        y = torch.randn(B, H, H, Cn, device=dev).to(bf)
        out = torch.empty(B, H, H, Cn, device=dev, dtype=bf)
        wd = (torch.randn(Cn, 9 * Cd, device=dev) / (3 * Cd ** 0.5)).to(bf)
        sc, sh, mu, isd = (torch.rand(Cn, device=dev) + 0.5 for _ in range(4))
        rows = K.conv_gemm_stat_rows(B, H, H, Cn)
        st = torch.zeros(rows, Cn, 2, device=dev)
        flops = 2.0 * B * H * H * Cn * 9 * Cd
        t0 = timeit(lambda: K.conv_gemm(K.A_CONV3X3, (B, H, H), dy, wd, out))
        t1 = timeit(lambda: K.conv_gemm(K.A_CONV3X3, (B, H, H), dy, wd, out, stats=st))
        t2 = timeit(lambda: K.conv_gemm(K.A_CONV3X3, (B, H, H), dy, wd, out, stats=st, bnb=(y, sc, sh, mu, isd)))
        print(f"{H}x{H} {Cd}->{Cn}: plain {t0:.3f} ms ({flops / t0 / 1e9:.0f} TF/s)  +stats {t1:.3f}  +bnb {t2:.3f} ms "
              f"({flops / t2 / 1e9:.0f} TF/s)", flush=True)
        del dy, y, out


if __name__ == "__main__":
    main()
