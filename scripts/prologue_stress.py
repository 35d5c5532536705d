"""Stress of the training-prologue kernels at the headline shapes: N repetitions of (conv+prologue, wgrad+prologue)
against the unfused launches, bit-identity checked every time, with an unrelated HBM-heavy kernel running on a second
stream to perturb the timing of the TMA / transform / MMA hand-offs."""
import sys

import torch

sys.path.insert(0, ".")
from selectivenet_for_semantic_segmentation_binary_b200 import kernels as K  # noqa: E402


def main(reps):
    dev = "cuda"
    bad = 0
    noise_src = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    noise_dst = torch.empty_like(noise_src)
    side = torch.cuda.Stream()
    for (B, H, W, Cin, Cout) in [(128, 64, 64, 256, 256), (32, 64, 64, 256, 256), (7, 64, 64, 256, 256)]:
        g = torch.Generator().manual_seed(B)
        y = torch.randn(B, H, W, Cin, generator=g).to(dev).to(torch.bfloat16)
        dy = torch.randn(B, H, W, Cout, generator=g).to(dev).to(torch.bfloat16)
        wf = (torch.randn(Cout, 9 * Cin, generator=g) / 50).to(dev).to(torch.bfloat16)
        scale = ((torch.rand(Cin, generator=g) + 0.5) * torch.where(torch.rand(Cin, generator=g) < 0.1, -1.0, 1.0)).to(dev)
        shift = (torch.randn(Cin, generator=g) * 0.3).to(dev)
        a = torch.empty_like(y)
        K.bn_relu_pool(y, scale, shift, a)
        grid = (B, H, W)
        rows = K.conv_gemm_stat_rows(B, H, W, Cout)
        st_a, st_p = torch.zeros(rows, Cout, 2, device=dev), torch.zeros(rows, Cout, 2, device=dev)
        out_a, out_p = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device=dev), torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device=dev)
        splits = K.wgrad_splits(grid, dy, K.A_CONV3X3, a)
        part_a, part_p = torch.empty(splits, 9, Cout, Cin, device=dev), torch.empty(splits, 9, Cout, Cin, device=dev)
        K.conv_gemm(K.A_CONV3X3, grid, a, wf, out_a, stats=st_a)
        K.wgrad_gemm(grid, dy, K.A_CONV3X3, a, part_a)
        for r in range(reps):
            out_p.fill_(float("nan"))
            part_p.fill_(float("nan"))
            if r % 2:
                with torch.cuda.stream(side):
                    noise_dst.copy_(noise_src)
            K.conv_gemm(K.A_CONV3X3, grid, y, wf, out_p, stats=st_p, pro=(scale, shift))
            K.wgrad_gemm(grid, dy, K.A_CONV3X3, y, part_p, b_pro=(scale, shift))
            torch.cuda.synchronize()
            ok = torch.equal(out_a.view(torch.int16), out_p.view(torch.int16)) and torch.equal(st_a, st_p) and \
                torch.equal(part_a.view(torch.int32), part_p.view(torch.int32))
            if not ok:
                bad += 1
                print(f"MISMATCH B={B} rep {r}", flush=True)
        print(f"B={B}: {reps} repetitions done, mismatches so far {bad}", flush=True)
    print("STRESS", "PASS" if bad == 0 else "FAIL")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]) if len(sys.argv) > 1 else 40) else 0)
