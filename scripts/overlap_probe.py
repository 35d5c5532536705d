"""Can an HBM-bound stream kernel (BN + ReLU apply) run BESIDE a tcgen05 GEMM (different streams) and hide behind it?
Times conv(A) + bn_relu(B) issued serially on one stream against the same two launches on two streams, per level.

    python scripts/overlap_probe.py [batch_per_half]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from selectivenet_for_semantic_segmentation_binary_b200 import kernels as K  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dev = "cuda"
    bf = torch.bfloat16
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for (H, C) in [(256, 64), (128, 128), (64, 256), (32, 512)]:
        x = torch.randn(B, H, H, C, device=dev).to(bf)
        w = (torch.randn(C, 9 * C, device=dev) / (3 * C ** 0.5)).to(bf)
        y = torch.empty(B, H, H, C, device=dev, dtype=bf)
        y2 = torch.randn(B, H, H, C, device=dev).to(bf)
        a2 = torch.empty_like(y2)
        dA = torch.randn(B, H, H, C, device=dev).to(bf)
        dy = torch.empty_like(y2)
        sc, sh, mu, isd = (torch.rand(C, device=dev) + 0.5 for _ in range(4))
        rows = K.conv_gemm_stat_rows(B, H, H, C)
        st = torch.zeros(rows, C, 2, device=dev)
        dg, db = torch.empty(C, device=dev), torch.empty(C, device=dev)
        ws = K.new_workspace(dev)

        def conv():
            K.conv_gemm(K.A_CONV3X3, (B, H, H), x, w, y, stats=st)

        def bn_fwd():
            K.bn_relu_pool(y2, sc, sh, a2, None)

        def bn_bwd():
            K.bn_bwd_apply(dA, y2, sc, sh, mu, isd, st, rows, dg, db, dy, ws)

        def timed(fn, iters=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters

        for name, ew in (("bn_relu", bn_fwd), ("bn_bwd_apply", bn_bwd)):
            t_conv, t_ew = timed(conv), timed(ew)

            def serial():
                conv()
                ew()

            def overlapped(first_gemm=True):
                cur = torch.cuda.current_stream()
                ev = torch.cuda.Event()
                ev.record(cur)
                s1.wait_event(ev)
                s2.wait_event(ev)
                if first_gemm:
                    with torch.cuda.stream(s1):
                        conv()
                    with torch.cuda.stream(s2):
                        ew()
                else:
                    with torch.cuda.stream(s2):
                        ew()
                    with torch.cuda.stream(s1):
                        conv()
                cur.wait_stream(s1)
                cur.wait_stream(s2)

            t_ser = timed(serial)
            t_ov = timed(lambda: overlapped(True))
            t_ov2 = timed(lambda: overlapped(False))
            print(f"{H}x{H}x{C} B{B} {name:13s}: conv {t_conv:.3f}  ew {t_ew:.3f}  serial {t_ser:.3f}  "
                  f"two streams (gemm first) {t_ov:.3f}  (ew first) {t_ov2:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
