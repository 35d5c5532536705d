"""Achieved HBM bandwidth of the stream kernels around the convolutions, at step-sized tensors.

    python scripts/ew_bw.py [batch]

Times each C-ABI call with CUDA events (tensors >> L2) and prints algorithmic GB/s
(bytes per element of each kernel) next to the measured HBM peak.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from selectivenet_for_semantic_segmentation_binary_b200 import kernels as K  # noqa: E402


def timeit(fn, iters=10):
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


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = "cuda"
    peak = 6533.8
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    ws = K.new_workspace(dev)
    bf = torch.bfloat16
    rows = []
    for (H, C) in [(256, 64), (128, 128), (64, 256), (32, 512)]:
        n = B * H * H * C
        y = (torch.randn(B, H, H, C, device=dev) * 1.5 + 0.3).to(bf)
        a = torch.empty_like(y)
        dA = torch.randn(B, H, H, C, device=dev).to(bf)
        dy = torch.empty_like(y)
        scale = torch.rand(C, device=dev) + 0.5
        shift = torch.randn(C, device=dev) * 0.2
        mean = torch.randn(C, device=dev) * 0.1
        invstd = torch.rand(C, device=dev) + 0.5
        dgamma, dbeta = torch.empty(C, device=dev), torch.empty(C, device=dev)
        t = timeit(lambda: K.bn_relu_pool(y, scale, shift, a, None))
        rows.append((f"bn_relu flat        {H}x{H}x{C}", 4 * n, t))
        t = timeit(lambda: K.bn_relu_pool_bwd(dA, None, y, scale, shift, mean, invstd, scale, dgamma, dbeta, dy, ws))
        rows.append((f"bn_bwd flat (r+a)   {H}x{H}x{C}", 10 * n, t))
        if C < 512:
            pooled = torch.empty(B, H // 2, H // 2, C, dtype=bf, device=dev)
            dP = torch.randn(B, H // 2, H // 2, C, device=dev).to(bf)
            dcat = torch.randn(B, H, H, 2 * C, device=dev).to(bf)
            t = timeit(lambda: K.bn_relu_pool(y, scale, shift, a, pooled))
            rows.append((f"bn_relu pool        {H}x{H}x{C}", int(4.5 * n), t))
            t = timeit(lambda: K.bn_relu_pool_bwd(dcat[..., C:], dP, y, scale, shift, mean, invstd, scale, dgamma,
                                                  dbeta, dy, ws))
            rows.append((f"bn_bwd pool (r+a)   {H}x{H}x{C}", 11 * n, t))
            del pooled, dP, dcat
        del y, a, dA, dy
    # heads: forward (y -> logits, activation not stored) and backward (y, dlogits -> dA + BN reduction rows)
    P = B * 256 * 256
    y = (torch.randn(B, 256, 256, 64, device=dev) * 1.5 + 0.3).to(bf)
    dA = torch.empty_like(y)
    sc, sh, mu, isd = (torch.rand(64, device=dev) + 0.5 for _ in range(4))
    wts = [torch.randn(1, 64, 1, 1, device=dev) for _ in range(3)]
    bss = [torch.randn(1, device=dev) for _ in range(3)]
    logits = torch.empty(3, P, device=dev)
    dl = torch.randn(3, P, device=dev)
    dws, dbs = [torch.empty_like(w) for w in wts], [torch.empty_like(b) for b in bss]
    part = torch.empty(K.heads_bwd_bn_rows(P), 64, 2, device=dev)
    t = timeit(lambda: K.bn_relu_heads(y, sc, sh, None, wts, bss, logits))
    rows.append(("bn_relu_heads (no a)  256x256x64", 2 * y.numel() + 12 * P, t))
    t = timeit(lambda: K.heads_bwd_bn(dl, y, sc, sh, mu, isd, wts, dA, dws, dbs, part, ws))
    rows.append(("heads_bwd_bn          256x256x64", 4 * y.numel() + 12 * P, t))
    a = torch.empty_like(y)
    t = timeit(lambda: K.heads_bwd(dl, a, wts, dA, dws, dbs, ws))
    rows.append(("heads_bwd (stored a)  256x256x64", 4 * y.numel() + 12 * P, t))
    # losses / metric / Adam (vectorised in round 2).  P2 = 512 patches of 256^2 so every array exceeds the 126 MB L2.
    P2 = 512 * 256 * 256
    lg = torch.randn(3, P2, device=dev)
    tgt = (torch.rand(P2, device=dev) < 0.4).float()
    lab8 = tgt.to(torch.uint8)
    sums = torch.zeros(4, dtype=torch.float64, device=dev)
    dlg = torch.empty(3, P2, device=dev)
    cnt = torch.zeros(6, dtype=torch.int64, device=dev)
    t = timeit(lambda: K.loss_sums(lg[0], lg[1], lg[2], tgt, sums, ws, pixels_out=sums[3:4]))
    rows.append(("loss_sums (3 logits + label f32)", 16 * P2, t))
    t = timeit(lambda: K.loss_bwd(lg[0], lg[1], lg[2], tgt, sums, P2, 2.0, 0.8, None, None, dlg[0], dlg[1], dlg[2]))
    rows.append(("loss_bwd (4 reads, 3 writes)", 28 * P2, t))
    t = timeit(lambda: K.metric_hist(lg[0], lg[1], lab8, 1e-7, 1e-7, True, cnt))
    rows.append(("metric_hist (2 logits + u8 label)", 9 * P2, t))
    t = timeit(lambda: K.metric_hist(lg[0], lg[1], tgt, 1e-7, 1e-7, True, cnt))
    rows.append(("metric_hist (2 logits + f32 label)", 12 * P2, t))
    del lg, dlg
    from selectivenet_for_semantic_segmentation_binary_b200.optim import Adam
    big = [torch.nn.Parameter(torch.randn(64 * 1024 * 1024, device=dev))]          # 268 MB per array
    big[0].grad = torch.randn_like(big[0])
    opt = Adam(big, lr=1e-3)
    t = timeit(lambda: opt.step())
    rows.append(("adam (one 64 Mi-element tensor)", 28 * big[0].numel(), t))
    for name, nbytes, t in rows:
        gbs = nbytes / t / 1e6
        print(f"{name:34s} {t:8.4f} ms  {gbs:8.1f} GB/s  {100 * gbs / peak:5.1f}% of {peak:.0f}", flush=True)


if __name__ == "__main__":
    main()
