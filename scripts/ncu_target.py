"""Small single-kernel workloads for `ncu --set full` (B200_PROFILING.md: keep the profiled command short).

    python scripts/ncu_target.py conv64|conv128|conv256|conv256pro|wgrad256pro|wgrad128|wgrad256|wgrad64|bnbwd|bnfwd|headsbwd|headsfwd|hist|
                                 losssums|lossbwd|adam

Shapes are the bench shapes (batch 128 of 256x256 patches) of the corresponding SUNet_B layer.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from selectivenet_for_semantic_segmentation_binary_b200 import kernels as K


def bf(*shape):
    return torch.randn(*shape, device="cuda").to(torch.bfloat16)


def conv(B, H, W, Cin, Cout, iters=3, pro=False):
    x = bf(B, H, W, Cin)
    kw = dict(pro=(torch.rand(Cin, device="cuda") + 0.5, torch.randn(Cin, device="cuda") * 0.3)) if pro else {}
    w = (torch.randn(Cout, 9 * Cin, device="cuda") / (3 * Cin ** 0.5)).to(torch.bfloat16)
    y = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device="cuda")
    rows = K.conv_gemm_stat_rows(B, H, W, Cout, K.A_CONV3X3)
    st = torch.zeros(rows, Cout, 2, device="cuda")
    for _ in range(iters):
        K.conv_gemm(K.A_CONV3X3, (B, H, W), x, w, y, stats=st, **kw)


def wgrad(B, H, W, Ca, Cb, iters=3, pro=False):
    dy, x = bf(B, H, W, Ca), bf(B, H, W, Cb)
    kw = dict(b_pro=(torch.rand(Cb, device="cuda") + 0.5, torch.randn(Cb, device="cuda") * 0.3)) if pro else {}
    splits = K.wgrad_splits((B, H, W), dy, K.A_CONV3X3, x)
    part = torch.empty(splits, 9, Ca, Cb, device="cuda")
    for _ in range(iters):
        K.wgrad_gemm((B, H, W), dy, K.A_CONV3X3, x, part, **kw)


def bn(B, H, W, C, bwd, iters=3):
    y, dA = bf(B, H, W, C), bf(B, H, W, C)
    sc, sh, mu, istd = (torch.rand(C, device="cuda") + 0.5 for _ in range(4))
    out = torch.empty_like(y)
    ws = K.new_workspace("cuda")
    dg, db = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    for _ in range(iters):
        if bwd:
            K.bn_relu_pool_bwd(dA, None, y, sc, sh, mu, istd, sc, dg, db, out, ws)
        else:
            K.bn_relu_pool(y, sc, sh, out, None)


def bnpool(B, H, W, C, iters=3):
    """forward BN + ReLU + 2x2 pool + winner capture, as the three pooled encoder blocks run it"""
    y = bf(B, H, W, C)
    sc, sh = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.3
    a = torch.empty_like(y)
    pooled = torch.empty(B, H // 2, W // 2, C, dtype=torch.bfloat16, device="cuda")
    ywin = torch.empty_like(pooled)
    for _ in range(iters):
        K.bn_relu_pool(y, sc, sh, a, pooled, ywin=ywin)


def heads(bwd, iters=3, B=128):
    P = B * 256 * 256
    y = (torch.randn(B, 256, 256, 64, device="cuda") * 1.5 + 0.3).to(torch.bfloat16)
    sc, sh, mu, isd = (torch.rand(64, device="cuda") + 0.5 for _ in range(4))
    wts = [torch.randn(1, 64, 1, 1, device="cuda") for _ in range(3)]
    bss = [torch.randn(1, device="cuda") for _ in range(3)]
    ws = K.new_workspace("cuda")
    if bwd:
        dA = torch.empty_like(y)
        dl = torch.randn(3, P, device="cuda")
        dws, dbs = [torch.empty_like(w) for w in wts], [torch.empty_like(b) for b in bss]
        part = torch.empty(K.heads_bwd_bn_rows(P), 64, 2, device="cuda")
        for _ in range(iters):
            K.heads_bwd_bn(dl, y, sc, sh, mu, isd, wts, dA, dws, dbs, part, ws)
    else:
        logits = torch.empty(3, P, device="cuda")
        for _ in range(iters):
            K.bn_relu_heads(y, sc, sh, None, wts, bss, logits)


def small(which, iters=3):
    P = 512 * 256 * 256                     # 302+ MB per launch: larger than the 126 MB L2
    lg = torch.randn(3, P, device="cuda")
    tgt = (torch.rand(P, device="cuda") < 0.4).float()
    ws = K.new_workspace("cuda")
    sums = torch.zeros(4, dtype=torch.float64, device="cuda")
    if which == "hist":
        lab8 = tgt.to(torch.uint8)
        cnt = torch.zeros(6, dtype=torch.int64, device="cuda")
        for _ in range(iters):
            K.metric_hist(lg[0], lg[1], lab8, 1e-7, 1e-7, True, cnt)
    elif which == "losssums":
        for _ in range(iters):
            K.loss_sums(lg[0], lg[1], lg[2], tgt, sums, ws, pixels_out=sums[3:4])
    elif which == "lossbwd":
        K.loss_sums(lg[0], lg[1], lg[2], tgt, sums, ws, pixels_out=sums[3:4])
        d = torch.empty(3, P, device="cuda")
        for _ in range(iters):
            K.loss_bwd(lg[0], lg[1], lg[2], tgt, sums, P, 2.0, 0.8, None, None, d[0], d[1], d[2])
    elif which == "adam":
        from selectivenet_for_semantic_segmentation_binary_b200.optim import Adam
        p = [torch.nn.Parameter(torch.randn(64 * 1024 * 1024, device="cuda"))]
        p[0].grad = torch.randn_like(p[0])
        opt = Adam(p, lr=1e-3)
        for _ in range(iters):
            opt.step()


MODES = {
    "headsbwd": lambda: heads(True),
    "headsfwd": lambda: heads(False),
    "hist": lambda: small("hist"),
    "losssums": lambda: small("losssums"),
    "lossbwd": lambda: small("lossbwd"),
    "adam": lambda: small("adam"),
    "conv64": lambda: conv(128, 256, 256, 64, 64),
    "conv512": lambda: conv(128, 32, 32, 512, 512),
    "conv128": lambda: conv(128, 128, 128, 128, 128),
    "conv256": lambda: conv(128, 64, 64, 256, 256),
    "conv256pro": lambda: conv(128, 64, 64, 256, 256, pro=True),
    "wgrad256pro": lambda: wgrad(128, 64, 64, 256, 256, pro=True),
    "wgrad128": lambda: wgrad(128, 128, 128, 128, 128),
    "wgrad256": lambda: wgrad(128, 64, 64, 256, 256),
    "wgrad64": lambda: wgrad(128, 256, 256, 64, 64),
    "bnbwd": lambda: bn(128, 256, 256, 64, True),
    "bnpool": lambda: bnpool(128, 256, 256, 64),
    "bnfwd": lambda: bn(128, 256, 256, 64, False),
}

if __name__ == "__main__":
    MODES[sys.argv[1]]()
    torch.cuda.synchronize()
    print("done", sys.argv[1])
