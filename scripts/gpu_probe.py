"""Kernel-by-kernel bring-up probe (run on a B200 through gpurun).

Each group runs in its own process (a trapped kernel poisons its CUDA context) and compares one
C-ABI entry point against plain PyTorch on the same bf16-rounded inputs.  This is a development
aid; the graded parity tests live in tests/.

    python scripts/gpu_probe.py            # run every group, each in a subprocess
    python scripts/gpu_probe.py g1_plain   # run one group in-process
"""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn.functional as F


def rel_err(got, ref):
    got = got.float()
    ref = ref.float()
    denom = ref.abs().max().clamp_min(1e-6)
    return ((got - ref).abs().max() / denom).item()


def report(name, got, ref, tol=2e-2):
    e = rel_err(got, ref)
    ok = e <= tol and torch.isfinite(got.float()).all().item()
    print(f"  [{'OK ' if ok else 'BAD'}] {name}: max-rel-err {e:.3e} (ref max {ref.abs().max().item():.3e})", flush=True)
    return ok


def nhwc(x):  # NCHW fp32 -> NHWC bf16
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw(x):  # NHWC bf16 -> NCHW fp32
    return x.float().permute(0, 3, 1, 2).contiguous()


def conv3x3_case(K, B, H, W, Cin, Cout, dual=False, stats=True):
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + H + Cin + Cout)
    x = torch.randn(B, Cin, H, W, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).to(dev)
    xb = nhwc(x)
    wf = torch.empty(Cout, 9 * Cin, dtype=torch.bfloat16, device=dev)
    wd = torch.empty(Cin, 9 * Cout, dtype=torch.bfloat16, device=dev)
    K.pack_conv3x3_weights(w, wf, wd)
    y = torch.full((B, H, W, Cout), float("nan"), dtype=torch.bfloat16, device=dev)
    rows = K.conv_gemm_stat_rows(B, H, W, Cout, K.A_CONV3X3)
    st = torch.zeros(rows, Cout, 2, device=dev)
    if dual:
        c0 = Cin // 2
        s0 = xb[..., :c0].contiguous()
        s1 = xb[..., c0:].contiguous()
        K.conv_gemm(K.A_CONV3X3, (B, H, W), s0, wf, y, src1=s1, stats=st if stats else None)
    else:
        K.conv_gemm(K.A_CONV3X3, (B, H, W), xb, wf, y, stats=st if stats else None)
    torch.cuda.synchronize()
    ref = F.conv2d(nchw(xb), w.to(torch.bfloat16).float(), padding=1)
    ok = report(f"conv3x3 fwd B{B} {H}x{W} {Cin}->{Cout}{' dual' if dual else ''}", nchw(y), ref)
    if stats:
        yb = y.float().reshape(-1, Cout)
        ok &= report("   stats sum", st[..., 0].sum(0), yb.sum(0), 1e-3)
        ok &= report("   stats sumsq", st[..., 1].sum(0), (yb * yb).sum(0), 1e-3)
    # dgrad through the same kernel with the flipped/transposed pack
    dy = torch.randn(B, Cout, H, W, generator=g).to(dev)
    dyb = nhwc(dy)
    dx = torch.full((B, H, W, Cin), float("nan"), dtype=torch.bfloat16, device=dev)
    K.conv_gemm(K.A_CONV3X3, (B, H, W), dyb, wd, dx)
    torch.cuda.synchronize()
    ref_dx = F.conv_transpose2d(nchw(dyb), w.to(torch.bfloat16).float(), padding=1)
    ok &= report("   dgrad", nchw(dx), ref_dx)
    return ok


def g1_bnb(K):
    """dgrad conv with the fused BatchNorm-backward reduction in its epilogue (conv_gemm bnb=...), followed by
    bn_bwd_apply: against the unfused pair (conv_gemm + bn_relu_pool_bwd) and against plain torch."""
    dev = "cuda"
    ok = True
    ws = K.new_workspace(dev)
    for (B, H, W, Cd, Cn) in [(2, 32, 32, 64, 64), (3, 16, 16, 128, 128), (1, 16, 48, 256, 256), (2, 16, 16, 512, 512),
                              (5, 16, 16, 64, 128), (9, 64, 64, 64, 64)]:
        g = torch.Generator().manual_seed(B * 100 + Cd + Cn)
        # dy of the producer layer (Cd channels) -> dA of the consumer block (Cn channels)
        w = (torch.randn(Cd, Cn, 3, 3, generator=g) / (3 * Cn ** 0.5)).to(dev)
        wf = torch.empty(Cd, 9 * Cn, dtype=torch.bfloat16, device=dev)
        wd = torch.empty(Cn, 9 * Cd, dtype=torch.bfloat16, device=dev)
        K.pack_conv3x3_weights(w, wf, wd)
        dyb = nhwc(torch.randn(B, Cd, H, W, generator=g).to(dev))
        yb = nhwc((torch.randn(B, Cn, H, W, generator=g) * 1.5 + 0.3).to(dev))
        scale = (torch.rand(Cn, generator=g) + 0.5).to(dev) * torch.where(torch.rand(Cn, generator=g) < 0.1, -1.0, 1.0).to(dev)
        shift = (torch.randn(Cn, generator=g) * 0.5).to(dev)
        mean = (torch.randn(Cn, generator=g) * 0.3).to(dev)
        invstd = (torch.rand(Cn, generator=g) + 0.5).to(dev)
        if not K.conv_gemm_bnb_supported(K.A_CONV3X3, (B, H, W), dyb, wd, torch.empty_like(yb)):
            print(f"   bnb unsupported for B{B} {H}x{W} {Cd}->{Cn}: skipped")
            continue
        rows = K.conv_gemm_stat_rows(B, H, W, Cn)
        st = torch.full((rows, Cn, 2), float("nan"), device=dev)
        dA = torch.full((B, H, W, Cn), float("nan"), dtype=torch.bfloat16, device=dev)
        K.conv_gemm(K.A_CONV3X3, (B, H, W), dyb, wd, dA, stats=st, bnb=(yb, scale, shift, mean, invstd))
        dA0 = torch.full_like(dA, float("nan"))
        K.conv_gemm(K.A_CONV3X3, (B, H, W), dyb, wd, dA0)
        torch.cuda.synchronize()
        tag = f"bnb B{B} {H}x{W} {Cd}->{Cn}"
        ok &= bool(torch.equal(dA, dA0))
        print(f"  [{'OK ' if torch.equal(dA, dA0) else 'BAD'}] {tag}: dA identical to the unfused dgrad")
        gm = dA.float() * ((yb.float() * scale + shift) > 0)
        xhat = (yb.float() - mean) * invstd
        ok &= report(f"   {tag} sum g", st[..., 0].sum(0), gm.reshape(-1, Cn).sum(0), 1e-3)
        ok &= report(f"   {tag} sum g*xhat", st[..., 1].sum(0), (gm * xhat).reshape(-1, Cn).sum(0), 1e-3)
        # end to end: fused rows + bn_bwd_apply  ==  bn_relu_pool_bwd on the same dA
        dg1, db1, dg2, db2 = (torch.empty(Cn, device=dev) for _ in range(4))
        dy1, dy2 = torch.empty_like(yb), torch.empty_like(yb)
        K.bn_bwd_apply(dA, yb, scale, shift, mean, invstd, st, rows, dg1, db1, dy1, ws)
        K.bn_relu_pool_bwd(dA0, None, yb, scale, shift, mean, invstd, scale, dg2, db2, dy2, ws)
        torch.cuda.synchronize()
        ok &= report(f"   {tag} dgamma", dg1, dg2, 1e-4)
        ok &= report(f"   {tag} dbeta", db1, db2, 1e-4)
        ok &= report(f"   {tag} dy", dy1.float(), dy2.float(), 4e-3)
    return ok


def g1_bnb_convT(K):
    """ConvTranspose backward-data GEMM (2x2 gather) with the fused BN-backward reduction."""
    dev = "cuda"
    ok = True
    ws = K.new_workspace(dev)
    for (B, h, w_, Cin, Cout) in [(2, 8, 8, 512, 256), (3, 16, 16, 256, 128), (5, 16, 32, 128, 64)]:
        g = torch.Generator().manual_seed(B + Cin)
        wT = (torch.randn(Cin, Cout, 2, 2, generator=g) / (2 * Cout ** 0.5)).to(dev)
        bias = torch.zeros(Cout, device=dev)
        wf = torch.empty(4 * Cout, Cin, dtype=torch.bfloat16, device=dev)
        wd = torch.empty(Cin, 4 * Cout, dtype=torch.bfloat16, device=dev)
        b4 = torch.empty(4 * Cout, device=dev)
        K.pack_convT_weights(wT, bias, wf, wd, b4)
        dup = nhwc(torch.randn(B, Cout, 2 * h, 2 * w_, generator=g).to(dev))
        yb = nhwc((torch.randn(B, Cin, h, w_, generator=g) * 1.5 + 0.3).to(dev))
        scale = (torch.rand(Cin, generator=g) + 0.5).to(dev)
        shift = (torch.randn(Cin, generator=g) * 0.5).to(dev)
        mean = (torch.randn(Cin, generator=g) * 0.3).to(dev)
        invstd = (torch.rand(Cin, generator=g) + 0.5).to(dev)
        dA = torch.full((B, h, w_, Cin), float("nan"), dtype=torch.bfloat16, device=dev)
        dA0 = torch.full_like(dA, float("nan"))
        if not K.conv_gemm_bnb_supported(K.A_GATHER2X2, (B, h, w_), dup, wd, dA):
            print(f"   convT bnb unsupported for {Cin}->{Cout}: skipped")
            continue
        rows = K.conv_gemm_stat_rows(B, h, w_, Cin, K.A_GATHER2X2)
        st = torch.full((rows, Cin, 2), float("nan"), device=dev)
        K.conv_gemm(K.A_GATHER2X2, (B, h, w_), dup, wd, dA, stats=st, bnb=(yb, scale, shift, mean, invstd))
        K.conv_gemm(K.A_GATHER2X2, (B, h, w_), dup, wd, dA0)
        torch.cuda.synchronize()
        tag = f"convT-dgrad bnb B{B} {h}x{w_} {Cin}<-{Cout}"
        same = bool(torch.equal(dA, dA0))
        print(f"  [{'OK ' if same else 'BAD'}] {tag}: dA identical to the unfused launch")
        ok &= same
        dg1, db1, dg2, db2 = (torch.empty(Cin, device=dev) for _ in range(4))
        dy1, dy2 = torch.empty_like(yb), torch.empty_like(yb)
        K.bn_bwd_apply(dA, yb, scale, shift, mean, invstd, st, rows, dg1, db1, dy1, ws)
        K.bn_relu_pool_bwd(dA0, None, yb, scale, shift, mean, invstd, scale, dg2, db2, dy2, ws)
        torch.cuda.synchronize()
        ok &= report(f"   {tag} dgamma", dg1, dg2, 1e-4)
        ok &= report(f"   {tag} dbeta", db1, db2, 1e-4)
        ok &= report(f"   {tag} dy", dy1.float(), dy2.float(), 4e-3)
    return ok


def g1_bnb_pool(K):
    """Pooled block: reduction split between the decoder dgrad ([d_up | d_skip], bnb_col0 = C) and the encoder dgrad
    (d_pool against ywin), folded by bn_pool_bwd_apply — against the un-fused bn_relu_pool_bwd."""
    dev = "cuda"
    ok = True
    ws = K.new_workspace(dev)
    for (B, H, W, Cc, Cd1, Cd2) in [(2, 32, 32, 64, 64, 128), (3, 32, 64, 128, 128, 256), (2, 32, 32, 256, 256, 512)]:
        g = torch.Generator().manual_seed(B * 7 + Cc)
        yb = nhwc((torch.randn(B, Cc, H, W, generator=g) * 1.5 + 0.3).to(dev))
        scale = (torch.rand(Cc, generator=g) + 0.5).to(dev) * torch.where(torch.rand(Cc, generator=g) < 0.1, -1.0, 1.0).to(dev)
        shift = (torch.randn(Cc, generator=g) * 0.5).to(dev)
        mean = (torch.randn(Cc, generator=g) * 0.3).to(dev)
        invstd = (torch.rand(Cc, generator=g) + 0.5).to(dev)
        a = torch.empty_like(yb)
        pooled = torch.empty(B, H // 2, W // 2, Cc, dtype=torch.bfloat16, device=dev)
        ywin = torch.full_like(pooled, float("nan"))
        K.bn_relu_pool(yb, scale, shift, a, pooled, ywin=ywin)
        a2, pooled2 = torch.empty_like(a), torch.empty_like(pooled)
        K.bn_relu_pool(yb, scale, shift, a2, pooled2)
        torch.cuda.synchronize()
        same = bool(torch.equal(a, a2)) and bool(torch.equal(pooled, pooled2))
        # ywin: y of the first maximum of the fp32 activation in each window
        act = (yb.float() * scale + shift).permute(0, 3, 1, 2)
        _, idx = F.max_pool2d(act, 2, return_indices=True)
        yref = yb.float().permute(0, 3, 1, 2).flatten(2).gather(2, idx.flatten(2)).view_as(idx).permute(0, 2, 3, 1)
        same &= bool(torch.equal(ywin.float(), yref))
        print(f"  [{'OK ' if same else 'BAD'}] bn_relu_pool + ywin B{B} {H}x{W} C{Cc}")
        ok &= same
        # producer 1: [d_up | d_skip] at level L
        wd1 = (torch.randn(2 * Cc, 9 * Cd1, generator=g) / (3 * Cd1 ** 0.5)).to(dev).to(torch.bfloat16)
        dy1 = nhwc(torch.randn(B, Cd1, H, W, generator=g).to(dev))
        dcat = torch.full((B, H, W, 2 * Cc), float("nan"), dtype=torch.bfloat16, device=dev)
        dcat0 = torch.full_like(dcat, float("nan"))
        # producer 2: d_pool at level L+1
        wd2 = (torch.randn(Cc, 9 * Cd2, generator=g) / (3 * Cd2 ** 0.5)).to(dev).to(torch.bfloat16)
        dy2 = nhwc(torch.randn(B, Cd2, H // 2, W // 2, generator=g).to(dev))
        dpool = torch.full((B, H // 2, W // 2, Cc), float("nan"), dtype=torch.bfloat16, device=dev)
        dpool0 = torch.full_like(dpool, float("nan"))
        if not (K.conv_gemm_bnb_supported(K.A_CONV3X3, (B, H, W), dy1, wd1, dcat) and
                K.conv_gemm_bnb_supported(K.A_CONV3X3, (B, H // 2, W // 2), dy2, wd2, dpool)):
            print("   unsupported shape: skipped")
            continue
        rows0 = K.conv_gemm_stat_rows(B, H, W, 2 * Cc)
        rows1 = K.conv_gemm_stat_rows(B, H // 2, W // 2, Cc)
        st0 = torch.full((rows0, 2 * Cc, 2), float("nan"), device=dev)
        st0p = torch.zeros(rows0, 2 * Cc, 2, device=dev)
        st1 = torch.full((rows1, Cc, 2), float("nan"), device=dev)
        K.conv_gemm(K.A_CONV3X3, (B, H, W), dy1, wd1, dcat, stats=st0, bnb=(yb, scale, shift, mean, invstd, Cc))
        K.conv_gemm(K.A_CONV3X3, (B, H, W), dy1, wd1, dcat0, stats=st0p)
        K.conv_gemm(K.A_CONV3X3, (B, H // 2, W // 2), dy2, wd2, dpool, stats=st1, bnb=(ywin, scale, shift, mean, invstd))
        K.conv_gemm(K.A_CONV3X3, (B, H // 2, W // 2), dy2, wd2, dpool0)
        torch.cuda.synchronize()
        tag = f"pool bnb B{B} {H}x{W} C{Cc}"
        same = bool(torch.equal(dcat, dcat0)) and bool(torch.equal(dpool, dpool0))
        print(f"  [{'OK ' if same else 'BAD'}] {tag}: gradients identical to the plain launches")
        ok &= same
        ok &= report(f"   {tag} d_up column sums untouched", st0[:, :Cc, 0].sum(0), st0p[:, :Cc, 0].sum(0), 1e-6)
        dg1, db1, dg2, db2 = (torch.empty(Cc, device=dev) for _ in range(4))
        o1, o2 = torch.empty_like(yb), torch.empty_like(yb)
        K.bn_pool_bwd_apply(dcat[..., Cc:], dpool, yb, scale, shift, mean, invstd, (st0, rows0, 2 * Cc, Cc),
                            (st1, rows1, Cc, 0), dg1, db1, o1, ws)
        K.bn_relu_pool_bwd(dcat0[..., Cc:], dpool0, yb, scale, shift, mean, invstd, scale, dg2, db2, o2, ws)
        torch.cuda.synchronize()
        ok &= report(f"   {tag} dgamma", dg1, dg2, 1e-4)
        ok &= report(f"   {tag} dbeta", db1, db2, 1e-4)
        ok &= report(f"   {tag} dy", o1.float(), o2.float(), 4e-3)
    return ok


def g1_plain(K):
    dev = "cuda"
    ok = True
    for (B, H, W, Cin, N) in [(1, 16, 16, 64, 64), (2, 32, 32, 128, 128), (1, 16, 32, 64, 256), (3, 8, 8, 64, 64)]:
        g = torch.Generator().manual_seed(1)
        x = torch.randn(B, H, W, Cin, generator=g).to(dev).to(torch.bfloat16)
        w = (torch.randn(N, Cin, generator=g) / Cin ** 0.5).to(dev).to(torch.bfloat16)
        y = torch.full((B, H, W, N), float("nan"), dtype=torch.bfloat16, device=dev)
        K.conv_gemm(K.A_PLAIN, (B, H, W), x, w, y)
        torch.cuda.synchronize()
        ref = x.float().reshape(-1, Cin) @ w.float().t()
        ok &= report(f"plain gemm B{B} {H}x{W} K{Cin} N{N}", y.float().reshape(-1, N), ref)
    return ok


def g1_conv(K):
    ok = True
    ok &= conv3x3_case(K, 2, 32, 32, 64, 64)
    ok &= conv3x3_case(K, 1, 16, 16, 128, 128)
    ok &= conv3x3_case(K, 2, 8, 8, 256, 256)
    ok &= conv3x3_case(K, 1, 8, 256, 64, 64)
    ok &= conv3x3_case(K, 3, 4, 4, 64, 128)
    ok &= conv3x3_case(K, 2, 16, 16, 128, 64, dual=True)
    ok &= conv3x3_case(K, 2, 16, 16, 256, 512)
    return ok


def g1_eval_epilogue(K):
    """Inference epilogue: dst = relu(acc * scale + shift) (BatchNorm in eval mode folded into the conv), and the
    plain 2x2 max-pool that follows it for the encoder outputs."""
    dev = "cuda"
    ok = True
    for (B, H, W, Cin, Cout, dual) in [(2, 32, 32, 64, 64, False), (1, 16, 16, 128, 128, False), (2, 16, 16, 256, 256, True),
                                       (2, 16, 16, 256, 512, False), (3, 8, 8, 64, 128, False)]:
        g = torch.Generator().manual_seed(B + Cin + Cout)
        x = torch.randn(B, Cin, H, W, generator=g).to(dev)
        w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).to(dev)
        scale = ((torch.rand(Cout, generator=g) + 0.5) * torch.where(torch.rand(Cout, generator=g) < 0.1, -1.0, 1.0)).to(dev)
        shift = (torch.randn(Cout, generator=g) * 0.3).to(dev)
        xb = nhwc(x)
        wf = torch.empty(Cout, 9 * Cin, dtype=torch.bfloat16, device=dev)
        K.pack_conv3x3_weights(w, wf, None)
        a = torch.full((B, H, W, Cout), float("nan"), dtype=torch.bfloat16, device=dev)
        if dual:
            c0 = Cin // 2
            K.conv_gemm(K.A_CONV3X3, (B, H, W), xb[..., :c0].contiguous(), wf, a, src1=xb[..., c0:].contiguous(),
                        ep=(scale, shift))
        else:
            K.conv_gemm(K.A_CONV3X3, (B, H, W), xb, wf, a, ep=(scale, shift))
        pooled = torch.full((B, H // 2, W // 2, Cout), float("nan"), dtype=torch.bfloat16, device=dev)
        K.maxpool2x2(a, pooled)
        torch.cuda.synchronize()
        ref = F.relu(F.conv2d(nchw(xb), w.to(torch.bfloat16).float(), padding=1) * scale.view(1, -1, 1, 1) +
                     shift.view(1, -1, 1, 1))
        ok &= report(f"conv+bn(eval)+relu epilogue B{B} {H}x{W} {Cin}->{Cout}{' dual' if dual else ''}", nchw(a), ref)
        same = bool(torch.equal(nchw(pooled), F.max_pool2d(nchw(a), 2)))
        print(f"  [{'OK ' if same else 'BAD'}]    maxpool2x2 exact")
        ok &= same
    return ok


def g1_prologue(K):
    """Training prologue: conv(relu(scale*y + shift)) computed from the RAW y with the transform applied to the staged
    halo tiles must be BIT-IDENTICAL (outputs and statistics rows) to the conv of the materialised activation."""
    dev = "cuda"
    ok = True
    seen = 0
    for (B, H, W, Cin, Cout, dual) in [(4, 64, 64, 256, 256, False), (2, 32, 32, 512, 512, False),
                                       (3, 16, 16, 128, 64, False), (2, 64, 64, 64, 128, False),
                                       (1, 16, 48, 256, 256, True), (5, 16, 16, 64, 64, False)]:
        g = torch.Generator().manual_seed(B + Cin + Cout)
        y = nhwc(torch.randn(B, Cin, H, W, generator=g).to(dev))
        w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)).to(dev)
        scale = ((torch.rand(Cin, generator=g) + 0.5) * torch.where(torch.rand(Cin, generator=g) < 0.1, -1.0, 1.0)).to(dev)
        shift = (torch.randn(Cin, generator=g) * 0.3).to(dev)
        wf = torch.empty(Cout, 9 * Cin, dtype=torch.bfloat16, device=dev)
        K.pack_conv3x3_weights(w, wf, None)
        out_a = torch.full((B, H, W, Cout), float("nan"), dtype=torch.bfloat16, device=dev)
        out_p = torch.full_like(out_a, float("nan"))
        c0 = Cin // 2 if dual else Cin
        y0 = y[..., :c0].contiguous()
        s1 = y[..., c0:].contiguous() if dual else None          # second source: already an activation
        if not K.conv_gemm_pro_supported(K.A_CONV3X3, (B, H, W), y0, wf, out_a, src1=s1):
            print(f"  [skip]   prologue B{B} {H}x{W} {Cin}->{Cout}: not served by the halo kernel")
            continue
        seen += 1
        a0 = torch.empty_like(y0)
        K.bn_relu_pool(y0, scale[:c0].contiguous(), shift[:c0].contiguous(), a0)
        rows = K.conv_gemm_stat_rows(B, H, W, Cout)
        st_a = torch.zeros(rows, Cout, 2, device=dev)
        st_p = torch.zeros(rows, Cout, 2, device=dev)
        K.conv_gemm(K.A_CONV3X3, (B, H, W), a0, wf, out_a, src1=s1, stats=st_a)
        K.conv_gemm(K.A_CONV3X3, (B, H, W), y0, wf, out_p, src1=s1, stats=st_p, pro=(scale, shift, 1))
        torch.cuda.synchronize()
        same = bool(torch.equal(out_a.view(torch.int16), out_p.view(torch.int16))) and bool(torch.equal(st_a, st_p))
        print(f"  [{'OK ' if same else 'BAD'}]    prologue conv B{B} {H}x{W} {Cin}->{Cout}{' dual' if dual else ''} bit-identical"
              f" (max diff {float((out_a.float() - out_p.float()).abs().max()):.3g})")
        ok &= same
        src = torch.cat([a0, s1], dim=3) if dual else a0
        ref = F.conv2d(nchw(src), w.to(torch.bfloat16).float(), padding=1)
        ok &= report(f"prologue conv vs torch B{B} {H}x{W} {Cin}->{Cout}", nchw(out_p), ref)
    return ok and seen >= 4


def g2_wgrad_prologue(K):
    """Weight gradient with the B-operand prologue: partial slabs bit-identical to the materialised-activation launch."""
    dev = "cuda"
    ok = True
    seen = 0
    for (B, H, W, Cin, Cout) in [(4, 64, 64, 256, 256), (2, 64, 64, 128, 256), (1, 16, 64, 256, 512), (3, 8, 128, 128, 256)]:
        g = torch.Generator().manual_seed(11 + B)
        y = nhwc(torch.randn(B, Cin, H, W, generator=g).to(dev))
        dyb = nhwc(torch.randn(B, Cout, H, W, generator=g).to(dev))
        scale = ((torch.rand(Cin, generator=g) + 0.5) * torch.where(torch.rand(Cin, generator=g) < 0.1, -1.0, 1.0)).to(dev)
        shift = (torch.randn(Cin, generator=g) * 0.3).to(dev)
        if not K.wgrad_pro_supported((B, H, W), dyb, K.A_CONV3X3, y):
            print(f"  [skip]   wgrad prologue B{B} {H}x{W} {Cin}->{Cout}: not served by the CTA-pair shifted-window kernel")
            continue
        seen += 1
        a = torch.empty_like(y)
        K.bn_relu_pool(y, scale, shift, a)
        splits = K.wgrad_splits((B, H, W), dyb, K.A_CONV3X3, a)
        part_a = torch.full((splits, 9, Cout, Cin), float("nan"), device=dev)
        part_p = torch.full_like(part_a, float("nan"))
        K.wgrad_gemm((B, H, W), dyb, K.A_CONV3X3, a, part_a)
        K.wgrad_gemm((B, H, W), dyb, K.A_CONV3X3, y, part_p, b_pro=(scale, shift))
        torch.cuda.synchronize()
        same = bool(torch.equal(part_a.view(torch.int32), part_p.view(torch.int32)))
        print(f"  [{'OK ' if same else 'BAD'}]    wgrad prologue B{B} {H}x{W} {Cin}->{Cout} splits={splits} bit-identical"
              f" (max diff {float((part_a - part_p).abs().max()):.3g})")
        ok &= same
        grad = torch.empty(Cout, Cin, 3, 3, device=dev)
        K.wgrad_reduce(part_p, splits, 9, Cout, Cin, 0, grad)
        wref = torch.zeros(Cout, Cin, 3, 3, device=dev, requires_grad=True)
        F.conv2d(nchw(a), wref, padding=1).backward(nchw(dyb))
        ok &= report(f"wgrad prologue vs torch B{B} {H}x{W} {Cin}->{Cout}", grad, wref.grad, 1e-2)
    return ok and seen >= 2


def g1_big(K):
    # many tiles per CTA: exercises the persistent schedule, phase wrap-around, TMEM double buffering
    ok = conv3x3_case(K, 8, 128, 128, 64, 64)
    ok &= conv3x3_case(K, 4, 64, 64, 128, 256)
    return ok


def g1_convT(K):
    dev = "cuda"
    ok = True
    for (B, h, w_, Cin, Cout) in [(2, 8, 8, 128, 64), (1, 16, 16, 256, 128), (2, 4, 4, 512, 256)]:
        g = torch.Generator().manual_seed(7)
        x = torch.randn(B, Cin, h, w_, generator=g).to(dev)
        wt = (torch.randn(Cin, Cout, 2, 2, generator=g) / Cin ** 0.5).to(dev)
        bias = torch.randn(Cout, generator=g).to(dev)
        xb = nhwc(x)
        wf = torch.empty(4 * Cout, Cin, dtype=torch.bfloat16, device=dev)
        wd = torch.empty(Cin, 4 * Cout, dtype=torch.bfloat16, device=dev)
        b4 = torch.empty(4 * Cout, device=dev)
        K.pack_convT_weights(wt, bias, wf, wd, b4)
        # write into the first half of a 2*Cout-wide buffer (as the decoder concat would)
        buf = torch.zeros(B, 2 * h, 2 * w_, 2 * Cout, dtype=torch.bfloat16, device=dev)
        up = buf[..., :Cout]
        K.conv_gemm(K.A_PLAIN, (B, h, w_), xb, wf, up, bias=b4, d_mode=K.D_SCATTER2X2)
        torch.cuda.synchronize()
        ref = F.conv_transpose2d(nchw(xb), wt.to(torch.bfloat16).float(), bias, stride=2)
        ok &= report(f"convT fwd B{B} {h}x{w_} {Cin}->{Cout}", nchw(up), ref)
        ok &= report("   untouched half", buf[..., Cout:].float(), torch.zeros_like(buf[..., Cout:]).float(), 0)
        # dgrad: gather 2x2
        dup = torch.randn(B, Cout, 2 * h, 2 * w_, generator=g).to(dev)
        dbuf = torch.zeros(B, 2 * h, 2 * w_, 2 * Cout, dtype=torch.bfloat16, device=dev)
        dbuf[..., :Cout] = nhwc(dup)
        dx = torch.full((B, h, w_, Cin), float("nan"), dtype=torch.bfloat16, device=dev)
        K.conv_gemm(K.A_GATHER2X2, (B, h, w_), dbuf[..., :Cout], wd, dx)
        torch.cuda.synchronize()
        ref_dx = F.conv2d(dbuf[..., :Cout].float().permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), stride=2)
        ok &= report("   convT dgrad", nchw(dx), ref_dx)
    return ok


def g2_wgrad(K):
    dev = "cuda"
    ok = True
    cases = [(2, 16, 16, 64, 64, False), (1, 32, 32, 128, 128, False), (2, 8, 8, 256, 128, False),
             (2, 16, 16, 128, 64, True), (4, 64, 64, 64, 64, False), (2, 4, 4, 512, 256, False),
             (2, 8, 64, 128, 64, True), (1, 4, 256, 64, 64, False), (3, 32, 32, 256, 512, False)]
    for (B, H, W, Cin, Cout, dual) in cases:
        g = torch.Generator().manual_seed(3)
        x = torch.randn(B, Cin, H, W, generator=g).to(dev)
        dy = torch.randn(B, Cout, H, W, generator=g).to(dev)
        xb, dyb = nhwc(x), nhwc(dy)
        if dual:
            b0 = xb[..., :Cin // 2].contiguous()
            b1 = xb[..., Cin // 2:].contiguous()
        else:
            b0, b1 = xb, None
        splits = K.wgrad_splits((B, H, W), dyb, K.A_CONV3X3, b0, b1)
        part = torch.full((splits, 9, Cout, Cin), float("nan"), device=dev)
        K.wgrad_gemm((B, H, W), dyb, K.A_CONV3X3, b0, part, b1)
        grad = torch.empty(Cout, Cin, 3, 3, device=dev)
        K.wgrad_reduce(part, splits, 9, Cout, Cin, 0, grad)
        torch.cuda.synchronize()
        xr = nchw(xb).requires_grad_(False)
        wref = torch.zeros(Cout, Cin, 3, 3, device=dev, requires_grad=True)
        F.conv2d(xr, wref, padding=1).backward(nchw(dyb))
        ok &= report(f"wgrad conv3x3 B{B} {H}x{W} {Cin}->{Cout}{' dual' if dual else ''} splits={splits}", grad,
                     wref.grad, 1e-2)
    # convT wgrad (gather): dW[ci][co][a][b]
    for (B, h, w_, Cin, Cout) in [(2, 8, 8, 128, 64), (1, 16, 16, 256, 128)]:
        g = torch.Generator().manual_seed(5)
        x = torch.randn(B, Cin, h, w_, generator=g).to(dev)
        dup = torch.randn(B, Cout, 2 * h, 2 * w_, generator=g).to(dev)
        xb = nhwc(x)
        dbuf = torch.zeros(B, 2 * h, 2 * w_, 2 * Cout, dtype=torch.bfloat16, device=dev)
        dbuf[..., :Cout] = nhwc(dup)
        dupv = dbuf[..., :Cout]
        splits = K.wgrad_splits((B, h, w_), xb, K.A_GATHER2X2, dupv)
        part = torch.full((splits, 4, Cin, Cout), float("nan"), device=dev)
        K.wgrad_gemm((B, h, w_), xb, K.A_GATHER2X2, dupv, part)
        grad = torch.empty(Cin, Cout, 2, 2, device=dev)
        K.wgrad_reduce(part, splits, 4, Cin, Cout, 1, grad)
        torch.cuda.synchronize()
        wref = torch.zeros(Cin, Cout, 2, 2, device=dev, requires_grad=True)
        F.conv_transpose2d(nchw(xb), wref, stride=2).backward(dupv.float().permute(0, 3, 1, 2))
        ok &= report(f"wgrad convT B{B} {h}x{w_} {Cin}->{Cout} splits={splits}", grad, wref.grad, 1e-2)
    # plain (first layer, im2col'ed input)
    B, H, W, Cin, Cout = 2, 16, 16, 3, 64
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, Cin, H, W, generator=g).to(dev)
    dy = torch.randn(B, Cout, H, W, generator=g).to(dev)
    dyb = nhwc(dy)
    col = torch.empty(B, H, W, 64, dtype=torch.bfloat16, device=dev)
    K.pack_input_im2col(x, col)
    splits = K.wgrad_splits((B, H, W), dyb, K.A_PLAIN, col)
    part = torch.full((splits, 1, Cout, 64), float("nan"), device=dev)
    K.wgrad_gemm((B, H, W), dyb, K.A_PLAIN, col, part)
    grad = torch.empty(Cout, Cin, 3, 3, device=dev)
    K.wgrad_reduce(part, splits, 1, Cout, 64, 2, grad, real_cin=Cin)
    torch.cuda.synchronize()
    wref = torch.zeros(Cout, Cin, 3, 3, device=dev, requires_grad=True)
    F.conv2d(x.to(torch.bfloat16).float(), wref, padding=1).backward(nchw(dyb))
    ok &= report("wgrad first layer (im2col)", grad, wref.grad, 1e-2)
    # paired-pixel first layer: im2col32 + PLAIN GEMM over pixel pairs with block-diagonal weights
    col32 = torch.empty(B, H, W, 32, dtype=torch.bfloat16, device=dev)
    K.pack_input_im2col32(x, col32)
    same = bool(torch.equal(col32[..., :27], col[..., :27])) and bool((col32[..., 27:] == 0).all())
    print(f"  [{'OK ' if same else 'BAD'}] im2col32 == first 27 channels of im2col, zero padded")
    ok &= same
    w1q = (torch.randn(Cout, Cin, 3, 3, generator=torch.Generator().manual_seed(10)) / 5).to(dev)
    w1pp = torch.empty(128, 64, dtype=torch.bfloat16, device=dev)
    K.pack_conv1_pair_weights(w1q, w1pp)
    ypair = torch.full((B, H, W, Cout), float("nan"), dtype=torch.bfloat16, device=dev)
    rows = K.conv_gemm_stat_rows(B, H, W // 2, 128, K.A_PLAIN)
    stp = torch.zeros(rows, 128, 2, device=dev)
    K.conv_gemm(K.A_PLAIN, (B, H, W // 2), col32.view(B, H, W // 2, 64), w1pp, ypair.view(B, H, W // 2, 128), stats=stp)
    torch.cuda.synchronize()
    refp = F.conv2d(x.to(torch.bfloat16).float(), w1q.to(torch.bfloat16).float(), padding=1)
    ok &= report("first layer fwd (paired pixels)", nchw(ypair), refp)
    yf = ypair.float().reshape(-1, Cout)
    st2 = stp.view(2 * rows, 64, 2)
    ok &= report("   stats sum (rows x2 view)", st2[..., 0].sum(0), yf.sum(0), 1e-3)
    ok &= report("   stats sumsq", st2[..., 1].sum(0), (yf * yf).sum(0), 1e-3)
    dy2 = dyb.view(B, H, W // 2, 128)
    sp = K.wgrad_splits((B, H, W // 2), dy2, K.A_PLAIN, col32.view(B, H, W // 2, 64))
    partp = torch.empty(sp, 1, 128, 64, device=dev)
    K.wgrad_gemm((B, H, W // 2), dy2, K.A_PLAIN, col32.view(B, H, W // 2, 64), partp)
    gradp = torch.empty(Cout, Cin, 3, 3, device=dev)
    K.wgrad_reduce(partp, sp, 1, 128, 64, 3, gradp, real_cin=Cin)
    torch.cuda.synchronize()
    ok &= report("wgrad first layer (paired pixels)", gradp, wref.grad, 1e-2)
    # first-layer forward through the im2col + plain GEMM
    w1 = (torch.randn(Cout, Cin, 3, 3, generator=g) / 5).to(dev)
    w1p = torch.empty(Cout, 64, dtype=torch.bfloat16, device=dev)
    K.pack_conv1_weights(w1, w1p)
    y = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device=dev)
    K.conv_gemm(K.A_PLAIN, (B, H, W), col, w1p, y)
    torch.cuda.synchronize()
    ref = F.conv2d(x.to(torch.bfloat16).float(), w1.to(torch.bfloat16).float(), padding=1)
    ok &= report("first layer fwd (im2col)", nchw(y), ref)
    return ok


def ew_bn(K):
    dev = "cuda"
    ok = True
    ws = K.new_workspace(dev)
    for (B, H, W, Cc, pool) in [(2, 16, 16, 64, True), (2, 8, 8, 128, False), (1, 32, 32, 256, True),
                                (3, 4, 4, 512, False)]:
        g = torch.Generator().manual_seed(11)
        y = torch.randn(B, Cc, H, W, generator=g).to(dev) * 1.5 + 0.3
        yb = nhwc(y)
        gamma = (torch.rand(Cc, generator=g) + 0.5).to(dev)
        beta = (torch.randn(Cc, generator=g) * 0.2).to(dev)
        cbias = (torch.randn(Cc, generator=g) * 0.1).to(dev)
        # statistics as G1 would produce them
        yf = yb.float().reshape(-1, Cc)
        st = torch.stack([yf.sum(0), (yf * yf).sum(0)], dim=-1).reshape(1, Cc, 2).contiguous()
        rm = torch.zeros(Cc, device=dev)
        rv = torch.ones(Cc, device=dev)
        nbt = torch.zeros((), dtype=torch.int64, device=dev)
        scale, shift, mean, invstd = (torch.empty(Cc, device=dev) for _ in range(4))
        count = B * H * W
        K.bn_finalize(st, 1, Cc, count, gamma, beta, cbias, rm, rv, nbt, 0.1, 1e-5, scale, shift, mean, invstd)
        a = torch.empty(B, H, W, Cc, dtype=torch.bfloat16, device=dev)
        pooled = torch.empty(B, H // 2, W // 2, Cc, dtype=torch.bfloat16, device=dev) if pool else None
        K.bn_relu_pool(yb, scale, shift, a, pooled)
        torch.cuda.synchronize()
        # torch reference on the bf16-rounded conv output (+ bias, which BN must cancel)
        yin = (nchw(yb) + cbias.view(1, -1, 1, 1)).requires_grad_(True)
        bn = torch.nn.BatchNorm2d(Cc).to(dev)
        with torch.no_grad():
            bn.weight.copy_(gamma)
            bn.bias.copy_(beta)
        bn.train()
        aref = F.relu(bn(yin))
        ok &= report(f"bn+relu fwd B{B} {H}x{W} C{Cc}", nchw(a), aref, 1e-2)
        ok &= report("   running_mean", rm, bn.running_mean, 1e-4)
        ok &= report("   running_var", rv, bn.running_var, 1e-4)
        ok &= (int(nbt.item()) == 1)
        dA = torch.randn(B, Cc, H, W, generator=g).to(dev)
        dAb = nhwc(dA)
        if pool:
            pref = F.max_pool2d(aref, 2)
            ok &= report("   pooled", nchw(pooled), pref, 1e-2)
            dP = torch.randn(B, Cc, H // 2, W // 2, generator=g).to(dev)
            dPb = nhwc(dP)
            # plain fp32 reference: the pooled gradient goes to the first maximum of the fp32 activations
            # (model.py:31 semantics); the kernel takes the argmax on fmaf(y, scale, shift) in fp32 too
            total = (aref * nchw(dAb)).sum() + (F.max_pool2d(aref, 2) * nchw(dPb)).sum()
            total.backward()
        else:
            dPb = None
            (aref * nchw(dAb)).sum().backward()
        dgamma, dbeta = torch.empty(Cc, device=dev), torch.empty(Cc, device=dev)
        dy = torch.empty(B, H, W, Cc, dtype=torch.bfloat16, device=dev)
        K.bn_relu_pool_bwd(dAb, dPb, yb, scale, shift, mean, invstd, gamma, dgamma, dbeta, dy, ws)
        torch.cuda.synchronize()
        ok &= report("   bwd dy", nchw(dy), yin.grad, 2e-2)
        ok &= report("   bwd dgamma", dgamma, bn.weight.grad, 1e-2)
        ok &= report("   bwd dbeta", dbeta, bn.bias.grad, 1e-2)
    return ok


def ew_heads_loss(K):
    dev = "cuda"
    ok = True
    ws = K.new_workspace(dev)
    B, H, W = 2, 32, 32
    P = B * H * W
    g = torch.Generator().manual_seed(13)
    a = torch.randn(B, H, W, 64, generator=g).to(dev).to(torch.bfloat16)
    ws_ = [(torch.randn(1, 64, 1, 1, generator=g) * 0.2).to(dev) for _ in range(3)]
    bs_ = [(torch.randn(1, generator=g) * 0.2).to(dev) for _ in range(3)]
    logits = torch.empty(3, P, device=dev)
    K.heads_fwd(a, ws_, bs_, logits)
    torch.cuda.synchronize()
    af = a.float().reshape(P, 64)
    ref = torch.stack([af @ w.reshape(64) + b for w, b in zip(ws_, bs_)])
    ok &= report("heads fwd", logits, ref, 1e-5)
    # fused BN+ReLU+heads must reproduce (bn_relu -> heads_fwd) exactly
    yraw = torch.randn(B, H, W, 64, generator=g).to(dev).to(torch.bfloat16)
    sc_ = (torch.rand(64, generator=g) + 0.5).to(dev)
    sh_ = (torch.randn(64, generator=g) * 0.3).to(dev)
    a1 = torch.empty_like(yraw)
    a2 = torch.empty_like(yraw)
    lg1, lg2 = torch.empty(3, P, device=dev), torch.empty(3, P, device=dev)
    K.bn_relu_pool(yraw, sc_, sh_, a1, None)
    K.heads_fwd(a1, ws_, bs_, lg1)
    K.bn_relu_heads(yraw, sc_, sh_, a2, ws_, bs_, lg2)
    torch.cuda.synchronize()
    same = bool(torch.equal(a1, a2)) and bool(torch.equal(lg1, lg2))
    print(f"  [{'OK ' if same else 'BAD'}] fused bn_relu_heads == bn_relu + heads_fwd (bitwise)")
    ok &= same
    # losses vs autograd of the reference formula (stable form)
    tgt = (torch.rand(P, generator=g) < 0.4).float().to(dev)
    lg = ref.clone().requires_grad_(True)
    out, sel, aux = lg[0], lg[1], lg[2]
    s = torch.sigmoid(sel)
    cov = s.mean()
    risk = (F.binary_cross_entropy_with_logits(out, tgt, reduction="none") * s).mean() / cov
    pen = torch.clamp(0.8 - cov, min=0) ** 2
    lamb = 2.0
    sl = risk + lamb * pen
    al = F.binary_cross_entropy_with_logits(aux, tgt)
    (sl + al).backward()
    sums = torch.zeros(3, dtype=torch.float64, device=dev)
    K.loss_sums(logits[0], logits[1], logits[2], tgt, sums, ws)
    res = torch.empty(4, device=dev)
    K.loss_finalize(sums, P, lamb, 0.8, res)
    dl = torch.empty(3, P, device=dev)
    K.loss_bwd(logits[0], logits[1], logits[2], tgt, sums, P, lamb, 0.8, None, None, dl[0], dl[1], dl[2])
    torch.cuda.synchronize()
    ok &= report("loss values", res, torch.stack([sl, cov, al, sl + al]).detach(), 1e-5)
    ok &= report("loss grads", dl, lg.grad, 1e-4)
    # heads bwd
    dA = torch.empty(B, H, W, 64, dtype=torch.bfloat16, device=dev)
    dws = [torch.empty_like(w) for w in ws_]
    dbs = [torch.empty_like(b) for b in bs_]
    K.heads_bwd(dl, a, ws_, dA, dws, dbs, ws)
    torch.cuda.synchronize()
    dref = sum(dl[h].reshape(P, 1) * ws_[h].reshape(1, 64) for h in range(3))
    ok &= report("heads bwd dA", dA.float().reshape(P, 64), dref, 1e-2)
    for h in range(3):
        ok &= report(f"heads bwd dw{h}", dws[h].reshape(64), dl[h] @ af, 1e-4)
        ok &= report(f"heads bwd db{h}", dbs[h], dl[h].sum().reshape(1), 1e-4)
    # heads bwd on y with the BN-backward reduction fused (activation never stored): identical head gradients
    # and dA as heads_bwd on a1 = relu(bn(yraw)); rows + bn_bwd_apply == bn_relu_pool_bwd on that dA
    mean_ = (torch.randn(64, generator=g) * 0.3).to(dev)
    istd_ = (torch.rand(64, generator=g) + 0.5).to(dev)
    dA1, dA2 = torch.empty_like(dA), torch.empty_like(dA)
    dws1, dbs1 = [torch.empty_like(w) for w in ws_], [torch.empty_like(b) for b in bs_]
    dws2, dbs2 = [torch.empty_like(w) for w in ws_], [torch.empty_like(b) for b in bs_]
    K.heads_bwd(dl, a1, ws_, dA1, dws1, dbs1, ws)
    lgn = torch.empty(3, P, device=dev)
    K.bn_relu_heads(yraw, sc_, sh_, None, ws_, bs_, lgn)          # a not stored
    rows = K.heads_bwd_bn_rows(P)
    part = torch.full((rows, 64, 2), float("nan"), device=dev)
    K.heads_bwd_bn(dl, yraw, sc_, sh_, mean_, istd_, ws_, dA2, dws2, dbs2, part, ws)
    torch.cuda.synchronize()
    same = bool(torch.equal(dA1, dA2)) and bool(torch.equal(lgn, lg2))
    print(f"  [{'OK ' if same else 'BAD'}] heads_bwd_bn: dA and logits identical to the stored-activation path")
    ok &= same
    for h in range(3):
        ok &= report(f"heads_bwd_bn dw{h}", dws2[h], dws1[h], 1e-5)
        ok &= report(f"heads_bwd_bn db{h}", dbs2[h], dbs1[h], 1e-5)
    dg1, db1, dg2, db2 = (torch.empty(64, device=dev) for _ in range(4))
    dy1, dy2 = torch.empty_like(yraw), torch.empty_like(yraw)
    K.bn_bwd_apply(dA2, yraw, sc_, sh_, mean_, istd_, part, rows, dg1, db1, dy1, ws)
    K.bn_relu_pool_bwd(dA1, None, yraw, sc_, sh_, mean_, istd_, sc_, dg2, db2, dy2, ws)
    torch.cuda.synchronize()
    ok &= report("heads_bwd_bn -> dgamma", dg1, dg2, 1e-4)
    ok &= report("heads_bwd_bn -> dbeta", db1, db2, 1e-4)
    ok &= report("heads_bwd_bn -> dy", dy1.float(), dy2.float(), 4e-3)
    # metric
    counts = torch.zeros(6, dtype=torch.int64, device=dev)
    K.metric_hist(logits[0], logits[1], tgt, 0.0, 0.0, True, counts)
    torch.cuda.synchronize()
    pred = (logits[0] >= 0).long()
    selm = logits[1] >= 0
    lab = tgt.long()
    cm = torch.bincount((lab * 2 + pred)[selm], minlength=4)
    exp = torch.cat([cm, selm.sum().reshape(1), torch.tensor([P], device=dev)])
    good = bool((counts == exp).all().item())
    print(f"  [{'OK ' if good else 'BAD'}] metric hist {counts.tolist()} vs {exp.tolist()}")
    ok &= good
    # adam
    import ctypes as C
    from selectivenet_for_semantic_segmentation_binary_b200 import _lib
    ps = [torch.randn(n, generator=g).to(dev) for n in (1000, 7, 64 * 64 * 9)]
    gs = [torch.randn(p.shape, generator=g).to(dev) for p in ps]
    refp = [p.clone().requires_grad_(True) for p in ps]
    opt = torch.optim.Adam(refp, lr=1e-3)
    ms = [torch.zeros_like(p) for p in ps]
    vs = [torch.zeros_like(p) for p in ps]
    tab = (_lib.AdamTensor * len(ps))()
    for i, p in enumerate(ps):
        tab[i].param, tab[i].grad, tab[i].exp_avg, tab[i].exp_avg_sq, tab[i].numel = (
            p.data_ptr(), gs[i].data_ptr(), ms[i].data_ptr(), vs[i].data_ptr(), p.numel())
    tab_dev = torch.frombuffer(bytearray(bytes(tab)), dtype=torch.uint8).to(dev)
    for step in (1, 2, 3):
        for rp, gg in zip(refp, gs):
            rp.grad = gg.clone()
        opt.step()
        K.adam_step(tab_dev, len(ps), max(p.numel() for p in ps), 1e-3, 0.9, 0.999, 1e-8, 0.0, step)
    torch.cuda.synchronize()
    for i in range(len(ps)):
        ok &= report(f"adam tensor {i}", ps[i], refp[i].detach(), 1e-5)
    return ok


def swizzle_exp(K):
    """How does tcgen05.mma derive the 128B-swizzle phase of an operand whose descriptor start address is
    NOT 1024-byte aligned (start shifted by whole 128-byte rows)?  From absolute smem address bits, or
    relative to the start (+ base_offset field)?  Decides whether 3x3 taps can be shifted windows of one
    halo tile in shared memory.  Prints errors for every (shift, base_offset) pair; no pass/fail."""
    dev = "cuda"
    B, H, W, Cin, N = 1, 8, 16, 64, 64          # exactly one 128-row tile, one K step
    g = torch.Generator().manual_seed(21)
    x = torch.randn(B, H, W, Cin, generator=g).to(dev).to(torch.bfloat16)
    w = (torch.randn(N, Cin, generator=g) / 8).to(dev).to(torch.bfloat16)
    full = x.float().reshape(-1, Cin) @ w.float().t()
    print("  G1 (K-major A operand): out[m] should equal A[m+shift] . W")
    for shift in (1, 2, 3, 8, 9):
        for boff in sorted({0, shift & 7}):
            os.environ["SUNET_DBG_SHIFT"], os.environ["SUNET_DBG_BOFF"] = str(shift), str(boff)
            y = torch.zeros(B, H, W, N, dtype=torch.bfloat16, device=dev)
            K.conv_gemm(K.A_PLAIN, (B, H, W), x, w, y)
            torch.cuda.synchronize()
            got = y.float().reshape(-1, N)[:128 - shift]
            e = rel_err(got, full[shift:])
            print(f"    shift {shift} base_offset {boff}: max-rel-err {e:.3e} {'<== MATCH' if e < 2e-2 else ''}", flush=True)
    # G2: MN-major B operand, shift along K (pixels).  dy is zero where the shifted window runs off the box.
    B, H, W, Ca, Cb = 2, 4, 32, 64, 64
    dy = torch.randn(B, H, W, Ca, generator=g).to(dev).to(torch.bfloat16)
    xin = torch.randn(B, H, W, Cb, generator=g).to(dev).to(torch.bfloat16)
    print("  G2 (MN-major B operand): D[m][n] should equal sum_p dy[p][m] * x[p+shift][n]")
    for shift in (1, 2, 3, 8, 9):
        dyz = dy.clone()
        dyz[:, :, W - shift:, :] = 0
        ref = torch.einsum("bhwm,bhwn->mn", dyz[:, :, :W - shift].float(), xin[:, :, shift:].float())
        for boff in sorted({0, shift & 7}):
            os.environ["SUNET_DBG_SHIFT"], os.environ["SUNET_DBG_BOFF"] = str(shift), str(boff)
            splits = K.wgrad_splits((B, H, W), dyz, K.A_PLAIN, xin)
            part = torch.zeros(splits, 1, Ca, Cb, device=dev)
            K.wgrad_gemm((B, H, W), dyz, K.A_PLAIN, xin, part)
            torch.cuda.synchronize()
            e = rel_err(part.sum(0)[0], ref)
            print(f"    shift {shift} base_offset {boff}: max-rel-err {e:.3e} {'<== MATCH' if e < 1e-2 else ''}", flush=True)
    os.environ["SUNET_DBG_SHIFT"], os.environ["SUNET_DBG_BOFF"] = "0", "0"
    return True


GROUPS = {"swizzle_exp": swizzle_exp, "g1_plain": g1_plain, "g1_conv": g1_conv, "g1_convT": g1_convT, "g1_big": g1_big, "g1_eval_epilogue": g1_eval_epilogue, "g1_bnb": g1_bnb, "g1_bnb_convT": g1_bnb_convT, "g1_bnb_pool": g1_bnb_pool, "g2_wgrad": g2_wgrad,
          "g1_prologue": g1_prologue, "g2_wgrad_prologue": g2_wgrad_prologue, "ew_bn": ew_bn, "ew_heads_loss": ew_heads_loss}


def main():
    if len(sys.argv) > 1:
        from selectivenet_for_semantic_segmentation_binary_b200 import kernels as K
        name = sys.argv[1]
        print(f"== {name}", flush=True)
        ok = GROUPS[name](K)
        print(f"== {name}: {'PASS' if ok else 'FAIL'}", flush=True)
        sys.exit(0 if ok else 1)
    results = {}
    for name in GROUPS:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name], timeout=240)
            results[name] = "PASS" if r.returncode == 0 else f"FAIL(rc={r.returncode})"
        except subprocess.TimeoutExpired:
            results[name] = "TIMEOUT"
        print(f"-- {name}: {results[name]} ({time.time() - t0:.1f}s)", flush=True)
    print("SUMMARY", results)


if __name__ == "__main__":
    main()
