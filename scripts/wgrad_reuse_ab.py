"""A/B of the A-collector reuse in the weight-gradient GEMMs (SUNET_WGRAD_A_REUSE is read per launch).
    python scripts/wgrad_reuse_ab.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from selectivenet_for_semantic_segmentation_binary_b200 import kernels as K


def run(B, H, W, Ca, Cb, label, mode=None, dual=False):
    dev = "cuda"
    mode = K.A_CONV3X3 if mode is None else mode
    if mode == K.A_GATHER2X2:
        a = torch.randn(B, H, W, Ca, device=dev).to(torch.bfloat16)              # ConvT input
        b0 = torch.randn(B, 2 * H, 2 * W, Cb, device=dev).to(torch.bfloat16)     # d_up
        taps = 4
    else:
        a = torch.randn(B, H, W, Ca, device=dev).to(torch.bfloat16)
        b0 = torch.randn(B, H, W, Cb, device=dev).to(torch.bfloat16)
        taps = 9
    b1 = torch.randn_like(b0) if dual else None
    nb = Cb * (2 if dual else 1)
    splits = K.wgrad_splits((B, H, W), a, mode, b0, b1)
    part = torch.empty(splits * taps * Ca * nb, device=dev)
    res = {}
    for reuse in ("0", "1"):
        os.environ["SUNET_WGRAD_A_REUSE"] = reuse
        for _ in range(3):
            K.wgrad_gemm((B, H, W), a, mode, b0, part, b1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            K.wgrad_gemm((B, H, W), a, mode, b0, part, b1)
        e1.record()
        torch.cuda.synchronize()
        res[reuse] = (e0.elapsed_time(e1) / 10, part.clone())
    fl = 2.0 * B * H * W * taps * Ca * nb
    same = torch.equal(res["0"][1], res["1"][1])
    print(f"{label}: plain {res['0'][0]:.3f} ms ({fl / res['0'][0] / 1e9:.0f} TF/s)   A reuse {res['1'][0]:.3f} ms "
          f"({fl / res['1'][0] / 1e9:.0f} TF/s)   splits={splits}   identical results: {same}", flush=True)


run(128, 128, 128, 128, 64, "enc2_1  128<-64   @128^2 (single CTA)")
run(128, 128, 128, 128, 128, "enc2_2  128<-128  @128^2 (single CTA)")
run(128, 128, 128, 128, 128, "dec2_2  128<-256  @128^2 (single CTA, two sources)", dual=True)
run(128, 64, 64, 256, 256, "enc3_2  256<-256  @64^2  (CTA pair)")
run(128, 64, 64, 256, 256, "dec3_2  256<-512  @64^2  (CTA pair, two sources)", dual=True)
run(128, 32, 32, 512, 512, "dec4_1  512<-512  @32^2  (CTA pair)")
run(128, 32, 32, 512, 256, "unpool3 ConvT 512->256 @32^2 (CTA pair, 4 taps)", mode=K.A_GATHER2X2)
run(128, 128, 128, 128, 64, "unpool1 ConvT 128->64 @128^2 (single CTA, 4 taps)", mode=K.A_GATHER2X2)
