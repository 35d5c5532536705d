"""Timing-only experiment: is the MN-major operand fetch of tcgen05.mma slower than K-major?
Runs the wgrad kernels on a bench-sized layer with the true MN-major descriptors and with the
instruction descriptor forced to K-major (garbage results, same instruction count / bytes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from selectivenet_for_semantic_segmentation_binary_b200 import kernels as K

def run(B, H, W, Ca, Cb, label):
    dev = "cuda"
    dy = torch.randn(B, H, W, Ca, device=dev).to(torch.bfloat16)
    x = torch.randn(B, H, W, Cb, device=dev).to(torch.bfloat16)
    splits = K.wgrad_splits((B, H, W), dy, K.A_CONV3X3, x)
    part = torch.empty(splits, 9, Ca, Cb, device=dev)
    for boff in ("0", "16", "32"):
        os.environ["SUNET_DBG_BOFF"] = boff
        for _ in range(3):
            K.wgrad_gemm((B, H, W), dy, K.A_CONV3X3, x, part)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            K.wgrad_gemm((B, H, W), dy, K.A_CONV3X3, x, part)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * B * H * W * 9 * Ca * Cb
        print(f"{label} { {'0': 'real', '16': 'no TMA loads after fill (MMA+smem only)', '32': 'no MMAs (TMA delivery only)'}[boff] }: {ms:.3f} ms {fl / ms / 1e9:.0f} TF/s splits={splits}", flush=True)
    os.environ["SUNET_DBG_BOFF"] = "0"

run(128, 128, 128, 128, 128, "128->128 @128^2")
run(128, 64, 64, 256, 256, "256->256 @64^2")
os.environ["SUNET_WGRAD_NO_STACK"] = "1"
run(128, 256, 256, 64, 64, "64->64 @256^2 (generic kernel, M half empty)")
for kp in ("32", "128"):
    os.environ["SUNET_WGRAD_KP"] = kp
    run(128, 128, 128, 128, 128, f"128->128 @128^2 kp={kp}")
