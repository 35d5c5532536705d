"""Sustained tcgen05.mma (M=128, N, K=16, bf16) rate from shared memory for K-major / MN-major operands.
Development probe; prints cycles per MMA and the fraction of the ideal N/2 cycles."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from selectivenet_for_semantic_segmentation_binary_b200 import _lib

lib = _lib.load()
# the probe kernels are a development tool and live outside the product library:
#   make probes   ->  scripts/probes/libsunet_probe.so
raw = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "probes", "libsunet_probe.so"))
raw.sunet_dbg_mma_probe.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_void_p]
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 148
iters = 4000
out = torch.zeros(grid, dtype=torch.int64, device="cuda")
for n in (64, 128, 192, 256):
    for a_mn, b_mn in ((0, 0), (1, 0), (0, 1), (1, 1)):
        for _ in range(2):
            rc = raw.sunet_dbg_mma_probe(n, a_mn, b_mn, iters, grid, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
            assert rc == 0, lib.sunet_last_error()
            torch.cuda.synchronize()
        cyc = out.double().mean().item() / (4 * iters)
        print(f"N={n:3d} A={'MN' if a_mn else 'K '} B={'MN' if b_mn else 'K '}: {cyc:7.1f} cycles/MMA  (ideal {n / 2:5.1f}) -> "
              f"{100 * (n / 2) / cyc:5.1f}% of tensor peak", flush=True)

raw.sunet_dbg_mma_probe_pair.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
pairs = grid // 2
print("CTA pairs (cta_group::2, M=256): ideal = n/2 cycles per MMA if both SMs run at full rate")
for n in (64, 128, 256):
    for _ in range(2):
        rc = raw.sunet_dbg_mma_probe_pair(n, iters, pairs, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.sunet_last_error()
        torch.cuda.synchronize()
    cyc = out[:pairs].double().mean().item() / (4 * iters)
    print(f"N={n:3d} pair: {cyc:7.1f} cycles/MMA (256 x {n} x 16)  (ideal {n / 2:5.1f}) -> {100 * (n / 2) / cyc:5.1f}% of tensor peak", flush=True)
