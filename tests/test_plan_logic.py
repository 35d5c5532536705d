"""Planning entry points of the C ABI that run without a GPU (no kernel is launched): the split-K plan of the
weight-gradient GEMMs and the availability of the training-prologue variants, at the layer shapes of the headline
step (batch 128, 256 x 256) and of one 8-GPU shard (batch 16)."""
import ctypes as C

import pytest

from selectivenet_for_semantic_segmentation_binary_b200 import _lib

SMS = 148        # csrc falls back to the B200's SM count when no device is present
A_CONV3X3, A_PLAIN, A_GATHER2X2, D_NHWC = 0, 1, 2, 0
_CH = {1: 64, 2: 128, 3: 256, 4: 512}
# (level, dY channels, input channels of source 0, of source 1) of the thirteen 3x3 conv layers after the first
LAYERS = [(1, 64, 64, 0), (2, 128, 64, 0), (2, 128, 128, 0), (3, 256, 128, 0), (3, 256, 256, 0), (4, 512, 256, 0),
          (4, 512, 512, 0), (3, 256, 256, 256), (3, 256, 256, 0), (2, 128, 128, 128), (2, 128, 128, 0), (1, 64, 64, 64),
          (1, 64, 64, 0)]


def _wgrad_args(batch, size, level, ca, c0, c1, mode=A_CONV3X3):
    w = _lib.WgradGemmArgs()
    w.batch, w.height, w.width = batch, size >> (level - 1), size >> (level - 1)
    w.a, w.a_channels, w.a_pix_stride = 1, ca, ca          # pointers only need to be non-null here
    w.b_mode = mode
    w.b0, w.b0_channels, w.b0_pix_stride = 1, c0, c0
    if c1:
        w.b1, w.b1_channels, w.b1_pix_stride = 1, c1, c1
    return w


@pytest.mark.parametrize("batch", [128, 16])
def test_split_k_plan_is_one_wave(batch):
    """One CTA per SM at a time: splits x (CTAs per split) must not exceed the SM count, and must fill most of it."""
    lib = _lib.load()
    for level, ca, c0, c1 in LAYERS:
        w = _wgrad_args(batch, 256, level, ca, c0, c1)
        splits = lib.sunet_wgrad_gemm_splits(C.byref(w))
        assert splits >= 1, (level, ca, c0, c1, lib.sunet_last_error())
        cin = c0 + c1
        if ca == 64:                       # 64-channel dY-block kernel: one CTA per 64 input channels, all nine taps
            per_split = cin // 64
        elif ca % 256 == 0 and cin % 128 == 0:      # CTA pairs: 256 dY channels x 128 input channels x one filter row
            per_split = (ca // 256) * (cin // 128) * 3 * 2
        else:                              # single CTA: 128 dY channels x (128 | 64) input channels x one filter row
            bnw = 128 if (c0 % 128 == 0 and c1 % 128 == 0) else 64
            per_split = ((ca + 127) // 128) * (cin // bnw) * 3
        ctas = splits * per_split
        assert ctas <= SMS, (level, ca, c0, c1, splits, per_split)
        assert ctas > SMS * 0.8, (level, ca, c0, c1, splits, per_split)


def test_prologue_variants_are_available_where_expected():
    lib = _lib.load()
    got = {}
    for level, ca, c0, c1 in LAYERS:
        if c1:
            continue
        w = _wgrad_args(128, 256, level, ca, c0, 0)
        a = _lib.ConvGemmArgs()
        a.batch, a.height, a.width = w.batch, w.height, w.width
        a.a_mode, a.d_mode = A_CONV3X3, D_NHWC
        a.src0, a.src0_channels, a.src0_pix_stride = 1, c0, c0
        a.weights, a.n_total, a.k_total = 1, ca, 9 * c0
        a.dst, a.dst_pix_stride = 1, ca
        got[(level, ca, c0)] = (bool(lib.sunet_conv_gemm_pro_supported(C.byref(a))),
                                bool(lib.sunet_wgrad_gemm_pro_supported(C.byref(w))))
    # the forward conv has the variant wherever the CTA-pair halo kernel runs (every level of a 256^2 patch) ...
    assert all(conv for conv, _ in got.values()), got
    # ... the weight gradient only on the CTA-pair shifted-window kernel: >= 256 dY channels and a 64-pixel row block
    assert {k for k, (_, wg) in got.items() if wg} == {(3, 256, 128), (3, 256, 256)}, got
    # two-source (concat) weight gradients never
    w = _wgrad_args(128, 256, 3, 256, 256, 256)
    assert not lib.sunet_wgrad_gemm_pro_supported(C.byref(w))
    # and the transposed-conv form never
    w = _wgrad_args(128, 256, 4, 512, 256, 0, mode=A_GATHER2X2)
    assert not lib.sunet_wgrad_gemm_pro_supported(C.byref(w))
