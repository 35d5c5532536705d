"""Host-side logic that needs no GPU: module layout, threshold bisection, shard bounds, gradient groups."""
import numpy as np
import pytest
import torch

from oracle import sunet_oracle as O


def test_module_matches_reference_layout(golden):
    from selectivenet_for_semantic_segmentation_binary_b200.model import UNet_B, CBR_2D
    torch.manual_seed(0)
    net = UNet_B("RGB", selective=True)
    sd = O.init_state_dict(0, "RGB", True)
    assert list(net.state_dict().keys()) == list(sd.keys())
    for k, v in net.state_dict().items():
        assert torch.equal(v, sd[k]) and v.dtype == sd[k].dtype, k
    assert [n for n, _ in net.named_parameters()] == [str(s) for s in golden["param_names"]]
    assert sum(p.numel() for p in net.parameters()) == 7703107
    assert len(UNet_B("RGB", selective=False).state_dict()) == 106
    assert UNet_B("GH").encoder_layer_1_1[0].weight.shape == (64, 2, 3, 3)       # model.py:24-27
    assert UNet_B("H_RGB").encoder_layer_1_1[0].weight.shape == (64, 3, 3, 3)
    blk = CBR_2D(8, 16)
    assert [type(m).__name__ for m in blk] == ["Conv2d", "BatchNorm2d", "ReLU"]
    with pytest.raises(RuntimeError):            # no CPU fallback
        net(torch.zeros(1, 3, 32, 32))


def test_param_order_and_gradient_groups():
    from selectivenet_for_semantic_segmentation_binary_b200.engine import FlatGrads, param_order
    sd = O.init_state_dict(0, "RGB", True)
    order = param_order(True)
    assert order == [k for k in sd if "running" not in k and "num_batches" not in k]
    fg = FlatGrads({n: tuple(sd[n].shape) for n in order}, order, "cpu")
    ranges = fg.group_ranges()
    # the groups tile the flat buffer without gaps, in reverse order of how backward finishes them
    tags = ["enc1", "enc2", "enc3", "dec4", "dec3", "dec2", "dec1"]
    pos = 0
    for t in tags:
        lo, hi = ranges[t]
        assert lo == pos and hi > lo
        pos = hi
    assert pos == fg.total
    for n in order:
        o, k = fg.offsets[n]
        assert o % 4 == 0 and fg.views[n].shape == sd[n].shape and fg.views[n].data_ptr() == fg.flat[o:].data_ptr()
    # every parameter of a group lies inside its range
    lo, hi = ranges["dec4"]
    for n in order:
        if n.startswith("decoder_layer_4"):
            assert lo <= fg.offsets[n][0] < hi
    lo, hi = ranges["dec1"]
    for n in ("unpool1.weight", "decoder_layer_1_1.0.weight", "conv_aux.bias"):
        assert lo <= fg.offsets[n][0] < hi


def test_chunk_bounds_follow_torch_chunk():
    from selectivenet_for_semantic_segmentation_binary_b200.trainer import chunk_bounds
    for total, world in ((128, 8), (128, 2), (100, 8), (5, 8), (16, 4)):
        sizes = [c.numel() for c in torch.arange(total).chunk(world)]
        got = [chunk_bounds(total, world, r) for r in range(world)]
        assert [hi - lo for lo, hi in got if hi > lo] == sizes
        assert got[0][0] == 0 and max(hi for _, hi in got) == total


def test_logit_threshold_matches_oracle_and_numpy():
    from selectivenet_for_semantic_segmentation_binary_b200.utils.compute_metric import logit_threshold
    for cut in (0.5, 0.3, 0.7, 0.05, 0.95):
        for path in ("train", "eval"):
            assert np.float32(logit_threshold(cut, path)) == O.logit_threshold(cut, path)
    assert np.float32(logit_threshold(0.5, "train")).view(np.uint32) == 0x25340000
    assert np.float32(logit_threshold(0.5, "eval")).view(np.uint32) == 0x34044623
    # raw-logit comparison (--output_scale None): x > cut  <=>  x >= nextafter(float32(cut))
    x = np.float32([0.5, np.nextafter(np.float32(0.5), np.float32(1)), 0.49999997, 0.7])
    thr = np.float32(logit_threshold(0.5, "eval", scale="None"))
    np.testing.assert_array_equal(x >= thr, x > 0.5)


def test_evaluator_metrics_formulas_without_gpu():
    """The derived metrics are the reference's float64 formulas; check them on the notebook example by
    seeding the host-side matrix (no kernel involved)."""
    from selectivenet_for_semantic_segmentation_binary_b200.utils.compute_metric import Evaluator
    ev = Evaluator(2, False)
    ev.confusion_matrix = np.array([[2., 1.], [2., 4.]])
    ref = O.Evaluator(2, False)
    ref.confusion_matrix = np.array([[2., 1.], [2., 4.]])
    assert ev.get_Pixel_Accuracy() == ref.get_Pixel_Accuracy()
    np.testing.assert_array_equal(ev.get_Precision(), ref.get_Precision())
    np.testing.assert_array_equal(ev.get_Recall(), ref.get_Recall())
    assert ev.get_mIoU() == ref.get_mIoU() and ev.get_FWIoU() == ref.get_FWIoU()
    np.testing.assert_array_equal(ev.get_Dice_Score(), ref.get_Dice_Score())
    np.testing.assert_array_equal(ev.get_F1_Score(ev.get_Precision(), ev.get_Recall()),
                                  ref.get_F1_Score(ref.get_Precision(), ref.get_Recall()))
    with pytest.raises(ValueError):
        Evaluator(3, False)
    ev.reset()
    assert ev.confusion_matrix.sum() == 0


def test_dp_bucket_plan_tiles_the_flat_gradient_buffer():
    """Every bucket plan must cover each gradient element exactly once, in arrival order (trainer.bucket_plan)."""
    from selectivenet_for_semantic_segmentation_binary_b200.engine import FlatGrads, param_order
    from selectivenet_for_semantic_segmentation_binary_b200.model import UNet_B
    from selectivenet_for_semantic_segmentation_binary_b200.trainer import GROUP_ORDER, bucket_plan
    net = UNet_B("RGB", selective=True)
    shapes = {n: tuple(p.shape) for n, p in net.named_parameters()}
    fg = FlatGrads(shapes, param_order(True), "cpu")
    ranges = fg.group_ranges()
    for spec in ("dec1,dec2,dec3,dec4,enc3,enc2,enc1", "enc1", "", "dec3,enc1", "dec2,dec4"):
        plan = bucket_plan(ranges, fg.total, spec)
        assert "enc1" in plan
        covered = torch.zeros(fg.total, dtype=torch.int32)
        prev_lo = fg.total
        for tag in GROUP_ORDER:
            if tag in plan:
                lo, hi = plan[tag]
                assert hi == prev_lo and lo == ranges[tag][0]
                covered[lo:hi] += 1
                prev_lo = lo
        assert bool((covered == 1).all())
    with pytest.raises(ValueError):
        bucket_plan(ranges, fg.total, "dec9")
