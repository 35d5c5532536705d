"""Input transforms (SURVEY.md §8(f) "next" #2): the oracle restatement against golden vectors produced by the
reference's own Normalization / RandomFlip / ToTensor classes (tests/golden/make_transform_golden.py), the package's
host-side mirror classes, and — on the GPU — the fused uint8 kernel against both, bit for bit."""
import os

import numpy as np
import pytest
import torch

from oracle import sunet_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tg():
    return np.load(os.path.join(ROOT, "tests", "golden", "transform_golden.npz"))


def test_oracle_transform_matches_reference_golden(tg):
    for i in range(tg["img_u8"].shape[0]):
        x, lab = O.input_transform(tg["img_u8"][i], tg["label_u8"][i], int(tg["flip"][i]))
        assert x.dtype == np.float32 and lab.dtype == np.int64
        assert np.array_equal(x, tg["x"][i]) and np.array_equal(lab, tg["label"][i]), i
    # the byte table reproduces the same values without flips
    lut = O.input_lut()
    x0, _ = O.input_transform(tg["img_u8"][3], tg["label_u8"][3], 0)
    assert np.array_equal(lut[tg["img_u8"][3]].transpose(2, 0, 1), x0)
    assert lut[0] == -1.0 and lut[255] == 1.0


def test_host_transform_classes_match_reference_golden(tg):
    from selectivenet_for_semantic_segmentation_binary_b200.utils.data_utils import Normalization, RandomFlip, ToTensor
    norm, flip, tot = Normalization(0.5, 0.5), RandomFlip(), ToTensor()
    for i in range(tg["img_u8"].shape[0]):
        inp, lab = tg["img_u8"][i] / 255.0, tg["label_u8"][i] / 255.0
        inp, lab = inp.astype(np.float32), lab.astype(np.uint8)
        np.random.seed(100 + i)                          # the seed make_transform_golden.py used for this sample
        d = tot(flip(norm({"input": inp, "label": lab})))
        assert torch.equal(d["input"], torch.from_numpy(tg["x"][i]))
        assert torch.equal(d["label"], torch.from_numpy(tg["label"][i]))
        np.random.seed(100 + i)
        assert RandomFlip.draw() == int(tg["flip"][i])


@pytest.mark.gpu
def test_u8_kernel_is_bit_identical_to_transform_then_im2col(tg):
    from selectivenet_for_semantic_segmentation_binary_b200 import kernels as K
    dev = "cuda"
    rng = np.random.default_rng(5)
    for (B, H, W) in [(6, 16, 16), (3, 32, 48)]:
        if (B, H, W) == (6, 16, 16):
            img, lab, flip = tg["img_u8"], tg["label_u8"], tg["flip"]
        else:
            img = rng.integers(0, 256, size=(B, H, W, 3), dtype=np.uint8)
            lab = rng.choice(np.array([0, 255, 7], dtype=np.uint8), size=(B, H, W))
            flip = np.array([0, 1, 2], dtype=np.uint8)
        xs, ls = zip(*(O.input_transform(img[i], lab[i], int(flip[i])) for i in range(B)))
        x = torch.from_numpy(np.stack(xs)).to(dev)
        ref = torch.empty(B, H, W, 32, dtype=torch.bfloat16, device=dev)
        K.pack_input_im2col32(x, ref)
        got = torch.full_like(ref, float("nan"))
        lut = torch.from_numpy(O.input_lut()).to(dev)
        K.pack_input_u8_im2col32(torch.from_numpy(img).to(dev), lut, torch.from_numpy(flip).to(dev), got)
        l32 = torch.full((B, H, W), -1.0, device=dev)
        K.pack_label_u8(torch.from_numpy(lab).to(dev), torch.from_numpy(flip).to(dev), l32)
        torch.cuda.synchronize()
        assert torch.equal(got.view(torch.int16), ref.view(torch.int16))
        assert torch.equal(l32.cpu(), torch.from_numpy(np.stack(ls)).float())


@pytest.mark.gpu
def test_step_u8_equals_step_on_transformed_tensors():
    from selectivenet_for_semantic_segmentation_binary_b200.model import UNet_B
    from selectivenet_for_semantic_segmentation_binary_b200.trainer import SUNetTrainer
    rng = np.random.default_rng(11)
    B, S = 4, 64
    img = rng.integers(0, 256, size=(B, S, S, 3), dtype=np.uint8)
    lab = rng.choice(np.array([0, 255], dtype=np.uint8), size=(B, S, S), p=[0.6, 0.4])
    flip = np.array([0, 1, 2, 3], dtype=np.uint8)
    xs, ls = zip(*(O.input_transform(img[i], lab[i], int(flip[i])) for i in range(B)))
    x = torch.from_numpy(np.stack(xs)).cuda()
    l32 = torch.from_numpy(np.stack(ls)).float().cuda()
    res = []
    for mode in ("f32", "u8"):
        torch.manual_seed(0)
        net = UNet_B("RGB", selective=True).cuda()
        net.train()
        tr = SUNetTrainer(net, lr=1e-3, s_lamb=2)
        for _ in range(4):                                   # crosses the eager -> CUDA-graph switch
            if mode == "f32":
                r = tr.step(x, l32)
            else:
                r = tr.step_u8(torch.from_numpy(img).cuda(), torch.from_numpy(lab).cuda(), torch.from_numpy(flip).cuda())
        torch.cuda.synchronize()
        res.append((r.clone(), {n: p.detach().clone() for n, p in net.named_parameters()}))
    assert torch.equal(res[0][0], res[1][0])                 # identical operand -> identical step
    for n in res[0][1]:
        assert torch.equal(res[0][1][n], res[1][1][n]), n
