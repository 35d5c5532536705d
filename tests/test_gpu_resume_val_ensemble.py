"""API rows finished in round 2, on the GPU through the C ABI:
  * optimizer state in the checkpoint (utils/net_utils.py:5-40): the fused Adam equals torch.optim.Adam step for
    step, its state_dict loads into torch.optim.Adam and back, net_save -> net_train_load resumes bit-exactly;
  * the validation loop body (train.py:275-331): eval-mode forward + losses without gradients == the oracle;
  * the evaluation ensemble (eval.py:209-222): mean of re-scaled member maps bit-identical to numpy;
  * one CUDA graph per batch shape (a ragged last batch, DataLoader drop_last=False) and uneven-shard pixel counts."""
import copy
import os

import numpy as np
import pytest
import torch

from oracle import sunet_oracle as O

pytestmark = pytest.mark.gpu


def _net(selective=True, seed=0):
    from selectivenet_for_semantic_segmentation_binary_b200.model import UNet_B
    torch.manual_seed(seed)
    return UNet_B("RGB", selective=selective).cuda()


def test_fused_adam_matches_torch_adam_and_state_round_trips(tmp_path):
    from selectivenet_for_semantic_segmentation_binary_b200.optim import Adam
    g = torch.Generator().manual_seed(0)
    shapes = [(64, 3, 3, 3), (64,), (128, 64, 3, 3), (1,), (7,), (512, 33)]
    ours_p = [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in shapes]
    ref_p = [torch.nn.Parameter(p.detach().clone()) for p in ours_p]
    ours = Adam(ours_p, lr=1e-3, weight_decay=5e-4)
    ref = torch.optim.Adam(ref_p, lr=1e-3, weight_decay=5e-4)
    for it in range(5):
        if it == 3:                       # a scheduler changes the learning rate through param_groups
            ours.param_groups[0]["lr"] = 5e-4
            ref.param_groups[0]["lr"] = 5e-4
        for a, b in zip(ours_p, ref_p):
            gr = torch.randn(a.shape, generator=g).cuda()
            a.grad, b.grad = gr.clone(), gr.clone()
        ours.step()
        ref.step()
    for a, b in zip(ours_p, ref_p):
        assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), (a - b).abs().max()
    # our state -> torch.optim.Adam (what the reference's net_train_load does) -> one more step each
    sd = ours.state_dict()
    ref2 = torch.optim.Adam(ref_p, lr=1.0)
    ref2.load_state_dict(sd)
    assert ref2.param_groups[0]["lr"] == 5e-4 and float(ref2.state_dict()["state"][0]["step"]) == 5.0
    # torch's state -> ours
    ours2 = Adam(ours_p, lr=1.0)
    ours2.load_state_dict(ref.state_dict())
    for a, b in zip(ours_p, ref_p):
        gr = torch.randn(a.shape, generator=g).cuda()
        a.grad, b.grad = gr.clone(), gr.clone()
    ours2.step()
    ref2.step()
    for a, b in zip(ours_p, ref_p):
        assert torch.allclose(a, b, rtol=4e-6, atol=4e-7)
    assert int(ours2.step_dev.item()) == 6


def test_checkpoint_resume_is_bit_exact(tmp_path):
    """net_save(ckpt_dir, net, trainer.optimizer, epoch) / net_train_load: two steps, save, one more step ==
    load into a fresh model + trainer, one step."""
    from selectivenet_for_semantic_segmentation_binary_b200.trainer import SUNetTrainer
    from selectivenet_for_semantic_segmentation_binary_b200.utils.net_utils import net_save, net_train_load
    x, label = O.synthetic_batch(2, 32, seed=5)
    x, label = x.cuda(), label.cuda()
    net = _net()
    net.train()
    tr = SUNetTrainer(net, lr=1e-3, s_lamb=2, use_cuda_graph=False)
    for _ in range(2):
        tr.step(x, label)
    ck = str(tmp_path / "ck")
    net_save(ck, net, tr.optimizer, 2)
    saved = torch.load(os.path.join(ck, "model_epoch2.pth"), map_location="cpu")
    assert set(saved) == {"net", "optim"} and len(saved["optim"]["state"]) == 68
    assert float(saved["optim"]["state"][0]["step"]) == 2.0
    # a stock torch.optim.Adam accepts it (the reference's net_train_load path)
    probe_net = _net(seed=1)
    torch.optim.Adam(probe_net.parameters()).load_state_dict(saved["optim"])
    res_a = tr.step(x, label).clone()
    net2 = _net(seed=123)
    net2.train()
    tr2 = SUNetTrainer(net2, lr=0.5, s_lamb=2, use_cuda_graph=False)
    net2, _, epoch = net_train_load(ck, net2, tr2.optimizer, device="cpu")
    assert epoch == 2 and int(tr2.step_dev.item()) == 2
    res_b = tr2.step(x, label).clone()
    torch.cuda.synchronize()
    assert torch.equal(res_a, res_b)
    for (n, a), (_, b) in zip(net.state_dict().items(), net2.state_dict().items()):
        assert torch.equal(a, b), n
    for a, b in zip(tr.exp_avg_sq, tr2.exp_avg_sq):
        assert torch.equal(a, b)


def test_validation_step_matches_oracle():
    """train.py:275-331 — net.eval() forward under no_grad, aux BCE + selective risk, float64-sigmoid thresholding."""
    from selectivenet_for_semantic_segmentation_binary_b200.trainer import SUNetTrainer
    from selectivenet_for_semantic_segmentation_binary_b200.utils.compute_metric import Evaluator
    net = _net()
    sd = O.init_state_dict(0, "RGB", True)
    # non-trivial running statistics: one training step on both sides first
    x0, l0 = O.synthetic_batch(2, 64, seed=7)
    net.train()
    ev = Evaluator(2, True, device=torch.device("cuda"))
    tr = SUNetTrainer(net, lr=0.0, s_lamb=2, val_evaluator=ev)
    tr.step(x0.cuda(), l0.cuda())
    O.unet_b_forward(sd, x0, True, True, update_running=True)
    before = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x, label = O.synthetic_batch(3, 64, seed=8)
    for _ in range(4):                    # eager, eager, capture + replay, replay: all must count the same batch
        res = tr.validate(x.cuda(), label.cuda()).clone()
    torch.cuda.synchronize()
    for k, v in net.state_dict().items():                 # no parameter / running-stat update in validation
        assert torch.equal(v, before[k]), k
    with torch.no_grad():
        out, sel, aux = O.unet_b_forward(sd, x, True, False)      # (selective, training=False)
        l_sel, cov = O.selective_risk_b(out, sel, label, lamb=2)
        l_aux = O.bce_with_logits_mean(aux, label)
    got = res.cpu().tolist()
    for g_, r_ in zip(got, [l_sel.item(), cov.item(), l_aux.item(), (l_sel + l_aux).item()]):
        assert abs(g_ - r_) <= 2e-2 * abs(r_), (got, r_)
    assert ev.total == 4 * 3 * 64 * 64                    # four validate() calls, nothing lost by the graph path
    ev.reset()
    tr.validate(x.cuda(), label.cuda())
    lg = tr.net._plan_for(x.cuda()).logits.view(3, 3, 64, 64).cpu().numpy()
    pred, selm = O.postprocess(lg[0], lg[1], path="train")
    oe = O.Evaluator(2, True)
    oe.add_batch(label.numpy().astype("uint8"), pred, selection=selm)
    assert (ev.confusion_matrix == oe.confusion_matrix).all()


@pytest.mark.parametrize("scale", ["None", "clip", "minmax"])
def test_ensemble_mean_is_bit_identical_to_numpy(scale):
    """eval.py:209-222: output = np.mean(np.asarray([scale(o) for o in outputs]), axis=0) in float32."""
    from selectivenet_for_semantic_segmentation_binary_b200 import kernels as K
    g = torch.Generator().manual_seed(3)
    maps = [(torch.randn(5, 64, 63, generator=g) * (1 + m)).contiguous() for m in range(3)]      # numel % 4 != 0 too
    dm = [m.cuda() for m in maps]
    ws = K.new_workspace("cuda")
    mm = None
    if scale == "minmax":
        mm = [torch.empty(2, device="cuda") for _ in dm]
        for m, e in zip(dm, mm):
            K.minmax_f32(m, e, ws)
        for m, e in zip(maps, mm):
            assert e.cpu().tolist() == [float(m.min()), float(m.max())]
    mean = torch.empty_like(dm[0])
    K.ensemble_mean(dm, scale, mean, minmax=mm)
    torch.cuda.synchronize()
    fn = {"None": lambda v: v, "clip": lambda v: np.clip(v, 0, 1),
          "minmax": lambda v: (v - v.min()) / (v.max() - v.min())}[scale]
    ref = np.mean(np.asarray([fn(m.numpy()) for m in maps]), axis=0)
    assert ref.dtype == np.float32
    np.testing.assert_array_equal(mean.cpu().numpy(), ref)


def test_ragged_batches_keep_their_graphs_and_global_pixel_count():
    """DataLoader(drop_last=False): full batches and a short tail alternate; each shape keeps its own captured graph,
    and every result equals the eager step of the same model state (uneven pixel counts come from device memory)."""
    from selectivenet_for_semantic_segmentation_binary_b200.trainer import SUNetTrainer
    x, label = O.synthetic_batch(6, 32, seed=9)
    x, label = x.cuda(), label.cuda()
    net_a, net_b = _net(), _net()
    net_a.train()
    net_b.train()
    ta = SUNetTrainer(net_a, lr=1e-3, s_lamb=2, use_cuda_graph=True)
    tb = SUNetTrainer(net_b, lr=1e-3, s_lamb=2, use_cuda_graph=False)
    for it in range(8):
        lo, hi = (0, 4) if it % 2 == 0 else (4, 6)
        ra = ta.step(x[lo:hi], label[lo:hi]).clone()
        rb = tb.step(x[lo:hi], label[lo:hi]).clone()
        assert torch.allclose(ra, rb, rtol=1e-5, atol=1e-7), (it, ra, rb)
    torch.cuda.synchronize()
    assert ta.graph_active("train") and len([k for k in ta._graphs if k[0] == "train"]) == 2
    for (n, a), (_, b) in zip(net_a.state_dict().items(), net_b.state_dict().items()):
        assert torch.allclose(a.float(), b.float(), rtol=1e-4, atol=1e-6), n
