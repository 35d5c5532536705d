"""The CPU oracle against (a) vectors produced by the real reference (tests/golden/make_golden.py)
and (b) the reference's own notebook known-answers.  No GPU, no reference import."""
import numpy as np
import pytest
import torch

from oracle import sunet_oracle as O


def test_state_dict_layout(golden):
    sd = O.init_state_dict(0, "RGB", True)
    assert len(sd) == int(golden["n_keys"]) == 110
    trainable = [k for k in sd if "running" not in k and "num_batches" not in k]
    assert trainable == [str(s) for s in golden["param_names"]]
    assert sum(sd[k].numel() for k in trainable) == int(golden["n_params"]) == 7703107
    assert len(O.init_state_dict(0, "RGB", False)) == 106   # 110 minus the two selective heads
    assert sum(v.numel() for k, v in O.init_state_dict(0, "RGB", False).items()
               if "running" not in k and "num_batches" not in k) == 7702977   # u-net_training.ipynb cell 1


def test_train_step_matches_reference(golden):
    sd = O.init_state_dict(0, "RGB", True)
    names = [str(s) for s in golden["param_names"]]
    for n in names:
        sd[n].requires_grad_(True)
    x, label = O.synthetic_batch(2, 32, seed=0)
    loss, aux = O.train_losses(sd, x, label, s_lamb=2, selective=True)
    loss.backward()
    np.testing.assert_allclose(aux["output"].detach().numpy(), golden["train_output"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(aux["selection"].detach().numpy(), golden["train_selection"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(aux["aux"].detach().numpy(), golden["train_aux"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(aux["aux_loss"].item(), golden["aux_loss"], rtol=1e-6)
    np.testing.assert_allclose(aux["select_loss"].item(), golden["select_loss"], rtol=1e-6)
    np.testing.assert_allclose(aux["coverage"].item(), golden["coverage"], rtol=1e-6)
    np.testing.assert_allclose(loss.item(), golden["loss"], rtol=1e-6)
    gn = np.array([sd[n].grad.norm().item() for n in names])
    np.testing.assert_allclose(gn, golden["grad_norm"], rtol=1e-4, atol=1e-9)
    gh = np.stack([np.resize(sd[n].grad.reshape(-1)[:8].numpy(), 8) for n in names])
    np.testing.assert_allclose(gh, golden["grad_head8"], rtol=1e-3, atol=1e-7)
    bn_sum = np.array([sd[str(k)].double().sum().item() for k in golden["bn_keys"]])
    np.testing.assert_allclose(bn_sum, golden["bn_sum"], rtol=1e-5, atol=1e-7)


def test_eval_and_postprocessing_match_reference(golden):
    sd = O.init_state_dict(0, "RGB", True)
    x, label = O.synthetic_batch(2, 32, seed=0)
    with torch.no_grad():
        out, sel, _ = O.unet_b_forward(sd, x, True, True)          # one training forward updates running stats
    pred, s = O.postprocess(out.numpy(), sel.numpy(), path="train")
    ev = O.Evaluator(2, True)
    ev.add_batch(label.numpy().astype("uint8"), pred, selection=s)
    np.testing.assert_array_equal(ev.confusion_matrix, golden["train_cm"])
    assert label.numel() - s.sum() == float(golden["train_reject"])
    assert ev.get_Pixel_Accuracy() == float(golden["train_acc"])
    assert ev.get_mIoU() == float(golden["train_miou"])
    x2, label2 = O.synthetic_batch(2, 32, seed=10)
    with torch.no_grad():
        o2, s2, _ = O.unet_b_forward(sd, x2, True, False)
    np.testing.assert_allclose(o2.numpy(), golden["eval_output"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(s2.numpy(), golden["eval_selection"], rtol=1e-5, atol=1e-6)
    # counts are compared on the reference's own logits so the check is exact
    for cut, scut, tag in ((0.5, 0.5, "a"), (0.3, 0.6, "b")):
        pred2, sel2 = O.postprocess(golden["eval_output"], golden["eval_selection"], path="eval", cut_off=cut,
                                    s_cut_off=scut)
        for selective, t2 in ((True, "sel"), (False, "all")):
            e = O.Evaluator(2, selective)
            e.add_batch(label2.numpy().astype("uint8"), pred2, selection=sel2 if selective else None)
            np.testing.assert_array_equal(e.confusion_matrix, golden[f"eval_cm_{tag}_{t2}"])
        assert pred2.size - sel2.sum() == float(golden[f"eval_reject_{tag}"])


def test_notebook_known_answers(golden):
    """chcek_losses.ipynb / check_metrics.ipynb printed values."""
    target = torch.tensor([[[1., 0., 1.], [1., 1., 1.], [0., 0., 1.]]])
    logit1 = torch.tensor([[[1., 0., 1.], [1., 1., 0.], [0., 0., 0.]]])
    assert abs(O.bce_with_logits_mean(logit1, target).item() - 0.5243) < 5e-5
    np.testing.assert_allclose(O.bce_with_logits_mean(logit1, target).item(), golden["nb_bce"], rtol=1e-6)
    selx = torch.tensor([[[2., -1., .5], [3., 0., -2.], [1., 1., -.5]]])
    l0, c0 = O.selective_risk_b(logit1, torch.zeros_like(logit1), target, lamb=2)
    l1, c1 = O.selective_risk_b(logit1, selx, target, lamb=2)
    np.testing.assert_allclose([l0.item(), c0.item(), l1.item(), c1.item()], golden["nb_sel_loss"], rtol=1e-6)
    np.testing.assert_allclose([l0.item(), c0.item()], [0.70430923, 0.5], rtol=1e-6)
    np.testing.assert_allclose([l1.item(), c1.item()], [0.5769160985946655, 0.5759591460227966], rtol=1e-6)
    # Evaluator: check_metrics.ipynb cells 1-5
    t = target.numpy().astype("uint8")
    pred = np.array([[[1, 1, 1], [1, 1, 0], [0, 0, 0]]], dtype="uint8")
    ev = O.Evaluator(2, False)
    ev.add_batch(t, pred)
    np.testing.assert_array_equal(ev.confusion_matrix, [[2, 1], [2, 4]])
    assert abs(ev.get_Pixel_Accuracy() - 2 / 3) < 1e-12
    np.testing.assert_allclose(ev.get_Precision(), [0.5, 0.8])
    np.testing.assert_allclose(ev.get_Recall(), [2 / 3, 2 / 3])
    np.testing.assert_allclose(ev.get_F1_Score(ev.get_Precision(), ev.get_Recall()), [0.5714285714, 0.7272727273])
    np.testing.assert_allclose(ev.get_mIoU(), 0.4857142857)
    np.testing.assert_allclose(ev.get_IoU_Class(), [0.4, 0.5714285714])
    evs = O.Evaluator(2, True)
    evs.add_batch(t, pred, selection=np.array([[[1, 0, 1], [1, 1, 0], [1, 1, 0]]], dtype="float64"))
    np.testing.assert_array_equal(evs.confusion_matrix, [[2, 0], [0, 4]])
    assert evs.get_Pixel_Accuracy() == 1.0 and evs.get_mIoU() == 1.0
    # numpy sigmoid goldens (check_metrics.ipynb cells 6-7)
    assert O.sigmoid_np(np.array([0.5]), np.float64)[0] == 0.6224593312018546
    assert O.sigmoid_np(np.array([0.0]), np.float64)[0] == 0.5


def test_logit_thresholds():
    """the host decision sigmoid(x) > 0.5 is x >= x*, not x > 0."""
    t_train = O.logit_threshold(0.5, "train")
    t_eval = O.logit_threshold(0.5, "eval")
    assert np.array([t_train], dtype=np.float32).view(np.uint32)[0] == 0x25340000
    assert np.array([t_eval], dtype=np.float32).view(np.uint32)[0] == 0x34044623
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(size=4000).astype(np.float32) * 3,
                        np.float32([0.0, -0.0, t_train, t_eval, np.nextafter(t_eval, np.float32(0)), 1e-8, 5e-8])])
    for path, thr in (("train", t_train), ("eval", t_eval)):
        pred, _ = O.postprocess(x, None, path=path)
        np.testing.assert_array_equal(pred, (x >= thr).astype("uint8"))
    for cut in (0.3, 0.7, 0.9):
        thr = O.logit_threshold(cut, "eval")
        pred, _ = O.postprocess(x, None, path="eval", cut_off=cut)
        np.testing.assert_array_equal(pred, (x >= thr).astype("uint8"))
