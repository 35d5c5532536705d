"""train.py / eval.py keep the reference's command line and run end to end on synthetic patches."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_train_then_eval_cli(tmp_path):
    model_dir = str(tmp_path / "model")
    cmd = [sys.executable, os.path.join(ROOT, "train.py"), "--model_arch", "UNet_B", "--selective", "1", "--s_lamb", "2",
           "--loss", "BCElogit", "--batch_size", "4", "--n_epoch", "2", "--patch_size", "64", "--local_rank", "0",
           "--model_dir", model_dir, "--synthetic", "16"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "train_loss" in r.stdout and "valid_loss" in r.stdout and "train_rejection" in r.stdout, r.stdout[-1500:]
    ckpt_dir = os.path.join(model_dir, "1-fold", "checkpoint")
    files = sorted(os.listdir(ckpt_dir))
    assert files == ["model_epoch1.pth", "model_epoch2.pth"]
    ck = torch.load(os.path.join(ckpt_dir, files[-1]), map_location="cpu")
    assert set(ck) == {"net", "optim"} and len(ck["net"]) == 110
    assert set(ck["optim"]) == {"state", "param_groups"} and len(ck["optim"]["state"]) == 68      # a real Adam state
    # the loss must go down over the two epochs (same 16 patches each epoch); validation runs every epoch
    lines = [l for l in r.stdout.splitlines() if l.startswith("train_loss")]
    losses = [float(l.split()[1]) for l in lines]
    vals = [float(l.split("valid_loss")[1].split()[0]) for l in lines]
    assert len(losses) == 2 and losses[1] < losses[0], losses
    assert all(np.isfinite(v) for v in vals), vals

    # eval.py on the saved checkpoint (single .pth in the directory)
    os.remove(os.path.join(ckpt_dir, files[0]))
    cmd = [sys.executable, os.path.join(ROOT, "eval.py"), "--model_dir", ckpt_dir, "--selective", "1", "--select_eval",
           "1", "--batch_size", "4", "--patch_size", "64", "--local_rank", "0", "--synthetic", "8", "--cut_off", "0.5",
           "--s_cut_off", "0.4"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for key in ("rejection ratio:", "Acc:", "Acc_class:", "Prec:", "mIoU:", "IoU_class:"):
        assert key in r.stdout, r.stdout[-1500:]


def test_eval_counts_match_oracle_on_same_logits():
    """eval.py's counting path: GPU forward in eval mode, then the same logits through the oracle's numpy
    float32 post-processing + Evaluator must give identical integer counts."""
    from oracle import sunet_oracle as O
    from selectivenet_for_semantic_segmentation_binary_b200.model import UNet_B
    from selectivenet_for_semantic_segmentation_binary_b200.utils.compute_metric import Evaluator
    torch.manual_seed(3)
    net = UNet_B("RGB", selective=True).cuda()
    net.train(False)
    x, label = O.synthetic_batch(4, 64, seed=21)
    with torch.no_grad():
        out, sel, _ = net(x.cuda())
    ev = Evaluator(2, True)
    ev.add_batch_from_logits(label.cuda().to(torch.uint8), out, sel, cut_off=0.45, s_cut_off=0.55, path="eval")
    pred, selm = O.postprocess(out.cpu().numpy(), sel.cpu().numpy(), path="eval", cut_off=0.45, s_cut_off=0.55)
    ref = O.Evaluator(2, True)
    ref.add_batch(label.numpy().astype("uint8"), pred, selection=selm)
    np.testing.assert_array_equal(ev.confusion_matrix, ref.confusion_matrix)
    assert ev.total_reject == int(pred.size - selm.sum())


def test_eval_ensemble_and_scales_cli(tmp_path):
    """eval.py with several checkpoints in --model_dir = the ensemble branch (eval.py:208-222) with --ens_scale."""
    from selectivenet_for_semantic_segmentation_binary_b200.model import UNet_B
    from selectivenet_for_semantic_segmentation_binary_b200.utils.net_utils import net_save

    class _NoOpt:
        def state_dict(self):
            return {}
    mdir = str(tmp_path / "ens")
    for i in range(3):
        torch.manual_seed(10 + i)
        net_save(mdir, UNet_B("RGB", selective=False), _NoOpt(), i + 1)
    for ens_scale in ("None", "minmax"):
        cmd = [sys.executable, os.path.join(ROOT, "eval.py"), "--model_dir", mdir, "--batch_size", "4", "--patch_size",
               "64", "--local_rank", "0", "--synthetic", "12", "--ens_scale", ens_scale, "--single_scale", "None",
               "--cut_off", "0.4"]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert "3 model(s)" in r.stdout and "mIoU:" in r.stdout, r.stdout[-1500:]
    # selective checkpoints cannot be ensembled (the reference's branch has no selection either)
    smdir = str(tmp_path / "ens_sel")
    for i in range(2):
        net_save(smdir, UNet_B("RGB", selective=True), _NoOpt(), i + 1)
    cmd = [sys.executable, os.path.join(ROOT, "eval.py"), "--model_dir", smdir, "--selective", "1", "--batch_size", "4",
           "--patch_size", "64", "--local_rank", "0", "--synthetic", "4"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "ensemble" in (r.stdout + r.stderr)


def test_train_cli_unet_cross_entropy(tmp_path):
    """The reference's default pair --model_arch UNet --loss CE (train.py:70-86) through the same CLI."""
    model_dir = str(tmp_path / "model_ce")
    cmd = [sys.executable, os.path.join(ROOT, "train.py"), "--model_arch", "UNet", "--loss", "CE", "--selective", "1",
           "--s_lamb", "2", "--batch_size", "4", "--n_epoch", "2", "--patch_size", "32", "--local_rank", "0",
           "--model_dir", model_dir, "--synthetic", "8", "--lr_sche", "StepLR", "--patience", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("train_loss")]
    losses = [float(l.split()[1]) for l in lines]
    assert len(losses) == 2 and losses[1] < losses[0], r.stdout[-1500:]
    assert "learning rate 0.001" in r.stdout and "learning rate 0.0005" in r.stdout          # StepLR(step 1, gamma .5)
    ck = torch.load(os.path.join(model_dir, "1-fold", "checkpoint", "model_epoch2.pth"), map_location="cpu")
    assert ck["net"]["conv1x1.weight"].shape == (2, 64, 1, 1) and len(ck["optim"]["state"]) == 68
