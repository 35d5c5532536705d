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
    assert "train | loss:" in r.stdout and "rejection ratio" in r.stdout
    ckpt_dir = os.path.join(model_dir, "1-fold", "checkpoint")
    files = sorted(os.listdir(ckpt_dir))
    assert files == ["model_epoch1.pth", "model_epoch2.pth"]
    ck = torch.load(os.path.join(ckpt_dir, files[-1]), map_location="cpu")
    assert set(ck) == {"net", "optim"} and len(ck["net"]) == 110
    # the loss must go down over the two epochs (same 16 patches each epoch)
    losses = [float(l.split("loss:")[1].split(",")[0]) for l in r.stdout.splitlines() if l.startswith("train | loss:")]
    assert len(losses) == 2 and losses[1] < losses[0], losses

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
