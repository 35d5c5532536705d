"""Checkpoint helpers with the reference's names, signatures and file layout.

Contract (from /root/reference/utils/net_utils.py:5-52 and its callers train.py:357, eval.py:139-147): a checkpoint is
``{'net': state_dict, 'optim': state_dict}`` written to ``<ckpt_dir>/model_epoch<N>.pth``; the newest file of a directory
is the one with the largest number in its name; ``'net'`` keys may carry DataParallel's ``module.`` prefix.  The
110-key fp32 state_dict of ``UNet_B`` is the model half of that contract; the optimizer half is
``optim.Adam.state_dict()``, which has ``torch.optim.Adam``'s own layout.
"""
import os
import re
from collections import OrderedDict

import torch

_FILE = 'model_epoch{}.pth'


def net_save(ckpt_dir, net, optim, epoch):
    os.makedirs(ckpt_dir, exist_ok=True)
    torch.save({'net': net.state_dict(), 'optim': optim.state_dict()}, os.path.join(ckpt_dir, _FILE.format(epoch)))


def remove_module(ckpt):
    """'net' state_dict of a checkpoint without the ``module.`` that nn.DataParallel adds to every key."""
    return OrderedDict((key.replace('module.', ''), value) for key, value in ckpt['net'].items())


def _number_in(name):
    return int(''.join(re.findall(r'\d', name)))


def _read(path, device):
    ckpt = torch.load(path, map_location=device) if device is not None else torch.load(path)
    if isinstance(ckpt, dict) and isinstance(ckpt.get('net'), dict):
        ckpt['net'] = remove_module(ckpt)
    return ckpt


def net_train_load(ckpt_dir, net, optim, device=None):
    """Resume from the newest checkpoint of `ckpt_dir`: returns (net, optim, epoch); epoch 0 if there is none."""
    if not os.path.exists(ckpt_dir):
        return net, optim, 0
    newest = max(os.listdir(ckpt_dir), key=_number_in)
    print('model: ', newest)
    ckpt = _read(os.path.join(ckpt_dir, newest), device if device is not None else 'cpu')
    net.load_state_dict(ckpt['net'])
    optim.load_state_dict(ckpt['optim'])
    return net, optim, int(newest.split('epoch')[1].split('.pth')[0])


def net_test_load(model_path, net, device=None):
    net.load_state_dict(_read(model_path, device)['net'])
    return net
