"""Checkpoint helpers with the reference's names and file layout.

Mirrors /root/reference/utils/net_utils.py:5-52: ``{'net': state_dict, 'optim': state_dict}``
saved as ``model_epoch%d.pth``; loaders accept keys with or without DataParallel's ``module.``
prefix.  The 110-key fp32 state_dict of ``UNet_B`` is the checkpoint contract.
"""
import os
from collections import OrderedDict

import torch


def net_save(ckpt_dir, net, optim, epoch):
    if not os.path.exists(ckpt_dir):
        os.makedirs(ckpt_dir)
    torch.save({'net': net.state_dict(), 'optim': optim.state_dict()}, '%s/model_epoch%d.pth' % (ckpt_dir, epoch))


def remove_module(ckpt):
    net_state_dict = OrderedDict()
    for k, v in ckpt['net'].items():
        net_state_dict[k.replace("module.", "")] = v
    return net_state_dict


def _latest(ckpt_dir):
    ckpt_lst = os.listdir(ckpt_dir)
    ckpt_lst.sort(key=lambda f: int(''.join(filter(str.isdigit, f))))
    return ckpt_lst[-1]


def net_train_load(ckpt_dir, net, optim, device=None):
    if not os.path.exists(ckpt_dir):
        epoch = 0
        return net, optim, epoch
    last = _latest(ckpt_dir)
    print('model: ', last)
    ckpt = torch.load('%s/%s' % (ckpt_dir, last), map_location=device if device is not None else 'cpu')
    try:
        ckpt['net'] = remove_module(ckpt)
    except Exception:
        pass
    net.load_state_dict(ckpt['net'])
    optim.load_state_dict(ckpt['optim'])
    epoch = int(last.split('epoch')[1].split('.pth')[0])
    return net, optim, epoch


def net_test_load(model_path, net, device=None):
    if device is not None:
        ckpt = torch.load(model_path, map_location=device)
    else:
        ckpt = torch.load(model_path)
    try:
        ckpt['net'] = remove_module(ckpt)
    except Exception:
        pass
    net.load_state_dict(ckpt['net'])
    return net
