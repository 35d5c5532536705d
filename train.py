"""train.py with the reference's command line (/root/reference/train.py:12-55), B200-native underneath.

    python3 train.py --model_arch UNet_B --selective 1 --s_lamb 2 --loss BCElogit --batch_size 128 \
                     --n_epoch 200 --local_rank 0 1 2 3 4 5 6 7 --synthetic 1024

Differences from the reference, all forced by the hardware mapping and documented in INTEGRATION.md:
  * ``--local_rank`` is still the list of GPU ids, but each id gets its own process (spawned here) with
    weights resident, the batch sharded with torch.chunk sizes and NCCL all-reduces — not nn.DataParallel;
  * the per-batch loop body is ``SUNetTrainer.step`` (no per-step .item(), no host numpy);
  * ``--synthetic N`` trains on N seeded synthetic 200x_256-shaped patches (the reference ships no data);
    without it the reference's ``PatchDataset`` layout under ``--data_dir`` is read (jpg/png, PIL only);
  * only the path the north star names is built: ``--model_arch UNet_B --loss BCElogit`` (``UNet``/``CE``
    raise), Adam (``--optim SGD`` raises).
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse_arguments(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--data_dir', type=str, help='WSI data directory', default='/data')
    parser.add_argument('--fold', type=int, default=1, help='which fold in 5-fold cv')
    parser.add_argument('--input_type', type=str, default='RGB')
    parser.add_argument('--patch_mag', type=int, default=200)
    parser.add_argument('--patch_size', type=int, default=256)
    parser.add_argument('--n_cls', type=int, default=2)
    parser.add_argument('--model_dir', type=str, help='directory where logs and models would be saved',
                        default='/model')
    parser.add_argument('--model_arch', type=str, default='UNet', choices=['UNet', 'UNet_B'])
    parser.add_argument('--selective', type=bool, default=False, help='Is the network based on SelectiveNet?')
    parser.add_argument('--s_lamb', type=int, default=2, help='degree to follow target coverage')
    parser.add_argument('--output_dim', type=str, default='NHW', choices=['NCHW', 'NHW'])
    parser.add_argument('--output_scale', type=str, default='sigmoid', choices=['None', 'clip', 'sigmoid', 'minmax'])
    parser.add_argument('--optim', type=str, default='Adam', choices=['Adam', 'SGD'])
    parser.add_argument('--momentum', type=float, default=0, choices=[0.9])
    parser.add_argument('--w_decay', type=float, default=0, choices=[5e-4])
    parser.add_argument('--lr', type=float, default=1e-3)
    parser.add_argument('--lr_sche', type=str, default=None, choices=['StepLR', 'ReduceLR', 'CosineAnnealingLR'])
    parser.add_argument('--patience', type=int, default=10)
    parser.add_argument('--factor', type=float, default=0.5)
    parser.add_argument('--lr_min', type=float, default=1e-5)
    parser.add_argument('--loss', type=str, default='CE', choices=['BCElogit', 'CE'])
    parser.add_argument('--batch_size', type=int, default=16)
    parser.add_argument('--n_epoch', type=int, default=100)
    parser.add_argument('--local_rank', type=int, nargs='+', default=[0], help='local rank')
    parser.add_argument('--log_img', type=bool, default=False)
    # additions
    parser.add_argument('--synthetic', type=int, default=0, help='train on this many synthetic patches per epoch')
    parser.add_argument('--master_port', type=int, default=29533)
    args = parser.parse_args(argv)
    print('')
    print('args={}\n'.format(args))
    return args


class SyntheticPatches:
    """Seeded stand-in for PatchDataset + Normalization(0.5,0.5) + ToTensor (utils/data_utils.py:94-236)."""

    def __init__(self, n, size, in_ch, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.x = torch.rand(n, in_ch, size, size, generator=g) * 2 - 1
        self.y = (torch.rand(n, size, size, generator=g) < 0.4).float()

    def batches(self, batch_size):
        for i in range(0, self.x.shape[0] - batch_size + 1, batch_size):
            yield self.x[i:i + batch_size], self.y[i:i + batch_size]


def _lr_at(args, epoch, base_lr, tr_loss_hist, state):
    """StepLR / CosineAnnealingLR / ReduceLROnPlateau(train loss) as configured at train.py:94-101."""
    if args.lr_sche == 'StepLR':
        return base_lr * (args.factor ** (epoch // args.patience))
    if args.lr_sche == 'CosineAnnealingLR':
        import math
        return args.lr_min + (base_lr - args.lr_min) * (1 + math.cos(math.pi * epoch / args.n_epoch)) / 2
    if args.lr_sche == 'ReduceLR':
        best, bad, lr = state.get('best', float('inf')), state.get('bad', 0), state.get('lr', base_lr)
        if tr_loss_hist:
            cur = tr_loss_hist[-1]
            if cur < best * (1 - 1e-4):
                best, bad = cur, 0
            else:
                bad += 1
            if bad > args.patience:
                lr, bad = max(lr * args.factor, args.lr_min), 0
        state.update(best=best, bad=bad, lr=lr)
        return lr
    return base_lr


def train_worker(rank, world, args, ckpt_dir):
    import torch.distributed as dist
    from selectivenet_for_semantic_segmentation_binary_b200.model import UNet_B
    from selectivenet_for_semantic_segmentation_binary_b200.trainer import SUNetTrainer, chunk_bounds
    from selectivenet_for_semantic_segmentation_binary_b200.utils.compute_metric import Evaluator
    from selectivenet_for_semantic_segmentation_binary_b200.utils.net_utils import net_save, remove_module

    gpu = args.local_rank[rank]
    torch.cuda.set_device(gpu)
    dev = torch.device('cuda', gpu)
    group = None
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', str(args.master_port))
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
        group = dist.group.WORLD
    torch.manual_seed(0)                       # every rank draws the same initial weights
    net = UNet_B(args.input_type, selective=args.selective)
    start_epoch = 0
    if os.path.exists(ckpt_dir) and [f for f in os.listdir(ckpt_dir) if f.endswith('.pth')]:
        ckpts = sorted(os.listdir(ckpt_dir), key=lambda f: int(''.join(filter(str.isdigit, f))))
        ckpt = torch.load(os.path.join(ckpt_dir, ckpts[-1]), map_location='cpu')
        try:
            ckpt['net'] = remove_module(ckpt)
        except Exception:
            pass
        net.load_state_dict(ckpt['net'])       # optimizer state is not restored (train.py:126)
        start_epoch = int(ckpts[-1].split('epoch')[1].split('.pth')[0])
    net = net.to(dev)
    net.train()
    evaluator = Evaluator(num_class=args.n_cls, selective=args.selective, device=dev)
    trainer = SUNetTrainer(net, lr=args.lr, s_lamb=args.s_lamb, weight_decay=args.w_decay, process_group=group,
                           world_size=world, evaluator=evaluator)
    in_ch = net.input_ch
    if args.synthetic > 0:
        data = SyntheticPatches(args.synthetic, args.patch_size, in_ch)
    else:
        from selectivenet_for_semantic_segmentation_binary_b200.utils.data_utils import PatchArrays
        data = PatchArrays(args.data_dir, args.fold, args.patch_mag, args.patch_size, args.input_type)
    sched_state, loss_hist = {}, []
    for epoch in range(start_epoch + 1, start_epoch + args.n_epoch + 1):
        lr = _lr_at(args, epoch - 1, args.lr, loss_hist, sched_state)
        trainer.set_lr(lr)
        if rank == 0:
            print(f'epoch {epoch} / {start_epoch + args.n_epoch}, learning rate {lr}')
        acc = torch.zeros(4, device=dev)
        nb = 0
        for xb, yb in data.batches(args.batch_size):
            lo, hi = chunk_bounds(xb.shape[0], world, rank)
            res = trainer.step(xb[lo:hi].to(dev, non_blocking=True), yb[lo:hi].to(dev, non_blocking=True))
            acc += res                                   # stays on the device: no per-step sync
            nb += 1
        counts = evaluator.counts_tensor().clone()
        if world > 1:
            dist.all_reduce(counts)
        if rank == 0:
            m = (acc / max(nb, 1)).tolist()
            c = counts.cpu().numpy().astype(np.float64)
            tr_acc = (c[0] + c[3]) / max(c[:4].sum(), 1)
            print('train | loss: %.4f, accuracy: %.4f' % (m[3], tr_acc))
            if args.selective:
                print('     aux loss: %.4f | selection loss: %.4f, coverage: %.4f, rejection ratio: %.3f'
                      % (m[2], m[0], m[1], (c[5] - c[4]) / max(c[5], 1)))
            loss_hist.append(m[3])

            class _Opt:                                   # checkpoint keeps the reference's {'net','optim'} layout
                def state_dict(self_inner):
                    return {'step': int(trainer.step_dev.item()), 'lr': lr, 'type': 'sunet_b200.Adam'}
            net_save(ckpt_dir, net, _Opt(), epoch)
        evaluator.reset()
    if world > 1:
        # graphs that captured NCCL collectives make destroy_process_group() hang: synchronise and leave
        torch.cuda.synchronize(dev)
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


def train(args, ckpt_dir):
    if args.model_arch != 'UNet_B' or 'BCE' not in args.loss:
        raise SystemExit('this CLI drives the fused UNet_B/BCElogit step; the UNet/CE variant is available through the '
                         'module API (model.UNet, selective_loss.CrossEntropyLoss / calc_selective_risk_image; '
                         'DESIGN.md §5c)')
    if args.optim != 'Adam':
        raise SystemExit('only --optim Adam is built (the reference default)')
    world = len(args.local_rank)
    if world == 1:
        train_worker(0, 1, args, ckpt_dir)
    else:
        import torch.multiprocessing as mp
        mp.spawn(train_worker, args=(world, args, ckpt_dir), nprocs=world, join=True)


if __name__ == '__main__':
    args = parse_arguments()
    ckpt_dir = os.path.join(args.model_dir, f'{args.fold}-fold', 'checkpoint')
    train(args, ckpt_dir)
