"""train.py with the reference's command line (/root/reference/train.py:12-55), B200-native underneath.

    python3 train.py --model_arch UNet_B --selective 1 --s_lamb 2 --loss BCElogit --batch_size 128 \
                     --n_epoch 200 --local_rank 0 1 2 3 4 5 6 7 --synthetic 1024

Same epoch structure as the reference (train.py:164-357): training loop, learning-rate scheduler step, validation
loop under eval-mode BatchNorm, the same printed lines and TensorBoard scalars, checkpoint ``{'net', 'optim'}``
with a real Adam state.  Differences, all forced by the hardware mapping and documented in INTEGRATION.md:
  * ``--local_rank`` is still the list of GPU ids, but each id gets its own process (spawned here) with
    weights resident, the batch sharded with torch.chunk sizes and NCCL all-reduces — not nn.DataParallel;
  * ``--model_arch UNet_B --loss BCElogit``: the per-batch loop body is ``SUNetTrainer.step`` / ``.validate``
    (one CUDA graph per batch shape, no per-step .item(), no host numpy);
    ``--model_arch UNet --loss CE`` (the reference defaults): the module API (``model.UNet``,
    ``selective_loss.CrossEntropyLoss`` / ``calc_selective_risk_image``, ``optim.Adam``) on the same kernels;
  * ``--synthetic N`` trains on N seeded synthetic 200x_256-shaped patches (the reference ships no data);
    without it the reference's fold lists and patch files under ``--data_dir`` are read (PIL, RGB only);
  * Adam only (``--optim SGD`` raises).
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
    parser.add_argument('--synthetic_val', type=int, default=-1,
                        help='synthetic validation patches per epoch (default: a quarter of --synthetic, at least one '
                             'batch; 0 = no validation loop)')
    parser.add_argument('--resume_optim', type=bool, default=False,
                        help='also restore the Adam state on resume (the reference saves it but never loads it, '
                             'train.py:126)')
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

    def __len__(self):
        return self.x.shape[0]

    def batches(self, batch_size):
        """drop_last=False, like the reference's DataLoader (train.py:379-380)."""
        for i in range(0, self.x.shape[0], batch_size):
            yield self.x[i:i + batch_size], self.y[i:i + batch_size]


def make_scheduler(args, lr_holder):
    """The reference's scheduler objects (train.py:94-101), run on a one-parameter holder optimizer whose
    ``param_groups[0]['lr']`` is copied to the device-side learning rate once per epoch."""
    if args.lr_sche == 'StepLR':
        return torch.optim.lr_scheduler.StepLR(lr_holder, step_size=args.patience, gamma=args.factor)
    if args.lr_sche == 'ReduceLR':
        return torch.optim.lr_scheduler.ReduceLROnPlateau(lr_holder, mode='min', patience=args.patience,
                                                          factor=args.factor)
    if args.lr_sche == 'CosineAnnealingLR':
        return torch.optim.lr_scheduler.CosineAnnealingLR(lr_holder, T_max=args.patience, eta_min=args.lr_min)
    return None


class _Writers:
    """SummaryWriter pair of train.py:158-159 when tensorboard is importable, otherwise silent."""

    def __init__(self, log_dir, enabled):
        self.train = self.val = None
        if not enabled:
            return
        try:
            from torch.utils.tensorboard import SummaryWriter
            self.train = SummaryWriter(log_dir=os.path.join(log_dir, 'train'))
            self.val = SummaryWriter(log_dir=os.path.join(log_dir, 'valid'))
        except Exception as e:  # noqa: BLE001 - tensorboard is optional in this image
            print(f'(tensorboard not available: {e!r}; scalars are printed only)')

    def scalar(self, which, tag, value, epoch):
        w = self.train if which == 'train' else self.val
        if w is not None:
            w.add_scalar(tag, value, epoch)

    def images(self, tag, img, epoch):
        if self.train is not None:
            self.train.add_images(tag, img, epoch, dataformats='NHWC')

    def close(self):
        for w in (self.train, self.val):
            if w is not None:
                w.flush()


def _load_latest(ckpt_dir):
    """train.py:111-129: the newest checkpoint by the digits in its file name; `net` keys without 'module.'."""
    from selectivenet_for_semantic_segmentation_binary_b200.utils.net_utils import remove_module
    if not os.path.exists(ckpt_dir):
        return None, 0
    ckpt_lst = [f for f in os.listdir(ckpt_dir) if f.endswith('.pth')]
    if not ckpt_lst:
        return None, 0
    ckpt_lst.sort(key=lambda f: int(''.join(filter(str.isdigit, f))))
    ckpt = torch.load(os.path.join(ckpt_dir, ckpt_lst[-1]), map_location='cpu')
    try:
        ckpt['net'] = remove_module(ckpt)
    except Exception:  # noqa: BLE001
        pass
    print('Load weights from', os.path.join(ckpt_dir, ckpt_lst[-1]))
    return ckpt, int(ckpt_lst[-1].split('epoch')[1].split('.pth')[0])


def _datasets(args, in_ch):
    """(train iterable, valid iterable or None)"""
    if args.synthetic > 0:
        n_val = args.synthetic_val if args.synthetic_val >= 0 else max(args.batch_size, args.synthetic // 4)
        tr = SyntheticPatches(args.synthetic, args.patch_size, in_ch, seed=0)
        va = SyntheticPatches(n_val, args.patch_size, in_ch, seed=1) if n_val > 0 else None
        return tr, va
    from selectivenet_for_semantic_segmentation_binary_b200.utils.data_utils import PatchArrays, construct_train_valid
    if not os.path.exists(f'{args.data_dir}/1-fold_tumorable_data.npy') and \
            not os.path.exists(f'{args.data_dir}/2-fold_tumorable_data.npy'):
        raise SystemExit(f'no fold lists under {args.data_dir}: pass --synthetic N to train on synthetic patches')
    train_list, valid_list = construct_train_valid(args.data_dir, test_fold=args.fold)
    tr = PatchArrays(args.data_dir, train_list, args.patch_mag, args.patch_size, args.input_type, train=True)
    va = PatchArrays(args.data_dir, valid_list, args.patch_mag, args.patch_size, args.input_type, train=False)
    return tr, va


def train_worker(rank, world, args, ckpt_dir, log_dir):
    import torch.distributed as dist
    from selectivenet_for_semantic_segmentation_binary_b200 import selective_loss as SL
    from selectivenet_for_semantic_segmentation_binary_b200.model import UNet, UNet_B
    from selectivenet_for_semantic_segmentation_binary_b200.optim import Adam
    from selectivenet_for_semantic_segmentation_binary_b200.trainer import SUNetTrainer, shard_bounds
    from selectivenet_for_semantic_segmentation_binary_b200.utils.compute_metric import Evaluator
    from selectivenet_for_semantic_segmentation_binary_b200.utils.net_utils import net_save

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
    fused = args.model_arch == 'UNet_B'
    if fused:
        net = UNet_B(args.input_type, selective=args.selective)        # BCE loss, outputs (N, H, W)
    else:
        net = UNet(args.input_type, args.n_cls, selective=args.selective)   # CE loss, outputs (N, C, H, W)
    ckpt, start_epoch = _load_latest(ckpt_dir)
    if ckpt is not None:
        net.load_state_dict(ckpt['net'])       # optimizer state is not restored unless --resume_optim (train.py:126)
    net = net.to(dev)
    evaluator = Evaluator(num_class=args.n_cls, selective=args.selective, device=dev)
    trainer = None
    if fused:
        trainer = SUNetTrainer(net, lr=args.lr, s_lamb=args.s_lamb, weight_decay=args.w_decay, process_group=group,
                               world_size=world, evaluator=evaluator)
        optim = trainer.optimizer
    else:
        optim = Adam(net.parameters(), lr=args.lr, weight_decay=args.w_decay)
        loss_A = SL.CrossEntropyLoss()
        loss_S = SL.calc_selective_risk_image
        if world > 1:
            SL.set_data_parallel_group(group, True)
    if ckpt is not None and args.resume_optim:
        optim.load_state_dict(ckpt['optim'])
    # learning-rate schedule: the reference's torch scheduler objects on a holder optimizer (identical on all ranks)
    holder = torch.optim.SGD([torch.zeros(1, requires_grad=True)], lr=optim.param_groups[0]['lr'])
    scheduler = make_scheduler(args, holder)
    data_tr, data_va = _datasets(args, net.input_ch)
    writers = _Writers(log_dir, rank == 0)

    def module_batch(xb, yb, train):
        """UNet / CE variant through the module API; returns [select_loss, coverage, aux_loss, total] on the device."""
        if args.selective:
            output, selection, aux = net(xb)
            aux_loss = loss_A(aux, yb.long())
            select_loss, coverage = loss_S(output, selection, target=yb.long(), lamb=args.s_lamb)
            loss = aux_loss + select_loss
            res = torch.stack([select_loss.detach(), coverage.detach(), aux_loss.detach(), loss.detach()])
        else:
            output, selection = net(xb), None
            loss = loss_A(output, yb.long())
            z = torch.zeros((), device=dev)
            res = torch.stack([z, z, loss.detach(), loss.detach()])
        if train:
            optim.zero_grad()
            loss.backward()
            if world > 1:                       # SUM: the losses already carry the global 1/P (selective_loss.py)
                for p in net.parameters():
                    dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group)
            optim.step()
        # argmax over the two channels == logit difference > 0 (train.py:216-217,224-225)
        d_out = (output[:, 1] - output[:, 0]).detach()
        d_sel = None if selection is None else (selection[:, 1] - selection[:, 0]).detach()
        evaluator.add_batch_from_logits(yb, d_out, d_sel, cut_off=0.0, s_cut_off=0.0, path='train', scale='None')
        return res

    def run_epoch(data, train):
        """One pass; returns (mean [sel, cov, aux, total] over batches, confusion/selection counters) — identical on
        every rank (losses are global-batch quantities, counters are all-reduced)."""
        acc = torch.zeros(4, device=dev)
        nb = 0
        net.train(train)
        ctx = torch.enable_grad() if train else torch.no_grad()
        with ctx:
            for xb, yb in data.batches(args.batch_size):
                if xb.shape[0] < world:      # a ragged tail smaller than the GPU count has no shard for every rank
                    if rank == 0:
                        print(f'(skipping a tail batch of {xb.shape[0]} < {world} GPUs)')
                    continue
                lo, hi = shard_bounds(xb.shape[0], world, rank)
                xs, ys = xb[lo:hi].to(dev, non_blocking=True), yb[lo:hi].to(dev, non_blocking=True)
                if fused:
                    res = trainer.step(xs, ys) if train else trainer.validate(xs, ys)
                else:
                    res = module_batch(xs, ys, train)
                acc += res                                   # stays on the device: no per-step sync
                nb += 1
        counts = evaluator.counts_tensor()
        counts = torch.zeros(6, dtype=torch.int64, device=dev) if counts is None else counts.clone()
        if world > 1:
            dist.all_reduce(counts)
        evaluator.reset()
        return (acc / max(nb, 1)).tolist(), counts.cpu().numpy().astype(np.float64)

    def accuracy(c):                 # Evaluator.get_Pixel_Accuracy on the all-reduced counters
        return (c[0] + c[3]) / c[:4].sum() if c[:4].sum() > 0 else float('nan')

    for epoch in range(start_epoch + 1, start_epoch + args.n_epoch + 1):
        current_lr = holder.param_groups[0]['lr']
        optim.set_lr(current_lr)
        writers.scalar('train', 'lr', current_lr, epoch)
        if rank == 0:
            print(f'epoch {epoch} / {start_epoch + args.n_epoch}, learning rate {current_lr}')
        tr, c_tr = run_epoch(data_tr, True)
        tr_acc = accuracy(c_tr)
        if scheduler is not None:            # every rank steps its own copy with the same (all-reduced) epoch loss
            holder.step()
            if args.lr_sche == 'ReduceLR':
                scheduler.step(tr[3])
            else:
                scheduler.step()
        writers.scalar('train', 'loss', tr[3], epoch)
        writers.scalar('train', 'accuracy', tr_acc, epoch)
        if args.selective:
            writers.scalar('train', 'aux loss', tr[2], epoch)
            writers.scalar('train', 'selection loss', tr[0], epoch)
            writers.scalar('train', 'rejection ratio', (c_tr[5] - c_tr[4]) / max(c_tr[5], 1), epoch)
        va, c_va, val_acc = None, None, float('nan')
        if data_va is not None and len(data_va) > 0:
            va, c_va = run_epoch(data_va, False)
            val_acc = accuracy(c_va)
            writers.scalar('valid', 'loss', va[3], epoch)
            writers.scalar('valid', 'accuracy', val_acc, epoch)
            if args.selective:
                writers.scalar('valid', 'aux loss', va[2], epoch)
                writers.scalar('valid', 'selection loss', va[0], epoch)
                writers.scalar('valid', 'rejection ratio', (c_va[5] - c_va[4]) / max(c_va[5], 1), epoch)
        writers.close()
        if rank == 0:
            print('train_loss %.05f train_acc %.04f | valid_loss %.05f valid_acc %.04f'
                  % (tr[3], tr_acc, va[3] if va else float('nan'), val_acc))
            if args.selective:
                print('train_aux_loss %.05f | train_select_loss %.05f | train_rejection %.03f'
                      % (tr[2], tr[0], (c_tr[5] - c_tr[4]) / max(c_tr[5], 1)))
                if va:
                    print('valid_aux_loss %.05f | valid_select_loss %.05f | valid_rejection %.03f'
                          % (va[2], va[0], (c_va[5] - c_va[4]) / max(c_va[5], 1)))
            net_save(ckpt_dir=ckpt_dir, net=net, optim=optim, epoch=epoch)
    if world > 1:
        from bench import leave_process_group      # graphs that captured NCCL collectives go first (bench.py)
        leave_process_group(torch, dist, dev, [h for h in (trainer, net) if h is not None])


def train(args, ckpt_dir, log_dir=None):
    if (args.model_arch == 'UNet_B') != ('BCE' in args.loss):
        raise SystemExit('--model_arch UNet_B goes with --loss BCElogit (outputs (N,H,W)); --model_arch UNet with '
                         '--loss CE (outputs (N,C,H,W)) — the pairs the reference supports (train.py:70-86)')
    if args.model_arch == 'UNet' and args.n_cls != 2:
        raise SystemExit('UNet: only --n_cls 2 is built (binary segmentation)')
    if args.optim != 'Adam':
        raise SystemExit('only --optim Adam is built (the reference default)')
    if log_dir is None:
        log_dir = os.path.join(os.path.dirname(ckpt_dir), 'log')
    world = len(args.local_rank)
    if args.batch_size < world:
        raise SystemExit(f'--batch_size {args.batch_size} < {world} GPUs: every rank needs at least one patch')
    if world == 1:
        train_worker(0, 1, args, ckpt_dir, log_dir)
    else:
        import torch.multiprocessing as mp
        mp.spawn(train_worker, args=(world, args, ckpt_dir, log_dir), nprocs=world, join=True)


if __name__ == '__main__':
    args = parse_arguments()
    ckpt_dir = f'{args.model_dir}/{args.fold}-fold/checkpoint'
    log_dir = f'{args.model_dir}/{args.fold}-fold/log'
    train(args, ckpt_dir, log_dir)
