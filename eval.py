"""eval.py with the reference's command line (/root/reference/eval.py:16-57), B200-native underneath.

    python3 eval.py --model_dir ./model --selective 1 --select_eval 1 --local_rank 0 1 2 3 4 5 6 7 --synthetic 10000

Forward in eval mode (BatchNorm folded from running statistics into the conv epilogues), thresholding with the
float32 numpy semantics of eval.py:175-179 (``--cut_off``, ``--s_cut_off``, ``--single_scale``) and the
coverage-masked confusion matrix, all on the GPU; patches are sharded across the ids of ``--local_rank`` (one process
per GPU) and the six integer counters are all-reduced once at the end.

Every ``*.pth`` in ``--model_dir`` is loaded (eval.py:116-123).  One checkpoint: the single-model branch
(eval.py:198-206).  Several: the ensemble branch (eval.py:208-222) — each net's output map is re-scaled with
``--ens_scale`` (None / clip / minmax over the batch tensor / sigmoid), the maps are averaged in numpy's float32 order
by one kernel (``sunet_ensemble_mean``) and the mean is thresholded like a single output.  As in the reference the
ensemble branch has no selection ("selective 불가", eval.py:208).  ``--single_scale``: only ``sigmoid`` changes the
decision (eval.py:232-233); ``None`` / ``clip`` / ``minmax`` compare the raw value with the cut-off, exactly like the
reference, whose main loop never applies the clip / minmax lambdas.
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
    parser.add_argument('--data_dir', type=str, default='./data')
    parser.add_argument('--test_fold', type=int, default=1, help='which fold in 5-fold cv')
    parser.add_argument('--input_type', type=str, default='RGB')
    parser.add_argument('--patch_mag', type=int, default=200)
    parser.add_argument('--patch_size', type=int, default=256)
    parser.add_argument('--n_cls', type=int, default=2)
    parser.add_argument('--no_graph', action='store_true', help='issue every batch eagerly (no CUDA graph)')
    parser.add_argument('--batch_size', type=int, default=16)
    parser.add_argument('--num_workers', type=int, default=16, help='Dataloader num_workers')
    parser.add_argument('--model_dir', type=str, default='*/model', help='network ckpt (.pth) directory')
    parser.add_argument('--model_arch', type=str, nargs='+', default=['UNet_B'], choices=['UNet_B'])
    parser.add_argument('--selective', type=bool, default=False, help='Is the network based on SelectiveNet?')
    parser.add_argument('--select_eval', type=bool, default=False, help='calculate metrics with/without selection')
    parser.add_argument('--output_dim', type=str, default='NHW', choices=['NCHW', 'NHW'])
    parser.add_argument('--single_scale', type=str, default='sigmoid', choices=['None', 'clip', 'sigmoid', 'minmax'])
    parser.add_argument('--ens_scale', type=str, default='None', choices=['None', 'clip', 'sigmoid', 'minmax'])
    parser.add_argument('--cut_off', type=float, default=0.5, help='prob > cut_off -> pred: 1')
    parser.add_argument('--s_cut_off', type=float, default=0.5, help='selection > cut_off -> select: 1')
    parser.add_argument('--local_rank', type=int, nargs='+', default=[0], help='local gpu ids')
    parser.add_argument('--info_print', type=bool, default=False)
    parser.add_argument('--save_dir', type=str, default='./output', help='saving results')
    # additions
    parser.add_argument('--synthetic', type=int, default=0, help='evaluate this many synthetic patches')
    parser.add_argument('--random_init', type=int, default=0,
                        help='no checkpoint: this many seeded random-weight models (1 = single model, >1 = ensemble)')
    parser.add_argument('--master_port', type=int, default=29534)
    args = parser.parse_args(argv)
    print('')
    print('args={}\n'.format(args))
    return args


def load_nets(args, dev):
    """eval.py:116-157: one net per checkpoint in --model_dir (sorted), eval mode."""
    from selectivenet_for_semantic_segmentation_binary_b200.model import UNet_B
    from selectivenet_for_semantic_segmentation_binary_b200.utils.net_utils import net_test_load
    nets = []
    if args.random_init:
        for i in range(int(args.random_init)):
            torch.manual_seed(i)
            nets.append(UNet_B(args.input_type, selective=args.selective))
    else:
        model_list = sorted([c for c in os.listdir(args.model_dir) if 'pth' in c])
        if not model_list:
            raise SystemExit(f'no checkpoint (*.pth) in {args.model_dir}')
        for name in model_list:
            model_path = os.path.join(args.model_dir, name)
            if args.info_print:
                print(f'    {model_path} - UNet_B / SelectiveNet: {args.selective}')
            net = UNet_B(args.input_type, selective=args.selective)
            nets.append(net_test_load(model_path, net, device='cpu'))
    nets = [n.to(dev) for n in nets]
    for n in nets:
        n.train(False)
        n._plans = nets[0]._plans        # same shapes: one set of activation buffers serves every member
    return nets


def eval_worker(rank, world, args, ret=None):
    import torch.distributed as dist
    from selectivenet_for_semantic_segmentation_binary_b200 import kernels as K
    from selectivenet_for_semantic_segmentation_binary_b200.trainer import chunk_bounds
    from selectivenet_for_semantic_segmentation_binary_b200.utils.compute_metric import Evaluator

    gpu = args.local_rank[rank]
    torch.cuda.set_device(gpu)
    dev = torch.device('cuda', gpu)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', str(args.master_port))
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    if args.select_eval and not args.selective:
        raise SystemExit('--select_eval needs --selective')
    nets = load_nets(args, dev)
    ensemble = len(nets) > 1
    if ensemble and args.selective:
        raise SystemExit('the ensemble branch has no selection (eval.py:208 "ensemble, selective 불가"): '
                         'use one checkpoint with --selective, or non-selective checkpoints for an ensemble')
    evaluator = Evaluator(num_class=args.n_cls, selective=args.select_eval, device=dev)
    ws = K.new_workspace(dev)

    in_ch, size, bs = nets[0].input_ch, args.patch_size, args.batch_size
    data = None
    if args.synthetic > 0:
        n_total = args.synthetic
    else:
        from selectivenet_for_semantic_segmentation_binary_b200.utils.data_utils import PatchArrays, construct_test
        if not os.path.exists(f'{args.data_dir}/{args.test_fold}-fold_tumorable_data.npy'):
            raise SystemExit(f'no fold lists under {args.data_dir}: pass --synthetic N to evaluate synthetic patches')
        test_list = construct_test(args.data_dir, test_fold=args.test_fold)
        data = PatchArrays(args.data_dir, test_list, args.patch_mag, args.patch_size, args.input_type, train=False)
        n_total = len(data)
        if args.info_print:
            print(f'    # of test dataset {n_total}')
    lo, hi = chunk_bounds(n_total, world, rank)
    print("Model Prediction...") if rank == 0 else None
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ens_buf = {}

    def run_batch(x, label):
        if not ensemble:
            if args.selective:
                output, selection, _ = nets[0](x)
            else:
                output, selection = nets[0](x), None
        else:
            # eval.py:209-222: re-scale every member's map, mean over members in float32 (one kernel)
            n = x.shape[0]
            if n not in ens_buf:
                ens_buf[n] = ([torch.empty(n, size, size, device=dev) for _ in nets],
                              [torch.empty(2, device=dev) for _ in nets], torch.empty(n, size, size, device=dev))
            maps, mm, mean = ens_buf[n]
            for net, m, e in zip(nets, maps, mm):
                m.copy_(net(x))
                if args.ens_scale == 'minmax':
                    K.minmax_f32(m, e, ws)
            K.ensemble_mean(maps, args.ens_scale, mean, minmax=mm if args.ens_scale == 'minmax' else None)
            output, selection = mean, None
        evaluator.add_batch_from_logits(label, output, selection if args.select_eval else None,
                                        cut_off=args.cut_off, s_cut_off=args.s_cut_off, path='eval',
                                        scale=args.single_scale)

    def make_batch(start, n):
        g = torch.Generator(device=dev).manual_seed(1000 + start)           # batch content depends only on its position
        x = torch.rand(n, in_ch, size, size, generator=g, device=dev) * 2 - 1
        label = (torch.rand(n, size, size, generator=g, device=dev) < 0.4).to(torch.uint8)
        return x, label

    def real_batches():
        idx = list(range(lo, hi))
        for i in range(0, len(idx), bs):
            xs, ys = zip(*(data._load(j) for j in idx[i:i + bs]))
            yield (torch.from_numpy(np.stack(xs)).to(dev), torch.from_numpy(np.stack(ys)).to(dev))

    # Full batches replay one CUDA graph (forward + thresholding + masked counting: ~65 launches per member that
    # would otherwise be issued from Python per batch); the ragged tail, if any, runs eagerly.
    graph, sx, sl = None, None, None
    n_full = (hi - lo) // bs
    with torch.no_grad():
        if n_full >= 2 and not args.no_graph:
            sx = torch.zeros(bs, in_ch, size, size, device=dev)
            sl = torch.zeros(bs, size, size, dtype=torch.uint8, device=dev)
            nets[0](sx)                                                      # plan allocation, not timed / not counted
            saved = evaluator._ensure(dev).clone()
            run_batch(sx, sl)                                                # lazy one-time work outside capture
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                run_batch(sx, sl)
            evaluator.counts_tensor().copy_(saved)                           # the warm-up batch was not data
        elif hi > lo:
            nets[0](torch.zeros(min(bs, hi - lo), in_ch, size, size, device=dev))
    torch.cuda.synchronize(dev)
    t0.record()
    with torch.no_grad():
        if data is None:
            batches = (make_batch(start, min(bs, hi - start)) for start in range(lo, hi, bs))
        else:
            batches = real_batches()
        for x, label in batches:
            if graph is not None and x.shape[0] == bs:
                sx.copy_(x)
                sl.copy_(label)
                graph.replay()
            else:
                run_batch(x, label)
    t1.record()
    torch.cuda.synchronize(dev)
    secs = t0.elapsed_time(t1) / 1e3
    if rank == 0:       # forward + thresholding + masked confusion-matrix counting, synthetic batches made on the device
        print(f'    throughput: {round((hi - lo) / max(secs, 1e-9), 1)} patches/s per GPU ({hi - lo} patches of '
              f'{size}x{size} in {round(secs, 3)} s, x{world} GPUs, {len(nets)} model(s))')
    counts = evaluator._ensure(dev).clone()
    if world > 1:
        dist.all_reduce(counts)
        dist.destroy_process_group()
    if rank == 0:
        c = counts.cpu().numpy()
        final = Evaluator(num_class=args.n_cls, selective=args.select_eval)
        final.confusion_matrix = c[:4].reshape(2, 2).astype(np.float64)
        CM = final.Confusion_Matrix()
        Acc = final.get_Pixel_Accuracy()
        Acc_class = final.get_Pixel_Accuracy_Class()
        Prec = final.get_Precision()
        Recall = final.get_Recall()
        F1_Score = final.get_F1_Score(Prec, Recall)
        mIoU = final.get_mIoU()
        IoU_class = final.get_IoU_Class()
        if args.select_eval:
            print(f'    rejection ratio: {round(float(c[5] - c[4]) / float(c[5]), 3)}')
        print(f'    Acc:{Acc}')
        print(f'    Acc_class:{Acc_class}')
        print(f'    Prec:{Prec}, Recall:{Recall}, F1_Score:{F1_Score}')
        print(f'    mIoU:{mIoU}')
        print(f'    IoU_class:{IoU_class}')
        if ret is not None:
            ret['counts'] = c.tolist()
            ret['patches_per_s_per_gpu'] = (hi - lo) / max(secs, 1e-9)
        return CM


def main(argv=None, ret=None):
    args = parse_arguments(argv)
    world = len(args.local_rank)
    if world == 1:
        return eval_worker(0, 1, args, ret)
    import torch.multiprocessing as mp
    mp.spawn(eval_worker, args=(world, args), nprocs=world, join=True)


if __name__ == '__main__':
    main()
