"""Data-parallel parity on real GPUs (run with torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/dp_parity.py

Every rank runs one fused SUNetTrainer step on its torch.chunk shard of a seeded global batch; rank 0 then
checks the all-reduced losses and gradients against the CPU oracle emulating the reference's
nn.DataParallel semantics: per-replica BatchNorm statistics, loss on the gathered global
batch, gradients summed over replicas.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist

from oracle import sunet_oracle as O


def main():
    from selectivenet_for_semantic_segmentation_binary_b200.model import UNet_B
    from selectivenet_for_semantic_segmentation_binary_b200.trainer import SUNetTrainer, chunk_bounds
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import datetime
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    total, size = 2 * world + 1, 64                      # uneven shards on purpose
    x, label = O.synthetic_batch(total, size, seed=11)
    torch.manual_seed(0)
    net = UNet_B("RGB", selective=True).to(dev)
    net.train()
    tr = SUNetTrainer(net, lr=0.0, s_lamb=2, process_group=dist.group.WORLD, world_size=world, use_cuda_graph=False)
    lo, hi = chunk_bounds(total, world, rank)
    # uneven shards: the trainer's P_global = P_local * world assumes equal shards, so pass the true count
    tr.global_pixels_override = total * size * size
    res = tr.step(x[lo:hi].to(dev), label[lo:hi].to(dev))
    torch.cuda.synchronize()
    flat = tr.fg.flat.clone()
    ok = True
    if rank == 0:
        sd = O.init_state_dict(0, "RGB", True)
        names = [n for n, _ in net.named_parameters()]
        for n in names:
            sd[n].requires_grad_(True)
        outs = []
        for r in range(world):                            # one replica per shard, own BN statistics
            a, b = chunk_bounds(total, world, r)
            outs.append(O.unet_b_forward(sd, x[a:b], True, True, update_running=(r == 0)))
        out, sel, aux = (torch.cat([o[i] for o in outs]) for i in range(3))
        l_sel, cov = O.selective_risk_b(out, sel, label, lamb=2)
        loss = l_sel + O.bce_with_logits_mean(aux, label)
        loss.backward()
        got = res.cpu().tolist()
        print(f"loss: oracle {loss.item():.6f} ours {got[3]:.6f} | coverage: oracle {cov.item():.6f} ours {got[1]:.6f}")
        ok &= abs(got[3] - loss.item()) / abs(loss.item()) < 2e-2 and abs(got[1] - cov.item()) / cov.item() < 2e-2
        worst = 1.0
        for n in names:
            if n.endswith(".0.bias"):
                continue
            o, k = tr.fg.offsets[n]
            g = flat[o:o + k].cpu().double()
            r_ = sd[n].grad.flatten().double()
            c = (g @ r_ / (g.norm() * r_.norm()).clamp_min(1e-30)).item()
            worst = min(worst, c)
        print(f"worst gradient cosine vs DataParallel-emulating oracle: {worst:.4f}")
        ok &= worst > 0.90
        print("DP PARITY", "PASS" if ok else "FAIL")
    # every rank must hold identical (all-reduced) gradients
    chk = flat.double().sum().reshape(1)
    lst = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    if rank == 0:
        same = all(torch.equal(lst[0], t) for t in lst)
        print("gradients identical on all ranks:", same)
        ok &= same
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
