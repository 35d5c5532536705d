#!/bin/bash
# round-2 2-GPU CLI pass: train.py / eval.py with --local_rank 0 1 (NCCL), both model variants
set -u
mkdir -p gpurun_out
rm -rf /tmp/m2 /tmp/m2ce
timeout 600 python train.py --model_arch UNet_B --selective 1 --s_lamb 2 --loss BCElogit --batch_size 10 --n_epoch 3 \
  --patch_size 64 --local_rank 0 1 --model_dir /tmp/m2 --synthetic 45 --lr_sche ReduceLR --patience 0 --factor 0.5 \
  > gpurun_out/r2n_train.log 2>&1
echo "train rc=$?"; grep -E "epoch|train_loss|valid_aux|skipping|Error|error" gpurun_out/r2n_train.log | head -20
ls /tmp/m2/1-fold/checkpoint
timeout 600 python train.py --model_arch UNet_B --selective 1 --s_lamb 2 --loss BCElogit --batch_size 10 --n_epoch 1 \
  --patch_size 64 --local_rank 0 1 --model_dir /tmp/m2 --synthetic 20 --resume_optim 1 > gpurun_out/r2n_resume.log 2>&1
echo "resume rc=$?"; grep -E "Load weights|epoch|train_loss" gpurun_out/r2n_resume.log | head
rm -f /tmp/m2/1-fold/checkpoint/model_epoch1.pth /tmp/m2/1-fold/checkpoint/model_epoch2.pth /tmp/m2/1-fold/checkpoint/model_epoch3.pth
timeout 600 python eval.py --model_dir /tmp/m2/1-fold/checkpoint --selective 1 --select_eval 1 --batch_size 8 --patch_size 64 \
  --local_rank 0 1 --synthetic 70 > gpurun_out/r2n_eval.log 2>&1
echo "eval rc=$?"; tail -9 gpurun_out/r2n_eval.log
timeout 600 python train.py --model_arch UNet --loss CE --selective 1 --batch_size 8 --n_epoch 2 --patch_size 32 \
  --local_rank 0 1 --model_dir /tmp/m2ce --synthetic 16 > gpurun_out/r2n_train_ce.log 2>&1
echo "train CE rc=$?"; grep -E "epoch|train_loss|Error|error" gpurun_out/r2n_train_ce.log | head
