run() { # tag, env...
  tag=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --batch 32 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/dp_$tag.log 2>&1
  echo "$tag rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/dp_$tag.log | head -1)"
}
python bench.py --batch 16 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/dp_b16_1gpu.log 2>&1; echo "1gpu b16 $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/dp_b16_1gpu.log | head -1)"
run b7 SUNET_DP_BUCKETS=dec1,dec2,dec3,dec4,enc3,enc2,enc1
run b1 SUNET_DP_BUCKETS=enc1
run b2 SUNET_DP_BUCKETS=dec3,enc1
run b7c4 SUNET_DP_BUCKETS=dec1,dec2,dec3,dec4,enc3,enc2,enc1 NCCL_MAX_CTAS=4
run b2c4 SUNET_DP_BUCKETS=dec3,enc1 NCCL_MAX_CTAS=4
