run() { # tag, env...
  tag=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/dp8_$tag.log 2>&1
  echo "$tag rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/dp8_$tag.log | head -1)"
}
run b7 SUNET_DP_BUCKETS=dec1,dec2,dec3,dec4,enc3,enc2,enc1
run b1 SUNET_DP_BUCKETS=enc1
run b2 SUNET_DP_BUCKETS=dec3,enc1
