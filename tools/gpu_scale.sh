set -u
mkdir -p gpurun_out
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 2 --warmup 1 > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err
  echo "N=$N rc=$?"; cut -c1-330 gpurun_out/r2_scale_n$N.json; tail -2 gpurun_out/r2_scale_n$N.err | cut -c1-300
done
