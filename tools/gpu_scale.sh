# bench.py on N GPUs of one box, launched like the driver does: tools/gpu_scale.sh <tag> <N> [extra bench flags]
set -u
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 "$@" > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > gpurun_out/${TAG}_n${N}.json 2> gpurun_out/${TAG}_n${N}.err
fi
echo "rc=$?"; tail -2 gpurun_out/${TAG}_n${N}.err
python - <<PY
import json
for ln in open('gpurun_out/${TAG}_n${N}.json'):
    if ln.startswith('{'):
        l=json.loads(ln); print('N=%d: %.1f ms/step (device), %.1f ms/step (e2e), %.4g draws/s, launches %d' % (l['n_gpus'], l['ms_per_step'], l['e2e']['ms_per_step'], l['value'], l['gpu_launches']))
PY
