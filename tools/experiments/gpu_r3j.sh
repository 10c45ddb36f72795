set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.load(open('gpurun_out/r2z_bench.json'))
print('value %.4g ms %.1f | e2e %.4g ms %.1f | frac %.4f | launches %d' % (l['value'], l['ms_per_step'], l['e2e']['value'], l['e2e']['ms_per_step'], l['roofline']['frac'], l['gpu_launches']))
for k,v in l['other_configs'].items(): print(k, v.get('ms_per_step'), v.get('value'), v.get('failed'))
print(l['cpu_baseline'])
PY
python tools/rows_timing.py --pick 11 --iters 3000 --ctas 1,16
