set -u
TAG=r2z
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/${TAG}_smoke.log; tail -1 $OUT/${TAG}_smoke.log
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "bench ref rc=$?"
python - <<'PY'
import json
l=json.load(open('gpurun_out/r2z_bench.json')); r=json.load(open('gpurun_out/r2z_bench_ref.json'))
print('value %.4g ms %.1f | e2e %.4g ms %.1f | frac %.4f | launches %d | ref %.4g (%d cores) | e2e/ref %.1f' % (l['value'], l['ms_per_step'], l['e2e']['value'], l['e2e']['ms_per_step'], l['roofline']['frac'], l['gpu_launches'], r['value'], r['cpu_baseline']['cores'], l['e2e']['value']/r['value']))
for k,v in l['other_configs'].items(): print(k, v.get('ms_per_step'), v.get('value'), v.get('failed'), (v.get('cpu_baseline') or {}).get('value'))
print(l['cpu_baseline'])
PY
