set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2y_bench.err
python - <<'PY'
import json
l=json.load(open('gpurun_out/r2y_bench.json'))
print('value %.4g ms %.1f e2e %.4g e2e_ms %.1f frac %.4f' % (l['value'], l['ms_per_step'], l['e2e']['value'], l['e2e']['ms_per_step'], l['roofline']['frac']))
print(l['e2e']['host_phases_s_rank0'])
for k,v in l.get('other_configs',{}).items(): print(k, {a:b for a,b in v.items() if a in ('ms_per_step','value','failed','loci_measured')})
print(l.get('cpu_baseline'))
PY
