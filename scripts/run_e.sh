timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_e_pytest.log
cat gpurun_out/r02_e_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_e_bench.json 2> gpurun_out/r02_e_bench.err
tail -3 gpurun_out/r02_e_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_e_bench.json').read().strip().splitlines()[-1])
p=d['pairs']
print('sketch ms_per_step %.3f e2e %.2f ms | pairs ms_per_step %.3f value %.3g kernels %s parity %s rows %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], p['ms_per_step'], p['value'], {k:round(v,4) for k,v in p['kernels_ms_per_step'].items()}, p.get('parity_checked'), p['workload_stats']['rows_per_step']))
PY
