timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_h_pytest.log
cat gpurun_out/r02_h_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_h_bench.json 2> gpurun_out/r02_h_bench.err
tail -3 gpurun_out/r02_h_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_h_bench.json').read().strip().splitlines()[-1])
p=d['pairs']; g=d['genomes']
print('sketch ms_per_step %.3f e2e %.2f ms %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['per_step_ms']))
print('pairs ms_per_step %.3f value %.3g kernels %s parity %s'%(p['ms_per_step'], p['value'], {k:round(v,4) for k,v in p['kernels_ms_per_step'].items()}, p.get('parity_checked')))
print('genomes value %.4g ms_per_step %.3f kernels %s parity %s'%(g['value'], g['ms_per_step'], g['kernels_ms_per_step'], g.get('parity_checked')))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file gpurun_out/r02_h_launches_genomes.csv python bench.py --workload genomes --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_h_ncu_genomes.log 2>&1
