timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_i_pytest.log
cat gpurun_out/r02_i_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_i_bench.json 2> gpurun_out/r02_i_bench.err
tail -3 gpurun_out/r02_i_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_i_bench.json').read().strip().splitlines()[-1])
p=d['pairs']; g=d['genomes']
print('sketch ms_per_step %.3f e2e %.2f ms %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['per_step_ms']))
print('pairs ms_per_step %.3f value %.3g kernels %s parity %s'%(p['ms_per_step'], p['value'], {k:round(v,4) for k,v in p['kernels_ms_per_step'].items()}, p.get('parity_checked')))
print('genomes value %.4g ms_per_step %.3f kernels %s parity %s'%(g['value'], g['ms_per_step'], g['kernels_ms_per_step'], g.get('parity_checked')))
PY
timeout 900 python bench.py --workload profile --samples 16 --steps 10 --warmup 3 > gpurun_out/r02_i_bench_profile16_n1.json 2> gpurun_out/r02_i_bench_profile16_n1.err
tail -2 gpurun_out/r02_i_bench_profile16_n1.err; head -c 1500 gpurun_out/r02_i_bench_profile16_n1.json
