# round 2, run O: bootstrap with register counts for values 4..7 and the pinned rejection test; genome batches of 250
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_o_pytest.log
cat gpurun_out/r02_o_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_o_bench.json 2> gpurun_out/r02_o_bench.err
tail -3 gpurun_out/r02_o_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_o_bench.json').read().strip().splitlines()[-1])
p=d['pairs']; g=d['genomes']
print('sketch ms_per_step %.3f e2e %.2f ms %s h2d %d'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['per_step_ms'], d['e2e']['h2d_bytes_per_step']))
print(d['e2e']['ingest'])
print('pairs ms_per_step %.3f value %.3g kernels %s parity %s'%(p['ms_per_step'], p['value'], {k:round(v,4) for k,v in p['kernels_ms_per_step'].items()}, p.get('parity_checked')))
print('genomes value %.4g ms_per_step %.3f kernels %s parity %s'%(g['value'], g['ms_per_step'], g['kernels_ms_per_step'], g.get('parity_checked')))
PY
timeout 900 python bench.py --workload profile --samples 16 --steps 10 --warmup 3 > gpurun_out/r02_o_bench_profile16_n1.json 2> gpurun_out/r02_o_bench_profile16_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_o_bench_profile16_n1.json').read().strip().splitlines()[-1])
print('profile16 ms_per_step %.3f value %.4g kernels %s parity %s'%(d['ms_per_step'], d['value'], {k:round(v,4) for k,v in d.get('pairs',d).get('kernels_ms_per_step',{}).items()}, d.get('pairs',d).get('parity_checked')))
PY
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_boot_iter_p --csv --log-file gpurun_out/r02_o_boot.csv python scripts/run_query_steps.py 2000 6666667 2 > gpurun_out/r02_o_ncu_boot.log 2>&1
grep k_boot_iter_p gpurun_out/r02_o_boot.csv | tail -2 | cut -c1-60,200-
