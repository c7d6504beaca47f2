cat /sys/fs/cgroup/cpu.max
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_g_pytest.log
cat gpurun_out/r02_g_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_g_bench.json 2> gpurun_out/r02_g_bench.err
tail -3 gpurun_out/r02_g_bench.err
SYL_DEBUG_TIMING=1 timeout 300 python scripts/e2e_probe.py one 2>&1 | grep -v "post-pass" | tail -8
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_g_bench.json').read().strip().splitlines()[-1])
p=d['pairs']; g=d['genomes']
print('sketch ms_per_step %.3f e2e %.2f ms %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['per_step_ms']))
print('pairs ms_per_step %.3f value %.3g kernels %s parity %s'%(p['ms_per_step'], p['value'], {k:round(v,4) for k,v in p['kernels_ms_per_step'].items()}, p.get('parity_checked')))
print('genomes value %.4g ms_per_step %.3f kernels %s parity %s cpu %s'%(g['value'], g['ms_per_step'], g['kernels_ms_per_step'], g.get('parity_checked'), g.get('cpu_baseline')))
PY
ncu --set full --clock-control none --import-source on -k regex:k_boot_iter_p -c 2 -o gpurun_out/r02_g_boot python scripts/run_query_steps.py 2000 6666667 3 > gpurun_out/r02_g_ncu_boot.log 2>&1
