# round 2, run N: tiled (hash range x sample) join mapping, bootstrap at 6 CTAs/SM, packer prefetch default
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_n_pytest.log
cat gpurun_out/r02_n_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_n_bench.json 2> gpurun_out/r02_n_bench.err
tail -3 gpurun_out/r02_n_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_n_bench.json').read().strip().splitlines()[-1])
p=d['pairs']; g=d['genomes']
print('sketch ms_per_step %.3f e2e %.2f ms %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['per_step_ms']))
print('pairs ms_per_step %.3f value %.3g kernels %s parity %s'%(p['ms_per_step'], p['value'], {k:round(v,4) for k,v in p['kernels_ms_per_step'].items()}, p.get('parity_checked')))
print('pairs roofline', {k:v for k,v in p['roofline'].items() if k not in ('note',)})
print('genomes value %.4g ms_per_step %.3f kernels %s parity %s'%(g['value'], g['ms_per_step'], g['kernels_ms_per_step'], g.get('parity_checked')))
PY
for mode in tiled plain; do
  if [ $mode = plain ]; then export SYL_JOIN_PLAIN=1; else unset SYL_JOIN_PLAIN; fi
  timeout 900 python bench.py --workload profile --samples 16 --steps 10 --warmup 3 > gpurun_out/r02_n_bench_profile16_n1_$mode.json 2> gpurun_out/r02_n_bench_profile16_n1_$mode.err
  python - $mode <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r02_n_bench_profile16_n1_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'profile16 ms_per_step %.3f value %.4g kernels %s parity %s'%(d['ms_per_step'], d['value'], {k:round(v,4) for k,v in d.get('pairs',d).get('kernels_ms_per_step',{}).items()}, d.get('pairs',d).get('parity_checked')))
PY
done
unset SYL_JOIN_PLAIN
timeout 600 python bench.py --workload genomes --batch-genomes 250 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); g=d.get('genomes',d); print('genomes x250 value %.4g ms_per_step %.3f'%(g['value'], g['ms_per_step']))"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_join_hist|k_join2_order|k_boot_iter_p|k_range_bounds' -s 4 -c 4 -o gpurun_out/r02_n_contain16 python scripts/run_query16.py 12500 6666667 2 > gpurun_out/r02_n_ncu_contain16.log 2>&1
tail -3 gpurun_out/r02_n_ncu_contain16.log
