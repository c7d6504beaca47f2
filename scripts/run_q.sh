# round 2, final 1-GPU run: smoke, GPU suite, both bench arms as the driver runs them, launch list + ncu captures of the same code
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_q_pytest.log
cat gpurun_out/r02_q_pytest.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_q_ref_sketch.json 2> gpurun_out/r02_q_ref_sketch.err
timeout 900 python bench.py --impl reference --workload profile --steps 3 --warmup 1 > gpurun_out/r02_q_ref_profile.json 2> gpurun_out/r02_q_ref_profile.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_q_bench.json 2> gpurun_out/r02_q_bench.err
tail -3 gpurun_out/r02_q_bench.err
python - <<'PY'
import json
for f in ('r02_q_ref_sketch','r02_q_ref_profile'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['metric'], '%.4g'%d['value'], d['unit'], d['cpu_baseline']['cores'], 'cores', d.get('ms_per_step'))
    except Exception as e: print(f,'ERR',e)
d=json.loads(open('gpurun_out/r02_q_bench.json').read().strip().splitlines()[-1])
p=d['pairs']; g=d['genomes']
print('sketch ms_per_step %.3f value %.4g e2e %.2f ms %s h2d %d'%(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['per_step_ms'], d['e2e']['h2d_bytes_per_step']))
print('pairs ms_per_step %.3f value %.3g kernels %s parity %s'%(p['ms_per_step'], p['value'], {k:round(v,4) for k,v in p['kernels_ms_per_step'].items()}, p.get('parity_checked')))
print('genomes value %.4g ms_per_step %.3f kernels %s parity %s'%(g['value'], g['ms_per_step'], g['kernels_ms_per_step'], g.get('parity_checked')))
PY
timeout 900 python bench.py --workload profile --samples 16 --steps 10 --warmup 3 > gpurun_out/r02_q_bench_profile16_n1.json 2> gpurun_out/r02_q_bench_profile16_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 1500 --csv --log-file gpurun_out/r02_q_launches_bench.csv python bench.py --steps 2 --warmup 1 --fixed-warmup --no-cpu > gpurun_out/r02_q_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_seed -s 2 -c 1 -o gpurun_out/r02_q_k_seed python scripts/time_seed.py 6666667 > gpurun_out/r02_q_ncu_seed.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_boot_iter_p|k_join_hist|k_join2_order' -s 6 -c 3 -o gpurun_out/r02_q_contain python scripts/run_query_steps.py 2000 6666667 2 > gpurun_out/r02_q_ncu_contain.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_seed|k_dups_insert|k_dups_flag|k_tile_compact' -s 4 -c 4 -o gpurun_out/r02_q_genomes python bench.py --workload genomes --steps 2 --warmup 3 --no-cpu --fixed-warmup > gpurun_out/r02_q_ncu_genomes.log 2>&1
ls -la gpurun_out/r02_q_*.ncu-rep
