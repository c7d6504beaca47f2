# round 2, run K: full GPU suite, default bench, 16-sample profile at N=1, ncu captures of the round-2 kernels
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_k_pytest.log
cat gpurun_out/r02_k_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_k_bench.json 2> gpurun_out/r02_k_bench.err
tail -3 gpurun_out/r02_k_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_k_bench.json').read().strip().splitlines()[-1])
p=d['pairs']; g=d['genomes']
print('sketch ms_per_step %.3f e2e %.2f ms %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['per_step_ms']))
print('pairs ms_per_step %.3f value %.3g kernels %s parity %s'%(p['ms_per_step'], p['value'], {k:round(v,4) for k,v in p['kernels_ms_per_step'].items()}, p.get('parity_checked')))
print('genomes value %.4g ms_per_step %.3f kernels %s parity %s'%(g['value'], g['ms_per_step'], g['kernels_ms_per_step'], g.get('parity_checked')))
PY
timeout 900 python bench.py --workload profile --samples 16 --steps 10 --warmup 3 > gpurun_out/r02_k_bench_profile16_n1.json 2> gpurun_out/r02_k_bench_profile16_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_k_bench_profile16_n1.json').read().strip().splitlines()[-1])
print('profile16 ms_per_step %.3f value %.4g kernels %s'%(d['ms_per_step'], d['value'], {k:round(v,4) for k,v in d.get('kernels_ms_per_step',{}).items()}))
PY
# ncu: launch lists (per-launch times, cold cache, serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file gpurun_out/r02_k_launches_profile.csv python scripts/run_query_steps.py 2000 6666667 2 > gpurun_out/r02_k_ncu_profile.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file gpurun_out/r02_k_launches_genomes.csv python bench.py --workload genomes --steps 2 --warmup 3 --no-cpu --fixed-warmup > gpurun_out/r02_k_ncu_genomes.log 2>&1
# ncu --set full: bootstrap + pass-2 kernels of the device-driven profile, genome seeding + post-pass
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_boot_iter_p|k_join2_order|k_join_hist' -s 6 -c 6 -o gpurun_out/r02_k_contain python scripts/run_query_steps.py 2000 6666667 2 > gpurun_out/r02_k_ncu_contain.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_seed|k_tile_sort_compact|k_genome_dups' -s 6 -c 6 -o gpurun_out/r02_k_genomes python bench.py --workload genomes --steps 2 --warmup 3 --no-cpu --fixed-warmup > gpurun_out/r02_k_ncu_genomes_full.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
