cat /sys/fs/cgroup/cpu.max /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null | head -3 > gpurun_out/r02_f_cgroup.log; nproc >> gpurun_out/r02_f_cgroup.log
timeout 900 python -m pytest tests/test_sketch_gpu.py -m gpu -x -q -k "genomes" 2>&1 | tail -6 > gpurun_out/r02_f_pytest_genomes.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/r02_f_dist_check_2gpu.log 2>&1
echo "dist_check rc=$?" >> gpurun_out/r02_f_dist_check_2gpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_f_bench_n2.json 2> gpurun_out/r02_f_bench_n2.err
cat gpurun_out/r02_f_cgroup.log gpurun_out/r02_f_pytest_genomes.log; grep -v "^$" gpurun_out/r02_f_dist_check_2gpu.log | tail -12; tail -5 gpurun_out/r02_f_bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_f_bench_n2.json').read().strip().splitlines()[-1])
    p=d['pairs']
    print('N=2 sketch value %.3g ms_per_step %.3f e2e %.2f ms | pairs ms_per_step %.3f value %.3g kernels %s parity %s stats %s'%(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], p['ms_per_step'], p['value'], {k:round(v,4) for k,v in p['kernels_ms_per_step'].items()}, p.get('parity_checked'), p['workload_stats']))
except Exception as e: print('ERR',e)
PY
