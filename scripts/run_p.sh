# round 2, final 8-GPU run: NCCL parity of the sharded paths, then the driver-style bench line (N=8)
cat /sys/fs/cgroup/cpu.max
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 scripts/dist_check.py > gpurun_out/r02_p_dist_check_8gpu.log 2>&1
echo "dist_check rc=$?" >> gpurun_out/r02_p_dist_check_8gpu.log
grep "equal\|rc=" gpurun_out/r02_p_dist_check_8gpu.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_p_bench_n8.json 2> gpurun_out/r02_p_bench_n8.err
tail -3 gpurun_out/r02_p_bench_n8.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_p_bench_n8.json').read().strip().splitlines()[-1])
    p=d['pairs']; g=d['genomes']
    print('N=8 sketch value %.4g ms_per_step %.3f e2e %.2f ms (%.4g) h2d %d'%(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e']['h2d_bytes_per_step']))
    print(d['e2e']['ingest'])
    print('pairs ms_per_step %.3f value %.4g kernels %s parity %s'%(p['ms_per_step'], p['value'], {k:round(v,4) for k,v in p['kernels_ms_per_step'].items()}, p.get('parity_checked')))
    print(p.get('parity_detail')); print(p['workload_stats'])
    print('genomes value %.4g ms_per_step %.3f'%(g['value'], g['ms_per_step']))
except Exception as e: print('ERR',e)
PY
