# 8 GPUs: does the N=8 sketch step still jitter with the in-process NVML clock sampler?  (two back-to-back runs)
for i in 1 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$i bench.py --gpus 8 --steps 20 --warmup 5 --no-pairs --no-cpu > gpurun_out/r02_r_bench_n8_$i.json 2> gpurun_out/r02_r_bench_n8_$i.err
  python - $i <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r02_r_bench_n8_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print('run', sys.argv[1], 'N=8 value %.4g ms_per_step %.3f kernel %.3f e2e %.2f ms'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step']), d['clocks'])
except Exception as e: print('ERR', e)
PY
done
tail -2 gpurun_out/r02_r_bench_n8_2.err
