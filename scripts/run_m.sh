# 8 GPUs: e2e of the sketch step, hybrid packed/ASCII ingest vs ASCII only (host memory system shared by 8 ranks)
cat /sys/fs/cgroup/cpu.max
PORT=29520
for mode in default ascii; do
  PORT=$((PORT+1))
  if [ $mode = ascii ]; then export SYL_HOST_INGEST=ascii; else unset SYL_HOST_INGEST; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 --steps 10 --warmup 3 --no-pairs --no-cpu > gpurun_out/r02_m_bench_n8_$mode.json 2> gpurun_out/r02_m_bench_n8_$mode.err
  python - $mode <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r02_m_bench_n8_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'N=8 value %.4g ms_per_step %.3f e2e %.2f ms (%.4g) %s'%(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e']['per_step_ms']))
except Exception as e: print('ERR', e)
PY
done
