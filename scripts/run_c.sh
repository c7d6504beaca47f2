./exp/hotloop_bench > gpurun_out/r02_c_hotloop.log 2>&1
timeout 900 python -m pytest tests/test_packed_gpu.py tests/test_seed_gpu.py tests/test_sketch_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_c_pytest.log
for v in "" b c d; do
  if [ -n "$v" ]; then export SYLPH_B200_LIB=$PWD/sylph_b200/libsylph_b200_$v.so; else unset SYLPH_B200_LIB; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --no-pairs --no-cpu > gpurun_out/r02_c_bench_v$v.json 2> gpurun_out/r02_c_bench_v$v.err
done
unset SYLPH_B200_LIB
timeout 600 python scripts/e2e_probe.py all > gpurun_out/r02_c_e2e_probe.log 2>&1
cat gpurun_out/r02_c_hotloop.log gpurun_out/r02_c_pytest.log gpurun_out/r02_c_e2e_probe.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_c_bench_v*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms_per_step %.3f kernel_ms %.3f e2e %.2f ms %s'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['e2e'].get('per_step_ms')))
    except Exception as e:
        print(f,'ERR',e)
PY
