./exp/hotloop_bench > gpurun_out/r02_d_hotloop.log 2>&1
timeout 1200 python -m pytest tests/test_packed_gpu.py tests/test_seed_gpu.py tests/test_sketch_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_d_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-pairs --no-cpu > gpurun_out/r02_d_bench.json 2> gpurun_out/r02_d_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file gpurun_out/r02_d_launches_sketch.csv python scripts/run_sketch_steps.py 6666667 3 > gpurun_out/r02_d_ncu.log 2>&1
for t in 16 32 48; do SYL_PACK_THREADS=$t timeout 300 python scripts/e2e_probe.py one; done > gpurun_out/r02_d_e2e_probe.log 2>&1
cat gpurun_out/r02_d_hotloop.log gpurun_out/r02_d_pytest.log gpurun_out/r02_d_e2e_probe.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_d_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms_per_step %.3f kernel_ms %.3f e2e %.2f ms seed-in-e2e %.3f %s'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['e2e']['seed_kernel_ms_per_step'], d['e2e'].get('per_step_ms')))
    except Exception as e:
        print(f,'ERR',e)
PY
