timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r2z_gpu_tests.log; cat gpurun_out/r2z_gpu_tests.log
python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; tail -c 600 gpurun_out/r2z_bench_n1.err
python bench.py --impl reference > gpurun_out/r2z_bench_reference_arm.json 2>> gpurun_out/r2z_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','roofline','kernel_ms_per_step','clocks','gpu_launches') if k in d})
print(d.get('latency'))
print(d.get('front_half'))
r=json.loads(open('gpurun_out/r2z_bench_reference_arm.json').read().strip().splitlines()[-1])
print(r.get('value'), r.get('cpu_baseline'))
PY
