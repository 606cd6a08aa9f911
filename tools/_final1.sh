timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r2z_gpu_tests.log; cat gpurun_out/r2z_gpu_tests.log
python bench.py --impl reference > gpurun_out/r2z_bench_reference_arm.json 2> gpurun_out/r2z_bench_n1.err
python bench.py > gpurun_out/r2z_bench_n1.json 2>> gpurun_out/r2z_bench_n1.err; tail -c 300 gpurun_out/r2z_bench_n1.err
python tools/sweeps.py > gpurun_out/r2z_sweeps.log 2>&1; ls gpurun_out/*sweeps*json | tail -2
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2z_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','kernel_ms_per_step','clocks','gpu_launches') if k in d})
print(d['e2e']['value'], d['roofline']['frac'], d['roofline']['achieved'], d.get('roofline_codec',{}).get('frac'))
for l in d['latency']: print({k:l[k] for k in ('sessions','p50_ms','p99_ms','max_ms','graphs_built_while_serving','met')})
print(d.get('front_half',{}).get('streams_front_plus_tail'))
print(d.get('strong'))
r=json.loads(open('gpurun_out/r2z_bench_reference_arm.json').read().strip().splitlines()[-1])
print(r.get('value'), r.get('cpu_baseline'))
PY
