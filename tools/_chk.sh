timeout 900 python -m pytest tests/test_gpu_sched.py tests/test_gpu_resblock.py tests/test_gpu_kernel_variants.py tests/test_gpu_config_sizes.py -x -q 2>&1 | tail -4
BARGS="--steps 10 --warmup 3 --no-cpu-baseline --no-sessions --no-front --no-strong"
for i in 1 2; do
python bench.py $BARGS > gpurun_out/r3n_$i.json 2>>gpurun_out/r3n.err
python - gpurun_out/r3n_$i.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("kernel_ms_per_step"), d["clocks"]["sm_mhz"], d["roofline"]["frac"])
for l in d['latency']:
    print({k:l[k] for k in ('sessions','p50_ms','p99_ms','max_ms','max_sub_batch','graphs_built','graphs_built_while_serving','loadgen_late_ticks','met')})
PY
done
