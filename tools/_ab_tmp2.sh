timeout 900 python -m pytest tests/test_gpu_resblock.py tests/test_gpu_tail.py tests/test_gpu_config_sizes.py tests/test_gpu_kernel_variants.py -x -q 2>&1 | tail -3
BARGS="--steps 10 --warmup 3 --no-cpu-baseline --no-latency --no-sessions --no-front --no-strong"
for i in 1 2; do
python bench.py $BARGS > gpurun_out/r3m_new_$i.json 2>>gpurun_out/r3m.err
done
python tools/rb_dbg.py 256 2> gpurun_out/rb_dbg_r3m.txt
for f in gpurun_out/r3m_*.json; do python - $f <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["value"], d["ms_per_step"], d.get("kernel_ms_per_step"), d["clocks"]["sm_mhz"], d["roofline"]["frac"])
PY
done
grep "dbg\] C=\|rbt dbg\] k=\|conv_post epi" gpurun_out/rb_dbg_r3m.txt | tail -12
