timeout 600 python -m pytest tests/test_gpu_resblock.py tests/test_gpu_tail.py tests/test_gpu_config_sizes.py -x -q 2>&1 | tail -3
BARGS="--steps 10 --warmup 3 --no-cpu-baseline --no-latency --no-sessions --no-front --no-strong"
B2_RB_T=0 python bench.py $BARGS > gpurun_out/r3l_old_1.json 2>gpurun_out/r3l.err
python bench.py $BARGS > gpurun_out/r3l_new_1.json 2>>gpurun_out/r3l.err
B2_RB_T=0 python bench.py $BARGS > gpurun_out/r3l_old_2.json 2>>gpurun_out/r3l.err
python bench.py $BARGS > gpurun_out/r3l_new_2.json 2>>gpurun_out/r3l.err
python tools/rb_dbg.py 256 2> gpurun_out/rb_dbg_r3l.txt
for f in gpurun_out/r3l_*.json; do python - $f <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["value"], d["ms_per_step"], d.get("kernel_ms_per_step"), d["clocks"]["sm_mhz"], d["roofline"]["frac"])
PY
done
grep "rbt dbg" gpurun_out/rb_dbg_r3l.txt | tail -20
