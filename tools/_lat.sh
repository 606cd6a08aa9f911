for i in 1 2 3; do
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sessions --no-front --no-strong --latency-sessions 5000 > gpurun_out/r2z_lat_$i.json 2>gpurun_out/r2z_lat.err
python - gpurun_out/r2z_lat_$i.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for l in d['latency']:
    print({k:l[k] for k in ('sessions','p50_ms','p99_ms','p999_ms','max_ms','queue_p99_ms','mean_sub_batch','max_sub_batch','graphs_built','steps','loadgen_late_ticks')})
PY
done
