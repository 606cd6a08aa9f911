for nr in 0 1; do
B2_RB_NORING=$nr timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:"^k_resblock$" --csv --log-file gpurun_out/r3w_noring_$nr.csv python tools/rb_dbg.py 1024 --nodbg > /dev/null 2>&1
python - gpurun_out/r3w_noring_$nr.csv <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if r]
hi=next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
h=rows[hi]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
print(sys.argv[1], [ (r[ki].split('(')[0][-22:], r[vi], r[ui]) for r in rows[hi+1:]])
PY
done
