set -x
cd $GRAFT_REPO_ROOT
M="gpu__time_duration.sum"
P=${NCU_PREFIX:-r2}          # file prefix under gpurun_out/ (NCU_PREFIX=r2z for the stacked-output build)
# (1) launch list of one timed bench step (42 launches per step: skip the warm-up step)
timeout 300 ncu --metrics $M --clock-control none -s 42 -c 42 --csv --log-file gpurun_out/${P}_launches_bench_1024sessions.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-latency --no-strong --no-front > gpurun_out/ncu1.log 2>&1
# (2) full capture of every tcgen05 launch of one warm call at 128 sessions
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_resblock|k_conv_umma|k_gemm_tc" -o gpurun_out/${P}_tail_full python tools/rb_dbg.py 128 --nodbg > gpurun_out/ncu2.log 2>&1
ncu -i gpurun_out/${P}_tail_full.ncu-rep --page raw --csv > gpurun_out/${P}_tail_full_raw.csv 2>> gpurun_out/ncu2.log
# (3) full capture of one decoder step at position 96, 1,024 sessions (NCU_SKIP_DEC=1: not this time)
if [ -z "$NCU_SKIP_DEC" ]; then
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_gemm_tc|k_attend|k_add_ln" -o gpurun_out/${P}_dec_full python tools/dec_prof.py 1024 96 > gpurun_out/ncu3.log 2>&1
ncu -i gpurun_out/${P}_dec_full.ncu-rep --page raw --csv > gpurun_out/${P}_dec_full_raw.csv 2>> gpurun_out/ncu3.log
fi
ls -la gpurun_out/*.ncu-rep
for f in gpurun_out/${P}_tail_full.ncu-rep gpurun_out/${P}_dec_full.ncu-rep; do s=$(stat -c %s $f 2>/dev/null || echo 0); if [ "$s" -gt 25000000 ]; then rm -f $f; fi; done
for f in gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log; do tail -n 3 $f; done
wc -l gpurun_out/${P}_launches_bench_1024sessions.csv gpurun_out/${P}_tail_full_raw.csv gpurun_out/${P}_dec_full_raw.csv
