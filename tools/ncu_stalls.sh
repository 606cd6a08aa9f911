# per-source-line stall samples of the fused ResBlock kernel's k=3 launches (C = 32 and C = 64), 128 sessions
set -x
cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_resblock -o gpurun_out/r2_rb_src python tools/rb_dbg.py 128 --nodbg > gpurun_out/ncu4.log 2>&1
ncu -i gpurun_out/r2_rb_src.ncu-rep --page source --csv --print-source cuda > gpurun_out/r2_rb_src_cuda.csv 2>> gpurun_out/ncu4.log
ncu -i gpurun_out/r2_rb_src.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_rb_src_sass.csv 2>> gpurun_out/ncu4.log
ncu -i gpurun_out/r2_rb_src.ncu-rep --page details --csv > gpurun_out/r2_rb_details.csv 2>> gpurun_out/ncu4.log
rm -f gpurun_out/r2_rb_src.ncu-rep
ls -la gpurun_out/r2_rb_*; tail -n 5 gpurun_out/ncu4.log
