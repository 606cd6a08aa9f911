# Round-end evidence of the build in the tree, one gpurun call on one GPU: GPU tests, both bench arms, config-4/5 sweeps, ncu launch list + full capture
# of the tcgen05 launches, memcheck of the kernels that are new.  P = file prefix under gpurun_out/ (copy what is to be judged into profiles/).
cd $GRAFT_REPO_ROOT
export PYTHONDONTWRITEBYTECODE=1
P=${P:-r3}
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/${P}_gpu_tests.log; cat gpurun_out/${P}_gpu_tests.log
timeout 600 python bench.py --impl reference > gpurun_out/${P}_bench_reference_arm.json 2> gpurun_out/${P}_bench.err
timeout 900 python bench.py > gpurun_out/${P}_bench_n1.json 2>> gpurun_out/${P}_bench.err; tail -c 300 gpurun_out/${P}_bench.err
timeout 600 python tools/sweeps.py > gpurun_out/${P}_sweeps.log 2>&1; ls -t gpurun_out/*sweeps*json | head -2
python - <<PY
import json
d=json.loads(open('gpurun_out/${P}_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','kernel_ms_per_step','clocks','gpu_launches') if k in d})
print(d['e2e']['value'], d['roofline']['frac'], d['roofline']['achieved'], d.get('roofline_codec',{}).get('frac'))
for l in d['latency']: print({k:l.get(k) for k in ('sessions','p50_ms','p99_ms','max_ms','graphs_built_while_serving','met')})
print(d.get('front_half',{}).get('streams_front_plus_tail'))
print(d.get('strong'))
r=json.loads(open('gpurun_out/${P}_bench_reference_arm.json').read().strip().splitlines()[-1])
print(r.get('value'), r.get('cpu_baseline'))
PY
# launch list: three steps' worth from the start; tools/ncu_summary.py is not needed for this one (the CSV is small)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file gpurun_out/${P}_launches_raw.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-latency --no-strong --no-front --no-sessions > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_resblock|k_conv_umma|k_gemm_tc" -o gpurun_out/${P}_tail_full python tools/rb_dbg.py 128 --nodbg > gpurun_out/ncu2.log 2>&1
ncu -i gpurun_out/${P}_tail_full.ncu-rep --page raw --csv > gpurun_out/${P}_tail_full_raw.csv 2>> gpurun_out/ncu2.log
s=$(stat -c %s gpurun_out/${P}_tail_full.ncu-rep 2>/dev/null || echo 0); if [ "$s" -gt 25000000 ]; then rm -f gpurun_out/${P}_tail_full.ncu-rep; fi
tail -n 2 gpurun_out/ncu1.log gpurun_out/ncu2.log; wc -l gpurun_out/${P}_launches_raw.csv gpurun_out/${P}_tail_full_raw.csv
# memcheck: the C = 256 pair kernel (per-layer tests) and one whole tail call (fused upsampler, pair kernel inside the step)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest -q --no-header -x -m gpu \
  "tests/test_gpu_conv.py::test_tensor_core_conv_matches_torch_on_bf16_operands" "tests/test_gpu_tail.py::test_tail_bf16_snr_and_state" "tests/test_gpu_decoder.py::test_decoder_two_calls_continuous_batching_and_sub_passes" > gpurun_out/${P}_sanitizer_memcheck.log 2>&1
echo "rc=$?" >> gpurun_out/${P}_sanitizer_memcheck.log; tail -n 6 gpurun_out/${P}_sanitizer_memcheck.log
