# compute-sanitizer memcheck over the round-2 kernels (decoder, TMA GEMM, scheduler/graphs, ragged decode, chunker on tensor cores, slot validation)
cd $GRAFT_REPO_ROOT
export PYTHONDONTWRITEBYTECODE=1
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest -q --no-header -x -m gpu \
  "tests/test_gpu_decoder.py::test_decoder_replays_the_real_modules_golden" \
  "tests/test_gpu_decoder.py::test_decoder_feeds_the_tail_without_leaving_the_device" \
  "tests/test_gpu_sched.py::test_scheduler_same_session_twice_in_flight_keeps_order" \
  "tests/test_gpu_codec.py::test_decode_many_ragged_and_uniform_bit_exact_vs_oracle" \
  "tests/test_gpu_config_sizes.py::test_device_side_slot_validation" \
  "tests/test_gpu_tail.py::test_tail_bf16_snr_and_state" \
  "tests/test_gpu_resblock.py" > gpurun_out/r2_sanitizer.log 2>&1
echo "rc=$?" >> gpurun_out/r2_sanitizer.log
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r2_sanitizer.log
tail -n 12 gpurun_out/r2_sanitizer.log
