# one sanitizer tool per call (see B200_PROFILING.md); plain run first
python -m pytest tests/test_gpu_decode.py tests/test_gpu_encode.py -m gpu -x -q -k "mutated or frame_level or fixtures or patterns or kats or frames_equal or capacity" 2>&1 | tail -2 && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 --print-limit 20 python -m pytest tests/test_gpu_decode.py tests/test_gpu_encode.py -m gpu -x -q -k "mutated or frame_level or fixtures or patterns or kats or frames_equal or capacity" > gpurun_out/memcheck.log 2>&1; echo memcheck_rc=$?
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/memcheck.log | head -20
