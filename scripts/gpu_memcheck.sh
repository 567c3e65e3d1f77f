# compute-sanitizer memcheck over the GPU parity tests that exercise malformed input and every block kind (both expansion
# kernels: the decode tests are parametrised), plus the LZVN and raw slices of the hostile-input matrix.
# One sanitizer tool per call (see B200_PROFILING.md); plain run first.
SEL="mutated or frame_level or fixtures or patterns or kats or frames_equal or capacity or ragged or mixed or (mutate_0 and (vxn or raw)) or (mutate_6_7 and (vxn or raw))"
FILES="tests/test_gpu_decode.py tests/test_gpu_encode.py tests/test_gpu_configs.py tests/test_gpu_hostile.py"
python -m pytest $FILES -m gpu -x -q -k "$SEL" 2>&1 | tail -2 && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 --print-limit 20 python -m pytest $FILES -m gpu -x -q -k "$SEL" > gpurun_out/memcheck.log 2>&1; echo memcheck_rc=$?
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/memcheck.log | head -20
