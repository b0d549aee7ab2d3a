timeout 600 python -m pytest tests/test_gpu_planner.py tests/test_cpp_mirror.py -m gpu -x -q 2>&1 | tail -3
timeout 100 python tools/profile_lanes.py > gpurun_out/s3i_lanes.txt 2>&1
KCBS_TRACE_SOLVE=1 timeout 100 python tools/profile_solve.py Benchmark_Empty32x32_30robots 2 3 > gpurun_out/s3i_solve_trace.txt 2>&1
compute-sanitizer --tool memcheck --print-limit 20 python tools/profile_solve.py Corridor10x10_4robots 2 2 > gpurun_out/s3i_memcheck.txt 2>&1
tail -5 gpurun_out/s3i_memcheck.txt
