timeout 600 python -m pytest tests/test_gpu_planner.py tests/test_gpu_pair.py -m gpu -x -q 2>&1 | tail -3
for s in Corridor10x10_4robots Congested10x10_7robots; do timeout 120 python tools/stress_solve.py $s 6 10; done > gpurun_out/s3n_stress.txt 2>&1
timeout 200 python tools/profile_solve.py Benchmark_Empty32x32_30robots 1 9 > gpurun_out/s3n_solve30.txt 2>&1
KCBS_B200_LIB=$PWD/build/libkcbs_trace.so KCBS_TRACE_SOLVE=1 python tools/profile_solve.py Benchmark_Empty32x32_30robots 3 1 > gpurun_out/s3n_trace.txt 2>&1
