timeout 600 python -m pytest tests/test_gpu_planner.py tests/test_cpp_mirror.py -m gpu -x -q 2>&1 | tail -5
KCBS_RRT_ROUNDS=stream timeout 600 python -m pytest tests/test_gpu_planner.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/profile_lanes.py 2>&1 | tail -8
KCBS_RRT_ROUNDS=stream timeout 200 python tools/profile_lanes.py 2>&1 | tail -8
timeout 200 python tools/profile_solve.py Benchmark_Empty32x32_30robots 2 3 2>&1 | tail -5
KCBS_RRT_ROUNDS=stream timeout 200 python tools/profile_solve.py Benchmark_Empty32x32_30robots 2 3 2>&1 | tail -5
