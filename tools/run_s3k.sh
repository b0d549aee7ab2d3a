for i in 1 2 3 4 5; do timeout 300 python -m pytest tests/test_gpu_planner.py -m gpu -x -q -k concurrent 2>&1 | tail -2; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/s3k_bench.json 2> gpurun_out/s3k_bench.err; echo "bench rc=$?"
timeout 200 python tools/solve_scenes.py > gpurun_out/s3k_solve_scenes.txt 2>&1
timeout 200 python tools/profile_solve.py Benchmark_Empty32x32_30robots 1 9 > gpurun_out/s3k_solve30.txt 2>&1
timeout 100 python tools/profile_rrt.py > gpurun_out/s3k_rrt.txt 2>&1
