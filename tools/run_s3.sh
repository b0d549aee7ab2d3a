set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s3_gputests.log 2>&1; echo "rc=$?" >> gpurun_out/s3_gputests.log
tail -5 gpurun_out/s3_gputests.log
timeout 400 python bench.py > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; echo rc=$?
timeout 300 python bench.py --impl reference > gpurun_out/s3_bench_ref.json 2> gpurun_out/s3_bench_ref.err; echo rc=$?
timeout 200 python tools/solve_scenes.py > gpurun_out/s3_solve_scenes.txt 2>&1
timeout 200 python tools/profile_solve.py Benchmark_Empty32x32_30robots 1 8 > gpurun_out/s3_solve30.txt 2>&1
timeout 100 python tools/profile_lanes.py > gpurun_out/s3_lanes.txt 2>&1
timeout 100 python tools/profile_rrt.py > gpurun_out/s3_rrt.txt 2>&1
tail -3 gpurun_out/s3_solve30.txt gpurun_out/s3_rrt.txt
