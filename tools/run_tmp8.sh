for i in 1 2; do for s in Corridor10x10_4robots Congested10x10_7robots; do timeout 120 python tools/stress_solve.py $s 8 10; done > gpurun_out/s3t_stress_$i.txt 2>&1; done
KCBS_RRT_ROUNDS=stream bash -c 'for s in Corridor10x10_4robots Congested10x10_7robots; do timeout 120 python tools/stress_solve.py $s 8 10; done' > gpurun_out/s3t_stress_3.txt 2>&1
for i in 1 2 3 4 5 6; do timeout 300 python -m pytest tests/test_gpu_planner.py -m gpu -x -q 2>&1 | tail -1; done
timeout 200 python tools/solve_scenes.py > gpurun_out/s3t_solve_scenes.txt 2>&1
