# Round-2 final pass (after the planner's device-driven rounds): launch lists of the bench and of the RRT kernels, the
# per-kernel device stamps of graph-driven rounds, the size sweep.  Every ncu command only after the same command exited 0.
set -x
python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/r2f_plain_bench.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/r2f_ncu_bench.log 2>&1
KCBS_RRT_ROUNDS=stream python tools/profile_rrt.py > gpurun_out/r2f_rrt_stream.txt 2>&1 || exit 1
KCBS_RRT_ROUNDS=stream ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r2f_rrt_launches.csv python tools/profile_rrt.py > /dev/null 2>&1
python tools/profile_rrt.py > gpurun_out/r2f_rrt_graph.txt 2>&1
KCBS_B200_LIB=$PWD/build/libkcbs_trace.so KCBS_TRACE_SOLVE=1 python tools/profile_rrt.py > gpurun_out/r2f_rrt_trace.txt 2>&1
python tools/sweep.py > gpurun_out/r2f_size_sweep.md 2> gpurun_out/r2f_sweep.err
python tools/profile_lanes.py > gpurun_out/r2f_lanes.txt 2>&1
