timeout 600 python -m pytest tests/test_gpu_planner.py tests/test_gpu_pair.py -m gpu -x -q 2>&1 | tail -3
for s in Corridor10x10_4robots Congested10x10_7robots; do timeout 120 python tools/stress_solve.py $s 6 10; done > gpurun_out/s3o_stress.txt 2>&1
for lib in "" build/libkcbs_u1.so build/libkcbs_u4.so; do
echo "== lib $lib"
KCBS_B200_LIB=${lib:+$PWD/$lib} python tools/profile_rrt.py 2>&1 | tail -2
KCBS_B200_LIB=${lib:+$PWD/$lib} timeout 200 python tools/profile_solve.py Benchmark_Empty32x32_30robots 1 9 2>&1 | grep seed | awk '{print $5}' | tr '\n' ' '; echo
done
