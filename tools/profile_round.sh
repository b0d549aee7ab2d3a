set -x
S=Empty32x32_10robots
python tools/profile_k1.py 655360 0 packed $S 5 0 > gpurun_out/r2_plain_k1_packed.log 2>&1
python tools/profile_k1.py 655360 0 dense $S 5 0 > gpurun_out/r2_plain_k1_dense.log 2>&1
python tools/profile_k1.py 65536 0 packed $S 5 2 > gpurun_out/r2_plain_k1_packed64k.log 2>&1
cat gpurun_out/r2_plain_k1_*.log
ncu --set full --clock-control none --import-source on -k regex:k_propagate -s 1 -c 1 -o gpurun_out/r2f_k1_packed -f python tools/profile_k1.py 655360 0 packed $S 5 0 > gpurun_out/r2_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_propagate -s 1 -c 1 -o gpurun_out/r2f_k1_dense -f python tools/profile_k1.py 655360 0 dense $S 5 0 > gpurun_out/r2_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_propagate -s 1 -c 1 -o gpurun_out/r2f_k1_packed64k -f python tools/profile_k1.py 65536 0 packed $S 5 2 > gpurun_out/r2_ncu_c.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_plain_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/r2_ncu_bench.log 2>&1
python tools/sweep.py > gpurun_out/r2_size_sweep.md 2> gpurun_out/r2_sweep.err
python tools/profile_rrt.py > gpurun_out/r2_rrt_final.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_rrt_launches_final.csv python tools/profile_rrt.py > /dev/null 2>&1
python tools/profile_kernels.py conflict > gpurun_out/r2_plain_conflict.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_conflict_keys -s 1 -c 1 -o gpurun_out/r2f_k3 -f python tools/profile_kernels.py conflict > gpurun_out/r2_ncu_d.log 2>&1
python tools/profile_kernels.py validity > gpurun_out/r2_plain_validity.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_validity -s 1 -c 1 -o gpurun_out/r2f_k2 -f python tools/profile_kernels.py validity > gpurun_out/r2_ncu_e.log 2>&1
ls -la gpurun_out/r2f_*
