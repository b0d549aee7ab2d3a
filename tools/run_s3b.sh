fails=0
for i in $(seq 1 30); do
SEGV_BT_FILE=$PWD/gpurun_out/s3d_bt_$i.txt LD_PRELOAD=$PWD/build/segv_bt.so timeout 300 python -m pytest tests/test_gpu_planner.py -m gpu -x -q -p no:faulthandler > gpurun_out/s3d_planner_$i.log 2>&1; rc=$?
if [ $rc -eq 0 ]; then grep -c "round loop graph not built" gpurun_out/s3d_planner_$i.log > /dev/null && echo "run $i fell back"; rm -f gpurun_out/s3d_bt_$i.txt gpurun_out/s3d_planner_$i.log; else fails=$((fails+1)); echo "run $i rc=$rc"; fi
done
echo "failures: $fails of 30"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
