# usage: bash tools/run_scale.sh N  -- bench.py under torchrun on N GPUs of one box (both arms), output under gpurun_out/
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/final_bench_n$N.json 2> gpurun_out/final_bench_n$N.err; echo "bench rc=$?"
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_sharded_scan.py -m gpu -x -q 2>&1 | tail -2; fi
