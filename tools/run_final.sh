# Round-end check on one GPU: the GPU suite, smoke(), both bench arms.
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "reference arm rc=$?"
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
