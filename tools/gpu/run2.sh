set -x
python -m pytest tests -m gpu -x -q -k "not c5_full and not c5s_full and not c4_batched" > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2b_pytest_gpu.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err; echo "bench rc=$?"; tail -5 gpurun_out/r2b_bench_c2.err
