set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
free -g | head -2; nproc
( time python -m pytest tests -m gpu -x -q --durations=25 ) > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/r2a_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err; echo "bench rc=$?"
python bench.py --config c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_c5.json 2> gpurun_out/r2a_bench_c5.err; echo "c5 rc=$?"
python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err; echo "c3 rc=$?"
