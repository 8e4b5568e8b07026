N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
( time timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-extras ) > gpurun_out/r2u_bench_c2_n$N.out 2> gpurun_out/r2u_bench_c2_n$N.err; echo "bench rc=$?"; tail -5 gpurun_out/r2u_bench_c2_n$N.err
grep '^{' gpurun_out/r2u_bench_c2_n$N.out > gpurun_out/r2u_bench_c2_n$N.json
