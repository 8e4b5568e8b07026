set -x
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for c in "c2 600000" "c3 400000"; do
  timeout 300 $TR tools/check_dist.py $c > gpurun_out/r2s_check_dist_n${N}_${c%% *}.log 2>&1; echo "check_dist $c rc=$?"; tail -1 gpurun_out/r2s_check_dist_n${N}_${c%% *}.log
done
( time timeout 800 $TR bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/r2s_bench_c2_n$N.out 2> gpurun_out/r2s_bench_c2_n$N.err; echo "bench rc=$?"; tail -5 gpurun_out/r2s_bench_c2_n$N.err
grep '^{' gpurun_out/r2s_bench_c2_n$N.out > gpurun_out/r2s_bench_c2_n$N.json
