for c in c2 c3; do
python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2p_plain_$c.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2p_launches_$c.csv python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2p_ncu_$c.log 2>&1
echo "ncu $c rc=$?"
done
