python bench.py --config c3 --min-count 2 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2r_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2r_launches_c3mc.csv python bench.py --config c3 --min-count 2 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2r_ncu.log 2>&1
echo rc=$?
