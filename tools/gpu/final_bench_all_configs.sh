P=gpurun_out/r2fin
for c in c1 c3 c4 c5s; do python bench.py --config $c --steps 10 --warmup 3 --no-extras > ${P}_bench_$c.json 2> ${P}_bench_$c.err; echo "bench $c rc=$?"; done
python bench.py --config c3 --min-count 2 --steps 10 --warmup 3 --no-extras > ${P}_bench_c3_min_count2.json 2> ${P}_bench_c3mc.err; echo "bench c3mc rc=$?"
python bench.py --config c3f --final --steps 10 --warmup 3 --no-extras > ${P}_bench_c3f_final.json 2> ${P}_bench_c3f.err; echo "bench c3f rc=$?"
python bench.py --config c5 --steps 5 --warmup 3 --no-extras > ${P}_bench_c5.json 2> ${P}_bench_c5.err; echo "bench c5 rc=$?"
