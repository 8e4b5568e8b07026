python -m pytest tests -m gpu -x -q > gpurun_out/r2v_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2v_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/r2v_bench_c2.json 2> gpurun_out/r2v_bench_c2.err; echo "bench rc=$?"
for c in c1 c3 c4 c5s c5; do python bench.py --config $c --steps 10 --warmup 3 --no-extras > gpurun_out/r2v_bench_$c.json 2> gpurun_out/r2v_bench_$c.err; echo "bench $c rc=$?"; done
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2v_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2v_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2v_ncu1.log 2>&1
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2v_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_expand|k_finalize|k_chunk_insert|k_masks|k_rec_|k_pack" -s 9 -c 9 -f -o gpurun_out/r2v_full_c2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2v_ncu2.log 2>&1
echo "ncu rc=$?"
