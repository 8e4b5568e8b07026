P=gpurun_out/r2fin
python -m pytest tests -m gpu -q > ${P}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 ${P}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > ${P}_bench_c2.json 2> ${P}_bench_c2.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_c2_reference_arm.json 2> ${P}_ref.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > ${P}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > ${P}_ncu1.log 2>&1
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > ${P}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_expand|k_finalize|k_chunk_insert|k_masks|k_rec_|k_pack" -s 9 -c 9 -f -o ${P}_full_c2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > ${P}_ncu2.log 2>&1
echo "ncu rc=$?"
