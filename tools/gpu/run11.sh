for c in c3 c2; do
python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2j_plain_$c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_expand|k_finalize|k_chunk_insert|k_masks|k_rec_" -s 6 -c 6 -f -o gpurun_out/r2j_full_$c python bench.py --config $c --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2j_ncu_$c.log 2>&1
echo "ncu $c rc=$?"
done
