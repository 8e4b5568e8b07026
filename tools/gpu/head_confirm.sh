# last confirmation of HEAD on one B200 (short: the round's GPU minutes are nearly spent)
P=gpurun_out/r2head
python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 150 python -m pytest tests/test_parity.py tests/test_abi.py tests/test_contention.py -x -q -m gpu 2>&1 | tail -4 > ${P}_pytest_subset.log; cat ${P}_pytest_subset.log
timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras > ${P}_bench_c2.json 2> ${P}_bench_c2.err; echo "bench rc=$?"
python -c "
import json
d=json.loads([l for l in open('${P}_bench_c2.json') if l.startswith('{')][-1])
print(d['value']/1e9, d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']/1e9, d['parity']['equal'] if d.get('parity') else None)"
# e2e against the slice size of the host upload (default 16 MB)
for mb in 4 8 32; do
DBG_B200_SLICE_BYTES=$((mb<<20)) timeout 90 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > ${P}_bench_c2_slice${mb}.json 2> ${P}_bench_c2_slice${mb}.err; echo "slice $mb rc=$?"
python -c "
import json
d=json.loads([l for l in open('${P}_bench_c2_slice${mb}.json') if l.startswith('{')][-1])
print('slice ${mb} MB: e2e ms', d['e2e']['ms_per_step'], 'h2d ms', d['e2e'].get('ms_h2d'), 'device step ms', d['ms_per_step'])"
done
