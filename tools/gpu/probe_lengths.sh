# probe lengths of the node / chunk table at C2, C3 and C5s (dbg_probe_stats through bench.py) + its gpu test
P=gpurun_out/r2probe
timeout 120 python -m pytest tests/test_degrees.py tests/test_abi.py -x -q -m gpu 2>&1 | tail -3
for c in c2 c3 c5s; do
timeout 150 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-extras > ${P}_bench_$c.json 2> ${P}_bench_$c.err; echo "bench $c rc=$?"
python -c "
import json
d=json.loads([l for l in open('${P}_bench_$c.json') if l.startswith('{')][-1])
print('$c', round(d['value']/1e9,2), 'G k-mers/s', round(d['ms_per_step'],3), 'ms', 'parity', (d.get('parity') or {}).get('equal')); print(json.dumps(d['probe_length']))"
done
