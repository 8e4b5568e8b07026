# end of round 2: HEAD on one B200 -- smoke, the gpu tests touched in the last session, and the DEFAULT bench line (extras, cpu baseline)
P=gpurun_out/r2end
python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 ${P}_smoke.log
timeout 200 python -m pytest tests/test_alphabets.py tests/test_degrees.py tests/test_abi.py tests/test_parity.py tests/test_strands.py -x -q -m gpu 2>&1 | tail -3 | tee ${P}_pytest_subset.log
timeout 300 python bench.py > ${P}_bench_c2_default.json 2> ${P}_bench_c2_default.err; echo "bench rc=$?"
python -c "
import json
d=json.loads([l for l in open('${P}_bench_c2_default.json') if l.startswith('{')][-1])
print(round(d['value']/1e9,2), 'G k-mers/s', round(d['ms_per_step'],3), 'ms; frac', round(d['roofline']['frac'],3), '; e2e', round(d['e2e']['value']/1e9,2), 'G; parity', d['parity']['equal'], '; probe', d['probe_length']['nodes']['mean'], '; keys', sorted(d))"
