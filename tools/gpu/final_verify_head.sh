P=gpurun_out/r2fin3
python -m pytest tests -m gpu -q > ${P}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 ${P}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 ${P}_smoke.log
python bench.py > ${P}_bench_c2.json 2> ${P}_bench_c2.err; echo "bench rc=$?"
python -c "
import json
d=json.loads([l for l in open('${P}_bench_c2.json') if l.startswith('{')][-1])
print(d['value']/1e9, d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['parity']['equal'], d['n3_concat'])"
