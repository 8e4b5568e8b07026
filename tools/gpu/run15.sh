python -m pytest tests -m gpu -x -q --deselect tests/test_fullsize.py -k "min_count or dist or real_caller or abi or generic or kat1 or edge or ii_multi or descriptors or invalid or threshold or reads_not" > gpurun_out/r2q_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2q_pytest_gpu.log
python bench.py --config c3 --min-count 2 --steps 10 --warmup 3 --no-extras > gpurun_out/r2q_bench_c3_mc2.json 2> gpurun_out/r2q_bench_c3_mc2.err; echo "bench rc=$?"
python - <<'PY'
import json
p=json.loads([l for l in open('gpurun_out/r2q_bench_c3_mc2.json') if l.startswith('{')][-1])
print('c3 mc2: step %.3f e2e %.3f nodes %d contigs %d parity %s' % (p['ms_per_step'], p['e2e']['ms_per_step'], p['graph']['nodes'], p['graph']['contigs'], p.get('parity',{}).get('equal')), p['stage_ms'])
PY
