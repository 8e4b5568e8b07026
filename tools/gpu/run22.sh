set -x
python -m pytest tests/test_concat.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r2aa_pytest_concat.log
python - <<'P' 2>&1 | tee gpurun_out/r2aa_n3.log
import sys, json
sys.path.insert(0, 'py-debruijn_b200')
import bench
for thr in (None, '128', '256', '1024'):
    import os
    if thr: os.environ['DBG_B200_CC_THREADS'] = thr
    print(thr, json.dumps(bench.n3_concat(None, None)))
P
