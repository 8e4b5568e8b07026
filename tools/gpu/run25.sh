python -m pytest tests/test_contention.py -x -q -m gpu -k "False-47" 2>&1 | tail -40 | tee gpurun_out/r2ad_contention.log
