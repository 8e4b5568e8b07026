python -m pytest tests/test_contention.py -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2af_contention.log
