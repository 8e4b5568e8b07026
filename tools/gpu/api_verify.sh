P=gpurun_out/r2api
timeout 120 python -m pytest tests/test_parity.py tests/test_abi.py tests/test_min_count.py -x -q -m gpu 2>&1 | tail -3 | tee ${P}_pytest.log
timeout 60 python tools/gpu/api_batch_timing.py 2>&1 | tail -2 | tee ${P}_batch3.log
