# the widened generic-alphabet engine on one B200: its own tests, the peptide / golden parity tests, the real II caller, smoke
P=gpurun_out/r2alpha
timeout 200 python -m pytest tests/test_alphabets.py tests/test_real_caller.py tests/test_parity.py -x -q -m gpu -k "alphabet or caller or generic or kat1 or golden or invalid or limits or lookups" 2>&1 | tail -4 > ${P}_pytest.log; cat ${P}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 ${P}_smoke.log
