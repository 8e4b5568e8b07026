"""bench.py is what the driver parses: it must import, and its reference arm must print ONE JSON line with
the contract's keys (CPU only; the product arm needs a GPU and is exercised on the B200 box)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_imports_and_lists_configs():
    sys.path.insert(0, ROOT)
    try:
        import bench
    finally:
        sys.path.pop(0)
    assert {'c1', 'c2', 'c3', 'c4', 'c5s'} <= set(bench.CONFIGS)
    bases, off = bench.make_reads(bench.CONFIGS['c1'])
    assert len(off) > 1000 and len(bases) == int(off[-1])
    cfg = dict(bench.CONFIGS['c4'], T=20)
    b, o, rb, bo = bench.make_reads_batched(cfg)
    assert len(rb) == len(o) - 1 and (rb[1:] >= rb[:-1]).all() and len(bo) == 21


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--config', 'c1',
                          '--steps', '2', '--warmup', '1'], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
                'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['value'] > 0 and line['cpu_baseline']['kind'] == 'port'
    assert line['e2e']['h2d_bytes_per_step'] == 0


def test_product_arm_refuses_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1'], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode != 0 and 'no CPU path' in (out.stdout + out.stderr)
