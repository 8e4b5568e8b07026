"""The real caller, unmodified: /root/reference/II_assembleFromReads.py run once against the reference's
own debruijn.py and once with py-debruijn_b200 first on sys.path (the drop-in module), on a synthetic
de-novo peptide run directory ({froot}/setting.json + one TSV with DENOVO / Score / PPM Difference,
II_assembleFromReads.py:32-47).  stdout and the FASTA II writes must be identical.

Skipped where /root/reference does not exist (the GPU box); `emu` runs the product's sources under the
CPU emulator, `gpu` the real library."""
import json
import os
import random
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, 'py-debruijn_b200')
II = '/root/reference/II_assembleFromReads.py'

BOOT = r'''
import os, runpy, sys
script, use_product, backend = sys.argv[1], sys.argv[2] == '1', sys.argv[3]
if use_product:
    for p in (%(tests)r, %(root)r, %(pkg)r):
        sys.path.insert(0, p)
    if backend == 'emu':
        from emu import build_emu
        from dbg_b200 import _native
        _native._lib = _native.Library(build_emu.build())
    import debruijn
    assert debruijn.__file__.startswith(%(pkg)r), debruijn.__file__
else:
    sys.path.insert(0, os.path.dirname(script))
    sys.setrecursionlimit(100000)
sys.argv = [script] + sys.argv[4:]
runpy.run_path(script, run_name='__main__')
''' % {'tests': HERE, 'root': ROOT, 'pkg': PKG}


def _make_run_dir(tmp, froot, seed, k_lo, k_hi, threshold):
    r = random.Random(seed)
    aa = 'ACDEFGHIKLMNPQRSTVWY'
    protein = ''.join(r.choice(aa) for _ in range(260))
    protein = protein[:120] + protein[40:70] + protein[120:]          # a repeat: branch nodes
    rows = ['DENOVO\tScore\tPPM Difference']
    for _ in range(700):
        ln = r.randrange(7, 22)
        p = r.randrange(0, len(protein) - ln + 1)
        s = list(protein[p:p + ln])
        for j in range(ln):
            if r.random() < 0.02:
                s[j] = r.choice(aa)
        rows.append('%s\t%.2f\t%.1f' % (''.join(s), r.choice([0.5, 0.85, 0.9, 0.99]), r.uniform(-60, 60)))
    d = os.path.join(tmp, froot)
    src = os.path.join(tmp, froot + '_src')
    os.makedirs(d)
    os.makedirs(src)
    with open(os.path.join(src, 'spectra.tsv'), 'w') as f:
        f.write('\n'.join(rows) + '\n')
    with open(os.path.join(d, 'setting.json'), 'w') as f:
        json.dump({'score_cut': 0.8, 'k_lowerlimit': k_lo, 'k_upperlimit': k_hi, 'threshold': threshold,
                   'source': froot + '_src'}, f)


def _run(tmp, froot, use_product, backend):
    out = subprocess.run([sys.executable, '-c', BOOT, II, '1' if use_product else '0', backend, '-froot', froot],
                         cwd=tmp, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    with open(os.path.join(tmp, froot, froot + '.fasta')) as f:
        return out.stdout, f.read()


@pytest.mark.parametrize('k_lo,k_hi,threshold', [(5, 8, 2), (6, 6, 3)])
def test_ii_assemble_from_reads_unmodified(backend, tmp_path, k_lo, k_hi, threshold):
    if not os.path.exists(II):
        pytest.skip('the reference tree is not on this box')
    pytest.importorskip('pandas')
    a, b = str(tmp_path / 'ref'), str(tmp_path / 'new')
    os.makedirs(a)
    os.makedirs(b)
    for d in (a, b):
        _make_run_dir(d, 'run_%d-%dmer' % (k_lo, k_hi), 7, k_lo, k_hi, threshold)
    froot = 'run_%d-%dmer' % (k_lo, k_hi)
    ref_out, ref_fasta = _run(a, froot, False, backend)
    new_out, new_fasta = _run(b, froot, True, backend)
    assert ref_fasta.count('>') > 3
    assert new_fasta == ref_fasta
    assert new_out == ref_out
