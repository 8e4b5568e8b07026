"""Golden vectors for III_concatenateOutpus.py:53-146, made by running the UNMODIFIED
script from /root/reference in this container  --  TEST INFRASTRUCTURE ONLY.

The script draws with the module-level `random.choice` (:65).  Each case runs it in a fresh interpreter that calls
`random.seed(seed)` first and then `runpy.run_path(script, run_name='__main__')` in a synthetic run directory
({froot}/{froot}.fasta, {froot}/setting.json, one TSV); stdout and the FASTA it appends are the expected outputs.
The score table the script builds from the TSV (:37-51) is rebuilt here with the same filters and the same order
of additions, so a case carries everything a restatement needs: contigs, score table, seed.

    python -m oracle.make_golden_concat        ->  tests/golden/concat.json.gz
"""
import gzip
import json
import os
import random
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
III = '/root/reference/III_concatenateOutpus.py'
AA = 'ACDEFGHIKLMNPQRSTVWY'

BOOT = r'''
import os, random, runpy, sys
script, seed = sys.argv[1], int(sys.argv[2])
sys.path.insert(0, os.path.dirname(script))
sys.argv = [script] + sys.argv[3:]
random.seed(seed)
runpy.run_path(script, run_name='__main__')
'''


def make_case(seed, n_contigs, plen, swap_rate, dup_rate):
    """Contigs = pieces of one random protein (so heads and tails overlap), some with two adjacent letters
    exchanged near an end (the swap-tolerant arms :85-131), some repeated, some short."""
    r = random.Random(seed)
    protein = ''.join(r.choice(AA) for _ in range(plen))
    contigs = []
    for _ in range(n_contigs):
        ln = r.choice([4, 7, 12, 18, 22, 25, 30, 40, 55])
        p = r.randrange(0, max(1, len(protein) - ln + 1))
        s = list(protein[p:p + ln])
        if len(s) >= 8 and r.random() < swap_rate:
            j = r.choice([r.randrange(0, min(6, len(s) - 1)), len(s) - 2 - r.randrange(0, min(6, len(s) - 1))])
            s[j], s[j + 1] = s[j + 1], s[j]
        contigs.append(''.join(s))
        if r.random() < dup_rate:
            contigs.append(''.join(s))
    rows = []
    for _ in range(6 * n_contigs):
        ln = r.randrange(5, 11)
        p = r.randrange(0, len(protein) - ln + 1)
        rows.append((protein[p:p + ln], r.choice([0.5, 0.85, 0.9, 0.99]), round(r.uniform(-60, 60), 1)))
    return contigs, rows


def score_table(rows, score_cut):
    """III_concatenateOutpus.py:43-51 on the rows of the one TSV."""
    t = {}
    for denovo, score, ppm in rows:
        if score >= score_cut and -50 < ppm < 50:
            t[denovo] = score + t[denovo] if denovo in t else score
    return t


def run_literal(contigs, rows, seed, score_cut=0.8, k_hi=8):
    with tempfile.TemporaryDirectory() as tmp:
        froot = 'run'
        os.makedirs(os.path.join(tmp, froot))
        os.makedirs(os.path.join(tmp, 'src'))
        with open(os.path.join(tmp, froot, froot + '.fasta'), 'w') as f:
            for i, c in enumerate(contigs):
                f.write('>SEQUENCE_%d\n%s\n' % (i, c))
        with open(os.path.join(tmp, 'src', 'spectra.tsv'), 'w') as f:
            f.write('DENOVO\tScore\tPPM Difference\n')
            for denovo, score, ppm in rows:
                f.write('%s\t%r\t%r\n' % (denovo, score, ppm))
        with open(os.path.join(tmp, froot, 'setting.json'), 'w') as f:
            json.dump({'score_cut': score_cut, 'k_upperlimit': k_hi, 'source': 'src'}, f)
        out = subprocess.run([sys.executable, '-c', BOOT, III, str(seed), '-froot', froot], cwd=tmp,
                             capture_output=True, text=True, timeout=3600)
        assert out.returncode == 0, out.stderr[-3000:]
        with open(os.path.join(tmp, froot, froot + '_concatenated.fasta')) as f:
            fasta = f.read()
    lines = out.stdout.split('\n')
    assert lines[0].startswith('{')                                      # print(setting), :35
    return '\n'.join(lines[1:]), fasta


def main():
    cases = []
    shapes = [(12, 80, 0.3, 0.2), (25, 120, 0.4, 0.1), (40, 150, 0.25, 0.1), (60, 200, 0.3, 0.05),
              (30, 60, 0.5, 0.3), (8, 40, 0.0, 0.5), (50, 400, 0.2, 0.0), (90, 250, 0.3, 0.05)]
    for n, (nc, plen, sw, du) in enumerate(shapes):
        for rep in range(3):
            seed = 100 * n + rep
            contigs, rows = make_case(seed, nc, plen, sw, du)
            stdout, fasta = run_literal(contigs, rows, seed)
            cases.append({'seed': seed, 'contigs': contigs, 'score_items': list(score_table(rows, 0.8).items()),
                          'k_upperlimit': 8, 'stdout': stdout, 'fasta': fasta})
            print('case', seed, len(contigs), 'contigs ->', fasta.count('>'), 'outputs,', stdout.count('\n'), 'lines')
    path = os.path.join(ROOT, 'tests', 'golden', 'concat.json.gz')
    with gzip.open(path, 'wt') as f:
        json.dump({'made_by': 'oracle/make_golden_concat.py', 'script': 'III_concatenateOutpus.py (unmodified, random.seed(seed) first)',
                   'cases': cases}, f)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
