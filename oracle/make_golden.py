"""Generate tests/golden/*.json.gz by running the LITERAL reference  --  TEST INFRASTRUCTURE ONLY.

Run in the build container (where /root/reference exists):

    python oracle/make_golden.py

Writes
  tests/golden/kat.json.gz    KAT-1 (the peptide example at II_assembleFromReads.py:55),
                              full snapshots for k=3,4, final False/True
  tests/golden/fuzz.json.gz   seeded fuzz corpus, full snapshots (inputs included)
  tests/golden/kat2.json.gz   KAT-2 (config-1 shape: 10 kb genome, 100 bp x 20x, k=31,
                              threshold 2), digest form: per-field sha256 + contig lengths

Every value in these files is an output of /root/reference/debruijn.py itself; the
restatements (oracle/dbg_oracle.py, oracle/dbg_oracle.c) and the CUDA path are all
compared against them.
"""
from __future__ import annotations

import gzip
import hashlib
import json
import os
import random
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import dbg_oracle, literal  # noqa: E402

FIELDS = ('nodes', 'edges', 'pull_out_read', 'branch', 'pulled', 'ect', 'contigs')


def digest(obj):
    return hashlib.sha256(json.dumps(obj, separators=(',', ':')).encode()).hexdigest()


def kat2_reads(seed=1, G=10000, L=100, cov=20):
    """SURVEY.md section 4 KAT-2 generator (pure `random`, so it is reproducible anywhere)."""
    from collections import Counter
    r = random.Random(seed)
    genome = ''.join(r.choice('ACGT') for _ in range(G))
    n = G * cov // L
    reads = []
    for _ in range(n):
        p = r.randrange(0, G - L + 1)
        reads.append(genome[p:p + L])
    return genome, list(Counter(reads).keys())


def fuzz_case(r: random.Random, idx: int):
    """One random small problem.  Mix of shapes that exercise branches, tips, cycles,
    duplicates, short reads, ties."""
    kind = idx % 8
    if kind == 7:
        alphabet = 'ACDEFGHIKLMNPQRSTVWY'      # peptides (reference's real inputs)
    elif kind == 6:
        alphabet = 'AC'                        # tiny alphabet -> repeats and cycles
    elif kind == 5:
        alphabet = 'ACG'
    else:
        alphabet = 'ACGT'
    k = r.choice([2, 3, 3, 4, 4, 5, 6, 7])
    G = r.choice([12, 20, 30, 50, 80, 120])
    genome = ''.join(r.choice(alphabet) for _ in range(G))
    if r.random() < 0.3:                       # plant a repeat -> branch / cycle
        a = r.randrange(0, G // 2)
        ln = r.randrange(k, k + 6)
        b = r.randrange(G // 2, max(G // 2 + 1, G - ln))
        genome = genome[:b] + genome[a:a + ln] + genome[b + ln:]
        genome = genome[:G]
    L = r.choice([k - 1, k, k + 1, k + 2, 8, 10, 15, 25, 40])
    L = max(1, min(L, G))
    n = r.choice([1, 2, 4, 8, 15, 30, 60])
    err = r.choice([0.0, 0.0, 0.02, 0.05, 0.1])
    reads = []
    for _ in range(n):
        ln = max(1, L + r.choice([0, 0, 0, -1, 1, 3, -3]))
        ln = min(ln, G)
        p = r.randrange(0, G - ln + 1)
        s = list(genome[p:p + ln])
        for j in range(len(s)):
            if r.random() < err:
                s[j] = r.choice([c for c in alphabet if c != s[j]])
        reads.append(''.join(s))
    if r.random() < 0.5:                       # II_assembleFromReads.py:52 dedups; not always
        from collections import Counter
        reads = list(Counter(reads).keys())
    threshold = r.choice([1, 2, 2, 3, 5, 1.5, 2.5])
    final = r.random() < 0.4
    return reads, k, threshold, final


def main():
    if not literal.available():
        raise SystemExit('literal reference not found; golden files can only be generated '
                         'where /root/reference exists')
    mod = literal.load()
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(out_dir, exist_ok=True)

    # ---- KAT-1
    kat = []
    reads = ['EVQLVE', 'QLVAPG', 'LVESGGAL', 'LVESGGGL']
    for k in (3, 4, 5):
        for final in (False, True):
            snap = literal.snapshot(reads, k, 2, final, mod)
            kat.append({'reads': reads, 'k': k, 'threshold': 2, 'final': final, 'expect': snap})
    with gzip.open(os.path.join(out_dir, 'kat.json.gz'), 'wt') as f:
        json.dump(kat, f, separators=(',', ':'))
    print('kat.json.gz:', len(kat), 'cases')

    # ---- fuzz corpus
    r = random.Random(20261018)
    cases = []
    stats = {'branch': 0, 'pulled': 0, 'skipped': 0, 'final_multi': 0}
    idx = 0
    t0 = time.time()
    while len(cases) < 600:
        reads, k, threshold, final = fuzz_case(r, idx)
        idx += 1
        try:
            mine = dbg_oracle.snapshot(reads, k, threshold, final, max_paths=300)
        except RuntimeError:
            stats['skipped'] += 1             # final=True path explosion: exponential in the
            continue                          # reference too
        snap = literal.snapshot(reads, k, threshold, final, mod)
        for fld in FIELDS:
            if mine[fld] != snap[fld]:
                raise SystemExit('restatement != literal on %s for case %r' %
                                 (fld, (reads, k, threshold, final)))
        if snap['pulled']:
            stats['pulled'] += 1
        if any(len(s) > 1 for _, s in snap['edges']):
            stats['branch'] += 1
            if final:
                stats['final_multi'] += 1
        cases.append({'reads': reads, 'k': k, 'threshold': threshold, 'final': final,
                      'expect': snap})
    with gzip.open(os.path.join(out_dir, 'fuzz.json.gz'), 'wt') as f:
        json.dump(cases, f, separators=(',', ':'))
    print('fuzz.json.gz:', len(cases), 'cases', stats, '%.1fs' % (time.time() - t0))

    # ---- KAT-2 (digest form; the literal walk takes ~1 min)
    genome, reads = kat2_reads()
    kat2 = {'generator': {'seed': 1, 'G': 10000, 'L': 100, 'cov': 20}, 'k': 31, 'threshold': 2,
            'n_reads': len(reads), 'reads_sha256': digest(reads), 'cases': []}
    for final in (False, True):
        t0 = time.time()
        snap = literal.snapshot(reads, 31, 2, final, mod)
        ent = {'final': final, 'stdout': snap['stdout'],
               'n_nodes': len(snap['nodes']), 'n_ect': len(snap['ect']),
               'n_branch': len(snap['branch']), 'n_pulled': len(snap['pulled']),
               'contig_lengths': [len(c) for c in snap['contigs']],
               'contig_total_chars': sum(len(c) for c in snap['contigs']),
               'sha256': {fld: digest(snap[fld]) for fld in FIELDS},
               'sha256_sorted_contigs': digest(sorted(snap['contigs']))}
        kat2['cases'].append(ent)
        print('kat2 final=%s: %d nodes, %d contigs, %d chars, %.1fs' %
              (final, ent['n_nodes'], len(ent['contig_lengths']), ent['contig_total_chars'],
               time.time() - t0))
    with gzip.open(os.path.join(out_dir, 'kat2.json.gz'), 'wt') as f:
        json.dump(kat2, f, separators=(',', ':'))


if __name__ == '__main__':
    main()
