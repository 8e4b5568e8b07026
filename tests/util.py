"""Shared helpers of the parity tests: snapshots in the oracle's JSON-able shape."""
import contextlib
import gzip
import io
import json
import os
import random

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, 'golden')
FIELDS = ('nodes', 'edges', 'pull_out_read', 'branch', 'pulled', 'ect', 'contigs')


def load_golden(name):
    with gzip.open(os.path.join(GOLDEN, name), 'rt') as f:
        return json.load(f)


def is_dna(reads):
    return all(set(r) <= set('ACGT') for r in reads)


def product_snapshot(reads, k, threshold, final, min_count=1):
    """Run the product's drop-in module exactly as II_assembleFromReads.py:58-63 does."""
    import debruijn as db
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        kw = {'min_count': min_count} if min_count != 1 else {}
        g, pull, branch, pulled, ect = db.construct_graph(list(reads), k, threshold=threshold, final=final, **kw)
        contigs = db.output_contigs(g, branch, pulled)
    V, E = g
    return {
        'nodes': [[v, n.indegree, n.outdegree] for v, n in V.items()],
        'edges': [[v, list(s)] for v, s in E.items()],
        'pull_out_read': list(pull),
        'branch': list(branch),
        'pulled': list(pulled),
        'ect': [[e, c] for e, c in ect.items()],
        'contigs': list(contigs),
        'stdout': buf.getvalue().splitlines(),
    }


def assert_same(got, exp, where=''):
    """Bit-exact comparison.  List ORDER is compared too, except edge_count_table whose order
    inside one node's successors is first-appearance order in the reference (dict equality is
    what callers can observe: II only indexes it)."""
    for f in FIELDS:
        if f == 'ect':
            assert sorted(map(tuple, got[f])) == sorted(map(tuple, exp[f])), '%s: ect differs' % where
        else:
            assert got[f] == exp[f], '%s: %s differs\n got %r\n exp %r' % (where, f, got[f][:8], exp[f][:8])
    if 'stdout' in exp and 'stdout' in got:
        assert got['stdout'] == exp['stdout'], '%s: stdout differs' % where


def random_dna_case(r: random.Random, big=False):
    """Small random problems with branches, tips, cycles, duplicates, short and empty reads."""
    alphabet = r.choice(['ACGT', 'ACGT', 'ACG', 'AC'])
    k = r.choice([1, 2, 3, 4, 5, 6, 7, 9, 12] if not big else [15, 21, 31, 32, 33, 47, 63])
    G = r.choice([12, 20, 30, 50, 80, 120, 300]) if not big else r.choice([200, 400, 800])
    genome = ''.join(r.choice(alphabet) for _ in range(G))
    if r.random() < 0.4:
        a = r.randrange(0, G // 2)
        ln = r.randrange(k, k + 8)
        b = r.randrange(G // 2, max(G // 2 + 1, G - ln))
        genome = (genome[:b] + genome[a:a + ln] + genome[b + ln:])[:G]
    L = r.choice([k - 1, k, k + 1, k + 2, 8, 10, 15, 25, 40, 70]) if not big else r.choice([k + 1, k + 5, 100, 150])
    L = max(1, min(L, G))
    n = r.choice([1, 2, 4, 8, 15, 30, 60, 120])
    err = r.choice([0.0, 0.0, 0.02, 0.05, 0.1])
    reads = []
    for _ in range(n):
        ln = max(0, min(G, L + r.choice([0, 0, 0, -1, 1, 3, -3])))
        if r.random() < 0.03:
            ln = 0
        p = r.randrange(0, G - ln + 1)
        s = list(genome[p:p + ln])
        for j in range(len(s)):
            if r.random() < err:
                s[j] = r.choice([c for c in alphabet if c != s[j]])
        reads.append(''.join(s))
    if r.random() < 0.5:
        reads = list(dict.fromkeys(reads))
    threshold = r.choice([1, 2, 2, 3, 5, 1.5, 2.5, 0.5])
    final = r.random() < 0.4
    return reads, k, threshold, final


def synth_reads(seed, G, L, cov, err=0.0, dedup=True):
    """SURVEY.md 8(d) D2 generator (pure `random`, reproducible anywhere)."""
    r = random.Random(seed)
    genome = ''.join(r.choice('ACGT') for _ in range(G))
    n = G * cov // L
    reads = []
    for _ in range(n):
        p = r.randrange(0, G - L + 1)
        s = genome[p:p + L]
        if err:
            s = ''.join(r.choice([c for c in 'ACGT' if c != ch]) if r.random() < err else ch for ch in s)
        reads.append(s)
    if dedup:
        reads = list(dict.fromkeys(reads))              # II_assembleFromReads.py:52
    return genome, reads
