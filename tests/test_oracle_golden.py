"""Pin the oracle (oracle/dbg_oracle.py and oracle/dbg_oracle.c) to outputs of the LITERAL
reference: tests/golden/*.json.gz were written by oracle/make_golden.py running
/root/reference/debruijn.py (the reference ships no tests of its own)."""
import hashlib
import json
import random

import pytest

import util
from oracle import c_oracle, dbg_oracle, literal


def digest(obj):
    return hashlib.sha256(json.dumps(obj, separators=(',', ':')).encode()).hexdigest()


def test_python_oracle_kat1():
    for case in util.load_golden('kat.json.gz'):
        got = dbg_oracle.snapshot(case['reads'], case['k'], case['threshold'], case['final'])
        for f in util.FIELDS:
            assert got[f] == case['expect'][f], (case['k'], case['final'], f)


def test_python_oracle_fuzz_corpus():
    for i, case in enumerate(util.load_golden('fuzz.json.gz')):
        got = dbg_oracle.snapshot(case['reads'], case['k'], case['threshold'], case['final'])
        for f in util.FIELDS:
            assert got[f] == case['expect'][f], (i, f)


def test_c_oracle_fuzz_corpus():
    n = 0
    for i, case in enumerate(util.load_golden('fuzz.json.gz')):
        if not util.is_dna(case['reads']):
            with pytest.raises(ValueError):
                c_oracle.snapshot(case['reads'], case['k'], case['threshold'], case['final'])
            continue
        got = c_oracle.snapshot(case['reads'], case['k'], case['threshold'], case['final'])
        util.assert_same(got, {f: case['expect'][f] for f in util.FIELDS}, 'fuzz %d' % i)
        n += 1
    assert n > 400


def _kat2_reads(gen):
    from collections import Counter
    r = random.Random(gen['seed'])
    genome = ''.join(r.choice('ACGT') for _ in range(gen['G']))
    reads = []
    for _ in range(gen['G'] * gen['cov'] // gen['L']):
        p = r.randrange(0, gen['G'] - gen['L'] + 1)
        reads.append(genome[p:p + gen['L']])
    return list(Counter(reads).keys())


@pytest.mark.parametrize('which', ['python', 'c'])
def test_oracles_kat2(which):
    kat2 = util.load_golden('kat2.json.gz')
    reads = _kat2_reads(kat2['generator'])
    assert len(reads) == kat2['n_reads'] and digest(reads) == kat2['reads_sha256']
    snap_fn = dbg_oracle.snapshot if which == 'python' else c_oracle.snapshot
    for case in kat2['cases']:
        got = snap_fn(reads, kat2['k'], kat2['threshold'], case['final'])
        assert len(got['nodes']) == case['n_nodes']
        assert [len(c) for c in got['contigs']] == case['contig_lengths']
        for f in util.FIELDS:
            assert digest(got[f]) == case['sha256'][f], (which, case['final'], f)


def test_c_oracle_vs_python_oracle_random():
    r = random.Random(77)
    for i in range(300):
        reads, k, thr, final = util.random_dna_case(r, big=(i % 5 == 4))
        try:
            exp = dbg_oracle.snapshot(reads, k, thr, final, max_paths=300)
        except RuntimeError:
            continue
        got = c_oracle.snapshot(reads, k, thr, final)
        util.assert_same(got, exp, 'random %d' % i)


@pytest.mark.skipif(not literal.available(), reason='literal reference only exists in the build container')
def test_literal_live():
    """Where /root/reference exists, re-check a fresh random sample against the reference itself."""
    mod = literal.load()
    r = random.Random(4242)
    done = 0
    while done < 60:
        reads, k, thr, final = util.random_dna_case(r)
        if thr == 0.5 and False:
            continue
        try:
            mine = dbg_oracle.snapshot(reads, k, thr, final, max_paths=200)
        except RuntimeError:
            continue
        ref = literal.snapshot(reads, k, thr, final, mod)
        for f in util.FIELDS:
            assert mine[f] == ref[f], (reads, k, thr, final, f)
        done += 1
