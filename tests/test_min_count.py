"""min_count -- an EXTENSION the reference does not have (SURVEY.md section 9 delta 5; BASELINE.json configs[2]
"abundance filter min-count=2").  No reference oracle exists for it: the definition is oracle/dbg_oracle.py
build(min_count) ("skip the windows whose (k+1)-mer is rarer"), the C restatement follows it independently, the
two are fuzzed against each other here, and the product is compared with them bit for bit, list order included.
min_count = 1 must be the reference (the golden tests cover that)."""
import random

import pytest

import util
from oracle import c_oracle, dbg_oracle


def _cases(seed, n, big=False):
    r = random.Random(seed)
    out = []
    while len(out) < n:
        reads, k, thr, final = util.random_dna_case(r, big=big)
        mc = r.choice([2, 2, 3, 5])
        try:
            exp = dbg_oracle.snapshot(reads, k, thr, final, max_paths=200, min_count=mc)
        except RuntimeError:
            continue
        out.append((reads, k, thr, final, mc, exp))
    return out


def test_restatements_agree():
    for reads, k, thr, final, mc, exp in _cases(9001, 120):
        got = c_oracle.snapshot(reads, k, thr, final, min_count=mc)
        util.assert_same(got, exp, 'C vs Python oracle, min_count=%d: %r' % (mc, (reads, k, thr, final)))
    # min_count = 1 is the unfiltered reference
    for reads, k, thr, final, mc, _ in _cases(9002, 20):
        try:
            assert dbg_oracle.snapshot(reads, k, thr, final, max_paths=200, min_count=1) == \
                dbg_oracle.snapshot(reads, k, thr, final, max_paths=200)
        except RuntimeError:
            pass                                                       # path explosion of the unfiltered final=True case


def test_filter_really_filters():
    reads = ['ACGTACGTTGCA', 'ACGTACGTTGCA', 'ACGTACGATGCA']
    full = dbg_oracle.snapshot(reads, 4, 2, False)
    filt = dbg_oracle.snapshot(reads, 4, 2, False, min_count=2)
    assert len(filt['nodes']) < len(full['nodes'])
    assert all(c >= 2 for _, c in filt['ect'])
    assert {e for e, c in full['ect'] if c >= 2} == {e for e, _ in filt['ect']}


def test_product_vs_oracle(backend):
    for reads, k, thr, final, mc, exp in _cases(9003, 60) + _cases(9004, 12, big=True):
        got = util.product_snapshot(reads, k, thr, final, min_count=mc)
        util.assert_same(got, exp, 'min_count=%d: %r' % (mc, (reads, k, thr, final)))


@pytest.mark.parametrize('k', [21, 41])
def test_product_vs_oracle_with_errors(backend, k):
    """1 % substitution errors at 12x: most error k-mers occur once and go; what is left is compared in full."""
    _, reads = util.synth_reads(31, 4000, 90, 12, err=0.01)
    for mc in (2, 3):
        exp = dbg_oracle.snapshot(reads, k, 2, False, min_count=mc)
        full = dbg_oracle.snapshot(reads, k, 2, False)
        assert len(exp['nodes']) < len(full['nodes'])
        got = util.product_snapshot(reads, k, 2, False, min_count=mc)
        util.assert_same(got, exp, 'errors, k=%d min_count=%d' % (k, mc))


def test_batched_with_min_count(backend):
    import contextlib
    import io
    import debruijn as db
    batches = []
    for b in range(6):
        _, reads = util.synth_reads(100 + b, 300 + 50 * b, 40, 8, err=0.02)
        batches.append(reads)
    with contextlib.redirect_stdout(io.StringIO()):
        res = db.construct_graph_batch(batches, 15, threshold=2, final=False, min_count=2)
        for b, (g, pull, branch, pulled, ect) in enumerate(res):
            exp = dbg_oracle.snapshot(batches[b], 15, 2, False, min_count=2)
            V, E = g
            got = {'nodes': [[v, n.indegree, n.outdegree] for v, n in V.items()],
                   'edges': [[v, list(s)] for v, s in E.items()], 'pull_out_read': list(pull),
                   'branch': list(branch), 'pulled': list(pulled), 'ect': [[e, c] for e, c in ect.items()],
                   'contigs': list(db.output_contigs(g, branch, pulled))}
            util.assert_same(got, exp, 'batched min_count, batch %d' % b)
