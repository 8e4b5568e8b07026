"""Contig concatenation + substring dedup (III_concatenateOutpus.py:53-146).

* the restatement (oracle/concat_oracle.py) against outputs of the UNMODIFIED script run under random.seed(s)
  (tests/golden/concat.json.gz, made by oracle/make_golden_concat.py);
* the product (one persistent kernel per k, Mersenne Twister on the device) against the same golden outputs and
  against the restatement on fresh random cases: printed lines, merged contigs, their order and the generator state
  afterwards, all identical."""
import gzip
import json
import os
import random

import pytest

from oracle import concat_oracle
from oracle.make_golden_concat import make_case, score_table

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'concat.json.gz')


def _golden():
    with gzip.open(GOLDEN, 'rt') as f:
        return json.load(f)['cases']


def _fasta(outputs, k_hi):
    return ''.join('>SEQUENCE_{}_{}mer\n{}\n'.format(i, k_hi, o) for i, o in enumerate(outputs))   # :148-149


def test_oracle_vs_golden():
    for c in _golden():
        lines = []
        outs = concat_oracle.concatenate_outputs(c['contigs'], dict(c['score_items']), random.Random(c['seed']), out=lines.append)
        assert _fasta(outs, c['k_upperlimit']) == c['fasta'], c['seed']
        assert '\n'.join(lines) + '\n' == c['stdout'], c['seed']


def test_product_vs_golden(backend):
    from dbg_b200.concat import concatenate_outputs
    cases = _golden()
    swaps = 0
    for c in cases:
        lines = []
        outs = concatenate_outputs(c['contigs'], dict(c['score_items']), rng=random.Random(c['seed']), out=lines.append)
        assert '\n'.join(lines) + '\n' == c['stdout'], c['seed']
        assert _fasta(outs, c['k_upperlimit']) == c['fasta'], c['seed']
        swaps += sum('   ' in ln for ln in lines)
    assert swaps > 0                                    # the swap-tolerant arms (:85-131) were exercised


def test_product_vs_oracle_random(backend):
    """Fresh cases, the module-level generator (as the script uses it), and the generator state afterwards."""
    from dbg_b200.concat import concatenate_outputs
    for seed in [11, 12, 13, 14, 15, 16]:
        contigs, rows = make_case(seed, 20 + 10 * (seed % 3), 90 + 40 * (seed % 4), 0.3, 0.15)
        table = score_table(rows, 0.8)
        a, b = [], []
        r = random.Random(seed)
        exp = concat_oracle.concatenate_outputs(contigs, table, r, out=a.append)
        random.seed(seed)
        got = concatenate_outputs(contigs, table, out=b.append)
        assert b == a and got == exp
        assert random.getstate() == r.getstate()


def test_small_ks_and_edges(backend):
    """k = 1 (no swap arm: range(0)), duplicates of current (the FIRST equal contig leaves, :134), a merged string
    that is already in nonconcatenatable, nothing concatenatable at all, a bounded iteration count."""
    from dbg_b200.concat import ConcatSession, concatenate
    cases = [
        (['AB', 'BC', 'AB', 'CA', 'A', ''], [1]),
        (['ABCD', 'CDEF', 'ABCD', 'EFAB', 'XY'], [2, 1]),
        (['AAAA', 'AAAA', 'AAA', 'AAAAA'], [3, 2, 1]),
        (['QWERT', 'YUIOP'], [4, 3]),
        (['ABCDE', 'DCEFG', 'EDABC', 'CDEAB'], [3, 2]),
    ]
    for contigs, ks in cases:
        for seed in range(4):
            for max_iter in (100000, 2):
                a, b = [], []
                exp = list(contigs)
                concat_oracle.concatenate(exp, {'AB': 1.5, 'CD': 2, 'DE': 0.25}, random.Random(seed), ks=ks, out=a.append, max_iter=max_iter)
                with ConcatSession(contigs) as s:
                    concatenate(s, {'AB': 1.5, 'CD': 2, 'DE': 0.25}, rng=random.Random(seed), ks=ks, out=b.append, max_iter=max_iter)
                    got = s.alive_strings()
                    keep = s.dedup()
                    ded = s.strings([i for i in range(len(contigs)) if keep[i]])
                assert (b, got) == (a, exp), (contigs, ks, seed, max_iter)
                assert ded == concat_oracle.dedup_substrings(exp)


def test_arena_growth(backend, monkeypatch):
    """A string arena that starts with no room to spare: the kernel stops inside an iteration, the host reallocates and
    the same iteration goes on (CC_NEED_ARENA)."""
    from dbg_b200.concat import concatenate_outputs
    monkeypatch.setenv('DBG_B200_CC_ARENA', '8')
    c = _golden()[3]
    lines = []
    outs = concatenate_outputs(c['contigs'], dict(c['score_items']), rng=random.Random(c['seed']), out=lines.append)
    assert '\n'.join(lines) + '\n' == c['stdout']
    assert _fasta(outs, c['k_upperlimit']) == c['fasta']


def test_empty_list_raises(backend):
    from dbg_b200.concat import concatenate_outputs
    with pytest.raises(ValueError):
        concatenate_outputs([], {})
