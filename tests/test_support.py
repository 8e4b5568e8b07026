"""findSupportReadScore / the IV sort (IV_sortOutputs.py:10-15, :54) on the GPU against the
restatement, which is itself pinned against the literal reference function when /root/reference is present."""
import importlib.util
import os
import random

import pytest

from oracle import support_oracle

REF = '/root/reference/IV_sortOutputs.py'
AA = 'ACDEFGHIKLMNPQRSTVWY'


def _case(r, alphabet, n_reads, n_contigs, float_scores):
    genome = ''.join(r.choice(alphabet) for _ in range(r.choice([40, 200, 800])))
    table = {}
    for _ in range(n_reads):
        L = r.choice([1, 2, 3, 5, 8, 8, 12, 20])
        if r.random() < 0.7 and len(genome) > L:
            p = r.randrange(0, len(genome) - L)
            s = genome[p:p + L]
        else:
            s = ''.join(r.choice(alphabet) for _ in range(L))
        sc = r.uniform(0.1, 99.9) if float_scores else r.randrange(1, 100)
        table[s] = table.get(s, 0) + sc                      # IV_sortOutputs.py:44-47: scores of equal reads add up
    if r.random() < 0.2:
        table[''] = 1.5 if float_scores else 2               # '' is in every contig
    contigs = []
    for _ in range(n_contigs):
        L = r.choice([0, 3, 10, 30, 120])
        if r.random() < 0.8 and len(genome) > L:
            p = r.randrange(0, len(genome) - L)
            contigs.append(genome[p:p + L])
        else:
            contigs.append(''.join(r.choice(alphabet) for _ in range(L)))
    return contigs, table


def test_literal_reference():
    if not os.path.exists(REF):
        pytest.skip('reference tree not present')
    src = open(REF).read().split("if __name__ == '__main__':")[0]
    src = src.replace('from debruijn import read_reads', '').replace('import pandas as pd', '').replace('from tqdm import trange', '')
    ns = {}
    exec(compile(src, REF, 'exec'), ns)                       # the function text of IV_sortOutputs.py:10-15, unmodified
    r = random.Random(1)
    for i in range(60):
        contigs, table = _case(r, AA if i % 2 else 'ACGT', 40, 12, i % 3 == 0)
        for c in contigs:
            assert ns['findSupportReadScore'](c, table) == support_oracle.find_support_read_score(c, table)


@pytest.mark.parametrize('float_scores', [False, True])
def test_support_scores_vs_oracle(backend, float_scores):
    import debruijn as db
    r = random.Random(7 + float_scores)
    for i in range(25):
        contigs, table = _case(r, AA if i % 2 else 'ACGT', r.choice([0, 1, 30, 200]), r.choice([1, 5, 40]), float_scores)
        exp = [support_oracle.find_support_read_score(c, table) for c in contigs]
        got = db.support_read_scores(contigs, table)
        assert got == [float(x) for x in exp], (i, contigs[:3])
        if contigs:
            assert db.findSupportReadScore(contigs[0], table) == float(exp[0])
        s_got, sc_got = db.sort_by_support(contigs, table)
        s_exp, sc_exp = support_oracle.sort_outputs(contigs, table)
        assert s_got == s_exp and sc_got == [float(x) for x in sc_exp]
