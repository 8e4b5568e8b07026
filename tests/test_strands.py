"""Both strands (DBG_BOTH_STRANDS / construct_graph(..., both_strands=True)): an extension the reference does not have
-- it has no reverse complement anywhere (SURVEY.md 9 delta 1) -- defined so that it HAS an oracle: the result is exactly
what the reference computes on reads + [reverse_complement(r) for r in reads].  Compared bit-exactly (list order
included) with the restatement on that read list; plus the properties of the unfolded canonical graph: every node has
its twin, twins carry mirrored edge counts, exactly one node of a non-palindromic pair is canonical."""
import contextlib
import io
import random

import pytest

import util
from oracle import dbg_oracle


def _rc(s):
    return s.translate(str.maketrans('ACGT', 'TGCA'))[::-1]


def _cases():
    rng = random.Random(11)
    out = []
    for i in range(14):
        reads, k, thr, final = util.random_dna_case(rng, big=i >= 10)
        if set(''.join(reads)) <= set('ACGT'):
            out.append((reads, k, thr, final))
    _, reads = util.synth_reads(5, 1500, 60, 6, err=0.01)
    out.append((reads, 21, 2, False))
    out.append((reads, 33, 2, True))
    return out


def test_both_strands_equals_reference_on_reads_plus_reverse_complements(backend):
    import debruijn as db
    for reads, k, thr, final in _cases():
        both = reads + [_rc(r) for r in reads]
        exp = dbg_oracle.snapshot(both, k, thr, final)
        with contextlib.redirect_stdout(io.StringIO()):
            g, pull, branch, pulled, ect = db.construct_graph(reads, k, threshold=thr, final=final, both_strands=True)
            contigs = db.output_contigs(g, branch, pulled)
        assert [[v, n.indegree, n.outdegree] for v, n in g[0].items()] == exp['nodes'], (k, reads[:3])
        assert dict(ect) == {e: c for e, c in exp['ect']}
        assert list(branch) == exp['branch'] and list(pulled) == exp['pulled']
        assert pull == exp['pull_out_read']
        assert contigs == exp['contigs']


def test_twin_nodes(backend):
    from dbg_b200 import Graph
    for reads, k, thr, final in _cases():
        if not any(len(r) > k for r in reads):
            continue
        g = Graph(k)
        g.build(reads, threshold=thr, both_strands=True)
        labels = list(g.labels())
        t = g.twin_nodes()
        pos = {v: i for i, v in enumerate(labels)}
        assert t['twin'].tolist() == [pos[_rc(v)] for v in labels]           # every node has its twin
        assert t['canonical'].tolist() == [v <= _rc(v) for v in labels]
        cnt = g.nodes()['cnt']
        for i, v in enumerate(labels):                                        # count[v + c] == count[rc(v + c)]
            for c in range(4):
                if cnt[i, c]:
                    e = v + 'ACGT'[c]
                    r = _rc(e)
                    assert cnt[pos[r[:-1]], 'ACGT'.index(r[-1])] == cnt[i, c]
        g.close()
        # a forward-only build may lack twins
        g = Graph(k)
        g.build(reads, threshold=thr)
        labels = list(g.labels())
        pos = {v: i for i, v in enumerate(labels)}
        assert g.twin_nodes()['twin'].tolist() == [pos.get(_rc(v), -1) for v in labels]
        g.close()


def test_both_strands_needs_dna(backend):
    import debruijn as db
    with pytest.raises(ValueError):
        with contextlib.redirect_stdout(io.StringIO()):
            db.construct_graph(['EVQLVESGG', 'QLVESGGGL'], 3, both_strands=True)
