"""True in-degree / out-degree / balance (dbg_degree_balance; north-star kernel 4, SURVEY.md 9 delta 4) against a brute
force over the reads: an EXTRA output -- the reference itself keeps outdegree = distinct successors and a 0/1 first-seen
indegree flag (debruijn.py:129-143), which the parity tests cover."""
import random

from util import random_dna_case


def _brute(reads, k):
    pred, succ, order = {}, {}, []
    for r in reads:
        if len(r) <= k:
            continue
        for i in range(len(r) - k + 1):
            v = r[i:i + k]
            if v not in pred:
                pred[v], succ[v] = set(), set()
                order.append(v)
            if i + k < len(r):
                w = r[i + 1:i + k + 1]
                succ[v].add(w)
        for i in range(len(r) - k):
            pred.setdefault(r[i + 1:i + k + 1], set()).add(r[i:i + k])
    return order, pred, succ


def test_degree_balance(backend):
    from dbg_b200 import Graph
    rng = random.Random(3)
    cases = [random_dna_case(rng)[:3] for _ in range(25)] + [random_dna_case(rng, big=True)[:3] for _ in range(6)]
    cases.append((['ACGTACGTACGTTTGACCA' * 3, 'TTGACCAGGT', 'ACGTACGA'], 21, 2))      # k >= 19: node table laid out by minimizer
    cases.append((['ACGTTGCATGCCGTAAGCTTGACGGTACCATGCA' * 2, 'CCGTAAGCTTGACGGTACCATGCATTTT'], 33, 2))   # 128-bit keys
    peptides = ['EVQLVESGGGLVQPGGSLRLSCAAS', 'LVQPGGSLRLSCAASGFTFS', 'GGSLRLSCAASGYTF', 'QPGGSLRL']
    for reads, k, thr in cases + [(peptides, 4, 2)]:
        if not any(len(r) > k for r in reads):
            continue
        alphabet = None if set(''.join(reads)) <= set('ACGT') else bytes(sorted(set(''.join(reads).encode())))
        g = Graph(k, alphabet=alphabet)
        g.build(reads, threshold=thr)
        order, pred, succ = _brute(reads, k)
        labels = list(g.labels())
        assert labels == order
        d = g.degree_balance()
        assert d['indegree'].tolist() == [len(pred[v]) for v in order]
        assert d['outdegree'].tolist() == [len(succ[v]) for v in order]
        assert d['balance'].tolist() == [len(succ[v]) - len(pred[v]) for v in order]
        g.close()


def test_probe_stats(backend):
    """dbg_probe_stats: every entry counted once, lengths consistent, and the mean where linear probing puts it
    for the table's load (successful search: (1 + 1 / (1 - load)) / 2), for one- and two-word keys."""
    import util
    from dbg_b200 import Graph
    _, reads = util.synth_reads(9, 20000, 100, 12, err=0.01)
    for k in (21, 47):
        g = Graph(k)
        g.build(reads, threshold=2)
        st = g.stats()
        p = g.probe_stats('nodes')
        assert p['entries'] == st.n_nodes == sum(p['hist']) and p['slots'] == st.table_slots
        assert p['p50'] <= p['p99'] <= min(p['max'], 64) and p['max'] < p['slots']
        theory = (1 + 1 / (1 - p['load'])) / 2
        assert 1.0 <= p['mean'] < 1.5 * theory + 0.5
        c = g.probe_stats('chunks')
        assert c['entries'] == sum(c['hist']) > 0 and c['entries'] <= st.chunk_slots and 1.0 <= c['mean'] < 4
        g.close()
    import pytest
    from dbg_b200._native import DbgError
    g = Graph(5, alphabet=b'ACDEFGHIKLMNPQRSTVWY')
    g.build(['ACDEFGHIKLMNPQRSTVWY'], threshold=2)
    with pytest.raises(DbgError):
        g.probe_stats('nodes')
    g.close()
