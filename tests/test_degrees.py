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
