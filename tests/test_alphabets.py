"""Any single-byte alphabet (debruijn.py treats reads as opaque strings): the generic engine picks its code width from
the number of distinct characters -- b = ceil(log2(n)) bits a symbol, b (k + 1) <= 125 -- so peptides keep k <= 24,
DNA with N / IUPAC codes reaches k = 40, text with up to 255 distinct byte values k = 14.  Compared bit for bit, list
order included, with the Python restatement (and with the literal reference where it is present)."""
import random

import pytest

import util
from oracle import dbg_oracle, literal


def _random_case(r, n_symbols, k, first=33):
    al = [chr(c) for c in range(first, first + n_symbols)]
    G = ''.join(r.choice(al) for _ in range(300))
    if r.random() < 0.7:
        G = G[:150] + G[20:20 + k + 5] + G[150:]                      # a repeat: branches, tips
    reads = [G[p:p + k + r.randrange(1, 30)] for p in (r.randrange(0, len(G) - k) for _ in range(40))]
    for i in range(len(reads)):
        if r.random() < 0.3:
            j = r.randrange(len(reads[i]))
            reads[i] = reads[i][:j] + r.choice(al) + reads[i][j + 1:]
    return reads, k, r.choice([1, 2, 3]), r.random() < 0.4


def _cases():
    r = random.Random(3)
    g = ''.join(r.choice('ACGTN') for _ in range(400))
    cases = [
        (['ACGTNACGTTGCA', 'CGTNACGTTGCAT', 'acgtacgtaa'], 4, 2, False),                       # N and lower case
        (['h\xe9llo w\xf6rld h\xe9llo', 'llo w\xf6rld h\xe9l'], 3, 2, False),                  # latin-1 above 0x7F
        (['A\tB\nC A\tB\nC', '\tB\nC A'], 2, 2, False),                                         # control characters
        ([''.join(chr(c) for c in range(0, 255)) + ''.join(chr(c) for c in range(0, 40))], 14, 2, False),   # NUL, 255 symbols
        (['AB' * 30 + 'A', 'BA' * 20], 40, 2, False),                                           # 1-bit codes
        ([g[i:i + 90] for i in range(0, 310, 7)], 31, 2, False),                                # DNA + N at the bench k
        ([g[i:i + 90] for i in range(0, 310, 7)], 40, 2, True),
    ]
    for n_symbols, k in ((2, 61), (3, 61), (4, 40), (8, 40), (9, 30), (16, 30), (17, 24), (33, 19), (64, 19), (65, 16),
                         (128, 16), (129, 14), (200, 14)):
        cases.append(_random_case(r, n_symbols, k))
    return cases


def test_any_single_byte_alphabet(backend):
    for reads, k, thr, final in _cases():
        exp = dbg_oracle.snapshot(reads, k, thr, final, max_paths=500)
        got = util.product_snapshot(reads, k, thr, final)
        util.assert_same(got, exp, '%d symbols, k = %d' % (len(set(''.join(reads))), k))


@pytest.mark.skipif(not literal.available(), reason='the literal reference is not on this machine')
def test_restatement_equals_the_literal_reference_on_these_alphabets():
    for reads, k, thr, final in _cases():
        exp = literal.snapshot(reads, k, thr, final)
        got = dbg_oracle.snapshot(reads, k, thr, final, max_paths=500)
        util.assert_same(got, exp, 'restatement vs literal, %d symbols, k = %d' % (len(set(''.join(reads))), k))


def test_limits_are_reported(backend):
    with pytest.raises(ValueError, match='3-bit codes: 1..40'):
        util.product_snapshot(['ACGTN' * 20], 41, 2, False)
    with pytest.raises(ValueError, match='distinct characters'):
        util.product_snapshot([''.join(chr(0x4E00 + c) for c in range(300))], 2, 2, False)
    with pytest.raises(ValueError, match='256 byte values'):
        util.product_snapshot([''.join(chr(c) for c in range(256)) * 2], 2, 2, False)


def test_lookups_on_a_wide_alphabet(backend):
    import debruijn as db
    import io
    import contextlib
    reads, k, thr, final = _random_case(random.Random(9), 70, 12)
    with contextlib.redirect_stdout(io.StringIO()):
        g, pull, branch, pulled, ect = db.construct_graph(reads, k, threshold=thr, final=False)
        contigs = db.output_contigs(g, branch, pulled)
    exp = dbg_oracle.snapshot(reads, k, thr, False, max_paths=500)
    want = dict(map(tuple, exp['ect']))
    for e, c in want.items():
        assert ect[e] == c
    graph = ect._graph
    keys = list(want)[:50] + ['~' * (k + 1), keys_absent(reads, k)]
    got = graph.lookup_edges(keys[:-2]).tolist()                                                  # the device lookup, not the dict
    assert got == [want[e] for e in keys[:-2]]
    with pytest.raises(KeyError):
        ect[keys[-1]]
    scores = [db.get_score(ect, c, k) for c in contigs]
    assert scores == [sum(want[c[i:i + k + 1]] for i in range(len(c) - k)) for c in contigs]


def test_unicode_reads_travel_as_private_bytes(backend):
    """Characters beyond latin-1 (the reference takes any str): one private byte per distinct character, every string
    that comes back (labels, branch / pulled lists, edge keys, contigs) is mapped back; lookups translate their query."""
    import contextlib
    import io
    import debruijn as db
    r = random.Random(5)
    al = '\u65e5\u672c\u8a9e\u306e\u30c6\u30ad\u30b9\u30c8\u03b1\u03b2\u03b3\u03b4\u20ac\xe9A'
    G = ''.join(r.choice(al) for _ in range(200))
    G = G[:100] + G[10:25] + G[100:]
    reads = [G[p:p + r.randrange(8, 30)] for p in (r.randrange(0, len(G) - 8) for _ in range(60))]
    for k, thr, final in ((3, 2, False), (5, 2, True), (7, 1, False)):
        exp = dbg_oracle.snapshot(reads, k, thr, final, max_paths=500)
        util.assert_same(util.product_snapshot(reads, k, thr, final), exp, 'unicode k = %d' % k)
    with contextlib.redirect_stdout(io.StringIO()):
        g, pull, branch, pulled, ect = db.construct_graph(reads, 3, 2, False)
        contigs = db.output_contigs(g, branch, pulled)
        per_batch = db.construct_graph_batch([reads, reads[:20]], 3, 2, False)
    want = dict(map(tuple, dbg_oracle.snapshot(reads, 3, 2, False)['ect']))
    assert dict(ect) == want and dict(per_batch[0][4]) == want
    assert ect._graph.lookup_edges(list(want)[:20]).tolist() == [want[e] for e in list(want)[:20]]
    assert all(db.get_score(ect, c, 3) == sum(want[c[i:i + 4]] for i in range(len(c) - 3)) for c in contigs)
    with pytest.raises(KeyError):
        ect['ZZZZ']


def keys_absent(reads, k):
    return reads[0][:k] + '\x7f'
