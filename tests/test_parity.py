"""Parity of the product (CUDA path behind include/dbg_b200.h, driven through the drop-in
``debruijn`` module) against the oracle and the golden vectors of the literal reference.

Bit-exact: all of this is integer / string work (BASELINE.json north_star).  Every test runs
on the ``backend`` fixture: ``gpu`` = the real library on a B200 (-m gpu), ``emu`` = the same
sources under the CPU emulator (see tests/conftest.py)."""
import hashlib
import json
import random

import numpy as np
import pytest

import util
from oracle import c_oracle, dbg_oracle


def digest(obj):
    return hashlib.sha256(json.dumps(obj, separators=(',', ':')).encode()).hexdigest()


def test_golden_fuzz_corpus(backend):
    """Outputs of /root/reference/debruijn.py itself: the DNA cases run on the 2-bit engine, the others
    (peptides, 2- and 3-letter alphabets count as DNA subsets) on the generic-alphabet engine."""
    cases = util.load_golden('fuzz.json.gz')
    n_dna = n_other = 0
    for i, case in enumerate(cases):
        got = util.product_snapshot(case['reads'], case['k'], case['threshold'], case['final'])
        util.assert_same(got, case['expect'], 'golden fuzz %d' % i)
        if util.is_dna(case['reads']):
            n_dna += 1
        else:
            n_other += 1
    assert n_dna > 400 and n_other >= 75


def test_kat1_peptides_of_the_reference(backend):
    """KAT-1: the only example input in the reference, II_assembleFromReads.py:55
    (['EVQLVE','QLVAPG','LVESGGAL','LVESGGGL']), outputs recorded from the literal debruijn.py."""
    cases = util.load_golden('kat.json.gz')
    assert cases
    for case in cases:
        got = util.product_snapshot(case['reads'], case['k'], case['threshold'], case['final'])
        util.assert_same(got, case['expect'], 'KAT-1 k=%d final=%s' % (case['k'], case['final']))


def test_generic_alphabet_vs_oracle(backend):
    """Random peptide-like problems (20 letters, k 2..12, errors, duplicates, branches, cycles, final=True):
    the generic-alphabet engine against the pure-Python restatement; plus DNA with N's, lower case, and the
    same DNA problem forced through the generic engine must equal the 2-bit engine."""
    import debruijn as db
    from dbg_b200 import Graph
    r = random.Random(20241019)
    aa = 'ACDEFGHIKLMNPQRSTVWY'
    n = 0
    while n < 120:
        alphabet = r.choice([aa, aa, aa[:6], 'ACGTN', 'acgtACGT', 'XY'])
        k = r.choice([2, 3, 4, 5, 6, 8, 10, 12])
        G = r.choice([15, 30, 60, 150, 400])
        genome = ''.join(r.choice(alphabet) for _ in range(G))
        if r.random() < 0.4 and G > 2 * k + 10:
            a = r.randrange(0, G // 2)
            ln = r.randrange(k, k + 6)
            b = r.randrange(G // 2, max(G // 2 + 1, G - ln))
            genome = (genome[:b] + genome[a:a + ln] + genome[b + ln:])[:G]
        L = max(1, min(G, r.choice([k, k + 1, k + 3, 12, 20, 35])))
        reads = []
        for _ in range(r.choice([1, 3, 8, 20, 60, 150])):
            ln = max(0, min(G, L + r.choice([0, 0, -1, 2, -3])))
            p = r.randrange(0, G - ln + 1)
            s_ = list(genome[p:p + ln])
            for j in range(len(s_)):
                if r.random() < r.choice([0.0, 0.03, 0.08]):
                    s_[j] = r.choice(alphabet)
            reads.append(''.join(s_))
        if r.random() < 0.5:
            reads = list(dict.fromkeys(reads))
        thr, final = r.choice([1, 2, 2, 3, 5, 1.5, 0.5]), r.random() < 0.4
        try:
            exp = dbg_oracle.snapshot(reads, k, thr, final, max_paths=300)
        except RuntimeError:
            continue
        got = util.product_snapshot(reads, k, thr, final)
        util.assert_same(got, exp, 'generic case %d: %r' % (n, (reads, k, thr, final)))
        n += 1
    # a DNA problem through both engines
    _, reads = util.synth_reads(9, 1500, 50, 12, err=0.02)
    a = util.product_snapshot(reads, 15, 2, False)
    g = Graph(15, alphabet=b'ACGT')
    g.build(reads, threshold=2, final=False)
    g.walk(stop_at_branch=True)
    assert g.contig_strings() == a['contigs']
    assert g.contig_scores().tolist() == [db.get_score(dict(map(tuple, a['ect'])), c, 15) for c in a['contigs']]
    g.close()
    # getScore through the device lookup (II_assembleFromReads.py:14-18) on a peptide graph
    pep = ['EVQLVESGGALVQPGG', 'VQLVESGGGLVQ', 'LVESGGALVQPGGSLR', 'ESGGGLVQPGGSLRLS']
    gr, pull, branch, pulled, ect = db.construct_graph(pep, 5, threshold=2, final=True)
    for c in db.output_contigs(gr, branch, pulled):
        assert db.get_score(ect, c, 5) == sum(dict(ect)[c[i:i + 6]] for i in range(len(c) - 5))
    with pytest.raises(KeyError):
        ect['WWWWWW']


def test_invalid_input_raises(backend):
    """No CPU fallback and no silent truncation: more than 255 distinct characters, or k beyond what the engine's
    keys hold, raise (tests/test_alphabets.py has the alphabets that do work)."""
    import debruijn as db
    with pytest.raises(ValueError):
        db.construct_graph([''.join(chr(0x4E00 + c) for c in range(300))], 3, 2)
    with pytest.raises(ValueError):
        db.construct_graph(['EVQLVESGGALVQPGGSLRLSCAASGFTFS' * 3], 31, 2)     # 12 letters, 4-bit codes: k <= 30
    with pytest.raises(ValueError):
        db.construct_graph(['ACDEFGHIKLMNPQRSTVWY' * 3], 25, 2)               # 20 letters, 5-bit codes: k <= 24
    with pytest.raises(ValueError):
        db.construct_graph(['ACGT' * 40], 64, 2)


def test_kat2_config1_shape(backend):
    """KAT-2 = BASELINE.json configs[0]: 10 kb genome, 100 bp x 20x, k=31, threshold 2."""
    from test_oracle_golden import _kat2_reads
    kat2 = util.load_golden('kat2.json.gz')
    reads = _kat2_reads(kat2['generator'])
    for case in kat2['cases']:
        got = util.product_snapshot(reads, kat2['k'], kat2['threshold'], case['final'])
        assert got['stdout'] == case['stdout']
        assert [len(c) for c in got['contigs']] == case['contig_lengths']
        for f in util.FIELDS:
            assert digest(got[f]) == case['sha256'][f], (case['final'], f)


@pytest.mark.parametrize('seed,big', [(1, False), (2, False), (3, True)])
def test_random_vs_oracle(backend, seed, big):
    r = random.Random(1000 + seed)
    n = 0
    while n < (150 if not big else 40):
        reads, k, thr, final = util.random_dna_case(r, big=big)
        try:
            exp = dbg_oracle.snapshot(reads, k, thr, final, max_paths=200)
        except RuntimeError:
            continue
        got = util.product_snapshot(reads, k, thr, final)
        util.assert_same(got, exp, 'random seed %d case %d: %r' % (seed, n, (reads, k, thr, final)))
        n += 1


def test_k1_every_base_is_a_node(backend):
    """k = 1: no non-node position separates the nodes of two reads, so the end of a read's last chunk
    must come from the read structure (the masks alone gave a phantom edge 'CG' for
    ['AC', 'GT'])."""
    r = random.Random(4242)
    for _ in range(60):
        alphabet = r.choice(['ACGT', 'AC', 'ACG'])
        reads = [''.join(r.choice(alphabet) for _ in range(r.choice([0, 1, 2, 3, 5, 9, 33, 40]))) for _ in range(r.choice([1, 2, 5, 20, 70]))]
        thr, final = r.choice([1, 2, 3, 0.5]), r.random() < 0.3
        try:
            exp = dbg_oracle.snapshot(reads, 1, thr, final, max_paths=200)
        except RuntimeError:
            continue
        got = util.product_snapshot(reads, 1, thr, final)
        util.assert_same(got, exp, 'k=1 %r' % ((reads, thr, final),))


def test_edge_cases(backend):
    import debruijn as db
    cases = [
        ([], 5, 2, False),                                  # no reads at all
        ([''], 5, 2, False),                                # one empty read
        (['AC', 'GT'], 1, 2, False),                        # k = 1: every base is a node; no edge joins two reads
        (['A', 'C', 'ACGT', 'T', 'TTTT', 'GA'], 1, 2, True),
        (['ACG', 'AC', ''], 3, 2, False),                   # every read has len <= k (debruijn.py:126)
        (['ACGT'], 3, 2, False),                            # exactly one window
        (['AAAAAAAAAA'], 3, 2, False),                      # self loop: the only start is on a cycle (:289)
        (['ACGTACGTACGTACGT'], 4, 2, False),                # pure cycle
        (['ACGTACGTACGTACGTT'], 4, 2, True),                # cycle with an exit
        (['ACGTTGCA'] * 5, 4, 2, False),                    # duplicates kept (II dedups, the module must not)
        (['A' * 70 + 'C', 'A' * 69 + 'CG'], 63, 2, False),  # k = 63: 126-bit keys
        (['ACGT' * 40], 33, 2, True),                       # 128-bit keys, cyclic
        (['ACGTAGCTAGCTAGGATCGATCGTAGCTAGCATCGA'], 31, 2, False),  # k = 31: 62-bit keys
        (['ACGTAGCTAGCTAGGATCGATCGTAGCTAGCATCGA'], 32, 2, False),  # k = 32: first 128-bit k
        (['ACGTA', 'ACGTC', 'ACGTG', 'ACGTT'], 4, 1, True),  # four-way tie at threshold 1
        (['ACGTA', 'ACGTC', 'ACGTC', 'ACGTG'], 4, 0.5, False),  # threshold < 1 drops ties
    ]
    for reads, k, thr, final in cases:
        exp = dbg_oracle.snapshot(reads, k, thr, final)
        got = util.product_snapshot(reads, k, thr, final)
        util.assert_same(got, exp, repr((reads, k, thr, final)))
    for bad_k in (0, 64, -1):
        with pytest.raises(ValueError):
            db.construct_graph(['ACGTACGT'], bad_k, 2)


def test_threshold_zero_raises_like_reference(backend):
    """debruijn.py:163 divides by threshold: ZeroDivisionError as soon as one node has two successors."""
    import debruijn as db
    with pytest.raises(ZeroDivisionError):
        db.construct_graph(['ACGTA', 'ACGTC'], 4, 0)
    g, pull, br, pulled, ect = db.construct_graph(['ACGTA'], 4, 0)     # no multi-successor node: fine
    assert db.output_contigs(g, br, pulled) == ['ACGTA']


def test_reads_not_mutated_and_repeatable(backend):
    import debruijn as db
    _, reads = util.synth_reads(7, 2000, 60, 12, err=0.02)
    before = list(reads)
    a = util.product_snapshot(reads, 15, 2, False)
    b = util.product_snapshot(reads, 15, 2, False)
    assert reads == before
    for f in util.FIELDS:
        assert a[f] == b[f]


def test_ii_multi_k_loop(backend):
    """The caller's loop, II_assembleFromReads.py:52-75: dedup, k loop, getScore sort,
    extend(pull_out_read), final=True on the last k."""
    import contextlib
    import io
    import debruijn as db
    _, reads = util.synth_reads(11, 1500, 50, 15, err=0.01)

    def run(mod, get_score):
        seqs = list(dict.fromkeys(reads))
        out = None
        with contextlib.redirect_stdout(io.StringIO()):
            for k in range(12, 16):
                final = k == 15
                g, pull, branch, pulled, ect = mod.construct_graph(seqs, k, threshold=2, final=final)
                contigs = mod.output_contigs(g, branch, pulled)
                contigs.sort(key=lambda x: get_score(ect, x, k), reverse=True)
                if final:
                    out = contigs
                    break
                seqs = contigs
                seqs.extend(pull)
        return out

    def ref_score(ect, contig, k):                          # II_assembleFromReads.py:14-18 verbatim semantics
        return sum(ect[contig[i:i + k + 1]] for i in range(len(contig) - k))

    class OracleMod:
        construct_graph = staticmethod(lambda r, k, threshold, final: dbg_oracle.construct_graph(r, k, threshold, final, verbose=False))
        output_contigs = staticmethod(lambda g, b, p: dbg_oracle.output_contigs(g, b, p, verbose=False))

    exp = run(OracleMod, ref_score)
    got = run(db, ref_score)
    assert got == exp
    got2 = run(db, db.get_score)                            # batched device lookup gives the same order
    assert got2 == exp


def test_descriptors_scores_and_lookup(backend):
    """C-ABI level: descriptors, device getScore and edge lookups against the C oracle."""
    from dbg_b200 import Graph
    _, reads = util.synth_reads(5, 3000, 80, 15, err=0.01)
    k = 21
    g = Graph(k)
    g.build(reads, threshold=2, final=False)
    g.walk(stop_at_branch=True)
    res = c_oracle.run(reads, k, 2, False)
    st = g.stats()
    assert (st.n_nodes, st.n_windows, st.n_starts) == (res.n_nodes, res.n_windows, res.n_starts)
    assert st.n_branch == res.n_branch_internal and st.n_pulled == len(res.pulled)
    nd = g.nodes()
    assert np.array_equal(nd['tag'], res.tag)
    assert np.array_equal(nd['key_lo'], res.key_lo) and np.array_equal(nd['key_hi'], res.key_hi)
    assert np.array_equal(nd['cnt'], res.cnt)
    desc = g.contig_descriptors()
    assert len(desc) == len(res.contigs)
    id_of_slot = {int(s): i for i, s in enumerate(nd['id'])}
    assert [id_of_slot[int(s)] for s in desc['start']] == res.contigs[:, 0].tolist()
    assert desc['n_nodes'].tolist() == res.contigs[:, 1].tolist()
    assert [id_of_slot[int(s)] for s in desc['terminal']] == res.contigs[:, 2].tolist()
    assert desc['kind'].tolist() == res.contigs[:, 3].tolist()
    strings = g.contig_strings()
    assert strings == res.contig_strings
    ect = {lab + 'ACGT'[c]: int(res.cnt[v, c]) for v, lab in enumerate(g.labels()) for c in range(4) if res.cnt[v, c]}
    exp_scores = [sum(ect[s[i:i + k + 1]] for i in range(len(s) - k)) for s in strings]
    assert g.contig_scores().tolist() == exp_scores
    qs = list(ect)[:500] + ['A' * (k + 1), 'ACGT' * 5 + 'AC']
    got = g.lookup_edges(qs).tolist()
    assert got == [ect.get(q, 0) for q in qs]
    g.close()


@pytest.mark.parametrize('grid', ['1', '3', ''])
def test_partition_passes(backend, monkeypatch, grid):
    """The table build replays the windows once per L2-sized table region; force many passes,
    and few blocks so that every warp carries its queues across many tiles."""
    from dbg_b200 import Graph
    monkeypatch.setenv('DBG_B200_L2_REGION_BYTES', '16384')
    monkeypatch.setenv('DBG_B200_REGIONS', '0')                 # the plain hash layout is the one that needs passes
    monkeypatch.setenv('DBG_B200_SLICE_BYTES', '8192')        # the host reads arrive in many slices
    if grid:
        monkeypatch.setenv('DBG_B200_INSERT_GRID', grid)
    _, reads = util.synth_reads(21, 3000, 70, 10, err=0.01)
    for k in ((19, 41) if backend == 'gpu' else (19,) if grid == '1' else (41,)):
        exp = dbg_oracle.snapshot(reads, k, 2, False)
        got = util.product_snapshot(reads, k, 2, False)
        util.assert_same(got, exp, 'partition passes k=%d' % k)
        g = Graph(k)
        g.build(reads, threshold=2)
        assert g.stats().n_insert_passes > 8
        g.close()


def test_batched_equals_per_batch_oracle(backend):
    """BASELINE.json configs[3] shape: many independent small read sets assembled in ONE pass
    (dbg_build_batched) must give, per batch, exactly what construct_graph gives on that batch alone
    -- including batches that share k-mers, empty batches and batches without any window."""
    import contextlib
    import io
    import debruijn as db
    r = random.Random(77)
    for k, final in ((9, False), (25, False), (25, True), (31, False), (40, False)):
        batches = []
        shared = ''.join(r.choice('ACGT') for _ in range(120))          # the same stretch in several batches
        for b in range(12):
            G = r.choice([60, 150, 300])
            genome = ''.join(r.choice('ACGT') for _ in range(G))
            if b % 3 == 0:
                genome = genome[:G // 2] + shared + genome[G // 2:]
            L = r.choice([k + 3, 50, 70])
            reads = []
            for _ in range(r.choice([0, 3, 20, 40])):
                p = r.randrange(0, max(1, len(genome) - L))
                s = list(genome[p:p + L])
                for j in range(len(s)):
                    if r.random() < 0.03:
                        s[j] = r.choice('ACGT')
                reads.append(''.join(s))
            if b == 5:
                reads = ['ACG', '']                                    # nothing longer than k
            batches.append(reads)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            res = db.construct_graph_batch(batches, k, threshold=2, final=final)
            got = []
            for (g, pull, branch, pulled, ect) in res:
                contigs = db.output_contigs(g, branch, pulled)
                V, E = g
                got.append({'nodes': [[v, n.indegree, n.outdegree] for v, n in V.items()],
                            'edges': [[v, list(s)] for v, s in E.items()], 'pull_out_read': list(pull),
                            'branch': list(branch), 'pulled': list(pulled), 'ect': [[e, c] for e, c in ect.items()],
                            'contigs': list(contigs)})
        for b, reads in enumerate(batches):
            exp = dbg_oracle.snapshot(reads, k, 2, final, max_paths=2000)
            util.assert_same(got[b], exp, 'batched k=%d final=%s batch %d' % (k, final, b))
        # get_score through the batched device lookup (II_assembleFromReads.py:14-18)
        g, pull, branch, pulled, ect = res[0]
        for c in got[0]['contigs'][:5]:
            assert db.get_score(ect, c, k) == sum(dict(map(tuple, got[0]['ect']))[c[i:i + k + 1]] for i in range(len(c) - k))
    from dbg_b200 import Graph
    with pytest.raises(ValueError):                                      # 5 batches need 3 bits, a 64-bit key leaves 2
        Graph(31).build_batched([['ACGT' * 20]] * 5)


@pytest.mark.parametrize('table_only', [False, True])
def test_undersized_tables_are_rebuilt(backend, monkeypatch, table_only):
    """Chunk table and node table both start far too small: the build notices (the node table's flag
    is only looked at by the next stage's counters) and rebuilds them larger, results unchanged."""
    from dbg_b200 import Graph
    monkeypatch.setenv('DBG_B200_CHUNK_SLOTS', '512')
    monkeypatch.setenv('DBG_B200_NODE_SLOTS', '4096')
    _, reads = util.synth_reads(23, 4000, 70, 8, err=0.01)
    for k in (21, 40):
        if table_only:
            g = Graph(k)
            g.build_table(reads)
            n = int(g.export_counts(1)[0])
            exp = dbg_oracle.snapshot(reads, k, 2, False)
            assert n == len(exp['nodes'])
            g.close()
        else:
            exp = dbg_oracle.snapshot(reads, k, 2, False)
            got = util.product_snapshot(reads, k, 2, False)
            util.assert_same(got, exp, 'undersized tables k=%d' % k)


def test_remembered_table_sizes(backend, monkeypatch):
    """A handle remembers the table sizes of its last build and sizes the next one from them without the
    sketch round trips (dbg_handle::hint); the result must be the same whether the sizes fit (same reads
    again), are too small (a read set of about the same length with many more distinct k-mers: the build
    notices the overflow afterwards and redoes itself cold) or too large, and with the hint switched off."""
    from dbg_b200 import Graph
    r = random.Random(77)
    _, clean = util.synth_reads(51, 1200, 150, 40, err=0.0)          # few distinct k-mers for its size
    noisy = [''.join(r.choice('ACGT') for _ in range(150)) for _ in range(len(clean) + 50)]   # all distinct, same size
    _, small = util.synth_reads(53, 900, 150, 42, err=0.01)
    monkeypatch.setenv('DBG_B200_L2_REGION_BYTES', '262144')         # several node-table passes: the self-hashing k_expand
    monkeypatch.setenv('DBG_B200_REGION_SLOTS', '256')               # (k = 41 lays the table out by minimizer: many small regions)
    for k in (19, 41):
        if k == 41:
            monkeypatch.setenv('DBG_B200_SLICE_BYTES', '16384')       # host reads arrive in slices: phase 1 runs behind the copies
        g = Graph(k)
        slots = []
        for reads in ((clean, clean, noisy, noisy, small, clean) if k == 19 or backend == 'gpu' else (clean, clean, noisy)):
            exp = dbg_oracle.snapshot(reads, k, 2, False)
            g.build(reads, threshold=2, final=False)
            g.walk(stop_at_branch=True)
            assert g.contig_strings() == exp['contigs']
            assert g.labels() == [n[0] for n in exp['nodes']]
            st = g.stats()
            slots.append((int(st.chunk_slots), int(st.table_slots)))
            assert st.n_windows == sum(max(0, len(x) - k) for x in reads)
            assert [r_ for r_, f in zip(reads, g.pull_read_flags().tolist()) if f] == exp['pull_out_read']
        assert slots[2][1] > 2 * slots[1][1]                          # the remembered sizes did not fit: rebuilt larger
        assert st.n_insert_passes > 1 or st.n_regions > 1
        monkeypatch.setenv('DBG_B200_NO_HINT', '1')
        g.build(clean, threshold=2, final=False)
        g.walk(stop_at_branch=True)
        assert g.contig_strings() == dbg_oracle.snapshot(clean, k, 2, False)['contigs']
        monkeypatch.delenv('DBG_B200_NO_HINT')
        g.close()


@pytest.mark.parametrize('region_slots', ['16', '512'])
def test_region_layout(backend, monkeypatch, region_slots):
    """Node table laid out by minimizer (TGeom): forced on for every k, with regions so small that crowded ones
    spill into their neighbours; every lookup (successor links, tips, junctions, edge counts) goes through the
    region home of its key.  Same results as the oracle, single build pass."""
    from dbg_b200 import Graph
    monkeypatch.setenv('DBG_B200_REGIONS', '1')
    monkeypatch.setenv('DBG_B200_REGION_SLOTS', region_slots)
    if region_slots == '16':
        monkeypatch.setenv('DBG_B200_VB_BITS', '3')                  # 8 buckets of 16: the order array is nearly all spill
    r = random.Random(4242)
    n = 0
    while n < 40:
        reads, k, thr, final = util.random_dna_case(r, big=(n % 4 == 3))
        try:
            exp = dbg_oracle.snapshot(reads, k, thr, final, max_paths=200)
        except RuntimeError:
            continue
        got = util.product_snapshot(reads, k, thr, final)
        util.assert_same(got, exp, 'region layout case %d: %r' % (n, (reads, k, thr, final)))
        n += 1
    _, reads = util.synth_reads(29, 5000, 80, 12, err=0.01)
    for k in (17, 31, 45):
        exp = dbg_oracle.snapshot(reads, k, 2, False)
        got = util.product_snapshot(reads, k, 2, False)
        util.assert_same(got, exp, 'region layout k=%d' % k)
        g = Graph(k)
        g.build(reads, threshold=2)
        st = g.stats()
        assert st.n_regions > 8 and st.n_insert_passes == 1
        ect = dict(map(tuple, exp['ect']))
        qs = list(ect)[:50] + ['A' * (k + 1)]
        assert g.lookup_edges(qs).tolist() == [ect.get(q, 0) for q in qs]
        g.close()


def test_handle_pool_reuses_handles_between_calls(backend, monkeypatch):
    """debruijn.construct_graph keeps the handle of a finished (garbage-collected) call for the next call with the same
    k and alphabet (dbg_create + dbg_destroy cost more than a small build).  A reused handle remembers buffers and table
    sizes of an unrelated read set: results must not depend on what it held before."""
    import gc
    import debruijn as db
    from dbg_b200 import graph as G
    r = random.Random(99)
    sizes = [200, 4000, 150, 30000, 300, 300, 12000, 100]
    created = []
    lib = G.N.get()
    real = lib.dbg_create_ex
    monkeypatch.setattr(lib, 'dbg_create_ex', lambda *a: (created.append(1), real(*a))[1])
    with G._POOL_LOCK:
        G._POOL.clear()
    for i, gsize in enumerate(sizes):
        _, reads = util.synth_reads(100 + i, gsize, 60, 12, err=r.choice([0.0, 0.02]))
        exp = dbg_oracle.snapshot(reads, 21, 2, False)
        got = util.product_snapshot(reads, 21, 2, False)
        util.assert_same(got, exp, 'pooled handle, call %d' % i)
        gc.collect()
    assert len(created) <= 2, 'every call created a handle: the pool is not used'
    pep = ['EVQLVESGGALVQPGG', 'VQLVESGGGLVQ', 'LVESGGALVQPGGSLR', 'ESGGGLVQPGGSLRLS']
    for _ in range(3):
        util.assert_same(util.product_snapshot(pep, 5, 2, False), dbg_oracle.snapshot(pep, 5, 2, False), 'pooled generic handle')
        gc.collect()
    # a caller that changes k with every call (II's loop): the pool never holds more than POOL_TOTAL handles, the oldest go
    _, reads = util.synth_reads(5, 2000, 60, 10)
    for k in list(range(30, 18, -1)) + [20, 20]:
        util.assert_same(util.product_snapshot(reads, k, 2, False), dbg_oracle.snapshot(reads, k, 2, False), 'k = %d' % k)
        gc.collect()
    assert sum(map(len, G._POOL.values())) <= G.POOL_TOTAL
    assert any(key[2] == 20 for key, v in G._POOL.items() if v) and not any(key[2] == 30 for key, v in G._POOL.items() if v)
    monkeypatch.setenv('DBG_B200_HANDLE_POOL', '0')
    with G._POOL_LOCK:
        G._POOL.clear()
    n0 = len(created)
    for i in range(3):
        util.product_snapshot(['ACGTACGGTCA', 'CGTACGGTCAT'], 4, 2, False)
        gc.collect()
    assert len(created) == n0 + 3 and not any(G._POOL.values())


def test_reused_handle_that_fails_is_replaced(backend, monkeypatch):
    """An internal error on a handle taken from the pool is answered with one more try on a new handle."""
    import gc
    import debruijn as db
    from dbg_b200 import graph as G
    from dbg_b200._native import DbgError
    reads = ['ACGTACGGTCATTGCA', 'CGTACGGTCATTGCAG', 'GTACGGTCATTGCAGT']
    exp = dbg_oracle.snapshot(reads, 5, 2, False)
    util.assert_same(util.product_snapshot(reads, 5, 2, False), exp, 'first call')
    gc.collect()
    assert any(key[2] == 5 for key, v in G._POOL.items() if v)
    real = G.Graph.build
    failed = []

    def build(self, *a, **kw):
        if self.from_pool and not failed:
            failed.append(1)
            raise DbgError(-5, 'injected')
        return real(self, *a, **kw)
    monkeypatch.setattr(G.Graph, 'build', build)
    util.assert_same(util.product_snapshot(reads, 5, 2, False), exp, 'call whose pooled handle failed')
    assert failed
