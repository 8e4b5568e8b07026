"""Drop-in for the reference's ``debruijn.py`` on the path ``II_assembleFromReads.py:58-63`` drives.

Put this directory on ``sys.path`` ahead of the reference's and ``import debruijn as db`` /
``from debruijn import read_reads`` (II_assembleFromReads.py:11-12, III:10, IV:8) resolve here.
Same names, argument meaning, return tuple and stdout lines as the reference:

    construct_graph(reads, k, threshold=3, final=False)
        -> ((vertices, edges), pull_out_read, branch_kmer, already_pull_out, edge_count_table)
    output_contigs(g, branch_kmer, already_pull_out) -> list[str]
    read_reads(fname) -> list[str]

All graph work runs in the CUDA library behind ``include/dbg_b200.h`` (sm_100a); this module
only packs the read list, calls the library and wraps the results.  ``vertices``, ``edges`` and
``edge_count_table`` are lazy mappings backed by the graph in HBM (II never inspects the first
two and only indexes the third); they compare equal to the reference's dicts when materialised.
There is no CPU path: without the built library or a CUDA device these calls raise.  DNA reads
(ACGT) run on the 2-bit engine, 1 <= k <= 63; reads over any other alphabet of single-byte characters (peptides,
DNA with N / IUPAC codes, text) on the generic engine, b = ceil(log2(distinct characters)) bits a symbol and
b (k + 1) <= 125: peptides k <= 24, ACGT + N k <= 40, 255 distinct characters k <= 14.  Reads with characters beyond
latin-1 (any Unicode text) travel as one private byte per distinct character (up to 255) and come back as they were.
"""
from __future__ import annotations

import numpy as np

from dbg_b200._native import AlphabetError, DbgError
from dbg_b200.graph import (BatchGraph, EdgeCountTable, Edges, Graph, KmerList, Node, NotSingleByteError, Vertices,  # noqa: F401
                            pack_reads)

__all__ = ['Node', 'read_reads', 'construct_graph', 'construct_graph_batch', 'output_contigs', 'get_score',
           'support_read_scores', 'findSupportReadScore', 'sort_by_support', 'concatenate_outputs']


def read_reads(fname):
    """debruijn.py:22-32: every line not starting with '>' is one read, right-stripped."""
    reads = []
    with open(fname, 'r') as f:
        for line in f:
            if line[0] != '>':
                reads.append(line.rstrip())
    return reads


_COMP = str.maketrans('ACGT', 'TGCA')


def reverse_complement(read):
    return read.translate(_COMP)[::-1]


def _build_pooled(make, **kw):
    """Graph.build on a handle from the pool; an internal error on a REUSED handle (never seen; the handle then
    carries an error and is destroyed, not kept) is answered with one more try on a new handle."""
    g = make(True)
    try:
        return g.build(**kw)
    except (DbgError, MemoryError):
        reused = g.from_pool
        g.close()
        if not reused:
            raise
    except BaseException:
        g.close()
        raise
    return make(False).build(**kw)


def _graph_for(reads, k, threshold, final, min_count=1, both_strands=False):
    """Build on the DNA engine; reads over another alphabet (the reference's real inputs are peptides,
    II_assembleFromReads.py:43-47) go to the generic-alphabet engine with the symbols they use."""
    try:
        packed = pack_reads(reads)
    except NotSingleByteError:                                                 # characters beyond latin-1: a private byte per character
        if both_strands:
            raise ValueError('dbg_b200: both_strands needs DNA reads (ACGT)') from None
        if min_count != 1:
            raise ValueError('dbg_b200: min_count exists for the DNA engine only') from None
        charset = sorted(set(''.join(reads)))
        if len(charset) > 255:
            raise ValueError('dbg_b200: the reads use %d distinct characters; the GPU path handles up to 255' % len(charset)) from None
        fwd = {ord(c): i for i, c in enumerate(charset)}
        packed = pack_reads([r.translate(fwd) for r in reads])
        return _build_pooled(lambda pooled: Graph(k, charset=charset, pooled=pooled), reads=reads, threshold=threshold,
                             final=final, packed=packed)
    try:
        return _build_pooled(lambda pooled: Graph(k, pooled=pooled), reads=reads, threshold=threshold, final=final,
                             packed=packed, min_count=min_count, both_strands=both_strands)
    except AlphabetError:
        pass
    if both_strands:
        raise ValueError('dbg_b200: both_strands needs DNA reads (ACGT)')
    symbols = np.flatnonzero(np.bincount(packed[0], minlength=256)).astype(np.uint8)
    if len(symbols) > 255:
        raise ValueError('dbg_b200: the reads use all 256 byte values; the GPU path handles alphabets of up to 255')
    return _build_pooled(lambda pooled: Graph(k, alphabet=symbols.tobytes(), pooled=pooled), reads=reads, threshold=threshold,
                         final=final, packed=packed, min_count=min_count)


def construct_graph(reads, k, threshold=3, final=False, min_count=1, both_strands=False):
    """debruijn.py:206-285 on the GPU.  `min_count` > 1 is an extension the reference does not have (an absolute
    abundance filter on the (k+1)-mers before the window loop, dbg_set_min_count in include/dbg_b200.h);
    `both_strands=True` is another (DBG_BOTH_STRANDS): the result is exactly
    construct_graph(reads + [reverse_complement(r) for r in reads], ...), made on the device -- every k-mer next to its
    reverse complement (Graph.twin_nodes), the canonical graph unfolded.  The defaults are the reference."""
    g = _graph_for(reads, k, threshold, final, min_count, both_strands)
    if both_strands:
        reads = _BothStrands(reads)
    st = g.stats()
    print('number of {}mer: '.format(k), st.n_nodes)                      # debruijn.py:224
    print('branch number: ', st.n_branch)                                 # debruijn.py:236
    if final:                                                              # debruijn.py:281-283
        pull_out_read = []
        branch_kmer = KmerList((), g)
    else:
        pull_out_read = [r for r, f in zip(reads, g.pull_read_flags().tolist()) if f] if st.n_pull_reads else []
        branch_kmer = KmerList(g.label_of_ids(g.branch_ids()), g)
    already_pull_out = KmerList(g.label_of_ids(g.pulled_ids()), g)
    return (Vertices(g), Edges(g)), pull_out_read, branch_kmer, already_pull_out, EdgeCountTable(g)


class _BothStrands:
    """reads + their reverse complements as a sequence (the second half made on demand)."""

    def __init__(self, reads):
        self.reads = reads

    def __len__(self):
        return 2 * len(self.reads)

    def __iter__(self):
        yield from self.reads
        for r in self.reads:
            yield reverse_complement(r)


def construct_graph_batch(batches, k, threshold=3, final=False, min_count=1):
    """construct_graph for MANY independent read lists at once (BASELINE.json configs[3]: one small
    assembly per template, V_FindTemplateGroups.py:152-190): `batches` is a list of read lists and
    the result is a list with, per batch, exactly the tuple construct_graph(batch, k, threshold, final)
    returns -- one pass over all reads on the GPU (dbg_build_batched: the batch id rides above the
    k-mer in every node key).  Prints the reference's two lines per batch, in batch order."""
    batch_bits = len(batches).bit_length()                              # ids 0 .. n-1 and the all-ones pattern stays free
    g = Graph(k, key_words=2 if 2 * k + batch_bits > 64 else 0)          # 128-bit keys when 64 bits cannot hold both
    try:
        g.set_min_count(min_count)
        g.build_batched(batches, threshold=threshold, final=final)
    except (AlphabetError, NotSingleByteError):                          # not DNA: one generic-alphabet build per batch
        g.close()
        return [construct_graph(b, k, threshold, final, min_count) for b in batches]
    flags = g.pull_read_flags().tolist() if (not final and g.stats().n_pull_reads) else None
    out = []
    for b, reads in enumerate(batches):
        bg = BatchGraph(g, b)
        st = bg.stats()
        print('number of {}mer: '.format(k), st.n_nodes)                  # debruijn.py:224
        print('branch number: ', st.n_branch)                             # debruijn.py:236
        if final:
            pull_out_read, branch_kmer = [], KmerList((), bg)
        else:
            r0 = int(g.batch_off[b])
            pull_out_read = [r for i, r in enumerate(reads) if flags[r0 + i]] if flags is not None else []
            branch_kmer = KmerList(bg.label_of_ids(bg.branch_ids()), bg)
        already_pull_out = KmerList(bg.label_of_ids(bg.pulled_ids()), bg)
        out.append(((Vertices(bg), Edges(bg)), pull_out_read, branch_kmer, already_pull_out, EdgeCountTable(bg)))
    return out


def output_contigs(g, branch_kmer, already_pull_out):
    """debruijn.py:326-347 on the GPU.  `g` must be the pair construct_graph returned; the walk
    stops at branch nodes iff `branch_kmer` is non-empty (it is emptied by final=True)."""
    graph = getattr(g[0], '_graph', None)
    if not isinstance(graph, (Graph, BatchGraph)):
        raise TypeError('output_contigs needs the graph returned by this module\'s construct_graph')
    st = graph.stats()
    print('Number of kmers have no income edges: ', st.n_starts)           # debruijn.py:336
    graph.walk(stop_at_branch=len(branch_kmer) > 0)
    return graph.contig_strings()


def get_score(edge_count_table, contig, k):
    """II_assembleFromReads.py:14-18 (kept here so callers can use the batched device lookup)."""
    graph = getattr(edge_count_table, '_graph', None)
    n = len(contig) - k
    if n <= 0:
        return 0
    if graph is None:
        return sum(edge_count_table[contig[i:i + k + 1]] for i in range(n))
    counts = graph.lookup_edges([contig[i:i + k + 1] for i in range(n)])
    if (counts == 0).any():
        raise KeyError(contig[int((counts == 0).argmax()):][:k + 1])
    return int(counts.sum())


def support_read_scores(contigs, score_table):
    """IV_sortOutputs.py:10-15 (`findSupportReadScore`) for a whole contig list in one GPU pass: list of float."""
    from dbg_b200.graph import support_read_scores as _srs
    return [float(x) for x in _srs(contigs, score_table)]


def findSupportReadScore(contig, score_table):
    """Same name and arguments as IV_sortOutputs.py:10 (one contig; use support_read_scores for many)."""
    return support_read_scores([contig], score_table)[0]


def sort_by_support(contigs, score_table):
    """IV_sortOutputs.py:54: `contigs.sort(key=findSupportReadScore, reverse=True)` -- a stable sort, so contigs of
    equal score keep their order.  Returns (sorted contigs, their scores)."""
    scores = support_read_scores(contigs, score_table)
    order = sorted(range(len(contigs)), key=lambda i: scores[i], reverse=True)
    return [contigs[i] for i in order], [scores[i] for i in order]


def concatenate_outputs(contigs, sequences_scores, rng=None, out=print):
    """III_concatenateOutpus.py:53-146 (the contig list read at :31 in, the list written at :148-149 out): the
    randomised greedy merge for k = 20 .. 6 and the substring dedup, on the GPU (dbg_b200/concat.py).  Prints the
    script's lines through `out`; draws from `rng` (default: the `random` module's generator, as the script does --
    after `random.seed(s)` the result is the unmodified script's result after `random.seed(s)`)."""
    from dbg_b200.concat import concatenate_outputs as _co
    return _co(contigs, sequences_scores, rng=rng, out=out)
