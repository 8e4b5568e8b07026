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
There is no CPU path: without the built library or a CUDA device these calls raise.  DNA
alphabet (ACGT) only, 1 <= k <= 63.
"""
from __future__ import annotations

from dbg_b200.graph import EdgeCountTable, Edges, Graph, KmerList, Node, Vertices  # noqa: F401

__all__ = ['Node', 'read_reads', 'construct_graph', 'output_contigs', 'get_score']


def read_reads(fname):
    """debruijn.py:22-32: every line not starting with '>' is one read, right-stripped."""
    reads = []
    with open(fname, 'r') as f:
        for line in f:
            if line[0] != '>':
                reads.append(line.rstrip())
    return reads


def construct_graph(reads, k, threshold=3, final=False):
    """debruijn.py:206-285 on the GPU."""
    g = Graph(k)
    g.build(reads, threshold=threshold, final=final)
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


def output_contigs(g, branch_kmer, already_pull_out):
    """debruijn.py:326-347 on the GPU.  `g` must be the pair construct_graph returned; the walk
    stops at branch nodes iff `branch_kmer` is non-empty (it is emptied by final=True)."""
    graph = getattr(g[0], '_graph', None)
    if not isinstance(graph, Graph):
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
