"""B200-native implementation of py-debruijn's assembly hot path (host side)."""
from . import _native                                                        # noqa: F401
from .graph import EdgeCountTable, Edges, Graph, KmerList, Node, Vertices, decode_keys, pack_reads  # noqa: F401
