"""ctypes binding of oracle/dbg_oracle.c  --  TEST INFRASTRUCTURE ONLY (see that file's header).

``build()`` compiles the C restatement with gcc into oracle/_build/ (git-ignored, travels
to the GPU box with the snapshot).  ``run()`` returns the same snapshot structure as
``dbg_oracle.snapshot`` (strings decoded from the 2-bit keys) or, with ``strings=False``,
numpy arrays for large inputs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'dbg_oracle.c')
OUT_DIR = os.path.join(HERE, '_build')
LIB = os.path.join(OUT_DIR, 'libdbg_oracle.so')

_lib = None


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(['gcc', '-O2', '-shared', '-fPIC', '-o', LIB, SRC])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.dbgo_run.restype = C.c_void_p
        L.dbgo_run.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64, C.c_int, C.c_double, C.c_int, C.c_int]
        L.dbgo_run_mc.restype = C.c_void_p
        L.dbgo_run_mc.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64, C.c_int, C.c_double, C.c_int, C.c_int, C.c_uint32]
        L.dbgo_free.argtypes = [C.c_void_p]
        L.dbgo_error.argtypes = [C.c_void_p]
        for name in ('n_nodes', 'n_windows', 'n_branch', 'n_branch_internal', 'n_pulled', 'n_starts',
                     'n_contigs', 'n_pull_reads'):
            fn = getattr(L, 'dbgo_' + name)
            fn.restype = C.c_uint64
            fn.argtypes = [C.c_void_p]
        L.dbgo_time.restype = C.c_double
        L.dbgo_time.argtypes = [C.c_void_p, C.c_int]
        L.dbgo_nodes.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.dbgo_adj.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.dbgo_branch.argtypes = [C.c_void_p, C.c_void_p]
        L.dbgo_pulled.argtypes = [C.c_void_p, C.c_void_p]
        L.dbgo_read_flags.argtypes = [C.c_void_p, C.c_void_p]
        L.dbgo_contigs.argtypes = [C.c_void_p, C.c_void_p]
        L.dbgo_contig_len.restype = C.c_uint64
        L.dbgo_contig_len.argtypes = [C.c_void_p, C.c_uint64]
        L.dbgo_contig_string.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        _lib = L
    return _lib


def pack_reads(reads):
    """list[str] -> (ASCII bytes, uint64 offsets[n+1])."""
    lens = np.fromiter((len(r) for r in reads), dtype=np.uint64, count=len(reads))
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    return ''.join(reads).encode('ascii'), off


def decode(key_lo, key_hi, k):
    v = (int(key_hi) << 64) | int(key_lo)
    return ''.join('ACGT'[(v >> (2 * (k - 1 - j))) & 3] for j in range(k))


class Result:
    """Arrays from one oracle run (all in first-occurrence node order)."""

    def __init__(self, h, k, n_reads, final):
        L = lib()
        self.k = k
        self.final = final
        n = L.dbgo_n_nodes(h)
        self.n_nodes = n
        self.n_windows = L.dbgo_n_windows(h)
        self.n_branch_internal = L.dbgo_n_branch_internal(h)
        self.n_starts = L.dbgo_n_starts(h)
        self.t_build, self.t_graph, self.t_walk = (L.dbgo_time(h, i) for i in range(3))
        self.key_lo = np.zeros(n, np.uint64)
        self.key_hi = np.zeros(n, np.uint64)
        self.tag = np.zeros(n, np.uint64)
        self.cnt = np.zeros((n, 4), np.uint32)
        self.indeg = np.zeros(n, np.uint8)
        self.outdeg = np.zeros(n, np.uint8)
        L.dbgo_nodes(h, *(a.ctypes.data for a in (self.key_lo, self.key_hi, self.tag, self.cnt,
                                                  self.indeg, self.outdeg)))
        self.n_kept = np.zeros(n, np.uint8)
        self.kept_base = np.zeros((n, 4), np.uint8)
        self.kept_idx = np.zeros((n, 4), np.uint32)
        self.pulled_flag = np.zeros(n, np.uint8)
        self.branch = np.zeros(0, np.uint32)
        self.pulled = np.zeros(0, np.uint32)
        self.read_flags = np.zeros(n_reads, np.uint8)
        self.contigs = np.zeros((0, 4), np.uint32)
        self.have_graph = False

    def load_graph(self, h, n_reads):
        L = lib()
        L.dbgo_adj(h, *(a.ctypes.data for a in (self.n_kept, self.kept_base, self.kept_idx,
                                                self.pulled_flag)))
        self.branch = np.zeros(L.dbgo_n_branch(h), np.uint32)
        L.dbgo_branch(h, self.branch.ctypes.data)
        self.pulled = np.zeros(L.dbgo_n_pulled(h), np.uint32)
        L.dbgo_pulled(h, self.pulled.ctypes.data)
        L.dbgo_read_flags(h, self.read_flags.ctypes.data)
        self.have_graph = True

    def load_contigs(self, h, strings):
        L = lib()
        nc = L.dbgo_n_contigs(h)
        self.contigs = np.zeros((nc, 4), np.uint32)     # start, n_nodes, terminal, kind
        L.dbgo_contigs(h, self.contigs.ctypes.data)
        self.contig_strings = None
        if strings:
            out = []
            for i in range(nc):
                ln = L.dbgo_contig_len(h, i)
                buf = C.create_string_buffer(ln)
                L.dbgo_contig_string(h, i, buf)
                out.append(buf.raw.decode('ascii'))
            self.contig_strings = out

    def label(self, v):
        return decode(self.key_lo[v], self.key_hi[v], self.k)


def run(reads, k, threshold, final, stages=3, strings=True, packed=None, min_count=1):
    """Run the C oracle.  `packed` = (bytes, offsets) skips re-packing of `reads`."""
    L = lib()
    data, off = packed if packed is not None else pack_reads(reads)
    n_reads = len(off) - 1
    h = L.dbgo_run_mc(data, off.ctypes.data, n_reads, k, float(threshold), int(bool(final)), stages, int(min_count))
    if not h:
        raise ValueError('k out of range for the C oracle (1..63)')
    try:
        err = L.dbgo_error(h)
        if err:
            raise ValueError('C oracle: non-ACGT input' if err == -2 else 'C oracle error %d' % err)
        res = Result(h, k, n_reads, final)
        if stages >= 2:
            res.load_graph(h, n_reads)
        if stages >= 3:
            res.load_contigs(h, strings)
        return res
    finally:
        L.dbgo_free(h)


def snapshot(reads, k, threshold, final, min_count=1):
    """Same structure as dbg_oracle.snapshot (DNA inputs only)."""
    res = run(reads, k, threshold, final, min_count=min_count)
    lab = [res.label(v) for v in range(res.n_nodes)]
    edges = []
    for v in range(res.n_nodes):
        if res.pulled_flag[v]:
            continue                                      # edges.pop(item), debruijn.py:254
        edges.append([lab[v], [lab[res.kept_idx[v, j]] for j in range(res.n_kept[v])]])
    ect = []
    # edge_count_table order: node order, then first appearance of each successor
    # (debruijn.py:214-222); first appearance is not exported, so the caller compares sorted.
    for v in range(res.n_nodes):
        for c in range(4):
            if res.cnt[v, c]:
                ect.append([lab[v] + 'ACGT'[c], int(res.cnt[v, c])])
    return {
        'nodes': [[lab[v], int(res.indeg[v]), int(res.outdeg[v])] for v in range(res.n_nodes)],
        'edges': edges,
        'pull_out_read': [r for r, f in zip(reads, res.read_flags) if f],
        'branch': [lab[v] for v in res.branch],
        'pulled': [lab[v] for v in res.pulled],
        'ect': ect,
        'contigs': res.contig_strings,
    }
