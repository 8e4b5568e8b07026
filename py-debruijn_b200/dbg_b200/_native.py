"""ctypes binding of ``libdbg_b200.so`` (the C ABI in ``include/dbg_b200.h``).

This is the whole FFI: plain pointers and sizes, no torch types.  ``get()`` opens the
nvcc-built library that lives in-tree at ``py-debruijn_b200/lib/libdbg_b200.so`` and nothing
else -- there is no CPU implementation to fall back to, so a missing library or a missing
CUDA device raises.  ``build()`` is the nvcc recipe (sm_100a only).
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))      # py-debruijn_b200/
REPO_DIR = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, 'csrc')
LIB_DIR = os.path.join(PKG_DIR, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libdbg_b200.so')
SOURCES = ('dbg_b200.cu',)
HEADERS = ('dbg_common.cuh', 'dbg_kernels.cuh', 'dbg_chunks.cuh', 'dbg_shard.cuh', 'dbg_generic.cuh', 'dbg_support.cuh',
           'dbg_concat.cuh', 'dbg_concat_host.cuh')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC']

# status codes (include/dbg_b200.h)
DBG_OK = 0
DBG_ERR_CUDA, DBG_ERR_ARG, DBG_ERR_ALPHABET, DBG_ERR_NOMEM = -1, -2, -3, -4
DBG_ERR_STATE, DBG_ERR_ZERODIV, DBG_ERR_OVERFLOW = -5, -6, -7
# build flags
DBG_FINAL, DBG_EXACT_ORDER, DBG_NO_PULL_READS, DBG_TABLE_ONLY, DBG_CHUNKS_ONLY, DBG_BOTH_STRANDS = 1, 2, 4, 8, 16, 32
# node flags
NODE_START, NODE_BRANCH, NODE_DEADEND, NODE_PULLED, NODE_MULTI = 0x02, 0x04, 0x08, 0x10, 0x20
# terminal kinds
TERM_DEADEND, TERM_BRANCH, TERM_PULLED = 0, 1, 2


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        'n_reads', 'n_bases', 'n_windows', 'n_nodes', 'n_edges', 'n_multi', 'n_branch', 'n_starts',
        'n_pulled', 'n_pull_reads', 'n_contigs', 'n_cycle_starts', 'table_slots', 'table_bytes',
        'hll_estimate', 'n_launches', 'n_chunk_instances', 'n_chunk_list', 'chunk_slots')] + [('key_words', C.c_uint32), ('jump_rounds', C.c_uint32), ('n_insert_passes', C.c_uint32),
                                           ('n_regions', C.c_uint32)] + [
        (n, C.c_float) for n in (
            'ms_h2d', 'ms_pack', 'ms_estimate', 'ms_table_init', 'ms_insert', 'ms_finalize',
            'ms_order', 'ms_lists', 'ms_tips', 'ms_pull_reads', 'ms_build_total', 'ms_walk_rank',
            'ms_walk_emit', 'ms_walk_total', 'ms_chunk_insert', 'ms_expand_sketch', 'ms_expand', 'ms_region_order')]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Contig(C.Structure):
    _fields_ = [('start', C.c_uint32), ('n_nodes', C.c_uint32), ('terminal', C.c_uint32),
                ('kind', C.c_uint32), ('seg_begin', C.c_uint64), ('seg_end', C.c_uint64)]


class ShardContig(C.Structure):
    _fields_ = [('tag', C.c_uint64), ('start_slot', C.c_uint32), ('start_rank', C.c_uint32),
                ('terminal_slot', C.c_uint32), ('terminal_rank', C.c_uint32), ('n_nodes', C.c_uint32),
                ('kind', C.c_uint32)]


class Segment(C.Structure):
    _fields_ = [('entry', C.c_uint32), ('n_nodes', C.c_uint32)]


class ConcatStatus(C.Structure):
    _fields_ = [('status', C.c_uint32), ('current', C.c_uint32), ('iterations', C.c_uint64), ('n_alive', C.c_uint32),
                ('max_len', C.c_uint32), ('n_pending', C.c_uint32), ('pad', C.c_uint32), ('arena_bytes', C.c_uint64)]


DBG_CONCAT_DONE, DBG_CONCAT_NEED_CHOICE = 1, 2


# every symbol include/dbg_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    'dbg_create': (C.c_int, [C.POINTER(_P), C.c_int, C.c_int]),
    'dbg_create_ex': (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int]),
    'dbg_create_alpha': (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, _P, C.c_int]),
    'dbg_destroy': (None, [_P]),
    'dbg_last_error': (C.c_char_p, [_P]),
    'dbg_set_stream': (C.c_int, [_P, _P]),
    'dbg_set_min_count': (C.c_int, [_P, C.c_uint32]),
    'dbg_support_scores': (C.c_int, [_P, _P, _P, C.c_uint64, _P, _P, _P, C.c_uint64, _P]),
    'dbg_version': (C.c_int, []),
    'dbg_build': (C.c_int, [_P, _P, _P, C.c_uint64, C.c_uint64, C.c_double, C.c_uint32]),
    'dbg_build_device': (C.c_int, [_P, _P, _P, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_uint32]),
    'dbg_build_batched': (C.c_int, [_P, _P, _P, C.c_uint64, _P, C.c_uint32, C.c_double, C.c_uint32]),
    'dbg_build_batched_device': (C.c_int, [_P, _P, _P, C.c_uint64, C.c_uint64, _P, C.c_uint32, C.c_double, C.c_uint32]),
    'dbg_stats': (C.c_int, [_P, C.POINTER(Stats)]),
    'dbg_graph_stage': (C.c_int, [_P, C.c_int, C.c_double, C.c_uint32]),
    'dbg_record_words': (C.c_int, [_P]),
    'dbg_order_words': (C.c_int, [_P]),
    'dbg_export_counts': (C.c_int, [_P, C.c_uint32, _P]),
    'dbg_export_records': (C.c_int, [_P, C.c_uint32, _P, _P]),
    'dbg_table_reset': (C.c_int, [_P, C.c_uint64]),
    'dbg_merge_records': (C.c_int, [_P, _P, C.c_uint64]),
    'dbg_chunk_record_words': (C.c_int, [_P]),
    'dbg_chunk_export_counts': (C.c_int, [_P, C.c_uint32, _P]),
    'dbg_chunk_export_records': (C.c_int, [_P, C.c_uint32, _P, _P]),
    'dbg_chunk_table_reset': (C.c_int, [_P, C.c_uint64]),
    'dbg_chunk_merge_records': (C.c_int, [_P, _P, C.c_uint64]),
    'dbg_chunk_expand': (C.c_int, [_P]),
    'dbg_order_export': (C.c_int, [_P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    'dbg_order_merge': (C.c_int, [_P, _P, C.c_uint64]),
    'dbg_shard_config': (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    'dbg_shard_exit_counts': (C.c_int, [_P, _P]),
    'dbg_shard_exit_keys': (C.c_int, [_P, _P]),
    'dbg_shard_mark_entries': (C.c_int, [_P, _P, C.c_uint64]),
    'dbg_shard_rank_local': (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    'dbg_shard_entry_ordinals': (C.c_int, [_P, _P]),
    'dbg_shard_reduced_records': (C.c_int, [_P, _P, _P, _P]),
    'dbg_shard_walk': (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]),
    'dbg_shard_contigs': (C.c_int, [_P, _P]),
    'dbg_export_nodes': (C.c_int, [_P] + [_P] * 7),
    'dbg_export_edges': (C.c_int, [_P] + [_P] * 5),
    'dbg_branch_nodes': (C.c_int, [_P, _P]),
    'dbg_pulled_nodes': (C.c_int, [_P, _P]),
    'dbg_pull_reads': (C.c_int, [_P, _P]),
    'dbg_node_keys': (C.c_int, [_P, _P, C.c_uint64, _P, _P]),
    'dbg_degree_balance': (C.c_int, [_P, _P, C.c_uint64, _P, _P, _P]),
    'dbg_probe_stats': (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P]),
    'dbg_twin_nodes': (C.c_int, [_P, _P, C.c_uint64, _P, _P]),
    'dbg_lookup_edges': (C.c_int, [_P, _P, C.c_uint64, _P]),
    'dbg_lookup_edges_batched': (C.c_int, [_P, _P, _P, C.c_uint64, _P]),
    'dbg_walk': (C.c_int, [_P, C.c_int, C.POINTER(C.c_uint64)]),
    'dbg_contigs': (C.c_int, [_P, _P]),
    'dbg_segments': (C.c_int, [_P, _P, C.POINTER(C.c_uint64)]),
    'dbg_materialize': (C.c_int, [_P, _P, C.c_uint64, _P, _P]),
    'dbg_contig_scores': (C.c_int, [_P, _P, C.c_uint64, _P]),
    'dbg_concat_begin': (C.c_int, [_P, _P, _P, C.c_uint64]),
    'dbg_concat_run': (C.c_int, [_P, C.c_int, C.c_uint64, _P, C.c_int, C.POINTER(ConcatStatus)]),
    'dbg_concat_pending': (C.c_int, [_P, _P, _P]),
    'dbg_concat_choose': (C.c_int, [_P, _P]),
    'dbg_concat_lengths': (C.c_int, [_P, _P, _P]),
    'dbg_concat_strings': (C.c_int, [_P, _P, C.c_uint64, _P, _P]),
    'dbg_concat_dedup': (C.c_int, [_P, _P]),
    'dbg_concat_end': (C.c_int, [_P]),
}


class AlphabetError(ValueError):
    """A read holds a byte outside the handle's alphabet (DBG_ERR_ALPHABET)."""


class DbgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__('dbg_b200 error %d: %s' % (code, msg))
        self.code = code


def _raise(code, msg):
    """Map the C status codes onto the exceptions the reference would raise."""
    if code == DBG_ERR_ALPHABET:
        raise AlphabetError('dbg_b200: %s' % msg)
    if code == DBG_ERR_ZERODIV:
        raise ZeroDivisionError('division by zero')            # debruijn.py:163 with threshold == 0
    if code == DBG_ERR_NOMEM:
        raise MemoryError('dbg_b200: ' + msg)
    if code == DBG_ERR_ARG:
        raise ValueError('dbg_b200: ' + msg)
    raise DbgError(code, msg)


class Library:
    """One loaded copy of the C ABI."""

    def __init__(self, path):
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(self.dll, name)              # AttributeError = the library lacks a symbol
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)

    def check(self, handle, code):
        if code != DBG_OK:
            msg = self.dbg_last_error(handle)
            _raise(code, msg.decode('utf-8', 'replace') if msg else '')


def build(force=False, verbose=False):
    """Compile the CUDA sources for sm_100a into ``lib/libdbg_b200.so`` (in-tree, git-ignored)."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.join(REPO_DIR, 'include', 'dbg_b200.h')]
    if not force and os.path.exists(LIB_PATH) and \
            all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found; cannot build %s' % LIB_PATH)
    os.makedirs(LIB_DIR, exist_ok=True)
    tmp = '%s.%d.tmp' % (LIB_PATH, os.getpid())        # several ranks may build at once: never expose a partial file
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', tmp] + srcs
    subprocess.check_call(cmd)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


_lib = None
_DEFAULT_LIB_PATH = LIB_PATH


def get():
    """The product library.  Raises if it has not been built; never substitutes anything.  (A missing
    in-tree library is compiled on the spot when nvcc is there -- that is building the product, not
    falling back.)"""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) and LIB_PATH == _DEFAULT_LIB_PATH and \
                (shutil.which('nvcc') or os.path.exists('/usr/local/cuda/bin/nvcc')):
            build()
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'dbg_b200: %s is missing. Build it with `python -c "import __graft_entry__ as g; '
                'g.build()"` (nvcc, sm_100a). There is no CPU fallback.' % LIB_PATH)
        _lib = Library(LIB_PATH)
    return _lib
