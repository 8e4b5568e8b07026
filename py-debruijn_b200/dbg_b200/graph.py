"""Host-side graph handle over the C ABI, plus the lazy Python views that stand in for the
reference's ``vertices`` / ``edges`` / ``edge_count_table`` dicts (debruijn.py:285).

Everything here is plumbing: packing ``list[str]`` into one byte buffer + offsets, calling the
library, and decoding 2-bit keys back to strings when somebody looks.  No stage of the
algorithm is computed in Python.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from collections.abc import Mapping

import numpy as np

from . import _native as N

_ALPHA = np.frombuffer(b'ACGT', dtype=np.uint8)
MAX_STRING_BYTES = 8 << 30          # refuse to materialise more contig characters than this at once


class Node:
    """debruijn.py:8-14."""

    __slots__ = ('label', 'indegree', 'outdegree')

    def __init__(self, lab):
        self.label = lab
        self.indegree = 0
        self.outdegree = 0


class NotSingleByteError(ValueError):
    """The reads hold characters beyond latin-1 (debruijn._graph_for then gives every character a private byte)."""


def pack_reads(reads):
    """list[str] -> (uint8[n_bases] ASCII, uint64[n+1] offsets); II hands over a plain list."""
    n = len(reads)
    lens = np.fromiter(map(len, reads), dtype=np.uint64, count=n)
    off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    try:
        data = ''.join(reads).encode('latin-1')
    except UnicodeEncodeError as e:
        raise NotSingleByteError('dbg_b200: reads must be strings of single-byte characters') from e
    bases = np.frombuffer(data, dtype=np.uint8)
    return bases, off


def decode_keys(key_lo, key_hi, k, bits=2, symbols=_ALPHA):
    """Packed keys (`bits` per symbol, first symbol most significant) -> list[str]."""
    n = len(key_lo)
    if n == 0:
        return []
    shifts = bits * (k - 1 - np.arange(k, dtype=np.int64))
    lo = np.asarray(key_lo, dtype=np.uint64)[:, None]
    hi = np.asarray(key_hi, dtype=np.uint64)[:, None]
    mask = np.uint64((1 << bits) - 1)
    codes = np.empty((n, k), dtype=np.uint8)
    low = shifts + bits <= 64                       # the symbol lies in the low word
    high = shifts >= 64                             # ... in the high word
    codes[:, low] = ((lo >> shifts[low].astype(np.uint64)) & mask).astype(np.uint8)
    if high.any():
        codes[:, high] = ((hi >> (shifts[high] - 64).astype(np.uint64)) & mask).astype(np.uint8)
    for j in np.nonzero(~low & ~high)[0]:           # straddles the two words (generic-alphabet keys only)
        sh = int(shifts[j])
        codes[:, j] = (((lo[:, 0] >> np.uint64(sh)) | (hi[:, 0] << np.uint64(64 - sh))) & mask).astype(np.uint8)
    chars = np.asarray(symbols, dtype=np.uint8)[codes]
    if symbols is _ALPHA:
        return chars.view('S%d' % k).ravel().astype('U%d' % k).tolist()
    text = chars.tobytes().decode('latin-1')        # any byte values: NUL, blanks and 0x80 .. 0xFF are symbols like the others
    return [text[i:i + k] for i in range(0, n * k, k)]


# Handles of finished construct_graph calls, kept for the next call with the same k / alphabet / device (debruijn.py's API
# has no state to hang a handle on, and dbg_create + dbg_destroy cost 1.7 ms -- more than building a small graph; a reused
# handle also keeps its buffers and remembers its table sizes).  Only handles whose last build was small and that never
# reported an error are kept; DBG_B200_HANDLE_POOL=0 switches the pool off.
_POOL = {}                               # key -> [(handle, age, library)], oldest first
_POOL_CLOCK = [0]
_POOL_LOCK = threading.Lock()
POOL_PER_KEY, POOL_TOTAL, POOL_MAX_BASES = 2, 8, 32 << 20


def _pool_take(key):
    with _POOL_LOCK:
        lst = _POOL.get(key)
        return lst.pop()[0] if lst else None


def _pool_put(key, h, lib):
    """Keep handle h for the next Graph of this shape; the oldest kept handle makes room when the pool is full (a
    caller that changes k with every call, as II's loop does, would otherwise fill it with handles nobody takes)."""
    if os.environ.get('DBG_B200_HANDLE_POOL', '1') == '0':
        return False
    evict = None
    with _POOL_LOCK:
        _POOL_CLOCK[0] += 1
        lst = _POOL.setdefault(key, [])
        if len(lst) >= POOL_PER_KEY:
            evict = lst.pop(0)
        elif sum(map(len, _POOL.values())) >= POOL_TOTAL:
            k_old = min((k for k, v in _POOL.items() if v), key=lambda k: _POOL[k][0][1])
            evict = _POOL[k_old].pop(0)
        lst.append((h, _POOL_CLOCK[0], lib))
    if evict is not None:
        evict[2].dbg_destroy(evict[0])
    return True


def _default_device():
    for var in ('DBG_B200_DEVICE', 'LOCAL_RANK'):
        v = os.environ.get(var)
        if v is not None and v != '':
            return int(v)
    return 0


class Graph:
    """One de Bruijn graph resident on one GPU (one ``dbg_t``).  Re-usable: ``build`` may be
    called again with new reads and keeps its HBM."""

    def __init__(self, k, device=None, lib=None, key_words=0, alphabet=None, charset=None, pooled=False):
        """`alphabet` = bytes of up to 255 distinct symbols: the generic-alphabet engine, b = ceil(log2(symbols)) bits a
        symbol and b (k + 1) <= 125 (peptides: 5 bits, k <= 24; ACGT + N: 3 bits, k <= 40; 8 bits: k <= 14).
        None = the DNA engine (ACGT, k <= 63).  `charset` = a list of up to 255 characters for reads that are not
        latin-1 (any Unicode text, as the reference takes): character i travels as byte value i, the alphabet is
        bytes(range(len(charset))) and every string that comes back (labels, contigs, edge keys) is mapped back.
        `pooled`: take the handle of an earlier, closed Graph of the same shape if there is one, and hand this one's
        on when it is closed (what debruijn.construct_graph does)."""
        self.lib = lib or N.get()
        self.k = int(k)
        self.device = _default_device() if device is None else int(device)
        self._chars_in = self._chars_out = None
        if charset is not None:
            charset = list(charset)
            alphabet = bytes(range(len(charset)))
            self._chars_in = {ord(c): i for i, c in enumerate(charset)}
            self._chars_out = {i: c for i, c in enumerate(charset)}
        self.alphabet = bytes(alphabet) if alphabet is not None else None
        self.sym_bits = max(1, (len(self.alphabet) - 1).bit_length()) if self.alphabet is not None else 2
        self._pool_key = (self.lib.path, self.device, self.k, int(key_words), self.alphabet) if pooled else None
        self._pool_ok = False
        h = _pool_take(self._pool_key) if pooled else None
        self.from_pool = h is not None
        if h is None:
            h = C.c_void_p()
            if self.alphabet is not None:
                rc = self.lib.dbg_create_alpha(C.byref(h), self.device, self.k, self.alphabet, len(self.alphabet))
            else:
                rc = self.lib.dbg_create_ex(C.byref(h), self.device, self.k, int(key_words))
            if rc != N.DBG_OK:
                msg = self.lib.dbg_last_error(None)
                N._raise(rc, msg.decode() if msg else '')
        self.h = h
        self.reads = None
        self.final = False
        self._reset_cache()

    def _reset_cache(self):
        self._batch_cache = None
        self._edges = None
        self._nodes = None
        self._labels = None
        self._contigs = None
        self._stats = None

    def close(self):
        h = getattr(self, 'h', None)
        if h:
            self.h = None
            if getattr(self, '_pool_key', None) is not None and self._pool_ok and not self.lib.dbg_last_error(h) \
                    and _pool_put(self._pool_key, h, self.lib):
                return
            self.lib.dbg_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:       # noqa: BLE001 - interpreter shutdown
            pass

    # ---- construct_graph -----------------------------------------------------------------
    def set_min_count(self, min_count):
        """EXTENSION (no reference line, include/dbg_b200.h dbg_set_min_count): the following builds skip every
        window whose (k+1)-mer occurs fewer than `min_count` times over all reads.  1 = the reference."""
        self.lib.check(self.h, self.lib.dbg_set_min_count(self.h, int(min_count)))
        return self

    def build(self, reads, threshold=3, final=False, exact_order=True, pull_reads=True,
              packed=None, read_idx_base=0, min_count=1, both_strands=False):
        """Host entry: `reads` is a list[str] (or pass packed=(uint8 bases, uint64 offsets)).  `both_strands`: an
        extension (DBG_BOTH_STRANDS) -- the graph of the reads followed by their reverse complements."""
        self._pool_ok = False
        self.set_min_count(min_count)
        bases, off = packed if packed is not None else pack_reads(reads)
        self.reads = reads
        self.final = bool(final)
        self._reset_cache()
        flags = (N.DBG_FINAL if final else 0) | (N.DBG_EXACT_ORDER if exact_order else 0) | \
                (0 if pull_reads else N.DBG_NO_PULL_READS) | (N.DBG_BOTH_STRANDS if both_strands else 0)
        self._keep = (bases, off)                      # keep the buffers alive during the call
        self.n_reads = len(off) - 1
        rc = self.lib.dbg_build(self.h, bases.ctypes.data if len(bases) else None, off.ctypes.data,
                                self.n_reads, int(read_idx_base), float(threshold), flags)
        self.lib.check(self.h, rc)
        self._pool_ok = len(bases) <= POOL_MAX_BASES   # (a handle that held a big graph keeps gigabytes of HBM: not kept)
        if both_strands:
            self.n_reads *= 2                          # read n + i = reverse complement of read i
        return self

    def build_batched(self, batches, threshold=3, final=False, exact_order=True, pull_reads=True):
        """MANY independent read sets in one pass (include/dbg_b200.h, dbg_build_batched):
        `batches` is a list of list[str]; batch b's reads are reads[batch_off[b]:batch_off[b+1]]."""
        reads = [r for b in batches for r in b]
        bases, off = pack_reads(reads)
        sizes = np.fromiter((len(b) for b in batches), dtype=np.int64, count=len(batches))
        self.batch_off = np.zeros(len(batches) + 1, np.int64)
        np.cumsum(sizes, out=self.batch_off[1:])
        read_batch = np.repeat(np.arange(len(batches), dtype=np.uint32), sizes)
        self.reads = reads
        self.final = bool(final)
        self.n_batches = len(batches)
        self._reset_cache()
        self._keep = (bases, off, read_batch)
        self.n_reads = len(off) - 1
        rc = self.lib.dbg_build_batched(self.h, bases.ctypes.data if len(bases) else None, off.ctypes.data,
                                        self.n_reads, read_batch.ctypes.data if len(read_batch) else None,
                                        self.n_batches, float(threshold), self._flags(final, exact_order, pull_reads))
        self.lib.check(self.h, rc)
        return self

    def build_batched_packed(self, bases, off, read_batch, n_batches, threshold=3, final=False, exact_order=True,
                             pull_reads=True, device_ptrs=None):
        """build_batched on packed arrays (uint8 bases, uint64 offsets, uint32 batch per read), or on
        device pointers: device_ptrs = (d_bases, d_off, d_read_batch, n_reads, n_bases)."""
        self.reads = None
        self.final = bool(final)
        self.n_batches = int(n_batches)
        self._reset_cache()
        flags = self._flags(final, exact_order, pull_reads)
        if device_ptrs is not None:
            d_bases, d_off, d_batch, n_reads, n_bases = device_ptrs
            self.n_reads = int(n_reads)
            rc = self.lib.dbg_build_batched_device(self.h, d_bases, d_off, self.n_reads, int(n_bases), d_batch,
                                                   self.n_batches, float(threshold), flags)
        else:
            self._keep = (bases, off, read_batch)
            self.n_reads = len(off) - 1
            rc = self.lib.dbg_build_batched(self.h, bases.ctypes.data if len(bases) else None, off.ctypes.data,
                                            self.n_reads, read_batch.ctypes.data if len(read_batch) else None,
                                            self.n_batches, float(threshold), flags)
        self.lib.check(self.h, rc)
        return self

    def build_device(self, d_bases_ptr, d_off_ptr, n_reads, n_bases, threshold=3, final=False,
                     exact_order=True, pull_reads=True, read_idx_base=0, min_count=1):
        """Device entry: raw device pointers (e.g. torch tensors' data_ptr()) already in HBM."""
        self.set_min_count(min_count)
        self.reads = None
        self.final = bool(final)
        self._reset_cache()
        flags = (N.DBG_FINAL if final else 0) | (N.DBG_EXACT_ORDER if exact_order else 0) | \
                (0 if pull_reads else N.DBG_NO_PULL_READS)
        self.n_reads = int(n_reads)
        rc = self.lib.dbg_build_device(self.h, d_bases_ptr, d_off_ptr, int(n_reads), int(n_bases),
                                       int(read_idx_base), float(threshold), flags)
        self.lib.check(self.h, rc)
        return self

    # ---- multi-GPU building blocks (include/dbg_b200.h) ------------------------------------
    def _flags(self, final, exact_order, pull_reads):
        return (N.DBG_FINAL if final else 0) | (N.DBG_EXACT_ORDER if exact_order else 0) | \
               (0 if pull_reads else N.DBG_NO_PULL_READS)

    def build_table(self, reads=None, packed=None, read_idx_base=0):
        """Window loop only (host buffers): leaves the node table, no graph stages."""
        bases, off = packed if packed is not None else pack_reads(reads)
        self.reads = reads
        self._reset_cache()
        self._keep = (bases, off)
        self.n_reads = len(off) - 1
        rc = self.lib.dbg_build(self.h, bases.ctypes.data if len(bases) else None, off.ctypes.data,
                                self.n_reads, int(read_idx_base), 3.0, N.DBG_TABLE_ONLY)
        self.lib.check(self.h, rc)
        return self

    def build_table_device(self, d_bases_ptr, d_off_ptr, n_reads, n_bases, read_idx_base=0):
        self.reads = None
        self._reset_cache()
        self.n_reads = int(n_reads)
        rc = self.lib.dbg_build_device(self.h, d_bases_ptr, d_off_ptr, int(n_reads), int(n_bases),
                                       int(read_idx_base), 3.0, N.DBG_TABLE_ONLY)
        self.lib.check(self.h, rc)
        return self

    def build_chunks(self, reads=None, packed=None, read_idx_base=0, device_ptrs=None):
        """Phase 1 of the chunk build only (DBG_CHUNKS_ONLY): the distinct super-k-mers of the reads.
        Host buffers, or device_ptrs = (d_bases, d_off, n_reads, n_bases)."""
        self._reset_cache()
        if device_ptrs is not None:
            d_bases, d_off, n_reads, n_bases = device_ptrs
            self.reads = None
            self.n_reads = int(n_reads)
            rc = self.lib.dbg_build_device(self.h, d_bases, d_off, self.n_reads, int(n_bases), int(read_idx_base), 3.0,
                                           N.DBG_CHUNKS_ONLY)
        else:
            bases, off = packed if packed is not None else pack_reads(reads)
            self.reads = reads
            self._keep = (bases, off)
            self.n_reads = len(off) - 1
            rc = self.lib.dbg_build(self.h, bases.ctypes.data if len(bases) else None, off.ctypes.data, self.n_reads,
                                    int(read_idx_base), 3.0, N.DBG_CHUNKS_ONLY)
        self.lib.check(self.h, rc)
        return self

    def chunk_record_words(self):
        return self.lib.dbg_chunk_record_words(self.h)

    def chunk_export_counts(self, n_parts):
        counts = np.zeros(n_parts, np.uint64)
        self.lib.check(self.h, self.lib.dbg_chunk_export_counts(self.h, n_parts, counts.ctypes.data))
        return counts

    def chunk_export_records(self, n_parts, counts, d_out_ptr):
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        self.lib.check(self.h, self.lib.dbg_chunk_export_records(self.h, n_parts, counts.ctypes.data, d_out_ptr))

    def chunk_table_reset(self, expected_chunks):
        self._reset_cache()
        self.lib.check(self.h, self.lib.dbg_chunk_table_reset(self.h, int(expected_chunks)))

    def chunk_merge_records(self, d_ptr, n):
        self.lib.check(self.h, self.lib.dbg_chunk_merge_records(self.h, d_ptr, int(n)))

    def chunk_expand(self):
        self.lib.check(self.h, self.lib.dbg_chunk_expand(self.h))

    def record_words(self):
        return self.lib.dbg_record_words(self.h)

    def order_words(self):
        return self.lib.dbg_order_words(self.h)

    def export_counts(self, n_parts):
        counts = np.zeros(n_parts, np.uint64)
        self.lib.check(self.h, self.lib.dbg_export_counts(self.h, n_parts, counts.ctypes.data))
        return counts

    def export_records(self, n_parts, counts, d_out_ptr):
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        self.lib.check(self.h, self.lib.dbg_export_records(self.h, n_parts, counts.ctypes.data, d_out_ptr))

    def table_reset(self, expected_nodes):
        self._reset_cache()
        self.lib.check(self.h, self.lib.dbg_table_reset(self.h, int(expected_nodes)))

    def merge_records(self, d_ptr, n):
        self.lib.check(self.h, self.lib.dbg_merge_records(self.h, d_ptr, int(n)))

    def graph_stage(self, stage, threshold=3, final=False, exact_order=True, pull_reads=True):
        self.final = bool(final)
        self._reset_cache()
        self.lib.check(self.h, self.lib.dbg_graph_stage(self.h, int(stage), float(threshold),
                                                        self._flags(final, exact_order, pull_reads)))

    def order_export(self, d_out_ptr, capacity):
        n = C.c_uint64()
        self.lib.check(self.h, self.lib.dbg_order_export(self.h, d_out_ptr, int(capacity), C.byref(n)))
        return int(n.value)

    def order_merge(self, d_ptr, n):
        self.lib.check(self.h, self.lib.dbg_order_merge(self.h, d_ptr, int(n)))

    # ---- sharded node table: owner-computes walk (include/dbg_b200.h, dbg_shard_*) ----------
    SHARD_CONTIG_DTYPE = np.dtype([('tag', np.uint64), ('start_slot', np.uint32), ('start_rank', np.uint32),
                                   ('terminal_slot', np.uint32), ('terminal_rank', np.uint32),
                                   ('n_nodes', np.uint32), ('kind', np.uint32)])

    def shard_config(self, rank, n_ranks):
        self._pool_key = None
        self.lib.check(self.h, self.lib.dbg_shard_config(self.h, int(rank), int(n_ranks)))

    def shard_exit_counts(self, n_ranks):
        counts = np.zeros(n_ranks, np.uint64)
        self.lib.check(self.h, self.lib.dbg_shard_exit_counts(self.h, counts.ctypes.data))
        return counts

    def shard_exit_keys(self, d_out_ptr):
        self.lib.check(self.h, self.lib.dbg_shard_exit_keys(self.h, d_out_ptr))

    def shard_mark_entries(self, d_keys_ptr, n):
        self.lib.check(self.h, self.lib.dbg_shard_mark_entries(self.h, d_keys_ptr, int(n)))

    def shard_rank_local(self):
        n = C.c_uint64()
        self.lib.check(self.h, self.lib.dbg_shard_rank_local(self.h, C.byref(n)))
        return int(n.value)

    def shard_entry_ordinals(self, d_out_ptr):
        self.lib.check(self.h, self.lib.dbg_shard_entry_ordinals(self.h, d_out_ptr))

    def shard_reduced_records(self, d_replies_ptr, rid_base, d_out_ptr):
        rid_base = np.ascontiguousarray(rid_base, dtype=np.uint64)
        self.lib.check(self.h, self.lib.dbg_shard_reduced_records(self.h, d_replies_ptr, rid_base.ctypes.data, d_out_ptr))

    def shard_walk(self, d_all_ptr, n_total, my_base):
        n = C.c_uint64()
        self.lib.check(self.h, self.lib.dbg_shard_walk(self.h, d_all_ptr, int(n_total), int(my_base), C.byref(n)))
        out = np.empty(n.value, dtype=self.SHARD_CONTIG_DTYPE)
        self.lib.check(self.h, self.lib.dbg_shard_contigs(self.h, out.ctypes.data if n.value else None))
        return out

    def set_stream(self, cuda_stream_ptr):
        """Run on the caller's stream.  0 / None is CUDA's legacy default stream (what
        torch.cuda.current_stream().cuda_stream returns by default), passed as cudaStreamLegacy;
        use own_stream() to go back to the handle's private stream."""
        self._pool_key = None
        self.lib.check(self.h, self.lib.dbg_set_stream(self.h, cuda_stream_ptr or 1))
        self._stream = int(cuda_stream_ptr or 0)

    def own_stream(self):
        self.lib.check(self.h, self.lib.dbg_set_stream(self.h, None))
        self._stream = None

    def shares_stream(self, cuda_stream_ptr):
        """True if the handle was put on exactly this stream with set_stream."""
        return getattr(self, '_stream', None) is not None and self._stream == int(cuda_stream_ptr or 0)

    def stats(self):
        s = N.Stats()
        self.lib.check(self.h, self.lib.dbg_stats(self.h, C.byref(s)))
        return s

    # ---- exports -------------------------------------------------------------------------
    def nodes(self):
        """Arrays of every node, sorted by first-occurrence tag (= the reference's dict order)."""
        if self._nodes is None:
            n = self.stats().n_nodes
            a = {
                'id': np.zeros(n, np.uint32), 'key_lo': np.zeros(n, np.uint64), 'key_hi': np.zeros(n, np.uint64),
                'tag': np.zeros(n, np.uint64), 'cnt': np.zeros((n, 4), np.uint32),
                'flags': np.zeros(n, np.uint8), 'kept': np.zeros(n, np.uint16),
            }
            rc = self.lib.dbg_export_nodes(self.h, *(a[f].ctypes.data for f in
                                                     ('id', 'key_lo', 'key_hi', 'tag', 'cnt', 'flags', 'kept')))
            self.lib.check(self.h, rc)
            t = a['tag']
            if len(t) > 1 and not bool((t[1:] > t[:-1]).all()):     # the library exports in tag order; an older one did not
                order = np.argsort(t, kind='stable')
                a = {f: v[order] for f, v in a.items()}
            self._nodes = a
            self._nodes['batch'] = self.batch_of_keys(self._nodes['key_lo'], self._nodes['key_hi'])
        return self._nodes

    def _batch_bounds(self, name, batch_fn):
        """First position of every batch id in a batch-major list (n_batches + 1 entries), computed once per list."""
        if self._batch_cache is None:
            self._batch_cache = {}
        key = ('bounds', name)
        if key not in self._batch_cache:
            batch = np.asarray(batch_fn())
            nb = int(getattr(self, 'n_batches', 0) or (int(batch.max()) + 1 if len(batch) else 0))
            self._batch_cache[key] = np.searchsorted(batch, np.arange(nb + 2, dtype=batch.dtype))
        return self._batch_cache[key]

    def batch_of_keys(self, key_lo, key_hi):
        """Batch id of node keys (the bits above the 2k key bits; 0 for a graph built without batches)."""
        sh = 2 * self.k
        lo = np.asarray(key_lo, dtype=np.uint64)
        hi = np.asarray(key_hi, dtype=np.uint64)
        if sh >= 64:
            b = hi >> np.uint64(sh - 64)
        else:
            b = (lo >> np.uint64(sh)) | (hi << np.uint64(64 - sh))
        return (b & np.uint64(0xFFFFFFFF)).astype(np.uint32)

    def text_in(self, s):
        """A caller's string as the graph holds it (identity unless the handle was made with `charset`); a character
        outside the charset becomes byte 0xFF, which is never a symbol."""
        if self._chars_in is None:
            return s
        m = self._chars_in
        return ''.join(chr(m.get(ord(c), 0xFF)) for c in s)

    def text_out(self, s):
        return s if self._chars_out is None else s.translate(self._chars_out)

    def _decode(self, lo, hi):
        if self.alphabet is not None:
            out = decode_keys(lo, hi, self.k, self.sym_bits, np.frombuffer(self.alphabet.ljust(256, b'?'), dtype=np.uint8))
            return out if self._chars_out is None else [x.translate(self._chars_out) for x in out]
        return decode_keys(lo, hi, self.k)

    def labels(self):
        if self._labels is None:
            nd = self.nodes()
            self._labels = self._decode(nd['key_lo'], nd['key_hi'])
        return self._labels

    def edges(self):
        """Generic-alphabet engine: the distinct (k+1)-mers grouped by source node in the reference's E order
        (dbg_export_edges); src / dst are positions in nodes() order."""
        if self._edges is None:
            nd = self.nodes()
            n = int(self.stats().n_edges)
            a = {'src': np.zeros(n, np.uint32), 'dst': np.zeros(n, np.uint32), 'symbol': np.zeros(n, np.uint8),
                 'count': np.zeros(n, np.uint32), 'kept': np.zeros(n, np.uint8)}
            self.lib.check(self.h, self.lib.dbg_export_edges(self.h, *(a[f].ctypes.data if n else None for f in
                                                                    ('src', 'dst', 'symbol', 'count', 'kept'))))
            slot_to_idx = np.full(int(self.stats().table_slots), -1, np.int64)
            slot_to_idx[nd['id']] = np.arange(len(nd['id']))
            a['src'] = slot_to_idx[a['src']]
            a['dst'] = slot_to_idx[a['dst']]
            self._edges = a
        return self._edges

    def probe_stats(self, which='nodes'):
        """Probe-length statistics of the node table ('nodes') or the chunk table ('chunks') of the last build
        (dbg_probe_stats): the number of slots a lookup of an entry's key touches (linear probing from its home slot)."""
        hist = np.zeros(64, np.uint64)
        n, total, longest, slots = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.lib.check(self.h, self.lib.dbg_probe_stats(self.h, {'nodes': 0, 'chunks': 1}[which], hist.ctypes.data, C.byref(n),
                                                        C.byref(total), C.byref(longest), C.byref(slots)))
        cum = np.cumsum(hist)
        q = lambda f: int(np.searchsorted(cum, f * n.value) + 1) if n.value else 0       # noqa: E731
        return {'entries': int(n.value), 'slots': int(slots.value), 'load': n.value / max(1, slots.value),
                'mean': total.value / max(1, n.value), 'p50': q(0.5), 'p99': q(0.99), 'max': int(longest.value),
                'hist': hist.astype(np.int64).tolist()}

    def degree_balance(self):
        """TRUE in-degree (distinct predecessors), out-degree (distinct successors) and balance = out - in of every
        node of the unpruned graph, in nodes() order (dbg_degree_balance).  An extra output: the reference keeps only
        `outdegree` and a 0/1 first-seen `indegree` flag (debruijn.py:129-143; SURVEY.md 9 delta 4)."""
        nd = self.nodes()
        n = len(nd['id'])
        if self.alphabet is not None:                                   # generic-alphabet engine: from its edge list
            e = self.edges()
            ind = np.bincount(e['dst'], minlength=n).astype(np.uint8)
            out = np.bincount(e['src'], minlength=n).astype(np.uint8)
            return {'indegree': ind, 'outdegree': out, 'balance': out.astype(np.int8) - ind.astype(np.int8)}
        ind, out, bal = np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(n, np.int8)
        if n:
            ids = np.ascontiguousarray(nd['id'], dtype=np.uint32)
            self.lib.check(self.h, self.lib.dbg_degree_balance(self.h, ids.ctypes.data, n, ind.ctypes.data, out.ctypes.data,
                                                               bal.ctypes.data))
        return {'indegree': ind, 'outdegree': out, 'balance': bal}

    def twin_nodes(self):
        """Per node, in nodes() order: position (in nodes() order) of the node that is its reverse complement, -1 if the
        graph does not hold it, and whether the node is the canonical one of the pair (k-mer <= its reverse complement).
        After a both_strands build every node has its twin (dbg_twin_nodes)."""
        nd = self.nodes()
        n = len(nd['id'])
        twin, canon = np.zeros(n, np.uint32), np.zeros(n, np.uint8)
        if n:
            ids = np.ascontiguousarray(nd['id'], dtype=np.uint32)
            self.lib.check(self.h, self.lib.dbg_twin_nodes(self.h, ids.ctypes.data, n, twin.ctypes.data, canon.ctypes.data))
        slot_to_idx = np.full(int(self.stats().table_slots) + 1, -1, np.int64)
        slot_to_idx[nd['id']] = np.arange(n)
        pos = np.where(twin == 0xFFFFFFFF, -1, slot_to_idx[np.minimum(twin, len(slot_to_idx) - 1)])
        return {'twin': pos, 'canonical': canon.astype(bool)}

    def label_of_ids(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        n = len(ids)
        lo = np.zeros(n, np.uint64)
        hi = np.zeros(n, np.uint64)
        if n:
            self.lib.check(self.h, self.lib.dbg_node_keys(self.h, ids.ctypes.data, n, lo.ctypes.data, hi.ctypes.data))
        return self._decode(lo, hi)

    def branch_ids(self):
        n = self.stats().n_branch
        out = np.zeros(n, np.uint32)
        self.lib.check(self.h, self.lib.dbg_branch_nodes(self.h, out.ctypes.data if n else None))
        return out

    def pulled_ids(self):
        n = self.stats().n_pulled
        out = np.zeros(n, np.uint32)
        self.lib.check(self.h, self.lib.dbg_pulled_nodes(self.h, out.ctypes.data if n else None))
        return out

    def pull_read_flags(self):
        out = np.zeros(self.n_reads, np.uint8)
        self.lib.check(self.h, self.lib.dbg_pull_reads(self.h, out.ctypes.data if self.n_reads else None))
        return out

    def lookup_edges(self, kplus1mers, batch=None):
        """edge_count_table values for a list of (k+1)-mer strings; 0 = absent.  `batch` = one
        batch id per query for a graph built by build_batched."""
        n = len(kplus1mers)
        out = np.zeros(n, np.uint32)
        if n == 0:
            return out
        if self._chars_in is not None:
            kplus1mers = [self.text_in(x) for x in kplus1mers]
        buf = ''.join(kplus1mers).encode('latin-1', 'replace')
        if len(buf) != n * (self.k + 1):
            raise ValueError('every query must have length k+1')
        if batch is None:
            self.lib.check(self.h, self.lib.dbg_lookup_edges(self.h, buf, n, out.ctypes.data))
        else:
            b = np.ascontiguousarray(np.broadcast_to(np.asarray(batch, dtype=np.uint32), (n,)))
            self.lib.check(self.h, self.lib.dbg_lookup_edges_batched(self.h, buf, b.ctypes.data, n, out.ctypes.data))
        return out

    # ---- output_contigs ------------------------------------------------------------------
    CONTIG_DTYPE = np.dtype([('start', np.uint32), ('n_nodes', np.uint32), ('terminal', np.uint32),
                             ('kind', np.uint32), ('seg_begin', np.uint64), ('seg_end', np.uint64)])

    def walk(self, stop_at_branch=True):
        n = C.c_uint64()
        self.lib.check(self.h, self.lib.dbg_walk(self.h, 1 if stop_at_branch else 0, C.byref(n)))
        out = np.empty(n.value, dtype=self.CONTIG_DTYPE)           # same layout as dbg_contig_t (32 bytes)
        self.lib.check(self.h, self.lib.dbg_contigs(self.h, out.ctypes.data if n.value else None))
        self._contigs = out
        return out

    def contig_descriptors(self):
        return self._contigs

    def contig_lengths(self):
        return self._contigs['n_nodes'].astype(np.int64) + (self.k - 1)

    def contig_strings(self, idx=None):
        """Materialise contig strings (all of them by default)."""
        if idx is None:
            idx = np.arange(len(self._contigs), dtype=np.uint64)
        idx = np.ascontiguousarray(idx, dtype=np.uint64)
        if len(idx) == 0:
            return []
        lens = self._contigs['n_nodes'][idx].astype(np.uint64) + np.uint64(self.k - 1)
        off = np.zeros(len(idx) + 1, np.uint64)
        np.cumsum(lens, out=off[1:])
        total = int(off[-1])
        if total > MAX_STRING_BYTES:
            raise MemoryError('dbg_b200: %d contig characters requested at once; use '
                              'contig_descriptors() / contig_strings(idx) in batches' % total)
        buf = np.empty(total, np.uint8)
        self.lib.check(self.h, self.lib.dbg_materialize(self.h, idx.ctypes.data, len(idx), buf.ctypes.data,
                                                        off.ctypes.data))
        raw = self.text_out(buf.tobytes().decode('latin-1'))
        o = off.tolist()
        return [raw[o[i]:o[i + 1]] for i in range(len(idx))]

    def contig_scores(self, idx=None):
        """getScore (II_assembleFromReads.py:14-18) of contigs, computed on the device."""
        if idx is None:
            idx = np.arange(len(self._contigs), dtype=np.uint64)
        idx = np.ascontiguousarray(idx, dtype=np.uint64)
        out = np.zeros(len(idx), np.uint64)
        if len(idx):
            self.lib.check(self.h, self.lib.dbg_contig_scores(self.h, idx.ctypes.data, len(idx), out.ctypes.data))
        return out


class _BatchStats:
    __slots__ = ('n_nodes', 'n_edges', 'n_branch', 'n_starts', 'n_pulled', 'n_pull_reads')


class BatchGraph:
    """One batch of a graph built by Graph.build_batched, with the part of Graph's interface the
    views and debruijn.output_contigs use.  Tags carry the global read index and reads are
    batch-major, so every per-batch list is a contiguous slice of the whole graph's list, already in
    the reference's order."""

    def __init__(self, parent, b):
        self.parent, self.b = parent, int(b)
        self.k, self.device = parent.k, parent.device
        bounds = parent._batch_bounds('nodes', lambda: parent.nodes()['batch'])   # nodes are sorted by tag = batch-major
        self._lo, self._hi = int(bounds[self.b]), int(bounds[self.b + 1])
        self._nodes = None
        self._labels = None

    def nodes(self):
        if self._nodes is None:
            nd = self.parent.nodes()
            self._nodes = {f: v[self._lo:self._hi] for f, v in nd.items()}
        return self._nodes

    def labels(self):
        if self._labels is None:
            nd = self.nodes()
            self._labels = decode_keys(nd['key_lo'], nd['key_hi'], self.k)
        return self._labels

    def stats(self):
        nd = self.nodes()
        s = _BatchStats()
        s.n_nodes = self._hi - self._lo
        s.n_edges = int(np.count_nonzero(nd['cnt']))
        s.n_branch = int(np.count_nonzero(nd['flags'] & N.NODE_BRANCH))
        s.n_starts = int(np.count_nonzero(nd['flags'] & N.NODE_START))
        s.n_pulled = int(np.count_nonzero(nd['flags'] & N.NODE_PULLED))
        return s

    def lookup_edges(self, kplus1mers):
        return self.parent.lookup_edges(kplus1mers, batch=self.b)

    def _split(self, name, ids_fn):
        """Slice of a whole-graph id list (reference order = batch-major) that belongs to this batch;
        the batch of every id is looked up once per graph and cached on the parent."""
        p = self.parent
        if p._batch_cache is None:
            p._batch_cache = {}
        if name not in p._batch_cache:
            ids = np.ascontiguousarray(ids_fn(), np.uint32)
            lo_k, hi_k = np.zeros(len(ids), np.uint64), np.zeros(len(ids), np.uint64)
            if len(ids):
                p.lib.check(p.h, p.lib.dbg_node_keys(p.h, ids.ctypes.data, len(ids), lo_k.ctypes.data, hi_k.ctypes.data))
            p._batch_cache[name] = (ids, p.batch_of_keys(lo_k, hi_k))
        ids, batch = p._batch_cache[name]
        bounds = p._batch_bounds(name, lambda: batch)
        return ids, int(bounds[self.b]), int(bounds[self.b + 1])

    def branch_ids(self):
        ids, a, z = self._split('branch', self.parent.branch_ids)
        return ids[a:z]

    def pulled_ids(self):
        ids, a, z = self._split('pulled', self.parent.pulled_ids)
        return ids[a:z]

    def label_of_ids(self, ids):
        return self.parent.label_of_ids(ids)

    def walk(self, stop_at_branch=True):
        """The whole graph is walked once, the way construct_graph's `final` implies, and every batch takes its
        slice.  A batch without branch nodes walks the same either way (debruijn.py:304); a batch WITH branch nodes
        whose caller asks for the other behaviour (an emptied / kept branch_kmer, debruijn.py:326) gets a walk of its
        own with that flag, its strings materialised at once."""
        p = self.parent
        default = not p.final
        self._strings = None
        if bool(stop_at_branch) != default and self.stats().n_branch:
            starts = p.walk(stop_at_branch=bool(stop_at_branch))['start']
            ids = np.ascontiguousarray(starts, np.uint32)
            lo_k, hi_k = np.zeros(len(ids), np.uint64), np.zeros(len(ids), np.uint64)
            if len(ids):
                p.lib.check(p.h, p.lib.dbg_node_keys(p.h, ids.ctypes.data, len(ids), lo_k.ctypes.data, hi_k.ctypes.data))
            a, z = np.searchsorted(p.batch_of_keys(lo_k, hi_k), [self.b, self.b + 1])
            self._contig_range = (int(a), int(z))
            self._strings = p.contig_strings(np.arange(a, z, dtype=np.uint64))
            if p._batch_cache:
                p._batch_cache.pop('contigs', None)                      # the handle no longer holds the default walk
                p._batch_cache.pop(('bounds', 'contigs'), None)
                p._batch_cache.pop('strings', None)
            return p.contig_descriptors()[a:z]
        _, a, z = self._split('contigs', lambda: p.walk(stop_at_branch=default)['start'])
        self._contig_range = (a, z)                                    # contigs come in start order = batch-major
        return p.contig_descriptors()[a:z]

    STRINGS_AT_ONCE = 1 << 28            # characters: below this the strings of ALL batches come over in one call

    def contig_strings(self):
        if getattr(self, '_strings', None) is not None:
            return self._strings
        a, z = self._contig_range
        p = self.parent
        cache = p._batch_cache
        if cache is not None and 'contigs' in cache:
            # the handle holds the default walk of the whole graph (_split('contigs')): one dbg_materialize for every
            # batch instead of one per batch (a call costs ~0.5 ms of copies and synchronisation whatever its size)
            if 'strings' not in cache:
                d = p.contig_descriptors()
                total = int(d['n_nodes'].sum()) + (self.k - 1) * len(d)
                cache['strings'] = p.contig_strings() if total <= self.STRINGS_AT_ONCE else None
            if cache['strings'] is not None:
                return cache['strings'][a:z]
        return p.contig_strings(np.arange(a, z, dtype=np.uint64))


# ---------------------------------------------------------------------------------------------
# lazy views handed back through the reference's return tuple (debruijn.py:285)
# ---------------------------------------------------------------------------------------------
class _View(Mapping):
    def __init__(self, graph):
        self._graph = graph
        self._d = None
        self._n = None

    def _dict(self):
        if self._d is None:
            self._d = self._materialise()
        return self._d

    def __getitem__(self, key):
        return self._dict()[key]

    def __iter__(self):
        return iter(self._dict())

    def __len__(self):
        if self._n is None:                              # one ctypes round trip, not one per call
            self._n = int(self._graph.stats().n_nodes)
        return self._n

    def __repr__(self):
        return '<%s of %d entries on cuda:%d>' % (type(self).__name__, len(self), self._graph.device)


class Vertices(_View):
    """label -> Node, first-occurrence order (debruijn.py:121-143)."""

    def _materialise(self):
        g = self._graph
        nd = g.nodes()
        labels = g.labels()
        indeg = ((nd['flags'] & N.NODE_START) == 0).astype(np.int64).tolist()
        if getattr(g, 'alphabet', None) is not None:                     # any number of distinct successors
            outdeg = np.bincount(g.edges()['src'], minlength=len(labels)).tolist()
        else:
            outdeg = ((nd['kept'] >> 12) & 7).tolist()
        d = {}
        for lab, i, o in zip(labels, indeg, outdeg):
            node = Node(lab)
            node.indegree = i
            node.outdegree = o
            d[lab] = node
        return d


class Edges(_View):
    """label -> kept successors in E order; pulled nodes are absent (debruijn.py:150-166, 254)."""

    def _materialise(self):
        g = self._graph
        nd = g.nodes()
        labels = g.labels()
        kept = nd['kept'].tolist()
        pulled = (nd['flags'] & N.NODE_PULLED).tolist()
        d = {}
        if getattr(g, 'alphabet', None) is not None:
            for lab, p in zip(labels, pulled):
                if not p:
                    d[lab] = []
            e = g.edges()
            for s_, t_, kp in zip(e['src'].tolist(), e['dst'].tolist(), e['kept'].tolist()):   # E order inside a node
                if kp and not pulled[s_]:
                    d[labels[s_]].append(labels[t_])
            return d
        for lab, w, p in zip(labels, kept, pulled):
            if p:
                continue
            tail = lab[1:]
            d[lab] = [tail + 'ACGT'[(w >> (2 * j)) & 3] for j in range((w >> 8) & 7)]
        return d

    def __len__(self):
        if self._n is None:
            s = self._graph.stats()
            self._n = int(s.n_nodes - s.n_pulled)
        return self._n


class EdgeCountTable(Mapping):
    """(k+1)-mer -> multiplicity (debruijn.py:213-222).  Small tables are pulled to the host
    once; large ones answer ``[]`` through dbg_lookup_edges."""

    HOST_LIMIT = 4_000_000

    def __init__(self, graph):
        self._graph = graph
        self._d = None

    def _dict(self):
        if self._d is None:
            g = self._graph
            nd = g.nodes()
            labels = g.labels()
            cnt = nd['cnt']
            d = {}
            if getattr(g, 'alphabet', None) is not None:
                e = g.edges()
                for s_, sym, c in zip(e['src'].tolist(), g.text_out(e['symbol'].tobytes().decode('latin-1')), e['count'].tolist()):
                    d[labels[s_] + sym] = c
                self._d = d
                return d
            rows, cols = np.nonzero(cnt)
            vals = cnt[rows, cols].tolist()
            for r, c, v in zip(rows.tolist(), cols.tolist(), vals):
                d[labels[r] + 'ACGT'[c]] = v
            self._d = d
        return self._d

    def __getitem__(self, key):
        g = self._graph
        if self._d is not None or g.stats().n_edges <= self.HOST_LIMIT:
            return self._dict()[key]
        if not isinstance(key, str) or len(key) != g.k + 1:
            raise KeyError(key)
        try:
            v = int(g.lookup_edges([key])[0])
        except ValueError:                               # a character outside the alphabet: not a key of the dict
            raise KeyError(key) from None
        if v == 0:
            raise KeyError(key)
        return v

    def __iter__(self):
        return iter(self._dict())

    def __len__(self):
        return int(self._graph.stats().n_edges)

    def __repr__(self):
        return '<EdgeCountTable of %d entries on cuda:%d>' % (len(self), self._graph.device)


class KmerList(list):
    """A plain list of k-mer strings (branch_kmer / already_pull_out) that remembers its graph."""

    def __init__(self, items=(), graph=None):
        super().__init__(items)
        self._graph = graph


def _pack_any(strings):
    """list[str] -> (uint8 bytes, uint64 offsets); any characters (latin-1), no alphabet check."""
    off = np.zeros(len(strings) + 1, np.uint64)
    if strings:
        np.cumsum(np.fromiter((len(x) for x in strings), dtype=np.uint64, count=len(strings)), out=off[1:])
    data = np.frombuffer(''.join(strings).encode('latin-1'), dtype=np.uint8) if strings else np.zeros(0, np.uint8)
    return np.ascontiguousarray(data), off


def support_read_scores(contigs, score_table, device=0, graph=None):
    """findSupportReadScore (IV_sortOutputs.py:10-15) for every contig of a list, on the GPU (include/dbg_b200.h,
    dbg_support_scores): float64 sums in the score_table's own key order, bit-identical with the reference's
    left-to-right `score +=`.  `score_table` is the reference's dict read -> score."""
    contigs = list(contigs)
    reads = list(score_table.keys())
    for x in contigs + reads:
        if any(ord(ch) > 255 for ch in x):
            raise ValueError('dbg_b200: support scores take latin-1 strings')
    out = np.zeros(len(contigs), np.float64)
    if not contigs:
        return out
    rb, ro = _pack_any(reads)
    cb, co = _pack_any(contigs)
    sc = np.ascontiguousarray(np.fromiter((float(score_table[r]) for r in reads), dtype=np.float64, count=len(reads)))
    g = graph if graph is not None else Graph(1, device=device)     # `graph`: a handle to borrow (its k does not matter)
    try:
        g.lib.check(g.h, g.lib.dbg_support_scores(g.h, rb.ctypes.data if len(rb) else None, ro.ctypes.data, len(reads),
                                                  sc.ctypes.data if len(sc) else None, cb.ctypes.data if len(cb) else None,
                                                  co.ctypes.data, len(contigs), out.ctypes.data))
    finally:
        if graph is None:
            g.close()
    return out
