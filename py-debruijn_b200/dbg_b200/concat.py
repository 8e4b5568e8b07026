"""III_concatenateOutpus.py:53-146 on the GPU (C ABI: the dbg_concat_* calls of
include/dbg_b200.h, kernels in csrc/dbg_concat.cuh).

The reference's randomised greedy merge runs as one persistent kernel per k; Python's Mersenne Twister state goes in
and comes back out, so after ``random.seed(s)`` the result -- the merged contigs, their order, every printed line --
is what the unmodified script produces after ``random.seed(s)``.  The host is only needed when an iteration meets a
swap-tolerant match (:85-131): it scores the two candidate strings with ``support_read_scores`` (the GPU
findSupportReadScore), prints the two lines the script prints there and lets the kernel go on.
"""
from __future__ import annotations

import ctypes as C
import random as _random

import numpy as np

from . import _native
from .graph import Graph, _pack_any, support_read_scores


class ConcatSession:
    """The contig list of III_concatenateOutpus.py in HBM: slot i starts as contig i, merges rewrite slots in
    place, removed slots stay behind as dead entries (the list order is the slot order of the live ones)."""

    def __init__(self, contigs, device=None):
        contigs = list(contigs)
        if not contigs:
            raise ValueError('max() arg is an empty sequence')          # what :53 raises on an empty FASTA
        for x in contigs:
            if any(ord(ch) > 255 for ch in x):
                raise ValueError('dbg_b200: contigs must be latin-1 strings')
        self.n = len(contigs)
        self._g = Graph(1, device=device)
        self.lib, self.h = self._g.lib, self._g.h
        data, off = _pack_any(contigs)
        self.lib.check(self.h, self.lib.dbg_concat_begin(self.h, data.ctypes.data if len(data) else None, off.ctypes.data, self.n))
        self.launches = 0

    def close(self):
        if self._g is not None:
            self.lib.dbg_concat_end(self.h)
            self._g.close()
            self._g = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def run(self, k, mt, max_iter, resume):
        st = _native.ConcatStatus()
        self.lib.check(self.h, self.lib.dbg_concat_run(self.h, int(k), int(max_iter), mt.ctypes.data, 1 if resume else 0, C.byref(st)))
        self.launches += 1
        return st

    def pending(self, n):
        slots, codes = np.zeros(n, np.uint32), np.zeros(n, np.uint8)
        self.lib.check(self.h, self.lib.dbg_concat_pending(self.h, slots.ctypes.data, codes.ctypes.data))
        return slots, codes

    def choose(self, choice):
        choice = np.ascontiguousarray(choice, np.uint8)
        self.lib.check(self.h, self.lib.dbg_concat_choose(self.h, choice.ctypes.data))

    def lengths(self):
        ln, alive = np.zeros(self.n, np.uint32), np.zeros(self.n, np.uint8)
        self.lib.check(self.h, self.lib.dbg_concat_lengths(self.h, ln.ctypes.data, alive.ctypes.data))
        return ln, alive

    def strings(self, slots, lengths=None):
        slots = np.ascontiguousarray(slots, np.uint32)
        if len(slots) == 0:
            return []
        if lengths is None:
            lengths = self.lengths()[0]
        off = np.zeros(len(slots) + 1, np.uint64)
        np.cumsum(lengths[slots].astype(np.uint64), out=off[1:])
        buf = np.zeros(max(int(off[-1]), 1), np.uint8)
        self.lib.check(self.h, self.lib.dbg_concat_strings(self.h, slots.ctypes.data, len(slots), buf.ctypes.data, off.ctypes.data))
        raw = buf.tobytes().decode('latin-1')
        return [raw[int(off[j]):int(off[j + 1])] for j in range(len(slots))]

    def alive_strings(self):
        ln, alive = self.lengths()
        return self.strings(np.flatnonzero(alive).astype(np.uint32), ln)

    def dedup(self):
        keep = np.zeros(self.n, np.uint8)
        self.lib.check(self.h, self.lib.dbg_concat_dedup(self.h, keep.ctypes.data))
        return keep


def _mt_words(rng):
    version, state, gauss = rng.getstate()
    if version != 3 or len(state) != 625:
        raise ValueError('dbg_b200: the generator must be a Mersenne Twister random.Random')
    return np.array(state, dtype=np.uint32), gauss


def concatenate(session, sequences_scores, rng=None, ks=None, out=print, max_iter=100000):
    """III_concatenateOutpus.py:57-137 on a ConcatSession.  `rng` is a random.Random (default: the `random` module's
    own generator, as in the script); it is advanced exactly as `random.choice` (:65) advances it."""
    inst = rng if rng is not None else _random._inst
    if ks is None:
        ks = [i for i in range(20, 5, -1)]                                           # :57
    mt, gauss = _mt_words(inst)
    try:
        for k in ks:
            st = session.run(k, mt, max_iter, resume=False)
            while st.status == _native.DBG_CONCAT_NEED_CHOICE:
                # swap-tolerant matches (:85-131): score result_1 / result_2 of each, print what the script prints
                slots, codes = session.pending(st.n_pending)
                ln = session.lengths()[0]
                strs = session.strings(np.concatenate([[st.current], slots]).astype(np.uint32), ln)
                current, contigs = strs[0], strs[1:]
                cand, where = [], []
                for j, (contig, cd) in enumerate(zip(contigs, codes.tolist())):
                    if cd == 3:
                        cand += [contig + current[k:], contig[:len(contig) - k] + current]       # :91-92
                        where.append(j)
                    elif cd == 4:
                        cand += [current + contig[k:], current[:len(current) - k] + contig]       # :114-115
                        where.append(j)
                scores = support_read_scores(cand, sequences_scores, graph=session._g)
                choice = np.zeros(len(slots), np.uint8)
                for t, j in enumerate(where):
                    first = scores[2 * t] > scores[2 * t + 1]                                      # :95, :118
                    choice[j] = 1 if first else 2
                    out('%s   %s' % (current, contigs[j]))                                         # :102, :126
                    out(cand[2 * t] if first else cand[2 * t + 1])                                 # :103, :127
                session.choose(choice)
                st = session.run(k, mt, max_iter, resume=True)
            out('max length after concat(k = {}):  {}'.format(k, st.max_len))                      # :137
            out(str(st.n_alive))                                                                   # :138
    finally:
        inst.setstate((3, tuple(int(x) for x in mt), gauss))
    return session


def concatenate_outputs(contigs, sequences_scores, rng=None, out=print, ks=None, max_iter=100000, device=None):
    """III_concatenateOutpus.py:53-146 without the file handling: the prints around the loop, the loop (:57-137),
    the substring dedup (:139-142).  Returns `outputs`, the list the script writes to {froot}_concatenated.fasta."""
    contigs = list(contigs)
    with ConcatSession(contigs, device=device) as s:
        out('max length before concat:  %d' % len(max(contigs, key=lambda x: len(x))))             # :53
        out(str(len(contigs)))                                                                     # :55
        concatenate(s, sequences_scores, rng=rng, ks=ks, out=out, max_iter=max_iter)
        keep = s.dedup()
        outputs = s.strings(np.flatnonzero(keep).astype(np.uint32))
    out(str(len(outputs)))                                                                         # :144
    out('max length after concat:  %d' % len(max(outputs, key=lambda x: len(x))))                  # :145
    return outputs
