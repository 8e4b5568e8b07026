"""Full-size parity check of the batched config (BASELINE.json configs[3]) on the GPU: ONE batched
pass over T templates against the C oracle run on single templates.

    python tools/check_batched.py [--T=10000] [--sample=300]

Checks, for a random sample of batches plus the first and last: node keys in dict order, tags
(relative to the batch's first read), edge counts, indegree flags, and the contig descriptors
(start, n_nodes, terminal, kind); and for ALL batches that node / contig counts add up.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'py-debruijn_b200')]
import bench                                   # noqa: E402
from dbg_b200 import Graph                     # noqa: E402
from oracle import c_oracle                    # noqa: E402


def main():
    T, sample = None, 300
    for a in sys.argv[1:]:
        if a.startswith('--T='):
            T = int(a[4:])
        if a.startswith('--sample='):
            sample = int(a[9:])
    cfg = dict(bench.CONFIGS['c4'])
    if T:
        cfg['T'] = T
    T = cfg['T']
    k, thr = cfg['k'], cfg['thr']
    bases, off, read_batch, batch_off = bench.make_reads_batched(cfg)
    print('c4: templates', T, 'reads', len(off) - 1, 'bases', len(bases), 'k', k, flush=True)
    g = Graph(k, key_words=2 if 2 * k + T.bit_length() > 64 else 0)
    for rep in range(2):
        t0 = time.time()
        g.build_batched_packed(bases, off, read_batch, T, threshold=thr, final=False)
        desc = g.walk(stop_at_branch=True)
        t1 = time.time()
    st = g.stats()
    nwin = bench.n_windows_of(off, k)
    print('gpu: %.3f s wall (second call), %.3g windows/s; ' % (t1 - t0, nwin / (t1 - t0)),
          {f: round(getattr(st, f), 3) if isinstance(getattr(st, f), float) else getattr(st, f) for f in
           ('n_windows', 'n_nodes', 'n_branch', 'n_starts', 'n_contigs', 'n_insert_passes', 'ms_build_total',
            'ms_estimate', 'ms_chunk_insert', 'ms_expand_sketch', 'ms_expand', 'ms_finalize', 'ms_lists', 'ms_walk_total')},
          flush=True)
    nd = g.nodes()
    ok = True
    nb_lo = np.searchsorted(nd['batch'], np.arange(T + 1))
    start_keys_lo = np.zeros(len(desc), np.uint64)
    start_keys_hi = np.zeros(len(desc), np.uint64)
    ids = np.ascontiguousarray(desc['start'], np.uint32)
    g.lib.check(g.h, g.lib.dbg_node_keys(g.h, ids.ctypes.data, len(ids), start_keys_lo.ctypes.data, start_keys_hi.ctypes.data))
    cb = g.batch_of_keys(start_keys_lo, start_keys_hi)
    c_lo = np.searchsorted(cb, np.arange(T + 1))
    slot_to_idx = np.full(int(st.table_slots), -1, np.int64)
    slot_to_idx[nd['id']] = np.arange(len(nd['id']))
    rng = np.random.default_rng(5)
    picks = sorted(set([0, T - 1] + rng.integers(0, T, sample).tolist()))
    kmask = np.uint64((1 << (2 * k)) - 1)
    t0 = time.time()
    tot_nodes = 0
    for b in picks:
        r0, r1 = int(batch_off[b]), int(batch_off[b + 1])
        sub_off = off[r0:r1 + 1] - off[r0]
        sub = bases[int(off[r0]):int(off[r1])]
        res = c_oracle.run(None, k, thr, False, strings=False, packed=(sub.tobytes(), sub_off))
        n0, n1 = int(nb_lo[b]), int(nb_lo[b + 1])
        good = (n1 - n0) == res.n_nodes
        if good:
            good &= np.array_equal(nd['key_lo'][n0:n1] & kmask, res.key_lo)
            good &= np.array_equal(nd['tag'][n0:n1] - (np.uint64(r0) << np.uint64(32)), res.tag)
            good &= np.array_equal(nd['cnt'][n0:n1], res.cnt)
            good &= np.array_equal(((nd['flags'][n0:n1] & 2) == 0).astype(np.uint8), res.indeg)
            a, z = int(c_lo[b]), int(c_lo[b + 1])
            good &= (z - a) == len(res.contigs)
            if good and z > a:
                d = desc[a:z]
                good &= np.array_equal(slot_to_idx[d['start']] - n0, res.contigs[:, 0].astype(np.int64))
                good &= np.array_equal(d['n_nodes'], res.contigs[:, 1])
                good &= np.array_equal(slot_to_idx[d['terminal']] - n0, res.contigs[:, 2].astype(np.int64))
                good &= np.array_equal(d['kind'], res.contigs[:, 3])
        if not good:
            print('  batch %d MISMATCH' % b, flush=True)
            ok = False
        tot_nodes += res.n_nodes
    print('checked %d batches against the C oracle in %.1f s' % (len(picks), time.time() - t0), flush=True)
    print('PARITY', 'OK' if ok else 'FAILED', flush=True)
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
