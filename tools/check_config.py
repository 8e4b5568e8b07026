"""Full-size parity check of one bench config on the GPU against the C oracle (descriptor level).

    python tools/check_config.py c3 [--final]

Compares: node count, sorted (key, tag, counts) table, branch list, pulled list, pull-out read
flags, and the contig descriptors (start key, n_nodes, terminal key, kind) in order.
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
    name = sys.argv[1]
    final = '--final' in sys.argv
    G = None
    for a in sys.argv[2:]:
        if a.startswith('--G='):
            G = int(a[4:])
    cfg = bench.CONFIGS[name]
    bases, off = bench.make_reads(cfg, G=G)
    k, thr = cfg['k'], cfg['thr']
    print('config', name, 'reads', len(off) - 1, 'bases', len(bases), 'k', k, 'final', final, flush=True)
    g = Graph(k)
    t0 = time.time()
    g.build(None, threshold=thr, final=final, packed=(bases, off))
    g.walk(stop_at_branch=not final)
    t1 = time.time()
    st = g.stats()
    print('gpu: %.3f s' % (t1 - t0), {f: getattr(st, f) for f in ('n_windows', 'n_nodes', 'n_edges', 'n_multi', 'n_branch', 'n_starts',
          'n_pulled', 'n_pull_reads', 'n_contigs', 'n_cycle_starts', 'n_insert_passes', 'ms_build_total', 'ms_insert',
          'ms_finalize', 'ms_order', 'ms_lists', 'ms_tips', 'ms_pull_reads', 'ms_walk_total')}, flush=True)
    t0 = time.time()
    res = c_oracle.run(None, k, thr, final, strings=False, packed=(bases.tobytes(), off))
    print('oracle: %.1f s' % (time.time() - t0), flush=True)
    ok = True

    def check(name, cond):
        nonlocal ok
        print('  %-28s %s' % (name, 'OK' if cond else 'MISMATCH'), flush=True)
        ok &= bool(cond)

    nd = g.nodes()
    check('n_nodes', st.n_nodes == res.n_nodes)
    check('n_windows', st.n_windows == res.n_windows)
    check('node keys (dict order)', np.array_equal(nd['key_lo'], res.key_lo) and np.array_equal(nd['key_hi'], res.key_hi))
    check('first-occurrence tags', np.array_equal(nd['tag'], res.tag))
    check('edge counts', np.array_equal(nd['cnt'], res.cnt))
    indeg = ((nd['flags'] & 2) == 0).astype(np.uint8)
    check('indegree flags', np.array_equal(indeg, res.indeg))
    check('outdegree', np.array_equal(((nd['kept'] >> 12) & 7).astype(np.uint8), res.outdeg))
    nk = ((nd['kept'] >> 8) & 7).astype(np.uint8)
    check('kept successor counts', np.array_equal(nk, res.n_kept))
    kb = np.stack([(nd['kept'] >> (2 * j)) & 3 for j in range(4)], axis=1).astype(np.uint8)
    msk = np.arange(4)[None, :] < nk[:, None]
    check('kept successors (E order)', np.array_equal(kb * msk, res.kept_base * msk))
    slot_to_idx = np.full(int(st.table_slots), -1, np.int64)
    slot_to_idx[nd['id']] = np.arange(len(nd['id']))
    check('branch list', np.array_equal(slot_to_idx[g.branch_ids()], res.branch.astype(np.int64)) if not final else True)
    check('pulled list', np.array_equal(slot_to_idx[g.pulled_ids()], res.pulled.astype(np.int64)))
    if not final:
        check('pull-out read flags', np.array_equal(g.pull_read_flags(), res.read_flags))
    d = g.contig_descriptors()
    check('n_contigs', len(d) == len(res.contigs))
    if len(d) == len(res.contigs):
        check('contig starts', np.array_equal(slot_to_idx[d['start']], res.contigs[:, 0].astype(np.int64)))
        check('contig n_nodes', np.array_equal(d['n_nodes'], res.contigs[:, 1]))
        check('contig terminals', np.array_equal(slot_to_idx[d['terminal']], res.contigs[:, 2].astype(np.int64)))
        check('contig kinds', np.array_equal(d['kind'], res.contigs[:, 3]))
    print('PARITY', 'OK' if ok else 'FAILED', flush=True)
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
