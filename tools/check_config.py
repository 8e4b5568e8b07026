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
sys.path[:0] = [ROOT, os.path.join(ROOT, 'py-debruijn_b200'), os.path.join(ROOT, 'tests')]
import bench                                   # noqa: E402
from dbg_b200 import Graph                     # noqa: E402
from oracle import c_oracle                    # noqa: E402
from oracle import compare as fullsize        # noqa: E402


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
    bad = fullsize.compare_with_oracle(g, res, final, report=lambda name, good: print('  %-28s %s' % (name, 'OK' if good else 'MISMATCH'), flush=True))
    ok = not bad
    print('PARITY', 'OK' if ok else 'FAILED', flush=True)
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
