import sys, os
sys.path[:0] = ['/root/repo', '/root/repo/py-debruijn_b200', '/root/repo/tests']
import numpy as np
from dbg_b200 import Graph
from test_contention import _hot_reads
from oracle import c_oracle, compare
for k, n, G, L in ((47, 30000, 300, 100), (47, 100000, 300, 100), (47, 300000, 300, 100), (63, 300000, 300, 100), (33, 300000, 300, 100),
                   (47, 300000, 3000, 100), (47, 300000, 300, 60)):
    bases, off = _hot_reads(1, G, L, n, 0.0)
    g = Graph(k)
    for rep in range(2):
        try:
            g.build(None, threshold=2, final=False, packed=(bases, off)); g.walk(stop_at_branch=True)
            res = c_oracle.run(None, k, 2, False, strings=False, packed=(bases.tobytes(), off))
            st = g.stats()
            print(k, n, G, L, 'rep', rep, compare.compare_with_oracle(g, res, False), st.n_nodes, st.table_slots, st.n_regions, st.hll_estimate,
                  st.n_chunk_instances, st.chunk_slots, st.n_insert_passes, flush=True)
        except Exception as e:
            print(k, n, G, L, 'rep', rep, 'ERR', e, flush=True)
    g.close()
