python -m pytest tests/test_strands.py tests/test_degrees.py -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2ai_strands.log
python - <<'P' 2>&1 | tee -a gpurun_out/r2ai_strands.log
import sys, time
sys.path[:0] = ['/root/repo', '/root/repo/py-debruijn_b200']
import numpy as np, bench
from dbg_b200 import Graph
cfg = bench.CONFIGS['c2']
bases, off = bench.make_reads(cfg)
g = Graph(cfg['k'])
for both in (False, True, True):
    t = time.perf_counter()
    g.build(None, threshold=2, packed=(bases, off), both_strands=both); g.walk()
    ms = (time.perf_counter() - t) * 1e3
    st = g.stats()
    print('c2 both_strands=%s: %.2f ms (host buffers), reads %d nodes %d contigs %d' % (both, ms, st.n_reads, st.n_nodes, st.n_contigs))
tw = g.twin_nodes()
print('every node has a twin:', bool((tw['twin'] >= 0).all()), 'canonical nodes:', int(tw['canonical'].sum()), 'of', len(tw['twin']))
P
