"""Where one construct_graph + output_contigs call of the C1 shape (1,799 reads, k=31) spends its 3 ms: pieces timed apart."""
import contextlib, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'py-debruijn_b200')]
import numpy as np
import bench
import debruijn as db
from dbg_b200 import Graph
from dbg_b200.graph import pack_reads
cfg = bench.CONFIGS['c1']
bases, off = bench.make_reads(cfg)
raw = bases.tobytes().decode('ascii')
reads = [raw[int(off[i]):int(off[i + 1])] for i in range(len(off) - 1)]
k, thr = cfg['k'], cfg['thr']
def best(f, n=20):
    b = 1e9
    for _ in range(n):
        t0 = time.perf_counter(); r = f(); b = min(b, time.perf_counter() - t0)
    return b * 1e3, r
def full():
    with contextlib.redirect_stdout(io.StringIO()):
        g, pull, branch, pulled, ect = db.construct_graph(reads, k, threshold=thr)
        return db.output_contigs(g, branch, pulled)
full(); full()
print('full call            %.3f ms' % best(full)[0])
print('pack_reads           %.3f ms' % best(lambda: pack_reads(reads))[0])
print('Graph() + close()    %.3f ms' % best(lambda: Graph(k).close())[0])
packed = pack_reads(reads)
def build_new():
    g = Graph(k); g.build(reads, threshold=thr, packed=packed); return g
ms, g = best(build_new); print('Graph() + build      %.3f ms' % ms)
print('build on a warm handle %.3f ms' % best(lambda: g.build(reads, threshold=thr, packed=packed))[0])
print('stats                %.3f ms' % best(lambda: g.stats())[0])
print('walk                 %.3f ms' % best(lambda: g.walk(stop_at_branch=True))[0])
print('contig_strings       %.3f ms' % best(lambda: g.contig_strings())[0])
print('branch+pulled labels %.3f ms' % best(lambda: (g.label_of_ids(g.branch_ids()), g.label_of_ids(g.pulled_ids())))[0])
