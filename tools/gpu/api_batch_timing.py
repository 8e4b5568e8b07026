"""construct_graph_batch + output_contigs through the Python API (bench.py e2e_api, c4_100_templates) -- wall clock, best of 5,
with the pieces timed separately."""
import contextlib, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'py-debruijn_b200')]
import bench
import debruijn as db
c4 = dict(bench.CONFIGS['c4']); c4['T'] = 100
b4, o4, rb, boff = bench.make_reads_batched(c4)
raw = b4.tobytes().decode('ascii'); L = c4['L']
batches = [[raw[i * L:(i + 1) * L] for i in range(int(boff[b]), int(boff[b + 1]))] for b in range(100)]
best = None
for it in range(6):
    with contextlib.redirect_stdout(io.StringIO()):
        t0 = time.perf_counter()
        res = db.construct_graph_batch(batches, c4['k'], threshold=c4['thr'], final=False)
        t1 = time.perf_counter()
        got = [db.output_contigs(g, branch, pulled) for (g, pull, branch, pulled, ect) in res]
        t2 = time.perf_counter()
    t = ((t2 - t0) * 1e3, (t1 - t0) * 1e3, (t2 - t1) * 1e3)
    if it and (best is None or t[0] < best[0]):
        best = t
print('c4_100_templates: %d reads, %d contigs, %d chars; best of 5: total %.2f ms = construct_graph_batch %.2f + output_contigs %.2f'
      % (len(o4) - 1, sum(map(len, got)), sum(len(c) for cs in got for c in cs), *best))
