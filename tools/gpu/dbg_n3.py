import sys, time, json
sys.path[:0] = ['/root/repo', '/root/repo/py-debruijn_b200']
import numpy as np
import bench
def n3(tag):
    t = time.perf_counter(); r = bench.n3_concat(None, None); print(tag, round(r['ms'], 1), 'ms', 'total call', round((time.perf_counter() - t) * 1e3), flush=True)
n3('fresh')
n3('again')
import torch
x = torch.zeros(1 << 20, device='cuda'); torch.cuda.synchronize()
n3('after torch init')
from dbg_b200 import Graph
cfg = bench.CONFIGS['c2']
bases, off = bench.make_reads(cfg)
g = Graph(cfg['k'], device=0)
g.build(None, threshold=2, final=False, packed=(bases, off)); g.walk()
n3('with a c2 handle alive')
hb = torch.from_numpy(bases).pin_memory()
n3('plus 125 MB pinned')
# where does the time go
from dbg_b200 import concat as cc
import random
orig_run = cc.ConcatSession.run
acc = {'run': 0.0, 'n': 0}
def timed_run(self, *a, **k):
    t = time.perf_counter(); r = orig_run(self, *a, **k); acc['run'] += time.perf_counter() - t; acc['n'] += 1; return r
cc.ConcatSession.run = timed_run
orig_srs = cc.support_read_scores
acc2 = {'t': 0.0, 'n': 0}
def timed_srs(*a, **k):
    t = time.perf_counter(); r = orig_srs(*a, **k); acc2['t'] += time.perf_counter() - t; acc2['n'] += 1; return r
cc.support_read_scores = timed_srs
n3('instrumented')
print('run()', acc, 'support', acc2)
g.close(); del hb
n3('after closing')
