"""Where the host-buffer step (bench.py `e2e`) spends its time at C2: per-stage CUDA-event times of dbg_stats_t after a
dbg_build from pinned host memory, next to the same stages of a device-resident build and the wall clock of both."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'py-debruijn_b200')]
import numpy as np
import torch
import bench
from dbg_b200 import Graph
from tools.tune import STAGES

cfg = bench.CONFIGS['c2']
bases, off = bench.make_reads(cfg)
dev = torch.device('cuda', 0)
hb = torch.from_numpy(bases).pin_memory(); ho = torch.from_numpy(off.view(np.int64)).pin_memory()
hb_np, ho_np = hb.numpy(), ho.numpy().view(np.uint64)
d_bases = hb.to(dev); d_off = ho.to(dev)
g = Graph(cfg['k'], device=0)
def dev_step():
    g.build_device(d_bases.data_ptr(), d_off.data_ptr(), len(off) - 1, len(bases), threshold=cfg['thr'])
    g.walk(stop_at_branch=True)
def host_step():
    g.build(None, threshold=cfg['thr'], final=False, packed=(hb_np, ho_np))
    g.walk(stop_at_branch=True)
for name, fn in (('device', dev_step), ('host', host_step), ('device', dev_step), ('host', host_step)):
    for _ in range(4): fn()
    acc = {}; wall = 0.0; n = 20
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); wall += (time.perf_counter() - t0) * 1e3 / n
        st = g.stats()
        for f in STAGES + ('ms_h2d',):
            acc[f] = acc.get(f, 0.0) + getattr(st, f, 0.0) / n
    print(name, 'wall %.3f ms' % wall, {f: round(v, 3) for f, v in acc.items() if v > 0.004}, flush=True)
