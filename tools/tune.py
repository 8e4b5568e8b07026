"""Kernel tuning harness (GPU box): time the build + walk stages of one bench config for a set of
library variants and environment settings.

    python tools/tune.py c2 [--steps 10] [--lib path/to/variant.so ...] [--env NAME=V1,V2 ...]

Every (lib, env combination) is run in a fresh subprocess (the library reads its tuning
environment variables at call time, but a variant is a different .so).  Prints one line each with
the per-stage CUDA-event times of dbg_stats_t, averaged over --steps builds after 3 warm-ups.
"""
import itertools
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'py-debruijn_b200')]

STAGES = ('ms_pack', 'ms_estimate', 'ms_table_init', 'ms_insert', 'ms_chunk_insert', 'ms_expand_sketch', 'ms_expand', 'ms_region_order',
          'ms_finalize', 'ms_order', 'ms_lists', 'ms_tips', 'ms_pull_reads', 'ms_build_total', 'ms_walk_rank',
          'ms_walk_emit', 'ms_walk_total')


def worker(cfg_name, steps, lib_path):
    import time

    import numpy as np
    import torch

    import bench
    from dbg_b200 import Graph, _native
    cfg = bench.CONFIGS[cfg_name]
    bases, off = bench.make_reads(cfg)
    dev = torch.device('cuda', 0)
    d_bases = torch.from_numpy(bases).to(dev)
    d_off = torch.from_numpy(off.view(np.int64)).to(dev)
    lib = _native.Library(lib_path) if lib_path else _native.get()
    g = Graph(cfg['k'], device=0, lib=lib)
    acc = {}
    wall = 0.0
    for i in range(3 + steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g.build_device(d_bases.data_ptr(), d_off.data_ptr(), len(off) - 1, len(bases), threshold=cfg['thr'])
        g.walk(stop_at_branch=True)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if i >= 3:
            st = g.stats()
            wall += (t1 - t0) * 1e3 / steps
            for f in STAGES:
                acc[f] = acc.get(f, 0.0) + getattr(st, f) / steps
    st = g.stats()
    out = {f: round(v, 4) for f, v in acc.items()}
    out.update(wall_ms=round(wall, 4), passes=int(st.n_insert_passes), nodes=int(st.n_nodes),
               chunk_slots=int(st.chunk_slots), launches=int(st.n_launches))
    print('RESULT ' + json.dumps(out), flush=True)


def main():
    args = sys.argv[1:]
    if args and args[0] == '--worker':
        worker(args[1], int(args[2]), args[3] if len(args) > 3 and args[3] != '-' else None)
        return
    cfg = args[0]
    steps, libs, envs = 10, [], []
    i = 1
    while i < len(args):
        if args[i] == '--steps':
            steps = int(args[i + 1]); i += 2
        elif args[i] == '--lib':
            libs.append(args[i + 1]); i += 2
        elif args[i] == '--env':
            name, vals = args[i + 1].split('=', 1)
            envs.append([(name, v) for v in vals.split(',')]); i += 2
        else:
            raise SystemExit('unknown argument ' + args[i])
    libs = libs or ['-']
    for lib in libs:
        for combo in itertools.product(*envs) if envs else [()]:
            env = dict(os.environ)
            for name, v in combo:
                if v == '':
                    env.pop(name, None)
                else:
                    env[name] = v
            p = subprocess.run([sys.executable, os.path.abspath(__file__), '--worker', cfg, str(steps), lib],
                               env=env, capture_output=True, text=True)
            res = [ln[7:] for ln in p.stdout.splitlines() if ln.startswith('RESULT ')]
            tag = '%s %s' % (os.path.basename(lib), ' '.join('%s=%s' % c for c in combo))
            print(tag, '|', res[0] if res else 'FAILED: ' + p.stderr[-400:], flush=True)


if __name__ == '__main__':
    main()
