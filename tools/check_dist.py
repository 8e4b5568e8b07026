"""torchrun --nproc-per-node N tools/check_dist.py [config] [G]: the N-GPU build (dist.build_sharded:
owner-computes walk on a sharded node table; replicated fallback when the graph has branch nodes) must
give exactly the single-GPU result -- node table in dict order (keys, tags, edge counts, flags, kept
successors) and contig descriptors by node key (dist.global_snapshot vs dist.single_snapshot)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'py-debruijn_b200')]
import bench                                   # noqa: E402
from dbg_b200 import Graph, dist as ddist      # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'c3'
    G = int(sys.argv[2]) if len(sys.argv) > 2 else 400_000
    rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(lr)
    dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
    cfg = dict(bench.CONFIGS[name])
    cfg['G'] = G
    bases, off = bench.make_reads(cfg)                      # every rank derives the same full read set
    n = len(off) - 1
    L = cfg['L']
    b = ddist.shard_bounds(n, world)
    lo, hi = b[rank], b[rank + 1]
    local = (bases[lo * L:hi * L].copy(), (off[lo:hi + 1] - off[lo]).copy())
    g = Graph(cfg['k'], device=lr)
    res = ddist.build_sharded(g, local, lo, threshold=cfg['thr'], final=False)
    snap = ddist.global_snapshot(res)
    ok = True
    if rank == 0:
        ref = Graph(cfg['k'], device=lr)
        ref.build(None, threshold=cfg['thr'], final=False, packed=(bases, off))
        ref.walk(stop_at_branch=True)
        exp = ddist.single_snapshot(ref)
        st = ref.stats()
        print('single GPU: nodes %d multi %d branch %d pulled %d starts %d contigs %d; %d ranks ran mode %s' %
              (st.n_nodes, st.n_multi, st.n_branch, st.n_pulled, st.n_starts, st.n_contigs, world, res.mode))
        for f_ in exp:
            same = np.array_equal(np.asarray(snap[f_]), np.asarray(exp[f_]))
            ok &= same
            print('  %-10s %s' % (f_, 'OK' if same else 'MISMATCH'))
        if res.mode == 'replicated':
            f = torch.from_numpy(res.pull_flags.astype(np.int64)).cuda().reshape(-1, 1)
        else:
            f = torch.zeros((hi - lo, 1), dtype=torch.int64, device='cuda')
    else:
        f = torch.from_numpy(res.pull_flags.astype(np.int64)).cuda().reshape(-1, 1) if res.mode == 'replicated' \
            else torch.zeros((hi - lo, 1), dtype=torch.int64, device='cuda')
    allf, _ = ddist._all_gather_var(f, None)
    if rank == 0:
        same = np.array_equal(allf.reshape(-1).cpu().numpy().astype(np.uint8), ref.pull_read_flags())
        ok &= same
        print('  %-10s %s' % ('pullflag', 'OK' if same else 'MISMATCH'))
        print('DIST PARITY (%d ranks, %s, G=%d, %s): %s' % (world, name, G, res.mode, 'OK' if ok else 'FAILED'))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
