"""torchrun --nproc-per-node N tools/check_dist.py [config] : the N-GPU distributed build must give
exactly the single-GPU result (node table, branch / pulled lists, pull-out flags, contig descriptors)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'py-debruijn_b200')]
import bench                                   # noqa: E402
from dbg_b200 import Graph, dist as ddist      # noqa: E402


def snapshot(g):
    nd = g.nodes()
    slot_to_idx = np.full(int(g.stats().table_slots), -1, np.int64)
    slot_to_idx[nd['id']] = np.arange(len(nd['id']))
    d = g.contig_descriptors()
    return {
        'key_lo': nd['key_lo'], 'key_hi': nd['key_hi'], 'tag': nd['tag'], 'cnt': nd['cnt'], 'flags': nd['flags'],
        'kept': nd['kept'], 'branch': slot_to_idx[g.branch_ids()], 'pulled': slot_to_idx[g.pulled_ids()],
        'c_start': slot_to_idx[d['start']], 'c_n': d['n_nodes'], 'c_term': slot_to_idx[d['terminal']], 'c_kind': d['kind'],
    }


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
    flags = ddist.build_distributed(g, local, lo, threshold=cfg['thr'], final=False)
    g.walk(stop_at_branch=True)
    mine = snapshot(g)
    f = torch.from_numpy(flags.astype(np.int64)).cuda().reshape(-1, 1)
    allf, _ = ddist._all_gather_var(f, None)
    ok = True
    if rank == 0:
        ref = Graph(cfg['k'], device=lr)
        ref.build(None, threshold=cfg['thr'], final=False, packed=(bases, off))
        ref.walk(stop_at_branch=True)
        exp = snapshot(ref)
        st = ref.stats()
        print('single GPU: nodes %d multi %d branch %d pulled %d starts %d contigs %d pull_reads %d' %
              (st.n_nodes, st.n_multi, st.n_branch, st.n_pulled, st.n_starts, st.n_contigs, st.n_pull_reads))
        for f_ in exp:
            same = np.array_equal(mine[f_], exp[f_])
            ok &= same
            print('  %-8s %s' % (f_, 'OK' if same else 'MISMATCH'))
        same = np.array_equal(allf.reshape(-1).cpu().numpy().astype(np.uint8), ref.pull_read_flags())
        ok &= same
        print('  %-8s %s' % ('pullflag', 'OK' if same else 'MISMATCH'))
        print('DIST PARITY (%d ranks, %s, G=%d): %s' % (world, name, G, 'OK' if ok else 'FAILED'))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
