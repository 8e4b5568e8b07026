"""Multi-GPU path on CPU: world_size 2 and 3 over gloo, each rank running the emulated library.
The distributed construct_graph must return exactly what the single-process one (and therefore
the oracle) returns, on every rank."""
import os
import sys

import pytest
import torch.multiprocessing as mp

import util

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, cases, out_q, mode='chunks'):
    import contextlib
    import io
    import torch.distributed as dist
    ROOT = os.path.dirname(HERE)
    for p in (ROOT, os.path.join(ROOT, 'py-debruijn_b200'), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    from emu import build_emu
    from dbg_b200 import _native, dist as ddist
    import debruijn as db
    _native._lib = _native.Library(build_emu.build())
    os.environ['DBG_B200_DIST'] = mode
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    results = []
    for reads, k, thr, final in cases:
        with contextlib.redirect_stdout(io.StringIO()) as buf:
            g, pull, branch, pulled, ect = ddist.construct_graph(list(reads), k, threshold=thr, final=final)
            contigs = db.output_contigs(g, branch, pulled)
        V, E = g
        results.append({
            'nodes': [[v, n.indegree, n.outdegree] for v, n in V.items()],
            'edges': [[v, list(s)] for v, s in E.items()],
            'pull_out_read': list(pull), 'branch': list(branch), 'pulled': list(pulled),
            'ect': [[e, c] for e, c in ect.items()], 'contigs': list(contigs),
            'stdout': buf.getvalue().splitlines(),
        })
    out_q.put((rank, results))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world,mode', [(2, 'chunks'), (3, 'chunks'), (2, 'records')])
def test_distributed_equals_oracle(world, mode):
    """mode = what the ranks exchange: distinct CHUNKS of their reads (default; owner = f(minimizer), only
    the owner expands them into nodes) or NODE records of a full local node table (the older path)."""
    import random
    from oracle import dbg_oracle
    r = random.Random(500 + world)
    cases = []
    _, reads = util.synth_reads(31, 3000, 60, 12, err=0.02)
    cases.append((reads, 17, 2, False))
    cases.append((reads, 35, 2, True))
    cases.append((['ACGTACGTAC'], 4, 2, False))                     # fewer reads than ranks
    cases.append(([], 5, 2, False))
    while len(cases) < 24:
        rd, k, thr, final = util.random_dna_case(r)
        try:
            dbg_oracle.snapshot(rd, k, thr, final, max_paths=100)
        except RuntimeError:
            continue
        cases.append((rd, k, thr, final))
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + world + (7 if mode == 'records' else 0)
    procs = [ctx.Process(target=_worker, args=(rk, world, port, cases, q, mode)) for rk in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rk, res = q.get(timeout=600)
        got[rk] = res
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for i, (rd, k, thr, final) in enumerate(cases):
        exp = dbg_oracle.snapshot(rd, k, thr, final)
        for rk in range(world):
            util.assert_same(got[rk][i], exp, 'world %d rank %d case %d %r' % (world, rk, i, (k, thr, final)))
