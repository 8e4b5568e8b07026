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


@pytest.mark.parametrize('world,mode', [(2, 'chunks'), (3, 'chunks')])
def test_distributed_equals_oracle(world, mode):
    """The ranks exchange the distinct CHUNKS of their reads (owner = f(minimizer), only the owner expands
    them into nodes); the shards are then gathered and every rank ends with the complete graph."""
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
    port = 29500 + (os.getpid() % 1000) + world
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


def _sharded_worker(rank, world, port, cases, out_q):
    import numpy as np
    import torch.distributed as dist
    ROOT = os.path.dirname(HERE)
    for p in (ROOT, os.path.join(ROOT, 'py-debruijn_b200'), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    from emu import build_emu
    from dbg_b200 import Graph, _native, dist as ddist
    from dbg_b200.graph import pack_reads
    _native._lib = _native.Library(build_emu.build())
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    results = []
    handles = {}                                   # one handle per k, reused: later builds are sized from the remembered hints
    for reads, k, thr, final in cases:
        b = ddist.shard_bounds(len(reads), world)
        g = handles.get(k)
        if g is None:
            g = handles[k] = Graph(k)
        res = ddist.build_sharded(g, pack_reads(reads[b[rank]:b[rank + 1]]), b[rank], threshold=thr, final=final)
        snap = ddist.global_snapshot(res)
        if rank == 0:
            ref = Graph(k)
            ref.build(reads, threshold=thr, final=final)
            ref.walk(stop_at_branch=not final)
            exp = ddist.single_snapshot(ref)
            diff = [f for f in exp if not np.array_equal(np.asarray(snap[f]), np.asarray(exp[f]))]
            results.append((res.mode, diff, int(len(snap['c_n'])), ddist.snapshot_digest(snap) == ddist.snapshot_digest(exp)))
            ref.close()
    for g in handles.values():
        g.close()
    out_q.put((rank, results))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_walk_equals_single(world):
    """Owner-computes path (dist.build_sharded): every rank keeps only its shard of the node table, the
    contig walk runs as a two-level list ranking across ranks; node table and contig descriptors (by node
    key) must equal the single-process build.  Graphs with branch nodes take the replicated fallback."""
    import random
    r = random.Random(900 + world)
    cases = []
    _, clean = util.synth_reads(41, 4000, 70, 10, err=0.0)
    cases.append((clean, 17, 2, False))                              # error-free: no branch nodes, sharded walk
    cases.append((clean, 17, 2, False))                              # the same again on the same handle: sized from its hints
    cases.append((clean[:len(clean) // 2] + [x[::-1] for x in clean], 17, 2, False))   # more distinct k-mers than remembered
    cases.append((clean, 35, 2, False))                              # 128-bit keys
    cases.append((clean, 17, 2, True))
    cases.append((['ACGTACGTACGTACGT'], 4, 2, False))                # a pure cycle: the start emits nothing
    cases.append((['ACGTACGTACGTACGTT', 'GGACGTAC'], 4, 2, False))   # a start that leads into a cycle
    cases.append((['ACGTTGCA'], 3, 2, False))                        # fewer reads than ranks
    cases.append(([], 5, 2, False))
    _, noisy = util.synth_reads(43, 3000, 60, 12, err=0.02)
    cases.append((noisy, 17, 2, False))                              # branch nodes: replicated fallback
    while len(cases) < 30:
        alphabet = r.choice(['ACGT', 'ACGT', 'AC'])
        G = r.choice([30, 80, 200, 600])
        genome = ''.join(r.choice(alphabet) for _ in range(G))
        k = r.choice([2, 3, 5, 8, 11, 21, 33])
        L = min(G, r.choice([k + 1, k + 3, 25, 60]))
        reads = [genome[p:p + L] for p in (r.randrange(0, G - L + 1) for _ in range(r.choice([1, 3, 10, 40])))]
        cases.append((reads, k, r.choice([1, 2, 3]), r.random() < 0.3))
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29800 + (os.getpid() % 1000) + world
    procs = [ctx.Process(target=_sharded_worker, args=(rk, world, port, cases, q)) for rk in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rk, res = q.get(timeout=900)
        got[rk] = res
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    modes = set()
    for i, (mode, diff, n_contigs, same_digest) in enumerate(got[0]):
        assert not diff and same_digest, 'world %d case %d (%s): %s differ' % (world, i, mode, diff)
        modes.add(mode)
    assert modes == {'sharded', 'replicated'}
    assert got[0][0][0] == 'sharded' and got[0][0][2] > 10
