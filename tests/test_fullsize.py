"""Full-size parity of every BASELINE.json config on the GPU (``-m gpu``): the product through the
C ABI against the C oracle (oracle/dbg_oracle.c, pinned to the literal reference by
tests/test_oracle_golden.py) on the WHOLE workload, compared row by row at descriptor level
(oracle/compare.py) -- strings cannot be materialised at these sizes (configs[1]
would be 9e10 characters).  Plus the size-independent properties the domain offers: every
window is counted exactly once (sum of edge counts == n_windows), node count == distinct
k-mers of a brute-force numpy count on a sample, idempotence of a rebuild on the same handle.

The same comparison runs on CPU (emulated library) on down-scaled copies of each shape, so the
test logic itself is exercised by ``-m "not gpu"``.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import bench
from oracle import compare as fullsize
from oracle import c_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_config(name, G=None, final=False):
    from dbg_b200 import Graph
    cfg = bench.CONFIGS[name]
    bases, off = bench.make_reads(cfg, G=G)
    k, thr = cfg['k'], cfg['thr']
    g = Graph(k)
    try:
        g.build(None, threshold=thr, final=final, packed=(bases, off))
        g.walk(stop_at_branch=not final)
        res = c_oracle.run(None, k, thr, final, strings=False, packed=(bases.tobytes(), off))
        bad = fullsize.compare_with_oracle(g, res, final)
        assert not bad, '%s%s: rows differ from the oracle: %s' % (name, ' final' if final else '', bad)
        st = g.stats()
        nd = g.nodes()
        # every window is counted exactly once (debruijn.py:132/136 appends once per iteration)
        assert int(nd['cnt'].astype(np.uint64).sum()) == st.n_windows == bench.n_windows_of(off, k)
        # idempotence: a second build on the same handle (buffers reused) gives the same table and contigs
        d1 = g.contig_descriptors().copy()
        tags1 = nd['tag'].copy()
        g.build(None, threshold=thr, final=final, packed=(bases, off))
        g.walk(stop_at_branch=not final)
        assert np.array_equal(g.nodes()['tag'], tags1)
        d2 = g.contig_descriptors()
        assert np.array_equal(d1['n_nodes'], d2['n_nodes']) and np.array_equal(d1['kind'], d2['kind'])
        return st
    finally:
        g.close()


# ---- CPU: the same comparison on down-scaled copies of each shape (emulated library) ----
@pytest.mark.parametrize('name,G,final', [('c2', 6000, False), ('c3', 6000, False), ('c3', 4000, True),
                                          ('c5s', 5000, False)])
def test_config_shapes_small(emu_library, monkeypatch, name, G, final):
    from dbg_b200 import _native
    monkeypatch.setattr(_native, '_lib', emu_library)
    _run_config(name, G=G, final=final)


def test_batched_shape_small(emu_library, monkeypatch):
    from dbg_b200 import _native
    monkeypatch.setattr(_native, '_lib', emu_library)
    _run_batched(T=12, sample=12)


def _run_batched(T, sample):
    from dbg_b200 import Graph
    cfg = dict(bench.CONFIGS['c4'])
    cfg['T'] = T
    k, thr = cfg['k'], cfg['thr']
    bases, off, read_batch, batch_off = bench.make_reads_batched(cfg)
    g = Graph(k, key_words=2 if 2 * k + T.bit_length() > 64 else 0)
    try:
        g.build_batched_packed(bases, off, read_batch, T, threshold=thr, final=False)
        desc = g.walk(stop_at_branch=True)
        rng = np.random.default_rng(5)
        picks = sorted(set([0, T - 1] + rng.integers(0, T, sample).tolist()))
        bad = fullsize.compare_batches_with_oracle(g, desc, bases, off, batch_off, k, thr, picks, c_oracle)
        assert not bad, 'batches differ from the per-template oracle: %s' % bad[:10]
        st = g.stats()
        assert int(g.nodes()['cnt'].astype(np.uint64).sum()) == st.n_windows == bench.n_windows_of(off, k)
    finally:
        g.close()


# ---- GPU: the BASELINE.json configs at full size ----
@pytest.mark.gpu
def test_c1_full(backend_gpu):
    """configs[0]: 10 kb template, 100 bp x 20x, k=31 (also covered string-exact by KAT-2)."""
    _run_config('c1')
    _run_config('c1', final=True)


@pytest.mark.gpu
def test_c2_full(backend_gpu):
    """configs[1]: 4.6 Mb genome, 150 bp x 30x, k=31: 99 M windows, 4.6 M nodes."""
    st = _run_config('c2')
    assert st.n_nodes > 4_000_000 and st.n_branch == 0


@pytest.mark.gpu
def test_c3_full(backend_gpu):
    """configs[2]: the same genome with 1 % substitutions, reference semantics (threshold 2):
    33 M nodes, 14 k branch nodes, tip pruning and pull-out reads at scale."""
    st = _run_config('c3')
    assert st.n_branch > 1000 and st.n_pulled > 0 and st.n_pull_reads > 0


@pytest.mark.gpu
def test_c3_final_subsample(backend_gpu):
    """configs[2] with final=True on the 50 kb sub-sample SURVEY.md D2 names (path enumeration
    through surviving branch nodes is exponential in the reference, debruijn.py:314)."""
    _run_config('c3', G=50_000, final=True)


@pytest.mark.gpu
def test_c4_batched_full(backend_gpu):
    """configs[3]: 10 k templates of 1-5 kb in ONE batched pass; 300 sampled templates (plus the
    first and last) against the oracle run per template."""
    _run_batched(T=bench.CONFIGS['c4']['T'], sample=300)


@pytest.mark.gpu
def test_c5s_full(backend_gpu):
    """configs[4] shape (k=63, 128-bit keys) at 20 Mb on one GPU."""
    st = _run_config('c5s')
    assert st.key_words == 2 and st.n_nodes > 15_000_000


@pytest.mark.gpu
def test_c5_full(backend_gpu):
    """configs[4] at its full size on ONE GPU: 100 Mb genome, 18 M reads, 1.6 G windows, 1e8 nodes
    of 126 bits (the C oracle needs ~2.5 min and ~25 GB for it)."""
    if os.environ.get('DBG_B200_SKIP_C5'):
        pytest.skip('DBG_B200_SKIP_C5 set')
    st = _run_config('c5')
    assert st.n_nodes > 90_000_000


@pytest.mark.gpu
def test_nccl_two_ranks_equal_one_gpu(backend_gpu):
    """The multi-GPU path over NCCL: 2 ranks (torchrun) must give exactly the 1-GPU result --
    node table, branch / pulled lists, pull-out flags, contig descriptors (tools/check_dist.py)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs on the box')
    port = 29700 + os.getpid() % 200
    for name, G in (('c3', 400_000), ('c2', 600_000), ('c5s', 300_000)):
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
               '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tools', 'check_dist.py'),
               name, str(G)]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and 'DIST PARITY' in out.stdout and 'FAILED' not in out.stdout, \
            out.stdout[-2000:] + out.stderr[-2000:]
