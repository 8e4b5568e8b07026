"""Lock-free paths under REAL concurrency (-m gpu; the CPU emulator runs kernels serially and cannot show a race):
hundreds of thousands of reads over a genome of a few hundred bases, so that every warp of k_chunk_insert /
k_expand hammers the same few table slots -- the 128-bit CAS on a chunk key's first half followed by the plain
store of its second half (NOT_READY retries, k > 31), the retry rings, atomicMin / atomicAdd on one slot from
thousands of lanes -- once with tables sized as usual and once with tables that start far too small
(growth / overflow / rebuild).  Everything is compared with the C oracle row by row; a lost update shows up as a
wrong count, a torn key as a missing or extra node."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import compare as fullsize


def _hot_reads(seed, G, L, n, err):
    rng = np.random.default_rng(seed)
    genome = rng.integers(0, 4, G, dtype=np.uint8)
    starts = rng.integers(0, G - L + 1, n)
    idx = starts[:, None] + np.arange(L)[None, :]
    reads = genome[idx]
    if err:
        flip = rng.random(reads.shape) < err
        reads = np.where(flip, (reads + rng.integers(1, 4, reads.shape, dtype=np.uint8)) & 3, reads).astype(np.uint8)
    bases = np.frombuffer(b'ACGT', dtype=np.uint8)[reads.reshape(-1)]
    off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
    return np.ascontiguousarray(bases), off


@pytest.mark.gpu
@pytest.mark.parametrize('k', [21, 31, 47, 63])
@pytest.mark.parametrize('small_tables', [False, True])
def test_hot_slots(backend_gpu, monkeypatch, k, small_tables):
    from dbg_b200 import Graph
    for seed, G, L, n, err, slots in [(1, 300, 100, 300_000, 0.0, 256), (2, 500, 150, 200_000, 0.002, 1 << 17)]:
        if small_tables:                                           # a few doublings away from what the data need
            monkeypatch.setenv('DBG_B200_CHUNK_SLOTS', '64')
            monkeypatch.setenv('DBG_B200_NODE_SLOTS', str(slots))
        bases, off = _hot_reads(seed, G, L, n, err)
        g = Graph(k)
        try:
            for rep in range(2):                                   # the second build takes the remembered-sizes path
                g.build(None, threshold=2, final=False, packed=(bases, off))
                g.walk(stop_at_branch=True)
                res = c_oracle.run(None, k, 2, False, strings=False, packed=(bases.tobytes(), off))
                bad = fullsize.compare_with_oracle(g, res, False)
                assert not bad, 'k=%d small=%s seed=%d rep=%d: rows differ from the oracle: %s' % (k, small_tables, seed, rep, bad)
                assert int(g.nodes()['cnt'].astype(np.uint64).sum()) == g.stats().n_windows
        finally:
            g.close()


def test_crowded_regions_fall_back(backend, monkeypatch):
    """Reads piled up on a 500-base genome with errors: tens of thousands of k-mer variants behind a few hundred
    minimizers.  The node table laid out by minimizer cannot hold them however large it gets (more slots = more
    regions nobody maps to); the build must notice and fall back to the plain hash layout -- for this build and the
    next one on the handle -- with the reference's result."""
    from dbg_b200 import Graph
    n = 60_000
    if backend == 'emu':                                   # the emulator is slow: small regions crowd with fewer reads
        monkeypatch.setenv('DBG_B200_REGION_SLOTS', '64')
        n = 4_000
    bases, off = _hot_reads(2, 500, 150, n, 0.002)
    k = 47
    g = Graph(k)
    try:
        for rep in range(2):
            g.build(None, threshold=2, final=False, packed=(bases, off))
            g.walk(stop_at_branch=True)
            res = c_oracle.run(None, k, 2, False, strings=False, packed=(bases.tobytes(), off))
            assert not fullsize.compare_with_oracle(g, res, False)
            assert g.stats().n_regions == 1
    finally:
        g.close()
