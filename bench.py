#!/usr/bin/env python
"""Benchmark of the assembly hot path (BASELINE.json): k-mers/s of construct_graph + contigs/s
of output_contigs on synthetic reads, with the HBM roofline of the dominant kernel and the
reference's CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2] [--impl dbg|reference]

One "step" = one pass of the hot path over the whole read set: dbg_build_device (pack, node
table, degrees/prune/branch flags, tip pruning, pull-out reads) + dbg_walk (list ranking, contig
descriptors).  `value` = windows (k-mer instances) of all ranks / step time with
the reads resident in HBM; `e2e` = the same through dbg_build with HOST (pinned) buffers, H2D and
the descriptor D2H inside the timed region.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'py-debruijn_b200')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# BASELINE.json configs.  seeds = 1000 + config number.
CONFIGS = {
    'c1': dict(G=10_000, L=100, cov=20, err=0.0, k=31, thr=2, seed=1001,
               name='configs[0]: 10 kb template, 100 bp error-free reads 20x, k=31'),
    'c2': dict(G=4_600_000, L=150, cov=30, err=0.0, k=31, thr=2, seed=1002,
               name='configs[1]: synthetic 4.6 Mb genome, 150 bp reads 30x, k=31'),
    'c3': dict(G=4_600_000, L=150, cov=30, err=0.01, k=31, thr=2, seed=1002,
               name='configs[2]: 4.6 Mb genome, 1% substitution errors, k=31 (reference semantics, threshold 2)'),
    'c3f': dict(G=50_000, L=150, cov=30, err=0.01, k=31, thr=2, seed=1002,
                name='configs[2] on a 50 kb sub-sample (SURVEY 8 D2: the shape final=True is run on; use with --final)'),
    'c4': dict(T=10_000, Gmin=1000, Gmax=5000, L=100, cov=20, err=0.0, k=25, thr=2, seed=1004, G=30_000_000,
               name='configs[3]: 10k templates of 1-5 kb, 100 bp reads 20x, k=25, every template its own graph '
                    '(one batched pass)'),
    'c5s': dict(G=20_000_000, L=150, cov=30, err=0.0, k=63, thr=2, seed=1005,
                name='configs[4] shape at 20 Mb: 150 bp reads 30x, k=63 (128-bit k-mers)'),
    'c5': dict(G=100_000_000, L=150, cov=30, err=0.0, k=63, thr=2, seed=1005,
               name='configs[4] shape on ONE GPU: 100 Mb genome, 150 bp reads 30x, k=63 (128-bit k-mers)'),
}
ASCII = np.frombuffer(b'ACGT', dtype=np.uint8)


def make_reads(cfg, seed_offset=0, G=None):
    """Synthetic forward-strand reads, deduped first-seen-order as II_assembleFromReads.py:52 does."""
    G = G or cfg['G']
    L, cov, err = cfg['L'], cfg['cov'], cfg['err']
    rng = np.random.default_rng(cfg['seed'] + seed_offset)
    genome = rng.integers(0, 4, G, dtype=np.uint8)
    n = G * cov // L
    starts = rng.integers(0, G - L + 1, n)
    mat = genome[starts[:, None] + np.arange(L)[None, :]]
    if err:
        flip = rng.random(mat.shape) < err
        mat = np.where(flip, (mat + rng.integers(1, 4, mat.shape, dtype=np.uint8)) & 3, mat).astype(np.uint8)
    w = rng.integers(1, 2**63, L, dtype=np.uint64) | np.uint64(1)
    h = (mat.astype(np.uint64) * w[None, :]).sum(axis=1, dtype=np.uint64)
    _, first = np.unique(h, return_index=True)
    mat = mat[np.sort(first)]
    bases = np.ascontiguousarray(ASCII[mat]).reshape(-1)
    off = (np.arange(mat.shape[0] + 1, dtype=np.uint64) * np.uint64(L))
    return bases, off


def make_reads_batched(cfg, T=None, seed_offset=0):
    """configs[3]: T templates with lengths uniform in [Gmin, Gmax], each its own read set (deduped
    per template as II_assembleFromReads.py:52 would for that template).  Returns the concatenated
    reads, offsets, the batch id of every read and the read range of every batch."""
    T = T or cfg['T']
    L, cov = cfg['L'], cfg['cov']
    rng = np.random.default_rng(cfg['seed'] + seed_offset)
    lens = rng.integers(cfg['Gmin'], cfg['Gmax'] + 1, T)
    g_off = np.zeros(T + 1, np.int64)
    np.cumsum(lens, out=g_off[1:])
    genome = rng.integers(0, 4, int(g_off[-1]), dtype=np.uint8)
    n_per = lens * cov // L
    batch = np.repeat(np.arange(T, dtype=np.int64), n_per)
    starts = (rng.random(len(batch)) * (lens[batch] - L + 1)).astype(np.int64)
    key = batch * (cfg['Gmax'] + 1) + starts                       # equal start in one template <=> equal read
    _, first = np.unique(key, return_index=True)
    keep = np.sort(first)                                          # first-seen order; batch-major because key is
    batch, starts = batch[keep], starts[keep]
    order = np.argsort(batch, kind='stable')
    batch, starts = batch[order], starts[order]
    mat = genome[(g_off[batch] + starts)[:, None] + np.arange(L)[None, :]]
    bases = np.ascontiguousarray(ASCII[mat]).reshape(-1)
    off = (np.arange(mat.shape[0] + 1, dtype=np.uint64) * np.uint64(L))
    batch_off = np.searchsorted(batch, np.arange(T + 1))
    return bases, off, batch.astype(np.uint32), batch_off


def make_reads_shard(cfg, rank, world, scale=None):
    """The distributed workload: ONE read set over a genome `scale` times the config's size (default
    scale = world: weak scaling, per-GPU work fixed; scale = 1: the fixed problem split over the ranks),
    deduped by start position, rank r holding the contiguous slice shard_bounds()[r] of the global read
    list.  Every rank derives the same global list."""
    G = cfg['G'] * (world if scale is None else scale)
    L, cov = cfg['L'], cfg['cov']
    rng = np.random.default_rng(cfg['seed'] + 77 * world)
    genome = rng.integers(0, 4, G, dtype=np.uint8)
    n = G * cov // L
    starts = rng.integers(0, G - L + 1, n)
    _, first = np.unique(starts, return_index=True)               # equal start <=> equal read (random genome)
    starts = starts[np.sort(first)]
    n_total = len(starts)
    lo, hi = n_total * rank // world, n_total * (rank + 1) // world
    mine = starts[lo:hi]
    del starts, first
    mat = np.empty((len(mine), L), np.uint8)
    step = 1 << 20                                                # bounded temporaries (configs[4]: 18 M reads)
    for a in range(0, len(mine), step):
        mat[a:a + step] = genome[mine[a:a + step, None] + np.arange(L)[None, :]]
    if cfg['err']:
        r2 = np.random.default_rng(cfg['seed'] + 1000 + rank)
        flip = r2.random(mat.shape) < cfg['err']
        mat = np.where(flip, (mat + r2.integers(1, 4, mat.shape, dtype=np.uint8)) & 3, mat).astype(np.uint8)
    bases = ASCII[mat].reshape(-1)
    off = (np.arange(mat.shape[0] + 1, dtype=np.uint64) * np.uint64(L))
    return bases, off, lo, n_total


def make_reads_dedup_by_start(cfg, world=1, scale=None):
    """The whole read list the `world` shards of make_reads_shard make up, in global read order."""
    parts = [make_reads_shard(cfg, r, world, scale)[0] for r in range(world)]
    bases = np.concatenate(parts) if len(parts) > 1 else parts[0]
    off = (np.arange(len(bases) // cfg['L'] + 1, dtype=np.uint64) * np.uint64(cfg['L']))
    return bases, off


def n_windows_of(off, k):
    lens = np.diff(off.astype(np.int64))
    return int(np.maximum(lens - k, 0).sum())


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for line in self.lines:
            p = [x.strip() for x in line.split(',')]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1])); power.append(float(p[2]))
            except ValueError:
                continue
            for nme, v in zip(names, p[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nme)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(power) if power else None, 'samples': len(sm), 'reasons': sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU legs (the oracle is the checker / the reported baseline, never the product)
# ---------------------------------------------------------------------------------------------
def cpu_python_port(cfg, G_sample):
    """The reference's pure-Python path, restated (oracle/dbg_oracle.py), 1 core, bounded sample."""
    from oracle import dbg_oracle
    bases, off = make_reads(cfg, G=G_sample)
    L = cfg['L']
    raw = bases.tobytes().decode('ascii')
    reads = [raw[i * L:(i + 1) * L] for i in range(len(off) - 1)]
    nwin = n_windows_of(off, cfg['k'])
    t0 = time.perf_counter()
    g, pull, branch, pulled, ect = dbg_oracle.construct_graph(reads, cfg['k'], cfg['thr'], False, verbose=False)
    t1 = time.perf_counter()
    contigs = dbg_oracle.output_contigs(g, branch, pulled, verbose=False)
    t2 = time.perf_counter()
    return {'windows': nwin, 'build_s': t1 - t0, 'walk_s': t2 - t1, 'contigs': len(contigs),
            'sample': '%d bp sub-genome of the same shape (%d reads, %d windows)' % (G_sample, len(reads), nwin)}


def cpu_literal(cfg, G_sample):
    """The LITERAL reference (debruijn.py, unmodified, recursion limit raised), 1 core, bounded sample of the C1
    shape -- only where its file is present (the build container; it does not travel to the GPU box)."""
    from oracle import literal
    if not literal.available():
        return None
    c1 = CONFIGS['c1']
    bases, off = make_reads(c1, G=G_sample)
    L = c1['L']
    raw = bases.tobytes().decode('ascii')
    reads = [raw[i * L:(i + 1) * L] for i in range(len(off) - 1)]
    nwin = n_windows_of(off, c1['k'])
    build_s, walk_s, n_contigs = literal.timed(reads, c1['k'], c1['thr'], False)
    return {'value': nwin / (build_s + walk_s), 'unit': 'k-mers/s', 'cores': 1, 'kind': 'literal',
            'sample': 'the unmodified debruijn.py on a %d bp sub-genome of the configs[0] shape (%d reads, %d windows), '
                      'construct_graph %.3f s + output_contigs %.3f s' % (G_sample, len(reads), nwin, build_s, walk_s),
            'contigs': n_contigs}


def cpu_c_port(cfg, bases, off, final=False):
    """The C restatement (oracle/dbg_oracle.c), 1 core, on the given reads."""
    from oracle import c_oracle
    t0 = time.perf_counter()
    res = c_oracle.run(None, cfg['k'], cfg['thr'], final, stages=3, strings=False, packed=(bases.tobytes(), off))
    t1 = time.perf_counter()
    return {'windows': int(res.n_windows), 'total_s': t1 - t0, 'build_s': res.t_build, 'graph_s': res.t_graph,
            'walk_s': res.t_walk, 'contigs': int(len(res.contigs)), 'nodes': int(res.n_nodes)}


def static_config(args, cfg, world):
    """`config` of the JSON line: a function of (config key, N) alone, so both arms print the same dict."""
    wl = cfg['name'] if world == 1 else \
        ('%s; genome scaled x%d (%d bp) so per-GPU work is fixed, ONE graph built by %d ranks' %
         (cfg['name'], world, cfg['G'] * world, world))
    batched = 'T' in cfg
    return {'workload': wl, 'config_key': args.config, 'k': cfg['k'], 'threshold': cfg['thr'],
            'genome_bp': cfg['G'] * world, 'read_len': cfg['L'], 'coverage': cfg['cov'], 'error_rate': cfg['err'],
            'final': bool(getattr(args, 'final', False)),
            'min_count': getattr(args, 'min_count', 1),
            'min_count_note': None if getattr(args, 'min_count', 1) <= 1 else
            'EXTENSION without a reference oracle (SURVEY.md section 9 delta 5): pinned by the two restatements only',
            'parallelism': 'single GPU' if world == 1 else
            ('templates sharded x%d by batch id, no collective (independent graphs)' % world) if batched else
            ('reads sharded x%d; distinct chunks hash-partitioned by minimizer owner, ONE NCCL all-to-all; degrees / '
             'prune / walk owner-computes on the node-table shards (two-level list ranking, 16 B per shard-boundary '
             'node all-gathered)' % world),
            'l2_policy': 'inputs (ASCII reads + node table) larger than the 126 MB L2; no explicit flush'}


def run_reference(args, cfg, rank, world):
    """--impl reference: the reference's CPU path on this box's host cores.  The reference is pure
    single-threaded Python that cannot finish configs[1] (its walk is O(path^2) and
    recursive), so each step is the C restatement of it (oracle/dbg_oracle.c, strictly faster than the
    reference: O(n) walk, packed keys), 1 thread -- the reference has no parallelism to use more.  At N=1
    every step is the FULL workload; at N>1 (workload = N x the genome) every step is one GPU's share of it,
    i.e. the N=1 workload -- a bounded sample, stated in cpu_baseline.sample."""
    if rank != 0:
        return
    G_sample = cfg['G']
    if args.ref_sample_bp:
        G_sample = min(cfg['G'], args.ref_sample_bp)
    bases, off = make_reads(cfg, G=G_sample)
    nwin = n_windows_of(off, cfg['k'])
    times, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_c_port(cfg, bases, off, final=bool(getattr(args, 'final', False)))
        if i >= args.warmup:
            times.append(last['total_s'])
    t = float(np.mean(times))
    value = nwin / t
    full = G_sample == cfg['G'] and world == 1
    line = {
        'impl': 'reference', 'metric': 'kmers_per_sec_graph_build_plus_contig_walk', 'value': value, 'unit': 'k-mers/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': t * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u64' if cfg['k'] <= 31 else 'u128',
        'data': 'synthetic', 'config': static_config(args, cfg, world),
        'contigs_per_sec': last['contigs'] / max(last['walk_s'], 1e-9),
        'cpu_baseline': {'value': value, 'unit': 'k-mers/s', 'cores': 1, 'kind': 'port',
                         'sample': '%s per step: %d bp genome of the workload\'s shape, %d reads, %d windows (C '
                                   'restatement oracle/dbg_oracle.c, construct_graph + output_contigs at descriptor '
                                   'level; the pure-Python reference cannot finish this config)' %
                                   ('FULL workload' if full else 'bounded sample (one GPU\'s share of the workload)'
                                    if world > 1 else 'bounded sample', G_sample, len(off) - 1, nwin)},
        'e2e': {'value': value, 'unit': 'k-mers/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'host_cores': os.cpu_count(),
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# the product arm
# ---------------------------------------------------------------------------------------------
STAGE_FIELDS = ('ms_pack', 'ms_estimate', 'ms_table_init', 'ms_insert', 'ms_finalize', 'ms_order', 'ms_lists',
                'ms_tips', 'ms_pull_reads', 'ms_build_total', 'ms_walk_rank', 'ms_walk_emit', 'ms_walk_total',
                'ms_chunk_insert', 'ms_expand_sketch', 'ms_expand')
NVLINK_GBS = 770.0        # measured peer copy per direction per GPU on this pool (B200_PROFILING.md); nominal 900


class Runner:
    """One process = one GPU: torch for buffers / events / collectives, the C ABI for everything else."""

    def __init__(self, rank, local_rank, world):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.local_rank, self.world = rank, local_rank, world
        self.dev = torch.device('cuda', local_rank)
        self.stream = torch.cuda.current_stream(self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def timed(self, fn, steps, after=None):
        """EXACTLY `steps` calls bracketed by barrier + synchronize, CUDA events on the launching stream,
        max over ranks."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for _ in range(steps):
            fn()
            if after is not None:
                after()
        e1.record(self.stream)
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def total(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t)
        return float(t.item())

    def to_dev(self, bases, off):
        torch = self.torch
        return torch.from_numpy(bases).to(self.dev), torch.from_numpy(off.view(np.int64)).to(self.dev)


def strong_c5(R, args):
    """The question north_star asks: the FIXED 100 Mb / k=63 problem (configs[4]) hash-partitioned over N ranks,
    T_N = max over ranks of the device time of one build_sharded step, and T_1 = the same reads on one GPU
    (rank 0 alone, dbg_build_device + dbg_walk), both in this run.  efficiency = T_1 / (N * T_N)."""
    import torch
    from dbg_b200 import Graph, dist as ddist
    cfg = CONFIGS['c5']
    k, thr = cfg['k'], cfg['thr']
    steps = max(2, min(args.steps, 3))
    out = {'workload': cfg['name'], 'steps': steps}
    g = Graph(k, device=R.local_rank)
    g.set_stream(R.stream.cuda_stream)
    if R.world > 1:
        bases, off, lo, n_total = make_reads_shard(cfg, R.rank, R.world, scale=1)
        d_bases, d_off = R.to_dev(bases, off)
        ex = {}

        def step():
            return ddist.build_sharded(g, None, lo, threshold=thr, final=False, stats=ex,
                                       local_device=(d_bases.data_ptr(), d_off.data_ptr(), len(off) - 1, len(bases)))
        for _ in range(2):
            res = step()
        ms = R.timed(step, steps) / steps
        tm = {}
        ddist.build_sharded(g, None, lo, threshold=thr, final=False, timings=tm,
                            local_device=(d_bases.data_ptr(), d_off.data_ptr(), len(off) - 1, len(bases)))
        nwin = R.total(n_windows_of(off, k))
        out.update({'n_gpus': R.world, 'ms_per_step': ms, 'kmers_per_sec': nwin / (ms * 1e-3), 'mode': res.mode,
                    'nodes': res.totals['n_nodes'], 'contigs': int(R.total(len(res.contigs) if res.contigs is not None else 0)),
                    'stage_ms_rank0': {kk: round(v, 3) for kk, v in tm.items()},
                    'exchange_bytes_out_rank0': ex.get('exchange_bytes_out'), 'walk_bytes_out_rank0': ex.get('walk_bytes_out')})
        del d_bases, d_off
        torch.cuda.empty_cache()
    # T_1 on rank 0, the other ranks wait at the barrier
    if R.rank == 0:
        bases, off = make_reads_dedup_by_start(cfg, R.world, scale=1)
        d_bases, d_off = R.to_dev(bases, off)
        g.shard_config(0, 0)

        def step1():
            g.build_device(d_bases.data_ptr(), d_off.data_ptr(), len(off) - 1, len(bases), threshold=thr, final=False)
            g.walk(stop_at_branch=True)
        for _ in range(2):
            step1()
        torch.cuda.synchronize(R.dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(R.stream)
        for _ in range(steps):
            step1()
        e1.record(R.stream)
        torch.cuda.synchronize(R.dev)
        t1 = e0.elapsed_time(e1) / steps
        st = g.stats()
        out['t1_ms'] = t1
        out['t1_kmers_per_sec'] = n_windows_of(off, k) / (t1 * 1e-3)
        out['t1_nodes'], out['t1_contigs'] = int(st.n_nodes), int(st.n_contigs)
        if R.world > 1:
            out['efficiency'] = t1 / (R.world * out['ms_per_step'])
            out['parity_counts_equal'] = (out['nodes'] == out['t1_nodes'] and out['contigs'] == out['t1_contigs'])
        del d_bases, d_off
    g.close()
    torch.cuda.empty_cache()
    R.barrier()
    return out


def c4_batched(R, args):
    """configs[3] at N ranks: templates sharded by batch id, no collective."""
    from dbg_b200 import Graph
    cfg = CONFIGS['c4']
    k, thr, T = cfg['k'], cfg['thr'], cfg['T']
    bases, off, read_batch, _ = make_reads_batched(cfg, seed_offset=R.rank)
    d_bases, d_off = R.to_dev(bases, off)
    d_batch = R.torch.from_numpy(read_batch.view(np.int32)).to(R.dev)
    g = Graph(k, device=R.local_rank, key_words=2 if 2 * k + T.bit_length() > 64 else 0)
    g.set_stream(R.stream.cuda_stream)

    def step():
        g.build_batched_packed(None, None, None, T, threshold=thr, final=False,
                               device_ptrs=(d_bases.data_ptr(), d_off.data_ptr(), d_batch.data_ptr(), len(off) - 1, len(bases)))
        g.walk(stop_at_branch=True)
    for _ in range(2):
        step()
    steps = max(2, min(args.steps, 5))
    ms = R.timed(step, steps) / steps
    nwin = R.total(n_windows_of(off, k))
    out = {'workload': cfg['name'], 'n_gpus': R.world, 'templates_per_gpu': T, 'steps': steps, 'ms_per_step': ms,
           'kmers_per_sec': nwin / (ms * 1e-3), 'contigs_per_gpu': int(g.stats().n_contigs)}
    g.close()
    return out


def e2e_api(R, args):
    """What II_assembleFromReads.py:58-63 calls, with Python objects on both sides: list[str] in,
    debruijn.construct_graph + output_contigs, list[str] out (configs[0]; 100 templates of configs[3] through
    construct_graph_batch), next to the pure-Python restatement of the reference on the same call."""
    import contextlib
    import io
    import debruijn as db
    from oracle import dbg_oracle
    out = {}
    cfg = CONFIGS['c1']
    bases, off = make_reads(cfg)
    L = cfg['L']
    raw = bases.tobytes().decode('ascii')
    reads = [raw[i * L:(i + 1) * L] for i in range(len(off) - 1)]

    def call():
        with contextlib.redirect_stdout(io.StringIO()):
            g, pull, branch, pulled, ect = db.construct_graph(reads, cfg['k'], threshold=cfg['thr'], final=False)
            return db.output_contigs(g, branch, pulled)
    for _ in range(2):
        contigs = call()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        contigs = call()
    ms = (time.perf_counter() - t0) / n * 1e3
    # the same call with a NEW handle every time (the module otherwise reuses the handle of the previous, finished call)
    os.environ['DBG_B200_HANDLE_POOL'] = '0'
    try:
        call()
        t0 = time.perf_counter()
        for _ in range(n):
            contigs = call()
        ms_fresh = (time.perf_counter() - t0) / n * 1e3
    finally:
        del os.environ['DBG_B200_HANDLE_POOL']
    t0 = time.perf_counter()
    g, pull, branch, pulled, ect = dbg_oracle.construct_graph(reads, cfg['k'], cfg['thr'], False, verbose=False)
    exp = dbg_oracle.output_contigs(g, branch, pulled, verbose=False)
    ms_port = (time.perf_counter() - t0) * 1e3
    out['c1'] = {'workload': cfg['name'], 'reads': len(reads), 'contigs': len(contigs), 'chars': sum(map(len, contigs)),
                 'ms_per_call': ms, 'ms_per_call_new_handle': ms_fresh,
                 'handle': 'ms_per_call: repeated calls reuse the handle of the previous finished call (pool in dbg_b200/graph.py); '
                           'ms_per_call_new_handle: DBG_B200_HANDLE_POOL=0',
                 'windows': n_windows_of(off, cfg['k']),
                 'kmers_per_sec': n_windows_of(off, cfg['k']) / (ms * 1e-3),
                 'python_port_ms': ms_port, 'strings_equal_port': contigs == exp,
                 'literal_reference_note': 'the literal debruijn.py needs 48 s for this call (BASELINE.md section 2: '
                                           'construct_graph 0.245 s + output_contigs 47.8 s, recursion limit raised)'}
    c4 = dict(CONFIGS['c4'])
    c4['T'] = 100
    b4, o4, rb, boff = make_reads_batched(c4)
    raw = b4.tobytes().decode('ascii')
    L = c4['L']
    batches = [[raw[i * L:(i + 1) * L] for i in range(int(boff[b]), int(boff[b + 1]))] for b in range(100)]

    def call4():
        with contextlib.redirect_stdout(io.StringIO()):
            res = db.construct_graph_batch(batches, c4['k'], threshold=c4['thr'], final=False)
            return [db.output_contigs(g, branch, pulled) for (g, pull, branch, pulled, ect) in res]
    call4()
    t0 = time.perf_counter()
    got = call4()
    ms4 = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    same = True
    for b in range(100):
        g, pull, branch, pulled, ect = dbg_oracle.construct_graph(batches[b], c4['k'], c4['thr'], False, verbose=False)
        same &= dbg_oracle.output_contigs(g, branch, pulled, verbose=False) == got[b]
    ms4_port = (time.perf_counter() - t0) * 1e3
    out['c4_100_templates'] = {'reads': len(o4) - 1, 'contigs': sum(map(len, got)), 'chars': sum(len(c) for cs in got for c in cs),
                               'ms_per_call': ms4, 'python_port_ms': ms4_port, 'strings_equal_port': bool(same)}
    return out


def n3_concat(R, args):
    """III_concatenateOutpus.py:53-146: the whole k = 20..6 loop + substring dedup on the GPU
    (dbg_concat_*: one persistent kernel per k, Python's Mersenne Twister on the device) next to the restatement of
    the script on the same contigs, score table and seed; outputs and printed lines compared."""
    import random
    from dbg_b200.concat import concatenate_outputs
    r = random.Random(5)
    aa = 'ACDEFGHIKLMNPQRSTVWY'
    protein = ''.join(r.choice(aa) for _ in range(2400))
    contigs, table = [], {}
    for _ in range(300):
        ln = r.choice([4, 7, 12, 18, 22, 25, 30, 40, 55])
        p = r.randrange(0, len(protein) - ln + 1)
        s = list(protein[p:p + ln])
        if ln >= 8 and r.random() < 0.2:
            j = r.randrange(0, 5) if r.random() < 0.5 else ln - 2 - r.randrange(0, 5)
            s[j], s[j + 1] = s[j + 1], s[j]
        contigs.append(''.join(s))
    for _ in range(1800):
        ln = r.randrange(5, 11)
        p = r.randrange(0, len(protein) - ln + 1)
        table[protein[p:p + ln]] = table.get(protein[p:p + ln], 0) + r.choice([0.85, 0.9, 0.99])
    a, b = [], []
    concatenate_outputs(contigs[:40], table, rng=random.Random(0), out=b.append)       # warm-up
    b.clear()
    import gc
    gc.collect()
    ms = None
    for _ in range(3):                                              # wall clock around a Python call: best of three
        b.clear()
        t0 = time.perf_counter()
        got = concatenate_outputs(contigs, table, rng=random.Random(1), out=b.append)
        t = (time.perf_counter() - t0) * 1e3
        ms = t if ms is None or t < ms else ms
    from oracle import concat_oracle                                # the CPU restatement, timed beside it
    t0 = time.perf_counter()
    exp = concat_oracle.concatenate_outputs(contigs, table, random.Random(1), out=a.append)
    ms_port = (time.perf_counter() - t0) * 1e3
    return {'workload': '%d peptide contigs cut from one 2400-letter sequence (adjacent swaps near the ends in 20 %%), '
                        'k = 20..6, random.seed(1)' % len(contigs),
            'outputs': len(got), 'ms': ms, 'python_port_ms': ms_port,
            'equal_port': bool(got == exp and a == b)}


def run_dbg(args, cfg, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from dbg_b200 import Graph, _native
    from dbg_b200 import dist as ddist

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product has no CPU path')
    _native.get()
    torch.cuda.set_device(local_rank)
    R = Runner(rank, local_rank, world)
    dev = R.dev
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    k, thr = cfg['k'], cfg['thr']
    fin = bool(args.final)
    read_base = 0
    batched = 'T' in cfg
    sharded = world > 1 and not batched
    if fin and (batched or sharded):
        raise SystemExit('bench.py: --final runs on one GPU, one graph')
    read_batch = None
    if batched:
        # independent templates: shard by batch id, no collective; rank r builds its own T templates
        bases, off, read_batch, _ = make_reads_batched(cfg, seed_offset=rank)
    elif world > 1:
        bases, off, read_base, _ = make_reads_shard(cfg, rank, world)
    else:
        bases, off = make_reads(cfg)
    n_reads, n_bases = len(off) - 1, len(bases)
    nwin = n_windows_of(off, k)
    d_bases, d_off = R.to_dev(bases, off)
    h_bases = torch.from_numpy(bases).pin_memory()
    h_off = torch.from_numpy(off.view(np.int64)).pin_memory()
    hb_np, ho_np = h_bases.numpy(), h_off.numpy().view(np.uint64)

    if batched:
        n_batches = cfg['T']
        d_batch = torch.from_numpy(read_batch.view(np.int32)).to(dev)
        h_batch = torch.from_numpy(read_batch.view(np.int32)).pin_memory()
        hbatch_np = h_batch.numpy().view(np.uint32)
    g = Graph(k, device=local_rank, key_words=2 if batched and 2 * k + cfg['T'].bit_length() > 64 else 0)
    stream = R.stream
    g.set_stream(stream.cuda_stream)
    last = {}
    ex_stats = {}

    def step_device():
        if batched:
            g.build_batched_packed(None, None, None, n_batches, threshold=thr, final=False,
                                   device_ptrs=(d_bases.data_ptr(), d_off.data_ptr(), d_batch.data_ptr(), n_reads, n_bases))
            g.walk(stop_at_branch=True)
        elif sharded:
            last['res'] = ddist.build_sharded(g, None, read_base, threshold=thr, final=False, stats=ex_stats,
                                              local_device=(d_bases.data_ptr(), d_off.data_ptr(), n_reads, n_bases))
        else:
            g.build_device(d_bases.data_ptr(), d_off.data_ptr(), n_reads, n_bases, threshold=thr, final=fin,
                           min_count=args.min_count)
            g.walk(stop_at_branch=not fin)

    def step_host():
        if batched:
            g.build_batched_packed(hb_np, ho_np, hbatch_np, n_batches, threshold=thr, final=False)
            g.walk(stop_at_branch=True)
        elif sharded:
            last['res'] = ddist.build_sharded(g, (hb_np, ho_np), read_base, threshold=thr, final=False)
        else:
            g.build(None, threshold=thr, final=fin, packed=(hb_np, ho_np), min_count=args.min_count)
            g.walk(stop_at_branch=not fin)

    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    stage = {}

    def collect():
        st_ = g.stats()
        for f in STAGE_FIELDS:
            stage[f] = stage.get(f, 0.0) + getattr(st_, f) / args.steps

    ms_total = R.timed(step_device, args.steps, collect)
    st = g.stats()
    # probe lengths of the two hash tables the timed step left (north-star evidence; outside every timed region)
    probe = None
    if not sharded:
        try:
            probe = {w: {f: v for f, v in g.probe_stats(w).items() if f != 'hist'} for w in ('nodes', 'chunks')}
            probe['what'] = 'slots a lookup of an entry touches (linear probing from its home slot); dbg_probe_stats'
        except Exception as e:                                           # e.g. a build that did not go through chunks
            probe = {'error': str(e)[:200]}
    launches = int(st.n_launches)
    mode = last['res'].mode if sharded else ('batched' if batched else 'single')
    n_contigs_local = len(last['res'].contigs) if sharded and last['res'].mode == 'sharded' else int(st.n_contigs)
    totals = last['res'].totals if sharded else None
    # the same step with the handle's remembered table sizes switched off (every build sketches its tables first,
    # as the very first build on a handle does): reported beside the headline, never instead of it
    cold = None
    if not sharded and not batched:
        os.environ['DBG_B200_NO_HINT'] = '1'
        try:
            step_device()
            cstage = {}

            def ccollect():
                st_ = g.stats()
                for f in STAGE_FIELDS:
                    cstage[f] = cstage.get(f, 0.0) + getattr(st_, f) / min(args.steps, 10)
            ms_cold = R.timed(step_device, min(args.steps, 10), ccollect) / min(args.steps, 10)
            cold = {'ms_per_step': ms_cold, 'value': nwin / (ms_cold * 1e-3),
                    'table_build_ms': cstage['ms_pack'] + cstage['ms_estimate'] + cstage['ms_insert']}
        finally:
            del os.environ['DBG_B200_NO_HINT']
        step_device()
    for _ in range(2):
        step_host()
    ms_e2e = R.timed(step_host, args.steps)
    h2d_stage = g.stats().ms_h2d
    clocks = sampler.stop() if rank == 0 else None
    # stage breakdown of the multi-GPU step: a SEPARATE pass (it synchronises per stage, so never inside the timed loop)
    dist_ms = {}
    if sharded:
        for _ in range(3):
            dist_ms.clear()
            # (the handle now holds THIS build: node ids of an earlier step's contigs would not match its table)
            last['res'] = ddist.build_sharded(g, None, read_base, threshold=thr, final=False, timings=dist_ms,
                                              local_device=(d_bases.data_ptr(), d_off.data_ptr(), n_reads, n_bases))

    ms_step = ms_total / args.steps
    total_windows = R.total(nwin)
    total_contigs = R.total(n_contigs_local) if sharded and mode == 'sharded' else float(n_contigs_local)
    value = total_windows / (ms_step * 1e-3)
    e2e_value = total_windows / (ms_e2e / args.steps * 1e-3)

    # ---- parity (outside every timed region): the N-rank result against a 1-GPU build of the same reads ----
    parity = None
    if sharded:
        snap = ddist.global_snapshot(last['res'])
        if rank == 0:
            fb, fo = make_reads_dedup_by_start(cfg, world)
            ref = Graph(k, device=local_rank)
            ref.build(None, threshold=thr, final=False, packed=(fb, fo))
            ref.walk(stop_at_branch=True)
            exp = ddist.single_snapshot(ref)
            diff = [f for f in exp if not np.array_equal(np.asarray(snap[f]), np.asarray(exp[f]))]
            parity = {'against': '1-GPU build (dbg_build + dbg_walk) of the same %d reads on rank 0' % (len(fo) - 1),
                      'rows': sorted(exp), 'rows_differing': diff, 'equal': not diff,
                      'sha256_n_ranks': ddist.snapshot_digest(snap), 'sha256_1_gpu': ddist.snapshot_digest(exp),
                      'nodes': int(len(exp['tag'])), 'contigs': int(len(exp['c_n'])), 'mode': mode}
            ref.close()
            del fb, fo
        R.barrier()

    extras = {}
    if not batched and not args.no_extras:
        del d_bases, d_off
        torch.cuda.empty_cache()
        if args.config == 'c2':
            extras['strong_c5'] = strong_c5(R, args)
            if world > 1:
                extras['c4_batched'] = c4_batched(R, args)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    W = 8 if k <= 31 else 16
    b_build = cfg['L'] / (cfg['L'] - k) + W + 16                     # SURVEY.md 8 D4, bytes per window
    # the table build = every kernel between the ASCII reads in HBM and the finished node table
    # (k_pack, k_masks, k_chunk_insert, k_expand sketch + passes, incl. the host round trips between them);
    # at N > 1 this is rank 0's LOCAL share (its reads' chunks + the expansion of the chunks it owns)
    ms_table = stage['ms_pack'] + stage['ms_estimate'] + stage['ms_insert']
    achieved = nwin * b_build / (ms_table * 1e-3) / 1e9
    n_nodes_global = totals['n_nodes'] if totals else int(st.n_nodes)
    walk_ms = stage['ms_walk_total'] if not sharded else (dist_ms.get('walk_ms') or float('nan'))
    walk_bytes = 16 * (n_nodes_global / world) + 32 * (total_contigs / world)            # SURVEY.md 8 D5, per GPU
    traffic = None
    traffic_file = os.path.join(ROOT, 'profiles', 'traffic_%s_n%d.json' % (args.config, world))
    if os.path.exists(traffic_file):                                  # measured for exactly this config and N, else null
        with open(traffic_file) as f:
            traffic = json.load(f).get('table_build_dram_bytes_per_step')
    line = {
        'metric': 'kmers_per_sec_graph_build_plus_contig_walk', 'value': value, 'unit': 'k-mers/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_step,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'u64' if k <= 31 else 'u128', 'data': 'synthetic',
        'config': static_config(args, cfg, world),
        'workload_size': {'reads_rank0': int(n_reads), 'windows_rank0': int(nwin), 'windows_total': int(total_windows)},
        'mode': mode,
        'build_kmers_per_sec': total_windows / ((ms_step - (walk_ms if walk_ms == walk_ms else 0.0)) * 1e-3),
        'contigs_per_sec': total_contigs / (walk_ms * 1e-3) if walk_ms == walk_ms and walk_ms > 0 else None,
        'dist_stage_ms': {kk: round(v, 4) for kk, v in dist_ms.items()} if sharded else None,
        'graph': {'nodes': int(n_nodes_global), 'edges': int(totals['n_edges'] if totals else st.n_edges),
                  'branch': int(totals['n_branch'] if totals else st.n_branch),
                  'starts': int(totals['n_starts'] if totals else st.n_starts), 'contigs': int(total_contigs),
                  'table_slots_rank0': int(st.table_slots), 'table_bytes_rank0': int(st.table_bytes),
                  'jump_rounds': int(st.jump_rounds), 'insert_passes': int(st.n_insert_passes),
                  'chunk_instances_rank0': int(st.n_chunk_instances), 'chunk_slots_rank0': int(st.chunk_slots)},
        'stage_ms': {f: round(v, 4) for f, v in stage.items()},
        'roofline': {'bound': 'hbm', 'kernel': 'table build: k_pack + k_masks + k_chunk_insert + k_expand (all passes)',
                     'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                     'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                     'bytes_per_window': b_build, 'kernel_ms': ms_table},
        'roofline_walk': {'bound': 'hbm', 'kernel': 'contig walk: k_link_init .. k_emit (list ranking + descriptors)',
                          'achieved': walk_bytes / (walk_ms * 1e-3) / 1e9 if walk_ms == walk_ms and walk_ms > 0 else None,
                          'peak': peak, 'unit': 'GB/s',
                          'frac': walk_bytes / (walk_ms * 1e-3) / 1e9 / peak if walk_ms == walk_ms and walk_ms > 0 else None,
                          'bytes': walk_bytes, 'kernel_ms': walk_ms,
                          'note': 'SURVEY.md 8 D5: 16 B per node + 32 B per contig; latency-bound pointer chasing'},
        'e2e': {'value': e2e_value, 'unit': 'k-mers/s', 'h2d_bytes_per_step': int(n_bases + 8 * (n_reads + 1)),
                'd2h_bytes_per_step': int(32 * n_contigs_local + 2 * 4096 * 4 + 4 * 128), 'ms_per_step': ms_e2e / args.steps,
                'ms_h2d': h2d_stage,
                'api': 'dbg_build (host pointers, pinned) + dbg_walk + dbg_contigs' if not sharded else
                       'dist.build_sharded over dbg_build(DBG_CHUNKS_ONLY, host pointers, pinned) ... dbg_shard_walk + dbg_shard_contigs'},
        'table_sizing': {
            'timed_steps': 'chunk and node table sized from the previous build on the same handle (grow-only buffers '
                           'remember their fill; checked after the build, redone with sketches when it does not fit)',
            'cold': None if cold is None else dict(
                cold, roofline_frac=nwin * b_build / (cold['table_build_ms'] * 1e-3) / 1e9 / peak,
                note='every build sizes its tables from HyperLogLog sketches first (DBG_B200_NO_HINT=1), as the first '
                     'build on a handle does')},
        'probe_length': probe,
        'gpu_launches': launches * args.steps,
        'gpu_launches_per_step': launches,
        'clocks': clocks,
        'host_cores': os.cpu_count(),
    }
    if sharded:
        ex_ms = dist_ms.get('exchange_ms')
        out_b = ex_stats.get('exchange_bytes_out')
        line['roofline_exchange'] = {
            'bound': 'nvlink', 'what': 'chunk-record all-to-all: bytes that leave rank 0 / time of the exchange stage '
                                       '(counts exchange + all_to_all_single + fence)',
            'bytes_out': out_b, 'ms': ex_ms, 'achieved': out_b / (ex_ms * 1e-3) / 1e9 if ex_ms and out_b else None,
            'peak': NVLINK_GBS, 'unit': 'GB/s', 'frac': out_b / (ex_ms * 1e-3) / 1e9 / NVLINK_GBS if ex_ms and out_b else None,
            'peak_source': 'measured peer copy per direction per GPU (B200_PROFILING.md; nominal 900)',
            'walk_bytes_out': ex_stats.get('walk_bytes_out')}
        line['parity'] = parity
    line.update(extras)
    if world == 1 and not args.no_cpu_baseline and not batched:
        py = cpu_python_port(cfg, args.cpu_sample_bp)
        cp = cpu_c_port(cfg, bases, off, final=fin)
        line['cpu_baseline'] = {
            'value': py['windows'] / (py['build_s'] + py['walk_s']), 'unit': 'k-mers/s', 'cores': 1, 'kind': 'port',
            'sample': 'pure-Python restatement of debruijn.py (oracle/dbg_oracle.py), construct_graph + '
                      'output_contigs on a ' + py['sample'],
            'build_kmers_per_sec': py['windows'] / py['build_s'], 'contigs_per_sec': py['contigs'] / py['walk_s'],
        }
        lit = cpu_literal(cfg, 2500)
        if lit is not None:
            line['cpu_baseline_literal'] = lit
        line['cpu_baseline_c'] = {
            'value': cp['windows'] / cp['total_s'], 'unit': 'k-mers/s', 'cores': 1, 'kind': 'port',
            'sample': 'C restatement (oracle/dbg_oracle.c) on the FULL workload of this run',
            'build_kmers_per_sec': cp['windows'] / (cp['build_s'] + cp['graph_s']),
            'contigs_per_sec': cp['contigs'] / max(cp['walk_s'], 1e-9),
        }
        # parity at N = 1: the product's result of the timed workload against the C oracle, row by row
        from oracle import c_oracle, compare
        g.shard_config(0, 0)
        g.build(None, threshold=thr, final=fin, packed=(bases, off), min_count=args.min_count)
        g.walk(stop_at_branch=not fin)
        res = c_oracle.run(None, k, thr, fin, strings=False, packed=(bases.tobytes(), off), min_count=args.min_count)
        rows = []
        bad = compare.compare_with_oracle(g, res, fin, report=lambda name, ok: rows.append(name))
        line['parity'] = {'against': 'C oracle (oracle/dbg_oracle.c) on the FULL workload of this run',
                          'rows': rows, 'rows_differing': bad, 'equal': not bad,
                          'nodes': int(res.n_nodes), 'contigs': int(len(res.contigs)), 'mode': mode}
        if not args.no_extras:
            line['e2e_api'] = e2e_api(R, args)
            line['n3_concat'] = n3_concat(R, args)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--config', default='c2', choices=sorted(CONFIGS))
    ap.add_argument('--impl', default='dbg', choices=['dbg', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-sample-bp', type=int, default=42_000)     # ~10 s of pure-Python work (the walk is super-linear)
    ap.add_argument('--ref-sample-bp', type=int, default=0)          # --impl reference: 0 = the full N=1 workload per step
    ap.add_argument('--min-count', type=int, default=1)              # EXTENSION (no reference oracle): abundance filter, configs[2] (ii)
    ap.add_argument('--final', action='store_true')                  # construct_graph(final=True): all paths enumerated (single GPU, e.g. --config c3f)
    ap.add_argument('--no-extras', action='store_true')              # skip strong_c5 / c4_batched / e2e_api / n3_concat
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    cfg = CONFIGS[args.config]
    if args.impl == 'reference':
        run_reference(args, cfg, rank, world)
        return
    run_dbg(args, cfg, rank, local_rank, world)


if __name__ == '__main__':
    main()
