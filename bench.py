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


def make_reads_shard(cfg, rank, world):
    """The distributed workload: ONE read set over a genome `world` times the config's size (weak
    scaling: per-GPU work fixed), deduped by start position, rank r holding the contiguous slice
    shard_bounds()[r] of the global read list.  Every rank derives the same global list."""
    G = cfg['G'] * world
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
    mat = genome[mine[:, None] + np.arange(L)[None, :]]
    if cfg['err']:
        r2 = np.random.default_rng(cfg['seed'] + 1000 + rank)
        flip = r2.random(mat.shape) < cfg['err']
        mat = np.where(flip, (mat + r2.integers(1, 4, mat.shape, dtype=np.uint8)) & 3, mat).astype(np.uint8)
    bases = np.ascontiguousarray(ASCII[mat]).reshape(-1)
    off = (np.arange(mat.shape[0] + 1, dtype=np.uint64) * np.uint64(L))
    return bases, off, lo, n_total


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


def cpu_c_port(cfg, bases, off, final=False):
    """The C restatement (oracle/dbg_oracle.c), 1 core, on the given reads."""
    from oracle import c_oracle
    t0 = time.perf_counter()
    res = c_oracle.run(None, cfg['k'], cfg['thr'], final, stages=3, strings=False, packed=(bases.tobytes(), off))
    t1 = time.perf_counter()
    return {'windows': int(res.n_windows), 'total_s': t1 - t0, 'build_s': res.t_build, 'graph_s': res.t_graph,
            'walk_s': res.t_walk, 'contigs': int(len(res.contigs)), 'nodes': int(res.n_nodes)}


def run_reference(args, cfg, rank):
    """--impl reference: the reference's CPU path on this box's host cores.  The reference is
    pure single-threaded Python that cannot finish configs[1], so each
    step is the C restatement of it (oracle/dbg_oracle.c, strictly faster than the reference:
    O(n) walk, packed keys) over the FULL workload, 1 thread -- the reference has no parallelism
    to use more."""
    if rank != 0:
        return
    # bound the whole run to ~2 min: the C port does ~1e7 windows/s on one core
    budget_s = 120.0 / (args.steps + args.warmup)
    per_base = cfg['cov'] * (cfg['L'] - cfg['k']) / cfg['L']
    G_sample = int(min(cfg['G'], max(10_000, budget_s * 1.0e7 / per_base)))
    bases, off = make_reads(cfg, G=G_sample)
    nwin = n_windows_of(off, cfg['k'])
    times, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_c_port(cfg, bases, off)
        if i >= args.warmup:
            times.append(last['total_s'])
    t = float(np.mean(times))
    value = nwin / t
    line = {
        'impl': 'reference', 'metric': 'kmers_per_sec_graph_build_plus_contig_walk', 'value': value, 'unit': 'k-mers/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': t * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u64' if cfg['k'] <= 31 else 'u128',
        'data': 'synthetic', 'config': workload_config(args, cfg, off, nwin, 1),
        'contigs_per_sec': last['contigs'] / max(last['walk_s'], 1e-9),
        'cpu_baseline': {'value': value, 'unit': 'k-mers/s', 'cores': 1, 'kind': 'port',
                         'sample': '%s per step: %d bp genome of the workload\'s shape, %d reads, %d windows (C '
                                   'restatement oracle/dbg_oracle.c; the pure-Python reference cannot finish this '
                                   'config)' % ('FULL workload' if G_sample == cfg['G'] else 'bounded sample',
                                                G_sample, len(off) - 1, nwin)},
        'e2e': {'value': value, 'unit': 'k-mers/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'host_cores': os.cpu_count(),
    }
    print(json.dumps(line), flush=True)


def workload_config(args, cfg, off, nwin, world):
    wl = cfg['name'] if world == 1 else \
        ('%s; genome scaled x%d (%d bp) so per-GPU work is fixed, ONE graph built by %d ranks' %
         (cfg['name'], world, cfg['G'] * world, world))
    return {'workload': wl, 'config_key': args.config, 'k': cfg['k'], 'threshold': cfg['thr'],
            'genome_bp': cfg['G'] * world, 'read_len': cfg['L'], 'coverage': cfg['cov'], 'error_rate': cfg['err'],
            'reads_per_gpu': int(len(off) - 1), 'windows_per_gpu': int(nwin), 'final': False,
            'parallelism': 'single GPU' if world == 1 else
            ('templates sharded x%d by batch id, no collective (independent graphs)' % world) if 'T' in cfg else
            ('reads sharded x%d; node records hash-partitioned by owner, NCCL all-to-all + all-gather; graph stages '
             'and contig walk replicated on every rank' % world),
            'l2_policy': 'inputs (ASCII reads + node table) larger than the 126 MB L2; no explicit flush'}


# ---------------------------------------------------------------------------------------------
# the product arm
# ---------------------------------------------------------------------------------------------
def run_dbg(args, cfg, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from dbg_b200 import Graph, _native

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product has no CPU path')
    _native.get()
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    k, thr = cfg['k'], cfg['thr']
    read_base = 0
    batched = 'T' in cfg
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
    d_bases = torch.from_numpy(bases).to(dev)
    d_off = torch.from_numpy(off.view(np.int64)).to(dev)
    h_bases = torch.from_numpy(bases).pin_memory()
    h_off = torch.from_numpy(off.view(np.int64)).pin_memory()
    hb_np, ho_np = h_bases.numpy(), h_off.numpy().view(np.uint64)

    if batched:
        n_batches = cfg['T']
        d_batch = torch.from_numpy(read_batch.view(np.int32)).to(dev)
        h_batch = torch.from_numpy(read_batch.view(np.int32)).pin_memory()
        hbatch_np = h_batch.numpy().view(np.uint32)
    g = Graph(k, device=local_rank, key_words=2 if batched and 2 * k + cfg['T'].bit_length() > 64 else 0)
    stream = torch.cuda.current_stream(dev)
    g.set_stream(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    dist_ms = {}

    def step_device():
        if batched:
            g.build_batched_packed(None, None, None, n_batches, threshold=thr, final=False,
                                   device_ptrs=(d_bases.data_ptr(), d_off.data_ptr(), d_batch.data_ptr(), n_reads, n_bases))
        elif world > 1:
            from dbg_b200 import dist as ddist
            ddist.build_distributed(g, None, read_base, threshold=thr, final=False, timings=dist_ms,
                                    local_device=(d_bases.data_ptr(), d_off.data_ptr(), n_reads, n_bases))
        else:
            g.build_device(d_bases.data_ptr(), d_off.data_ptr(), n_reads, n_bases, threshold=thr, final=False)
        g.walk(stop_at_branch=True)

    def step_host():
        if batched:
            g.build_batched_packed(hb_np, ho_np, hbatch_np, n_batches, threshold=thr, final=False)
        elif world > 1:
            from dbg_b200 import dist as ddist
            ddist.build_distributed(g, (hb_np, ho_np), read_base, threshold=thr, final=False)
        else:
            g.build(None, threshold=thr, final=False, packed=(hb_np, ho_np))
        return g.walk(stop_at_branch=True)

    def timed(fn, steps, collect=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
            if collect is not None:
                collect(g.stats())
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    stage = {}

    def collect(st):
        for f in ('ms_pack', 'ms_estimate', 'ms_table_init', 'ms_insert', 'ms_finalize', 'ms_order', 'ms_lists',
                  'ms_tips', 'ms_pull_reads', 'ms_build_total', 'ms_walk_rank', 'ms_walk_emit', 'ms_walk_total',
                  'ms_chunk_insert', 'ms_expand_sketch', 'ms_expand'):
            stage[f] = stage.get(f, 0.0) + getattr(st, f) / args.steps

    dist_ms.clear()
    ms_total = timed(step_device, args.steps, collect)
    dist_stage = {kk: round(v / args.steps, 4) for kk, v in dist_ms.items()}
    st = g.stats()
    launches = int(st.n_launches)
    for _ in range(2):
        step_host()
    ms_e2e = timed(step_host, args.steps)
    h2d_stage = g.stats().ms_h2d
    clocks = sampler.stop() if rank == 0 else None

    ms_step = ms_total / args.steps
    if world > 1:
        tw = torch.tensor([nwin], device=dev, dtype=torch.float64)
        dist.all_reduce(tw)
        total_windows = float(tw.item())
    else:
        total_windows = float(nwin)
    value = total_windows / (ms_step * 1e-3)
    e2e_value = total_windows / (ms_e2e / args.steps * 1e-3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    W = 8 if k <= 31 else 16
    b_build = cfg['L'] / (cfg['L'] - k) + W + 16                     # SURVEY.md 8 D4, bytes per window
    # the table build = every kernel between the ASCII reads in HBM and the finished node table
    # (k_pack, k_masks, k_chunk_insert, k_expand sketch + passes, incl. the host round trips between them)
    ms_table = stage['ms_pack'] + stage['ms_estimate'] + stage['ms_insert']
    achieved = nwin * b_build / (ms_table * 1e-3) / 1e9
    line = {
        'metric': 'kmers_per_sec_graph_build_plus_contig_walk', 'value': value, 'unit': 'k-mers/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_step,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'u64' if k <= 31 else 'u128', 'data': 'synthetic',
        'config': workload_config(args, cfg, off, nwin, world),
        'build_kmers_per_sec': total_windows / ((ms_step - stage['ms_walk_total']) * 1e-3),
        'contigs_per_sec': int(st.n_contigs) / (stage['ms_walk_total'] * 1e-3),
        'dist_stage_ms': dist_stage if world > 1 else None,
        'graph': {'nodes': int(st.n_nodes), 'edges': int(st.n_edges), 'branch': int(st.n_branch),
                  'starts': int(st.n_starts), 'contigs': int(st.n_contigs), 'table_slots': int(st.table_slots),
                  'table_bytes': int(st.table_bytes), 'jump_rounds': int(st.jump_rounds),
                  'insert_passes': int(st.n_insert_passes), 'chunk_instances': int(st.n_chunk_instances),
                  'chunk_list': int(st.n_chunk_list), 'chunk_slots': int(st.chunk_slots)},
        'stage_ms': {f: round(v, 4) for f, v in stage.items()},
        'roofline': {'bound': 'hbm', 'kernel': 'table build: k_pack + k_masks + k_chunk_insert + k_expand (all passes)',
                     'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                     'frac': achieved / peak, 'traffic': None, 'peak_source': peak_src,
                     'bytes_per_window': b_build, 'kernel_ms': ms_table},
        'e2e': {'value': e2e_value, 'unit': 'k-mers/s', 'h2d_bytes_per_step': int(n_bases + 8 * (n_reads + 1)),
                'd2h_bytes_per_step': int(32 * st.n_starts + 2 * 4096 * 4 + 4 * 128), 'ms_per_step': ms_e2e / args.steps,
                'ms_h2d': h2d_stage, 'api': 'dbg_build (host pointers, pinned) + dbg_walk + dbg_contigs'},
        'gpu_launches': launches * args.steps,
        'gpu_launches_per_step': launches,
        'clocks': clocks,
        'host_cores': os.cpu_count(),
    }
    traffic_file = os.path.join(ROOT, 'profiles', 'traffic_%s.json' % args.config)
    if os.path.exists(traffic_file):
        with open(traffic_file) as f:
            line['roofline']['traffic'] = json.load(f).get('table_build_dram_bytes_per_step')
    if world == 1 and not args.no_cpu_baseline and not batched:
        py = cpu_python_port(cfg, args.cpu_sample_bp)
        cp = cpu_c_port(cfg, bases, off)
        line['cpu_baseline'] = {
            'value': py['windows'] / (py['build_s'] + py['walk_s']), 'unit': 'k-mers/s', 'cores': 1, 'kind': 'port',
            'sample': 'pure-Python restatement of debruijn.py (oracle/dbg_oracle.py), construct_graph + '
                      'output_contigs on a ' + py['sample'],
            'build_kmers_per_sec': py['windows'] / py['build_s'], 'contigs_per_sec': py['contigs'] / py['walk_s'],
        }
        line['cpu_baseline_c'] = {
            'value': cp['windows'] / cp['total_s'], 'unit': 'k-mers/s', 'cores': 1, 'kind': 'port',
            'sample': 'C restatement (oracle/dbg_oracle.c) on the FULL workload of this run',
            'build_kmers_per_sec': cp['windows'] / (cp['build_s'] + cp['graph_s']),
            'contigs_per_sec': cp['contigs'] / max(cp['walk_s'], 1e-9),
            'parity': {'nodes_equal': cp['nodes'] == int(st.n_nodes), 'contigs_equal': cp['contigs'] == int(st.n_contigs)},
        }
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
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    cfg = CONFIGS[args.config]
    if args.impl == 'reference':
        run_reference(args, cfg, rank)
        return
    run_dbg(args, cfg, rank, local_rank, world)


if __name__ == '__main__':
    main()
