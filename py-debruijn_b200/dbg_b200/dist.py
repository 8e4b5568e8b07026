"""Multi-GPU construct_graph + output_contigs: one process per GPU, reads sharded by index, the
distinct super-k-mers ("chunks") of every rank's reads exchanged ONCE with torch.distributed (NCCL
over NVLink on GPUs; gloo on CPU in the tests), everything after that owner-computes.

The window loop (debruijn.py:123-144) is data parallel over reads.  What a rank's slice of the reads
leaves behind is its set of distinct chunks {string | min tag | multiplicity}, whose merge is associative
(min, sum) -- tags carry the GLOBAL read index, so the reference's dict order and its indegree flag are
partition invariant.  A chunk's owner is a function of its minimizer, which all its
nodes share and every occurrence of those nodes has, so everything about one node meets on one rank.

    rank r:  reads[lo_r:hi_r] --dbg_build(CHUNKS_ONLY)--> distinct chunks of the local reads
             --dbg_chunk_export_records(P)--> chunk records grouped by owner
             ==all-to-all==> the chunks this rank owns --merge, dbg_chunk_expand--> its SHARD of the node table

    build_sharded (default for large graphs):
             degrees / pruningEdges / tie order / branch flags / start list on the shard (stages 0-2)
             contig walk = two-level list ranking (csrc/dbg_shard.cuh): exits' successor keys ==all-to-all==>
             owners mark entries, rank their shard, answer 4-byte ordinals ==all-to-all==> reduced list
             ==all-gather (16 B per entry/start)==> ranked on every rank for its own starts.
             No node record leaves its owner.  A graph with branch nodes (tip pruning, debruijn.py:241-254, is
             order dependent and needs whole neighbourhoods) falls back to:
    build_distributed (complete graph on every rank; what the drop-in twin `construct_graph` returns):
             shards ==all-gather==> full node table everywhere, graph stages + walk replicated.
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.distributed as dist

from .graph import Graph, pack_reads


def _device_of(graph, group):
    backend = dist.get_backend(group)
    if backend == 'nccl':
        return torch.device('cuda', graph.device)
    return torch.device('cpu')


def _all_gather_var(t, group, ns=None):
    """All-gather of int64 tensors [n_i, words] with different n_i; returns the concatenation.
    `ns` = the n_i of every rank if the caller already knows them (saves a collective + a host wait)."""
    world = dist.get_world_size(group)
    if ns is None:
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        ns = torch.zeros(world, dtype=torch.int64, device=t.device)
        dist.all_gather_into_tensor(ns, n, group=group)
        ns = [int(x) for x in ns.tolist()]
    m = max(max(ns), 1)
    words = t.shape[1]
    if t.shape[0] < m:
        pad = torch.zeros((m, words), dtype=t.dtype, device=t.device)
        pad[:t.shape[0]] = t
    else:
        pad = t.contiguous()
    out = torch.empty((world * m, words), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    if all(c == m for c in ns):
        return out, ns
    return torch.cat([out[r * m:r * m + c] for r, c in enumerate(ns)], dim=0).contiguous(), ns


def shard_bounds(n_reads, world):
    """Contiguous, balanced read ranges: rank r gets [b[r], b[r+1])."""
    return [n_reads * r // world for r in range(world + 1)]


class _Clock:
    """Per-stage host timing (with a device sync per stage) -- only when the caller asks for it."""

    def __init__(self, timings, dev):
        self.t, self.dev = timings, dev
        self.t0 = time.perf_counter()

    def sub(self, name):
        """Finer breakdown inside a stage: recorded under `name` AND left in the enclosing stage's total."""
        if self.t is None:
            return
        if self.dev.type == 'cuda':
            torch.cuda.synchronize(self.dev)
        now = time.perf_counter()
        self.t[name] = self.t.get(name, 0.0) + (now - getattr(self, 'ts', self.t0)) * 1e3
        self.ts = now

    def tick(self, name):
        if self.t is None:
            return
        if self.dev.type == 'cuda':
            torch.cuda.synchronize(self.dev)
        now = time.perf_counter()
        self.t[name] = self.t.get(name, 0.0) + (now - self.t0) * 1e3
        self.t0 = now
        self.ts = now


def _fence(dev, graph=None):
    """A handle that runs on its own stream must not read what a torch collective wrote before the
    collective is done: make the host wait for it.  A handle set to torch's current stream
    (Graph.set_stream) is ordered behind the collective by the stream itself -- no host wait."""
    if dev.type == 'cuda' and not (graph is not None and graph.shares_stream(torch.cuda.current_stream(dev).cuda_stream)):
        torch.cuda.current_stream(dev).synchronize()


def _exchange_counts(send_counts, dev, group):
    sc = torch.tensor(np.asarray(send_counts, dtype=np.int64), dtype=torch.int64, device=dev)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc, group=group)
    return [int(x) for x in rc.tolist()]


def build_shard(graph: Graph, local_packed, read_idx_base, group=None, clock=None, local_device=None, stats=None):
    """The one real exchange: chunks of the local reads -> all-to-all by owner -> this rank's shard of
    the node table.  `stats`, if a dict, receives the bytes this rank sent to OTHER ranks."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = _device_of(graph, group)
    graph.shard_config(0, 0)
    if local_device is not None:
        graph.build_chunks(device_ptrs=local_device, read_idx_base=read_idx_base)
    else:
        graph.build_chunks(packed=local_packed, read_idx_base=read_idx_base)
    if clock:
        clock.tick('local_chunks_ms')
    cw = graph.chunk_record_words()
    counts = graph.chunk_export_counts(world)
    send = torch.empty((int(counts.sum()), cw), dtype=torch.int64, device=dev)
    graph.chunk_export_records(world, counts, send.data_ptr())
    rc = _exchange_counts(counts, dev, group)
    recv = torch.empty((sum(rc), cw), dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv, send, output_split_sizes=rc, input_split_sizes=[int(c) for c in counts], group=group)
    _fence(dev, graph)
    if stats is not None:
        stats['exchange_bytes_out'] = int(counts.sum() - counts[rank]) * cw * 8
    if clock:
        clock.tick('exchange_ms')
    graph.chunk_table_reset(max(1, recv.shape[0]))
    graph.chunk_merge_records(recv.data_ptr(), recv.shape[0])
    if clock:
        clock.sub('expand.merge_ms')
    graph.chunk_expand()
    if clock:
        clock.tick('owner_expand_ms')
    return recv                                     # keep alive until the caller is done with the stream


def _replicate(graph: Graph, threshold, final, exact_order, pull_reads, group, clock):
    """From one shard per rank to the complete graph on every rank (all-gather of node records)."""
    world = dist.get_world_size(group)
    dev = _device_of(graph, group)
    graph.shard_config(0, 0)
    rw = graph.record_words()
    c1 = graph.export_counts(1)
    shard = torch.empty((int(c1[0]), rw), dtype=torch.int64, device=dev)
    graph.export_records(1, c1, shard.data_ptr())
    everything, ns = _all_gather_var(shard, group)
    _fence(dev, graph)
    if clock:
        clock.tick('gather_ms')
    graph.table_reset(max(1, sum(ns)))
    graph.merge_records(everything.data_ptr(), everything.shape[0])
    del shard, everything
    if clock:
        clock.tick('full_merge_ms')
    # graph stages; the tie-order scan runs over the chunks each rank owns and is merged by key
    graph.graph_stage(0, threshold, final, exact_order, pull_reads)
    st = graph.stats()
    if exact_order and st.n_multi:
        graph.graph_stage(1, threshold, final, exact_order, pull_reads)
        ow = graph.order_words()
        cap = int(st.n_multi) + 8
        mine = torch.empty((cap, ow), dtype=torch.int64, device=dev)
        n = graph.order_export(mine.data_ptr(), cap)
        allrec, _ = _all_gather_var(mine[:n].contiguous(), group)
        _fence(dev, graph)
        graph.order_merge(allrec.data_ptr(), allrec.shape[0])
    graph.graph_stage(2, threshold, final, exact_order, pull_reads)
    if clock:
        clock.tick('graph_ms')
    return graph.pull_read_flags()


def build_distributed(graph: Graph, local_packed, read_idx_base, threshold=3, final=False, exact_order=True,
                      pull_reads=True, group=None, timings=None, local_device=None):
    """Distributed construct_graph that leaves the COMPLETE graph on every rank (the handle then answers
    everything a single-GPU handle does).  `local_packed` = (uint8 bases, uint64 offsets) of this rank's
    reads in HOST memory, or `local_device` = (d_bases_ptr, d_off_ptr, n_reads, n_bases) for reads already
    in HBM; `read_idx_base` is the global index of the rank's first read.  Returns the LOCAL pull-out
    flags (uint8 per local read)."""
    dev = _device_of(graph, group)
    clock = _Clock(timings, dev) if timings is not None else None
    keep = build_shard(graph, local_packed, read_idx_base, group, clock, local_device)
    flags = _replicate(graph, threshold, final, exact_order, pull_reads, group, clock)
    del keep
    return flags


class ShardedResult:
    """What build_sharded leaves on a rank: its shard (graph handle), the contig records of its own
    starts, and the global counters."""

    def __init__(self, graph, mode, contigs, totals, rank, world, pull_flags=None):
        self.graph, self.mode, self.contigs, self.totals = graph, mode, contigs, totals
        self.rank, self.world = rank, world
        self.pull_flags = pull_flags


def build_sharded(graph: Graph, local_packed, read_idx_base, threshold=3, final=False, exact_order=True,
                  group=None, timings=None, local_device=None, stats=None):
    """Owner-computes construct_graph + output_contigs (module docstring).  Returns a ShardedResult:
    mode 'sharded' -> .contigs = this rank's dbg_shard_contig_t records (sorted by tag; all ranks' records
    merged by tag are the reference's contig order), .graph holds only the shard;
    mode 'replicated' (the graph has branch nodes) -> .graph is complete and walked on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = _device_of(graph, group)
    clock = _Clock(timings, dev) if timings is not None else None
    keep = build_shard(graph, local_packed, read_idx_base, group, clock, local_device, stats)
    # ---- degrees / prune / branch flags on the shard ----
    graph.shard_config(rank, world)
    graph.graph_stage(0, threshold, final, exact_order, False)
    if clock:
        clock.sub('graph.stage0_ms')
    st = graph.stats()
    names = ('n_nodes', 'n_edges', 'n_multi', 'n_branch', 'n_starts', 'n_windows')
    counts = np.zeros(world, np.uint64)
    if st.n_branch == 0:
        # no branch node on this shard: finish the graph stages and list the exits now, so that ONE collective carries
        # the global counters AND everybody's exit counts (a rank that does have branch nodes sends zeros; then some
        # rank reports n_branch > 0 and all of them take the replicated path below)
        if exact_order and st.n_multi:
            graph.graph_stage(1, threshold, final, exact_order, False)
        graph.graph_stage(2, threshold, final, exact_order, False)
        counts = graph.shard_exit_counts(world)
    vec = torch.tensor([int(getattr(st, f)) for f in names] + [int(c) for c in counts], dtype=torch.int64, device=dev)
    allv = torch.empty(world * (len(names) + world), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allv, vec, group=group)
    m = allv.view(world, len(names) + world).tolist()
    totals = {f: sum(m[r][i] for r in range(world)) for i, f in enumerate(names)}
    if totals['n_branch']:
        # ordered tip pruning needs whole neighbourhoods on one rank: complete graph everywhere
        flags = _replicate(graph, threshold, final, exact_order, True, group, clock)
        graph.walk(stop_at_branch=not final)
        if clock:
            clock.tick('walk_ms')
        del keep
        return ShardedResult(graph, 'replicated', None, totals, rank, world, pull_flags=flags)
    del keep
    if clock:
        clock.tick('graph_ms')
    # ---- contig walk: exits -> owners, local ranking, reduced list ----
    W = int(st.key_words)
    n_exit = int(counts.sum())
    exit_keys = torch.empty((n_exit, W), dtype=torch.int64, device=dev)
    graph.shard_exit_keys(exit_keys.data_ptr())
    if clock:
        clock.sub('walk.exit_keys_ms')
    rc = [int(m[r][len(names) + rank]) for r in range(world)]           # what every rank sends to this one
    sc = [int(c) for c in counts]
    req = torch.empty((sum(rc), W), dtype=torch.int64, device=dev)
    dist.all_to_all_single(req, exit_keys, output_split_sizes=rc, input_split_sizes=sc, group=group)
    _fence(dev, graph)
    if clock:
        clock.sub('walk.exchange_keys_ms')
    graph.shard_mark_entries(req.data_ptr(), req.shape[0])
    if clock:
        clock.sub('walk.mark_entries_ms')
    n_rv = graph.shard_rank_local()
    if clock:
        clock.sub('walk.rank_local_ms')
    nrv = torch.tensor([n_rv], dtype=torch.int64, device=dev)
    all_nrv = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_nrv, nrv, group=group)
    all_nrv = [int(x) for x in all_nrv.tolist()]
    rid_base = np.zeros(world, np.uint64)
    rid_base[1:] = np.cumsum(all_nrv)[:-1]
    ords = torch.empty((req.shape[0],), dtype=torch.int32, device=dev)
    graph.shard_entry_ordinals(ords.data_ptr())
    replies = torch.empty((n_exit,), dtype=torch.int32, device=dev)
    dist.all_to_all_single(replies, ords, output_split_sizes=sc, input_split_sizes=rc, group=group)
    _fence(dev, graph)
    if clock:
        clock.sub('walk.replies_ms')
    mine = torch.empty((n_rv, 2), dtype=torch.int64, device=dev)
    graph.shard_reduced_records(replies.data_ptr(), rid_base, mine.data_ptr())
    everything, _ = _all_gather_var(mine, group, all_nrv)
    _fence(dev, graph)
    if clock:
        clock.sub('walk.gather_reduced_ms')
    if stats is not None:
        stats['walk_bytes_out'] = (n_exit - sc[rank]) * W * 8 + (sum(rc) - rc[rank]) * 4 + n_rv * 16 * (world - 1)
    contigs = graph.shard_walk(everything.data_ptr(), everything.shape[0], int(rid_base[rank]))
    totals['n_contigs_local'] = len(contigs)
    if clock:
        clock.tick('walk_ms')
    return ShardedResult(graph, 'sharded', contigs, totals, rank, world)


# ---------------------------------------------------------------------------------------------
# snapshots for parity checks (outside any timed region): everything keyed by node KEY, so a
# sharded result, a replicated result and a single-GPU result compare row by row
# ---------------------------------------------------------------------------------------------
NODE_FLAG_MASK = 0x3F                    # F_OCC .. F_MULTI; the shard-local exit / entry bits are not graph content


def _gather_arrays(arrs, group, dev):
    """list of equally long 1-d/2-d numpy arrays -> per array the concatenation over ranks (every rank)."""
    out = []
    for a in arrs:
        a = np.ascontiguousarray(a)
        flat = a.reshape(len(a), -1) if len(a) else a.reshape(0, int(np.prod(a.shape[1:])) if a.ndim > 1 else 1)
        as64 = flat.astype(np.uint64).view(np.int64)
        t = torch.from_numpy(np.ascontiguousarray(as64)).to(dev)
        g, _ = _all_gather_var(t, group)
        res = g.cpu().numpy().view(np.uint64).astype(a.dtype)
        out.append(res.reshape((-1,) + a.shape[1:]))
    return out


def global_snapshot(res: ShardedResult, group=None):
    """Node table in the reference's dict order (sorted by tag) and the contig descriptors in the
    reference's order with node ids replaced by node keys -- identical on every rank."""
    g = res.graph
    dev = _device_of(g, group)
    nd = g.nodes()
    if res.mode == 'replicated':
        d = g.contig_descriptors()
        slot_to_idx = np.full(int(g.stats().table_slots), -1, np.int64)
        slot_to_idx[nd['id']] = np.arange(len(nd['id']))
        si, ti = slot_to_idx[d['start']], slot_to_idx[d['terminal']]
        return {'key_lo': nd['key_lo'], 'key_hi': nd['key_hi'], 'tag': nd['tag'], 'cnt': nd['cnt'],
                'flags': nd['flags'] & NODE_FLAG_MASK, 'kept': nd['kept'],
                'c_start_lo': nd['key_lo'][si], 'c_start_hi': nd['key_hi'][si], 'c_n': d['n_nodes'],
                'c_term_lo': nd['key_lo'][ti], 'c_term_hi': nd['key_hi'][ti], 'c_kind': d['kind']}
    rank_col = np.full(len(nd['id']), res.rank, np.uint32)
    key_lo, key_hi, tag, cnt, flags, kept, ids, ranks = _gather_arrays(
        [nd['key_lo'], nd['key_hi'], nd['tag'], nd['cnt'], nd['flags'], nd['kept'], nd['id'], rank_col], group, dev)
    order = np.argsort(tag, kind='stable')
    c = res.contigs
    (c_tag, c_ss, c_sr, c_ts, c_tr, c_n, c_kind) = _gather_arrays(
        [c['tag'], c['start_slot'], c['start_rank'], c['terminal_slot'], c['terminal_rank'], c['n_nodes'], c['kind']], group, dev)
    corder = np.argsort(c_tag, kind='stable')
    # (rank, slot) -> row of the gathered node arrays
    gid = (ranks.astype(np.uint64) << np.uint64(32)) | ids.astype(np.uint64)
    gorder = np.argsort(gid)
    gsorted = gid[gorder]

    def rows(rk, sl):
        q = (rk.astype(np.uint64) << np.uint64(32)) | sl.astype(np.uint64)
        pos = np.searchsorted(gsorted, q)
        assert len(q) == 0 or (pos < len(gsorted)).all() and (gsorted[np.minimum(pos, len(gsorted) - 1)] == q).all(), \
            'a contig names a node no shard holds'
        return gorder[pos]
    si, ti = rows(c_sr[corder], c_ss[corder]), rows(c_tr[corder], c_ts[corder])
    return {'key_lo': key_lo[order], 'key_hi': key_hi[order], 'tag': tag[order], 'cnt': cnt[order],
            'flags': flags[order] & NODE_FLAG_MASK, 'kept': kept[order],
            'c_start_lo': key_lo[si], 'c_start_hi': key_hi[si], 'c_n': c_n[corder],
            'c_term_lo': key_lo[ti], 'c_term_hi': key_hi[ti], 'c_kind': c_kind[corder]}


def single_snapshot(g: Graph):
    """The same structure from a single-GPU handle that has been built and walked."""
    return global_snapshot(ShardedResult(g, 'replicated', None, None, 0, 1))


def snapshot_digest(snap):
    import hashlib
    h = hashlib.sha256()
    for f in sorted(snap):
        a = np.ascontiguousarray(snap[f])
        h.update(f.encode())
        h.update(str(a.dtype).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def construct_graph(reads, k, threshold=3, final=False, group=None, device=None):
    """Distributed twin of debruijn.construct_graph: call it on every rank with the SAME read
    list; returns the same tuple on every rank (complete graph everywhere)."""
    import debruijn as db
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    b = shard_bounds(len(reads), world)
    local = reads[b[rank]:b[rank + 1]]
    g = Graph(k, device=device)
    flags = build_distributed(g, pack_reads(local), b[rank], threshold, final, group=group)
    dev = _device_of(g, group)
    f = torch.from_numpy(flags.astype(np.int64)).to(dev).reshape(-1, 1)
    allf, _ = _all_gather_var(f, group)
    allf = allf.reshape(-1).cpu().numpy()
    st = g.stats()
    print('number of {}mer: '.format(k), st.n_nodes)
    print('branch number: ', st.n_branch)
    if final:
        pull_out_read, branch_kmer = [], db.KmerList((), g)
    else:
        pull_out_read = [r for r, fl in zip(reads, allf.tolist()) if fl]
        branch_kmer = db.KmerList(g.label_of_ids(g.branch_ids()), g)
    already = db.KmerList(g.label_of_ids(g.pulled_ids()), g)
    return (db.Vertices(g), db.Edges(g)), pull_out_read, branch_kmer, already, db.EdgeCountTable(g)
