"""Multi-GPU construct_graph: one process per GPU, reads sharded by index, node records
exchanged with torch.distributed (NCCL over NVLink on GPUs; gloo on CPU in the tests).

The window loop (debruijn.py:123-144) is data parallel over reads.  What a rank's slice of the
reads leaves behind is a set of node records {key | first-occurrence tag | 4 successor counts}
whose merge is associative (min tag, summed counts) -- tags carry the GLOBAL read index, so the
reference's dict order and its indegree flag are partition invariant.

    rank r:  reads[lo_r:hi_r] --dbg_build(CHUNKS_ONLY)--> the distinct super-k-mers (chunks) of the local reads
             --dbg_chunk_export_records(P)--> chunk records grouped by owner = f(minimizer)
             ==all-to-all==> chunks of the minimizers this rank owns --merge, dbg_chunk_expand--> node shard
             (DBG_B200_DIST=records: the older exchange of NODE records from a full local node table)
             ==all-gather==> every shard --merge--> the full node table, on every rank
             graph stage 0 (degrees / prune / branch flags)
             graph stage 1 (tie-order scan of the LOCAL reads) ==all-gather order records==> merge
             graph stage 2 (order apply, lists, tip pruning, pull-out flags of the LOCAL reads)
             ==all-gather read flags==>

Every rank ends with the same complete graph handle, so output_contigs / contig descriptors
are available everywhere without further communication.  The only collectives are on the data
path's one real exchange (records) plus the small gathers; nothing is reduced on the host.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .graph import Graph, pack_reads


def _device_of(graph, group):
    backend = dist.get_backend(group)
    if backend == 'nccl':
        return torch.device('cuda', graph.device)
    return torch.device('cpu')


def _all_gather_var(t, group):
    """All-gather of int64 tensors [n_i, words] with different n_i; returns the concatenation."""
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    ns = [int(x.item()) for x in ns]
    m = max(ns)
    words = t.shape[1]
    pad = torch.zeros((m, words), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:c] for o, c in zip(outs, ns)], dim=0).contiguous(), ns


def shard_bounds(n_reads, world):
    """Contiguous, balanced read ranges: rank r gets [b[r], b[r+1])."""
    return [n_reads * r // world for r in range(world + 1)]


def build_distributed(graph: Graph, local_packed, read_idx_base, threshold=3, final=False, exact_order=True,
                      pull_reads=True, group=None, timings=None, local_device=None):
    """Run the distributed construct_graph on `graph` (one per rank).  `local_packed` =
    (uint8 bases, uint64 offsets) of this rank's reads in HOST memory, or pass
    `local_device` = (d_bases_ptr, d_off_ptr, n_reads, n_bases) for reads already in HBM;
    `read_idx_base` is the global index of the rank's first read.  Returns the LOCAL pull-out
    flags (uint8 per local read)."""
    world = dist.get_world_size(group)
    dev = _device_of(graph, group)
    cuda = dev.type == 'cuda'

    def tick(name, t0):
        if timings is not None:
            if cuda:
                torch.cuda.synchronize(dev)
            import time
            timings[name] = timings.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
            return time.perf_counter()
        return t0

    def fence():
        """The library runs on its own stream: make the host wait for torch's collectives before
        the library reads what they wrote (the library itself returns only after its stream is idle)."""
        if cuda:
            torch.cuda.current_stream(dev).synchronize()

    import os
    import time
    t0 = time.perf_counter()
    if os.environ.get('DBG_B200_DIST', 'chunks') == 'chunks':
        # 1. phase 1 of the chunk build on the local reads: their distinct super-k-mers, no node table
        if local_device is not None:
            graph.build_chunks(device_ptrs=local_device, read_idx_base=read_idx_base)
        else:
            graph.build_chunks(packed=local_packed, read_idx_base=read_idx_base)
        t0 = tick('local_table_ms', t0)
        cw = graph.chunk_record_words()
        # 2. chunk records grouped by owner = f(minimizer), 3. all-to-all
        counts = graph.chunk_export_counts(world)
        send = torch.empty((int(counts.sum()), cw), dtype=torch.int64, device=dev)
        graph.chunk_export_records(world, counts, send.data_ptr())
        send_counts = torch.tensor(counts.astype(np.int64), dtype=torch.int64, device=dev)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=group)
        rc = [int(x) for x in recv_counts.tolist()]
        recv = torch.empty((sum(rc), cw), dtype=torch.int64, device=dev)
        dist.all_to_all_single(recv, send, output_split_sizes=rc, input_split_sizes=[int(c) for c in counts],
                               group=group)
        fence()
        t0 = tick('exchange_ms', t0)
        # 4. merge the chunks this rank owns and expand them: its shard of the node table
        graph.chunk_table_reset(max(1, recv.shape[0]))
        graph.chunk_merge_records(recv.data_ptr(), recv.shape[0])
        graph.chunk_expand()
    else:
        # 1. local window loop -> local node table
        if local_device is not None:
            graph.build_table_device(*local_device, read_idx_base=read_idx_base)
        else:
            graph.build_table(packed=local_packed, read_idx_base=read_idx_base)
        t0 = tick('local_table_ms', t0)
        rw = graph.record_words()
        # 2. node records grouped by owner, 3. all-to-all
        counts = graph.export_counts(world)
        send = torch.empty((int(counts.sum()), rw), dtype=torch.int64, device=dev)
        graph.export_records(world, counts, send.data_ptr())
        send_counts = torch.tensor(counts.astype(np.int64), dtype=torch.int64, device=dev)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=group)
        rc = [int(x) for x in recv_counts.tolist()]
        recv = torch.empty((sum(rc), rw), dtype=torch.int64, device=dev)
        dist.all_to_all_single(recv, send, output_split_sizes=rc, input_split_sizes=[int(c) for c in counts],
                               group=group)
        fence()
        t0 = tick('exchange_ms', t0)
        # 4. merge what this rank owns
        graph.table_reset(max(1, recv.shape[0]))
        graph.merge_records(recv.data_ptr(), recv.shape[0])
    rw = graph.record_words()
    c1 = graph.export_counts(1)
    shard = torch.empty((int(c1[0]), rw), dtype=torch.int64, device=dev)
    graph.export_records(1, c1, shard.data_ptr())
    t0 = tick('shard_merge_ms', t0)
    # 5. every shard everywhere -> the full table (one all-gather into one buffer, merged slice by slice)
    n_mine = torch.tensor([shard.shape[0]], dtype=torch.int64, device=dev)
    ns_t = [torch.zeros_like(n_mine) for _ in range(world)]
    dist.all_gather(ns_t, n_mine, group=group)
    ns = [int(x.item()) for x in ns_t]
    m = max(max(ns), 1)
    if shard.shape[0] < m:
        pad = torch.empty((m, rw), dtype=torch.int64, device=dev)
        pad[:shard.shape[0]] = shard
    else:
        pad = shard
    everything = torch.empty((world * m, rw), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(everything, pad, group=group)
    fence()
    t0 = tick('gather_ms', t0)
    graph.table_reset(max(1, sum(ns)))
    for r in range(world):
        if ns[r]:
            graph.merge_records(everything[r * m:r * m + ns[r]].data_ptr(), ns[r])
    del send, recv, shard, pad, everything
    t0 = tick('full_merge_ms', t0)
    # 6. graph stages; the tie-order scan runs over the local reads and is merged by key
    graph.graph_stage(0, threshold, final, exact_order, pull_reads)
    st = graph.stats()
    if exact_order and st.n_multi:
        graph.graph_stage(1, threshold, final, exact_order, pull_reads)
        ow = graph.order_words()
        cap = int(st.n_multi) + 8
        mine = torch.empty((cap, ow), dtype=torch.int64, device=dev)
        n = graph.order_export(mine.data_ptr(), cap)
        allrec, _ = _all_gather_var(mine[:n].contiguous(), group)
        fence()
        graph.order_merge(allrec.data_ptr(), allrec.shape[0])
    graph.graph_stage(2, threshold, final, exact_order, pull_reads)
    t0 = tick('graph_ms', t0)
    return graph.pull_read_flags()


def construct_graph(reads, k, threshold=3, final=False, group=None, device=None):
    """Distributed twin of debruijn.construct_graph: call it on every rank with the SAME read
    list; returns the same tuple on every rank."""
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
