// dbg_shard.cuh -- owner-computes graph stages and contig walk over a SHARDED node table.
//
// Multi-GPU layout: after the chunk exchange every rank holds the shard of the node
// table whose MINIMIZERS it owns (dbg_chunks.cuh: chunk_owner).  A node's owner is a function of the
// node alone (its minimizer m-mer), so any rank can tell where a successor lives without asking.
// Degrees / pruningEdges / branch flags (debruijn.py:129-166, 230-236) are per-node work on the shard.
// The contig walk (debruijn.py:288-347) is a list ranking whose chains cross ranks only where the
// minimizer changes (about every 8th node):
//
//   exit     chain node (one kept successor, not pulled) whose successor is on another rank
//   entry    node that is the successor of another rank's exit
//   local    the shard is ranked alone with exits as roots and entries + starts as splitters (the
//            single-GPU kernels, unchanged); every entry and every start then knows the exit or real
//            terminal its chain reaches inside the shard and the hop count to it
//   reduced  list = one vertex per entry / start of EVERY rank: (next reduced vertex, hops) or
//            (terminal node, kind, hops).  It is ~1/8 of the nodes, all-gathered (16 B a vertex) and
//            ranked on every rank for that rank's own starts; one contig record per start.
//
// What crosses NVLink: the successor keys of the exits (to the owner), 4-byte ordinals back, and the
// reduced list.  No node record leaves its owner.
#pragma once
#include "dbg_kernels.cuh"
#include "dbg_chunks.cuh"

// k-mer key -> the same k bases left-aligned in two 2-bit words (the layout of a chunk string);
// bits above the k-mer (batch id) are dropped
template <int W> DBG_D void key_left_align(typename KeyOf<W>::type key, int k, u64 *cw);
template <> DBG_D void key_left_align<1>(u64 key, int k, u64 *cw) {
    cw[0] = (key & key_mask<1>(k)) << (64 - 2 * k);                     // k <= 31: shift >= 2
    cw[1] = 0;
}
template <> DBG_D void key_left_align<2>(K128 key, int k, u64 *cw) {
    const K128 m = key_mask<2>(k);
    const u64 lo = key.lo & m.lo, hi = key.hi & m.hi;
    const int sh = 128 - 2 * k;                                          // 2 .. 126
    if (sh >= 64) { cw[0] = sh == 64 ? lo : lo << (sh - 64); cw[1] = 0; }
    else { cw[0] = (hi << sh) | (lo >> (64 - sh)); cw[1] = lo << sh; }
}

// owner of a node = owner of the chunks it occurs in (same minimizer, same hash, leftmost on ties)
template <int W> DBG_D u32 node_owner(typename KeyOf<W>::type key, int k, int w, int m, u32 n_parts) {
    u64 cw[2];
    key_left_align<W>(key, k, cw);
    return chunk_owner<W>(cw, w, m, 0u, n_parts);
}

DBG_D bool is_exit(u32 f, u32 kept_word) {               // chain node whose successor lives elsewhere
    return (f & F_OCC) && (f & F_EXIT) && !(f & F_PULLED) && kept_count(kept_word) == 1;
}

// ---- exits: (slot, successor key) grouped by the successor's owner -------------------------------
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_exit_count(const Slot<W> *__restrict__ table, u64 n_slots,
                                                          const u8 *__restrict__ flags, const u16 *__restrict__ kept,
                                                          int k, int w, int m, u32 n_parts, u64 *counts) {
    typedef typename KeyOf<W>::type K;
    __shared__ u32 s_cnt[DBG_MAX_PARTS];
    for (u32 p = threadIdx.x; p < n_parts; p += blockDim.x) s_cnt[p] = 0;
    __syncthreads();
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_slots && is_exit(flags[i], kept[i])) {
        const K sk = key_roll(table[i].key, (u32)kept_base(kept[i], 0), key_mask<W>(k));
        atomicAdd(&s_cnt[node_owner<W>(sk, k, w, m, n_parts)], 1u);
    }
    __syncthreads();
    for (u32 p = threadIdx.x; p < n_parts; p += blockDim.x)
        if (s_cnt[p]) atomicAdd(&counts[p], (u64)s_cnt[p]);
}

template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_exit_scatter(const Slot<W> *__restrict__ table, u64 n_slots,
                                                            const u8 *__restrict__ flags, const u16 *__restrict__ kept,
                                                            int k, int w, int m, u32 n_parts, u64 *cursors,
                                                            u32 *__restrict__ exit_slot, u64 *__restrict__ exit_keys) {
    typedef typename KeyOf<W>::type K;
    __shared__ u32 s_cnt[DBG_MAX_PARTS];
    __shared__ u64 s_base[DBG_MAX_PARTS];
    for (u32 p = threadIdx.x; p < n_parts; p += blockDim.x) s_cnt[p] = 0;
    __syncthreads();
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const bool have = i < n_slots && is_exit(flags[i], kept[i]);
    K sk = zero_key<W>();
    u32 owner = 0, local = 0;
    if (have) {
        sk = key_roll(table[i].key, (u32)kept_base(kept[i], 0), key_mask<W>(k));
        owner = node_owner<W>(sk, k, w, m, n_parts);
        local = atomicAdd(&s_cnt[owner], 1u);
    }
    __syncthreads();
    for (u32 p = threadIdx.x; p < n_parts; p += blockDim.x)
        if (s_cnt[p]) s_base[p] = atomicAdd(&cursors[p], (u64)s_cnt[p]);
    __syncthreads();
    if (have) {
        const u64 o = s_base[owner] + local;
        exit_slot[o] = (u32)i;
        key_to_words(sk, exit_keys + o * W);
    }
}

// ---- entries: received successor keys -> slots of this shard, flagged ----------------------------
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_mark_entries(const Slot<W> *__restrict__ table, TGeom tg,
                                                            const u64 *__restrict__ keys, u64 n, u8 *flags,
                                                            u32 *__restrict__ req_slot, DevCounters *ctr) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 idx = table_find<W>(table, tg, key_from_words<W>(keys + i * W));
    if (idx == NONE64) { raise_fault(ctr, FAULT_SUCC_MISSING, (u32)i); req_slot[i] = NONE32; return; }
    req_slot[i] = (u32)idx;
    // the flag bytes of four slots share a word; several exits may point at one entry
    atomicOr(reinterpret_cast<u32 *>(flags + (idx & ~3ull)), (u32)F_ENTRY << (8 * (u32)(idx & 3)));
}

// reduced vertices of this shard: every entry and every start, compacted; ord[slot] = position
__global__ void __launch_bounds__(DBG_BLOCK) k_rv_collect(const u8 *__restrict__ flags, u64 n_slots,
                                                          u32 *__restrict__ rv_slot, u32 *__restrict__ rv_ord,
                                                          u64 *cursor) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const bool hit = i < n_slots && (flags[i] & F_OCC) && (flags[i] & (F_ENTRY | F_START));
    __shared__ u32 s_cnt;
    __shared__ u64 s_base;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const u32 lane = threadIdx.x & 31;
    const u32 b = __ballot_sync(0xFFFFFFFFu, hit);
    u32 w_off = 0;
    if (lane == 0 && b) w_off = atomicAdd(&s_cnt, (u32)__popc(b));
    w_off = __shfl_sync(0xFFFFFFFFu, w_off, 0);
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) s_base = atomicAdd(cursor, (u64)s_cnt);
    __syncthreads();
    if (hit) {
        const u64 o = s_base + w_off + __popc(b & ((1u << lane) - 1));
        rv_slot[o] = (u32)i;
        rv_ord[i] = (u32)o;
    }
}

__global__ void __launch_bounds__(DBG_BLOCK) k_entry_ordinals(const u32 *__restrict__ req_slot, u64 n,
                                                              const u32 *__restrict__ rv_ord, u32 *__restrict__ out) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = req_slot[i] == NONE32 ? NONE32 : rv_ord[req_slot[i]];
}

// the owner's answers come back in exit order: succ[exit] = global reduced id of the remote entry
__global__ void __launch_bounds__(DBG_BLOCK) k_exit_store(const u32 *__restrict__ exit_slot, u64 n_exit,
                                                          const u64 *__restrict__ owner_off, u32 n_parts,
                                                          const u64 *__restrict__ rid_base, const u32 *__restrict__ replies,
                                                          u32 *succ) {
    const u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_exit) return;
    u32 lo = 0, hi = n_parts - 1;                            // last owner with owner_off[owner] <= e
    while (lo < hi) {
        const u32 mid = (lo + hi + 1) >> 1;
        if (owner_off[mid] <= e) lo = mid; else hi = mid - 1;
    }
    succ[exit_slot[e]] = (u32)(rid_base[lo] + replies[e]);
}

// ---- reduced list -------------------------------------------------------------------------------
// record = 2 words: w0 = next reduced id (NONE32: the chain ends inside this shard) | hops << 32
//                   w1 = terminal slot | (terminal rank << 8 | kind) << 32        (kind T_*: where it ends)
DBG_HD u64 rrec_w0(u32 next, u32 dist) { return (u64)next | ((u64)dist << 32); }
DBG_HD u64 rrec_w1(u32 slot, u32 rank, u32 kind) { return (u64)slot | ((u64)((rank << 8) | kind) << 32); }

__global__ void __launch_bounds__(DBG_BLOCK) k_reduced_build(const u32 *__restrict__ rv_slot, u64 n_rv,
                                                             const u64 *__restrict__ link, const u8 *__restrict__ flags,
                                                             const u16 *__restrict__ kept, const u32 *__restrict__ succ,
                                                             u32 my_rank, u64 *__restrict__ out) {
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_rv) return;
    const u32 v = rv_slot[j];
    u32 t = v, d = 0;
    u64 w0, w1;
    if (!resolve_chain(link, v, &t, &d)) { w0 = rrec_w0(NONE32, 0); w1 = rrec_w1(v, my_rank, T_CYCLE); }
    else {
        const u32 f = flags[t];
        if (is_exit(f, kept[t])) { w0 = rrec_w0(succ[t], link_dist_add(d, 1)); w1 = rrec_w1(t, my_rank, T_CYCLE); }
        else {
            const u32 kind = (f & F_PULLED) ? T_PULLED : (f & F_DEADEND) ? T_DEADEND : T_BRANCH;
            w0 = rrec_w0(NONE32, d);
            w1 = rrec_w1(t, my_rank, kind);
        }
    }
    out[2 * j] = w0;
    out[2 * j + 1] = w1;
}

// links over the gathered reduced list.  Splitters: this rank's own vertices (its starts are among them)
// plus the usual hash sample; roots: vertices whose chain ends inside their shard.
__global__ void __launch_bounds__(DBG_BLOCK) k_rlink_init(const u64 *__restrict__ recs, u64 n_total, u64 my_base,
                                                          u64 n_mine, u64 *link, u32 *__restrict__ split_list,
                                                          u64 split_cap, u64 *cursor, u32 *__restrict__ super_list,
                                                          u64 super_cap, u64 *super_cursor, DevCounters *ctr) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool split = false;
    if (i < n_total) {
        const u64 w0 = recs[2 * i];
        const u32 next = (u32)w0, dist = (u32)(w0 >> 32);
        const bool chain = next != NONE32;
        if (chain && next >= n_total) raise_fault(ctr, FAULT_JUMP_RANGE, (u32)i);
        split = chain && ((i >= my_base && i < my_base + n_mine) || (hash32((u32)i) & ((1u << WALK_SPLIT_LOG2) - 1)) == 0);
        link[i] = chain ? (link_pack(next, dist > LINK_DIST_MAX ? LINK_DIST_MAX : dist) | (split ? LINK_SPLIT : 0))
                        : link_pack((u32)i, 0);
    }
    __shared__ u32 s_cnt[2];
    __shared__ u64 s_base[2];
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const bool sup = split && super_hash((u32)i);
    const u32 lane = threadIdx.x & 31;
    const u32 b = __ballot_sync(0xFFFFFFFFu, split), bs = __ballot_sync(0xFFFFFFFFu, sup);
    u32 w_off = 0, ws_off = 0;
    if (lane == 0) {
        if (b) w_off = atomicAdd(&s_cnt[0], (u32)__popc(b));
        if (bs) ws_off = atomicAdd(&s_cnt[1], (u32)__popc(bs));
    }
    w_off = __shfl_sync(0xFFFFFFFFu, w_off, 0);
    ws_off = __shfl_sync(0xFFFFFFFFu, ws_off, 0);
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt[0]) s_base[0] = atomicAdd(cursor, (u64)s_cnt[0]);
    if (threadIdx.x == 1 && s_cnt[1]) s_base[1] = atomicAdd(super_cursor, (u64)s_cnt[1]);
    __syncthreads();
    if (split) {
        const u64 o = s_base[0] + w_off + __popc(b & ((1u << lane) - 1));
        if (o < split_cap) split_list[o] = (u32)i;
    }
    if (sup) {
        const u64 o = s_base[1] + ws_off + __popc(bs & ((1u << lane) - 1));
        if (o < super_cap) super_list[o] = (u32)i;
    }
}

struct ShardContig {            // same layout as dbg_shard_contig_t (include/dbg_b200.h)
    u64 tag;                    // first-occurrence tag of the start: sorting all ranks' records by it gives the reference's order
    u32 start_slot, start_rank;
    u32 terminal_slot, terminal_rank;
    u32 n_nodes, kind;
};

template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_emit_reduced(const Slot<W> *__restrict__ table,
                                                            const u32 *__restrict__ starts, u64 n_starts,
                                                            const u32 *__restrict__ rv_ord, u64 my_base, u32 my_rank,
                                                            const u64 *__restrict__ link, const u64 *__restrict__ recs,
                                                            ShardContig *__restrict__ out, u64 *n_dropped) {
    // one record per start, in start order (the start list is sorted by tag); n_nodes == 0 = nothing emitted
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool dropped = false;
    if (i < n_starts) {
        const u32 s = starts[i];
        const u32 r0 = (u32)(my_base + rv_ord[s]);
        u32 R = r0, h = 0;
        ShardContig c;
        c.tag = table[s].tag;
        c.start_slot = s; c.start_rank = my_rank;
        c.terminal_slot = NONE32; c.terminal_rank = 0; c.n_nodes = 0; c.kind = T_CYCLE;
        if (resolve_chain(link, r0, &R, &h)) {
            const u64 w0 = recs[2 * (u64)R], w1 = recs[2 * (u64)R + 1];
            const u32 info = (u32)(w1 >> 32), kind = info & 0xFFu;
            if (kind != T_CYCLE) {
                const u32 total = link_dist_add(h, (u32)(w0 >> 32));
                c.terminal_slot = (u32)w1; c.terminal_rank = info >> 8;
                c.kind = kind;
                c.n_nodes = kind == T_PULLED ? total : total + 1;
            }
        }
        dropped = c.n_nodes == 0;
        out[i] = c;
    }
    const u32 b = __ballot_sync(0xFFFFFFFFu, dropped);
    if (b && (threadIdx.x & 31) == 0) atomicAdd(n_dropped, (u64)__popc(b));
}
