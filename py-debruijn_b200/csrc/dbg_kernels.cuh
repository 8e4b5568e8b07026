// dbg_kernels.cuh -- device kernels of the assembly hot path (sm_100a).
//
// Stage map (reference file:line -> kernel):
//   read packing                       (new: 2-bit stream)          k_pack
//   window loop   debruijn.py:123-144  (nodes, successor counts)    dbg_chunks.cuh: k_masks, k_chunk_insert, k_expand
//   degrees/prune debruijn.py:129-143, 150-166, 230-236             k_finalize
//   tie order     debruijn.py:159-164  (Counter.most_common)        k_side_fill, k_order_chunks, k_order_apply
//   tip pruning   debruijn.py:169-186  (candidate paths)            k_tip_paths
//   pull-out      debruijn.py:274-278                                k_pull_reads
//   contig walk   debruijn.py:288-316, 326-347                      k_link_init, k_sublist, k_jump, k_emit
//   strings       debruijn.py:298-300                                k_materialize
//   getScore      II_assembleFromReads.py:14-18                      k_scores, k_lookup_edges
#pragma once
#include "dbg_common.cuh"

#define DBG_BLOCK 256
#define DBG_WPT 16                         // consecutive windows per thread
#define DBG_TILE (DBG_BLOCK * DBG_WPT)     // base positions per block
#define HLL_P 12
#define HLL_M (1 << HLL_P)

// internal terminal kinds (0..2 are the public DBG_TERM_*)
#define T_DEADEND 0u
#define T_BRANCH 1u
#define T_PULLED 2u
#define T_CYCLE 3u

struct DevCounters {
    u64 n_windows;
    u64 n_nodes;
    u64 n_edges;
    u64 n_multi;
    u64 n_branch;
    u64 n_starts;
    u64 n_list_a;       // append cursors for k_collect
    u64 n_list_b;
    u64 n_dropped;      // k_emit: starts that emit nothing
    u32 bad_alphabet;
    u32 overflow;       // table full
    u32 too_long;       // read longer than 2^32-1
    u32 changed;        // pointer jumping made progress
    u32 fault;          // first internal-consistency failure seen by a kernel (FAULT_* code), 0 = none
    u32 fault_info;
};

// internal-consistency guards: a kernel that would index out of range records a code instead
#define FAULT_SUCC_MISSING 1u     // k_finalize / k_order_apply: a kept successor is not in the table
#define FAULT_TIP_MISSING 2u      // k_tip_paths: ditto
#define FAULT_JUMP_RANGE 3u       // k_jump / k_emit: link points outside the table
#define FAULT_JUNCTION_MISSING 4u // k_junctions
DBG_D void raise_fault(DevCounters *ctr, u32 code, u32 info) {
    if (atomicCAS(&ctr->fault, 0u, code) == 0u) ctr->fault_info = info;
}

// ---------------------------------------------------------------------------------------------
// k_pack: ASCII -> 2-bit, 32 bases per u64 word, 128-bit loads (2 x uint4 per thread).
// A=0 C=1 G=2 T=3 via ((c>>1)&3) ^ ((c>>2)&1); every byte is checked against "ACGT"[code].
// ---------------------------------------------------------------------------------------------
DBG_D u32 pack4(u32 x, bool &bad) {
    u32 t = ((x >> 1) & 0x03030303u) ^ ((x >> 2) & 0x01010101u);
    u32 sel = (t & 3u) | (((t >> 8) & 3u) << 4) | (((t >> 16) & 3u) << 8) | (((t >> 24) & 3u) << 12);
    bad |= (__byte_perm(0x54474341u, 0u, sel) != x);
    return ((t & 3u) << 6) | (((t >> 8) & 3u) << 4) | (((t >> 16) & 3u) << 2) | ((t >> 24) & 3u);
}
DBG_D u32 pack16(uint4 v, bool &bad) {
    return (pack4(v.x, bad) << 24) | (pack4(v.y, bad) << 16) | (pack4(v.z, bad) << 8) | pack4(v.w, bad);
}

__global__ void __launch_bounds__(DBG_BLOCK) k_pack(const u8 *__restrict__ bases, u64 n_bases,
                                                    u64 *__restrict__ packed, u64 w_begin, u64 w_end,
                                                    int aligned16, DevCounters *ctr) {
    u64 w = w_begin + (u64)blockIdx.x * blockDim.x + threadIdx.x;   // words [w_begin, w_end): one slice of the stream
    if (w >= w_end) return;
    u64 b0 = w * 32;
    u64 word = 0;
    bool bad = false;
    if (aligned16 && b0 + 32 <= n_bases) {
        const uint4 *src = reinterpret_cast<const uint4 *>(bases + b0);
        uint4 a = __ldg(src), b = __ldg(src + 1);
        word = ((u64)pack16(a, bad) << 32) | pack16(b, bad);
    } else {
        for (int j = 0; j < 32; j++) {
            u64 i = b0 + j;
            u32 code = 0;
            if (i < n_bases) {
                u32 c = bases[i];
                code = ((c >> 1) & 3u) ^ ((c >> 2) & 1u);
                bad |= (c != ((0x54474341u >> (8 * code)) & 0xFFu));
            }
            word = (word << 2) | code;
        }
    }
    packed[w] = word;
    if (bad) ctr->bad_alphabet = 1;
}

// ---------------------------------------------------------------------------------------------
// k_read_stats: n_windows = sum(max(0, len - k)) and the 32-bit position guard.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DBG_BLOCK) k_read_stats(const u64 *__restrict__ off, u64 n_reads, int k,
                                                          DevCounters *ctr) {
    // grid-stride, one atomic per block (one per warp on the same word serialises in L2)
    u64 win = 0;
    bool too_long = false;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += (u64)gridDim.x * blockDim.x) {
        u64 len = __ldg(off + r + 1) - __ldg(off + r);
        if (len > (u64)k) win += len - k;
        too_long |= len > 0xFFFFFFFFULL;
    }
    for (int d = 16; d > 0; d >>= 1) win += __shfl_down_sync(0xFFFFFFFFu, win, d);
    __shared__ u64 s_win;
    if (threadIdx.x == 0) s_win = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && win) atomicAdd(&s_win, win);
    __syncthreads();
    if (threadIdx.x == 0 && s_win) atomicAdd(&ctr->n_windows, s_win);
    if (too_long) ctr->too_long = 1;
}

// ---------------------------------------------------------------------------------------------
// k_tile_first: tile t covers base positions [t*TILE, (t+1)*TILE); first[t] = index of the read
// holding its first position (last r with off[r] <= p).  n_tiles + 1 entries.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DBG_BLOCK) k_tile_first(const u64 *__restrict__ off, u64 n_reads, u64 n_bases,
                                                          u64 n_tiles, u32 *__restrict__ first) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    u64 p = t * DBG_TILE;
    if (p > n_bases - 1) p = n_bases - 1;
    u64 lo = 0, hi = n_reads - 1;
    while (lo < hi) {
        u64 mid = (lo + hi + 1) >> 1;
        if (__ldg(off + mid) <= p) lo = mid; else hi = mid - 1;
    }
    first[t] = (u32)lo;
}

// ---------------------------------------------------------------------------------------------
// Window iteration.  Each thread owns DBG_WPT consecutive base positions of the concatenated
// read stream and rolls the key along them (key = (key << 2 | next) & mask).  For every
// position that is a node occurrence of its read it calls f(key, read, pos, has_next, next):
//   BUILD  : reads longer than k, pos <= len - k            (debruijn.py:126-128)
//   !BUILD : reads of length >= k (substring test of pull_out_read, debruijn.py:274-278)
// has_next is false only for the last k-mer of a read (a node occurrence without an edge).
// ---------------------------------------------------------------------------------------------
template <int W, bool BUILD, typename F>
DBG_D void for_windows(const u64 *__restrict__ packed, const u64 *__restrict__ off,
                       const u32 *__restrict__ tile_first, u64 n_bases, int k, F f) {
    typedef typename KeyOf<W>::type K;
    const u64 tile = blockIdx.x;
    const u64 p0 = tile * DBG_TILE + (u64)threadIdx.x * DBG_WPT;
    if (p0 >= n_bases) return;
    u32 rlo = __ldg(tile_first + tile), rhi = __ldg(tile_first + tile + 1);
    while (rlo < rhi) {
        u32 mid = rlo + ((rhi - rlo + 1) >> 1);
        if (__ldg(off + mid) <= p0) rlo = mid; else rhi = mid - 1;
    }
    u32 r = rlo;
    u64 rs = __ldg(off + r), re = __ldg(off + r + 1);
    const K mask = key_mask<W>(k);
    K key = extract_key<W>(packed, p0, k);
    const u32 nbw = extract_bases(packed, p0 + k, DBG_WPT);      // the DBG_WPT bases that follow the first k-mer
    // everything below is 32-bit arithmetic relative to p0
    const int n_here = (int)(n_bases - p0 < (u64)DBG_WPT ? n_bases - p0 : (u64)DBG_WPT);
    int end = (int)(re - p0 < (u64)(1 << 30) ? re - p0 : (u64)(1 << 30));      // read end, relative
    u32 pos0 = (u32)(p0 - rs);                                   // position in read of relative 0 (wraps if the
    bool long_enough = re - rs > (u64)k;                         // read starts later: fixed when it does)
#pragma unroll
    for (int j = 0; j < DBG_WPT; j++) {
        if (j >= n_here) break;
        while (j >= end) {                                       // next read (also skips empty reads)
            r++;
            const u64 ne = __ldg(off + r + 1);
            long_enough = ne - re > (u64)k;
            pos0 = (u32)0 - (u32)end;                            // read starts at relative `end`
            end = (int)(ne - p0 < (u64)(1 << 30) ? ne - p0 : (u64)(1 << 30));
            re = ne;
        }
        const u32 nb = (nbw >> (2 * (DBG_WPT - 1 - j))) & 3u;
        const bool ok = BUILD ? (long_enough && j + k <= end) : (j + k <= end);
        if (ok) f(key, r, pos0 + (u32)j, j + k < end, nb);
        key = key_roll(key, nb, mask);
    }
}

// ---------------------------------------------------------------------------------------------
// table init
// ---------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_table_init(Slot<W> *table, u64 n_slots) {
    // grid-stride: a small grid on a side stream clears the table in the background of the packing kernels
    // (a full grid would take every block slot of every SM and serialise with them)
    Slot<W> s;
    s.key = empty_key<W>();
    s.tag = ~0ULL;
    s.cnt[0] = s.cnt[1] = s.cnt[2] = s.cnt[3] = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (u64)gridDim.x * blockDim.x) table[i] = s;
}

#define DBG_WARPS (DBG_BLOCK / 32)
#define DBG_UNIT 1024                     // stream positions per warp work unit (k_chunk_insert)

// ---------------------------------------------------------------------------------------------
// k_finalize: per slot, degrees + pruningEdges + branch flag + the single surviving successor.
// A warp sweeps FIN_SLOTS slots per step and hands the OCCUPIED ones to its lanes densely (half the slots
// are empty: one slot per lane left half the lanes idle through the whole body).  The successor is looked
// for in the node's OWN region first -- it has the same minimizer, hence the same region, unless the
// minimizer changes between the two (1 node in ~8) or this node was pushed over a region border; a key
// that is found is the key wherever it is found -- and only what that misses goes to a list that
// k_finalize_far resolves through the successor's own minimizer (16 hashed m-mers a node).
// ---------------------------------------------------------------------------------------------
#define FIN_SLOTS 56
#define FIN_NEAR_PROBES 6
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_finalize(const Slot<W> *__restrict__ table, TGeom tg, int k,
                                                        double threshold, u8 *__restrict__ flags,
                                                        u16 *__restrict__ kept, u32 *__restrict__ succ,
                                                        int sharded, u32 *__restrict__ far_list, DevCounters *ctr) {
    typedef typename KeyOf<W>::type K;
    __shared__ u32 s_far[DBG_WARPS][64];                                 // far list, staged per warp: one global atomic per 32 entries
    u32 n_far = 0;                                                       // (warp-uniform)
    const u32 wib = threadIdx.x >> 5;
    // the five counters are summed in registers and leave the block as ONE atomic each
    // (one atomic per warp on the same five words used to serialise in L2)
    u32 c_nodes = 0, c_edges = 0, c_starts = 0, c_branch = 0, c_multi = 0;
    const u64 n_slots = tg.n_slots;
    const K mask = key_mask<W>(k);
    const u32 lane = threadIdx.x & 31;
    const u64 warp0 = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 base = warp0 * FIN_SLOTS; base < n_slots; base += n_warps * FIN_SLOTS) {
        const u64 i0 = base + lane, i1 = base + 32 + lane;
        const bool in0 = i0 < n_slots, in1 = lane < FIN_SLOTS - 32 && i1 < n_slots;
        // (a key whose tag is still ~0 is no node: min_count left no window that touches it)
        u64 tg0 = 0, tg1 = 0;
        const bool o0 = in0 && load_key_tag(&table[i0], &tg0) != empty_key<W>() && tg0 != ~0ULL;
        const bool o1 = in1 && load_key_tag(&table[i1], &tg1) != empty_key<W>() && tg1 != ~0ULL;
        if (in0 && !o0) { flags[i0] = 0; kept[i0] = 0; succ[i0] = NONE32; }
        if (in1 && !o1) { flags[i1] = 0; kept[i1] = 0; succ[i1] = NONE32; }
        const u32 occ0 = __ballot_sync(0xFFFFFFFFu, o0), occ1 = __ballot_sync(0xFFFFFFFFu, o1);
        const u32 n0 = (u32)__popc(occ0), total = n0 + (u32)__popc(occ1);
        for (u32 gb = 0; gb < total; gb += 32) {
          const u32 g = gb + lane;
          bool far = false;
          u64 i = 0;
          if (g < total) {
            i = g < n0 ? base + nth_set_bit(occ0, g) : base + 32 + nth_set_bit(occ1, g - n0);
            u64 tag;
            const K key = load_key_tag(&table[i], &tag);
            const uint4 c4 = __ldcg(reinterpret_cast<const uint4 *>(&table[i].cnt[0]));
            const u32 cnt[4] = {c4.x, c4.y, c4.z, c4.w};
            const u32 word = prune_successors(cnt, (const u64 *)0, threshold);
            const int nk = kept_count(word), m = distinct_count(word);
            u32 f = F_OCC | ((u32)tag == 0 ? F_START : 0) | (nk > 1 ? F_BRANCH : 0) |
                    (m == 0 ? F_DEADEND : 0) | (m > 1 ? F_MULTI : 0);
            u32 sx = NONE32;
            if (nk == 1) {
                const K sk = key_roll(key, (u32)kept_base(word, 0), mask);
                const u32 hs = hash32_key(sk);
                bool found = false, settled = false;
                u64 idx;
                if (tg.n_regions > 1) idx = (i / tg.region_slots) * tg.region_slots + (((u64)hs * tg.region_slots) >> 32);
                else idx = slot_home(hs, n_slots);
                const u32 limit = tg.n_regions > 1 ? FIN_NEAR_PROBES : (u32)n_slots;
                for (u32 p = 0; p < limit; p++) {
                    const K cur = load_key(&table[idx]);
                    if (cur == sk) { found = true; break; }
                    if (cur == empty_key<W>()) { settled = tg.n_regions <= 1; break; }   // (own region: the real home may be elsewhere)
                    if (++idx == n_slots) idx = 0;
                }
                if (found) sx = (u32)idx;
                else if (tg.n_regions > 1 && !settled) far = true;       // not near: resolve through the successor's own minimizer
                else {                                                 // plain layout: the probe sequence ended at EMPTY (or the table is full)
                    if (sharded) f |= F_EXIT; else raise_fault(ctr, FAULT_SUCC_MISSING, (u32)i);
                }
            }
            c_nodes++; c_edges += (u32)m;
            c_starts += (f & F_START) ? 1u : 0u; c_branch += nk > 1 ? 1u : 0u; c_multi += m > 1 ? 1u : 0u;
            flags[i] = (u8)f;
            kept[i] = (u16)word;
            succ[i] = sx;
          }
          const u32 fm = __ballot_sync(0xFFFFFFFFu, far);
          if (fm) {
              if (far) s_far[wib][n_far + __popc(fm & ((1u << lane) - 1))] = (u32)i;
              n_far += (u32)__popc(fm);
              __syncwarp();
              if (n_far >= 32) {
                  u64 o = 0;
                  if (lane == 0) o = atomicAdd(&ctr->n_list_a, 32ULL);
                  o = __shfl_sync(0xFFFFFFFFu, o, 0);
                  far_list[o + lane] = s_far[wib][lane];
                  const u32 rest = s_far[wib][32 + lane];
                  __syncwarp();
                  s_far[wib][lane] = rest;
                  n_far -= 32;
                  __syncwarp();
              }
          }
        }
    }
    if (n_far) {
        u64 o = 0;
        if (lane == 0) o = atomicAdd(&ctr->n_list_a, (u64)n_far);
        o = __shfl_sync(0xFFFFFFFFu, o, 0);
        if (lane < n_far) far_list[o + lane] = s_far[wib][lane];
    }
    __shared__ u32 s_sum[5];
    if (threadIdx.x < 5) s_sum[threadIdx.x] = 0;
    __syncthreads();
    u32 v[5] = {c_nodes, c_edges, c_starts, c_branch, c_multi};
#pragma unroll
    for (int q = 0; q < 5; q++) {
        u32 x = v[q];
        for (int d = 16; d > 0; d >>= 1) x += __shfl_down_sync(0xFFFFFFFFu, x, d);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(&s_sum[q], x);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_sum[0]) atomicAdd(&ctr->n_nodes, (u64)s_sum[0]);
        if (s_sum[1]) atomicAdd(&ctr->n_edges, (u64)s_sum[1]);
        if (s_sum[2]) atomicAdd(&ctr->n_starts, (u64)s_sum[2]);
        if (s_sum[3]) atomicAdd(&ctr->n_branch, (u64)s_sum[3]);
        if (s_sum[4]) atomicAdd(&ctr->n_multi, (u64)s_sum[4]);
    }
}

// the successors k_finalize did not find near their predecessor
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_finalize_far(const Slot<W> *__restrict__ table, TGeom tg, int k,
                                                            const u16 *__restrict__ kept, const u32 *__restrict__ far_list,
                                                            u8 *flags, u32 *__restrict__ succ, int sharded, DevCounters *ctr) {
    typedef typename KeyOf<W>::type K;
    const u64 n = __ldcg(&ctr->n_list_a);
    const K mask = key_mask<W>(k);
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (u64)gridDim.x * blockDim.x) {
        const u32 i = far_list[e];
        const K sk = key_roll(load_key(&table[i]), (u32)kept_base(kept[i], 0), mask);
        const u64 fi = table_find<W>(table, tg, sk);
        if (fi == NONE64) {                                              // sharded table: the successor is another rank's node
            if (sharded) flags[i] = (u8)(flags[i] | F_EXIT); else raise_fault(ctr, FAULT_SUCC_MISSING, i);
        }
        succ[i] = (u32)fi;
    }
}

// ---------------------------------------------------------------------------------------------
// min_count (EXTENSION, no reference line: SURVEY.md section 9 delta 5, BASELINE configs[2]).  The window loop
// skips every window whose (k+1)-mer occurs fewer than min_count times over all reads (oracle/dbg_oracle.py
// build()).  The unfiltered table holds exactly those totals, so the filtered graph is made FROM it:
//   k_mc_filter  zeroes the counters below min_count and forgets every first-occurrence tag
//   k_mc_retag   (dbg_chunks.cuh) walks the DISTINCT chunks: a surviving window at offset j of a chunk tagged
//                (read, pos) is the first-occurrence candidate (read, pos + j) of its node and (read, pos + j + 1)
//                of its successor -- the chunk's tag is the minimum over its instances, which all hold the same
//                windows; nodes no surviving window touches keep tag ~0 and are no nodes (k_finalize skips them)
//   k_mc_starts  a node first touched at position i > 0 was created as a successor (the reference's indegree
//                1) only if window i - 1 of that read survived; otherwise it was created as the source of
//                window i: indegree 0, a contig start
// ---------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_mc_filter(Slot<W> *table, u64 n_slots, u32 min_count) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (u64)gridDim.x * blockDim.x) {
        if (load_key(&table[i]) == empty_key<W>()) continue;
#pragma unroll
        for (int c = 0; c < 4; c++) if (table[i].cnt[c] < min_count) table[i].cnt[c] = 0;
        table[i].tag = ~0ULL;
    }
}

template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_mc_starts(const u64 *__restrict__ packed, const u64 *__restrict__ off,
                                                         const Slot<W> *__restrict__ table, TGeom tg, int k,
                                                         u64 read_idx_base, u8 *flags, DevCounters *ctr) {
    typedef typename KeyOf<W>::type K;
    const K mask = key_mask<W>(k);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < tg.n_slots; i += (u64)gridDim.x * blockDim.x) {
        const u32 f = flags[i];
        if (!(f & F_OCC) || (f & F_START)) continue;
        const K key = load_key(&table[i]);
        const u64 tag = table[i].tag;
        const u64 r = (tag >> 32) - read_idx_base;
        const u32 pos = (u32)tag;                                        // > 0 (pos == 0 is a start already)
        const u32 pb = base_at(packed, __ldg(off + r) + pos - 1);
        // predecessor in that read: previous base + the node's first k - 1 bases (batch bits stay)
        K pk = key;
        {
            // shift right by one base inside the 2k key bits
            const K low = key_and_mask(key, mask);
            pk = key_or(key_andnot_mask(key, mask), key_shr2_put(low, pb, k));
        }
        const u64 sp = table_find<W>(table, tg, pk);
        const bool came_in = sp != NONE64 && __ldcg(&table[sp].cnt[key_last_base(key)]) != 0u;
        if (!came_in) {
            flags[i] = (u8)(f | F_START);
            atomicAdd(&ctr->n_starts, 1ULL);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Side sets.  Two later stages re-scan every window but only care about a small subset of the
// nodes: the exact successor order (multi-successor nodes, debruijn.py:159-164) and pull_out_read
// (branch nodes, debruijn.py:274-278).  Probing the big node table per window would be a random
// DRAM access each; instead the subset goes into a small open-addressing set keyed by the node
// key, fronted by a bitmap over the top bits of the slot hash, and both stay in L2 / L1.
// ---------------------------------------------------------------------------------------------
template <int W> struct SideEntry {
    typename KeyOf<W>::type key;   // all ones = empty
    u64 slot;                      // slot in the node table
    u64 first[4];                  // first-appearance tag per successor base (~0 = never)
};

template <int W>
DBG_D u64 side_locate(SideEntry<W> *st, u64 cap, typename KeyOf<W>::type key, bool insert) {
    typedef typename KeyOf<W>::type K;
    const K EMPTY = empty_key<W>();
    u64 idx = ((u64)(hash32_key(key) * 0x9E3779B1u) * cap) >> 32;        // re-mixed: independent of the bitmap bits
    for (u64 probes = 0; probes < cap; ++probes) {
        K cur = load_key(reinterpret_cast<const Slot<W> *>(&st[idx]));   // key is the first member of both
        if (cur == key) return idx;
        if (cur == EMPTY) {
            if (!insert) return NONE64;
            K old = cas_key(&st[idx].key, EMPTY, key);
            if (old == EMPTY || old == key) return idx;
        }
        if (++idx == cap) idx = 0;
    }
    return NONE64;
}
DBG_D bool bitmap_test(const u32 *__restrict__ bm, int bits, u32 h) {
    const u32 b = h >> (32 - bits);
    return (__ldg(bm + (b >> 5)) >> (b & 31)) & 1u;
}

// every slot whose flags hit `want` goes into the side set
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_side_fill(const Slot<W> *__restrict__ table, const u8 *__restrict__ flags,
                                                         u64 n_slots, u32 want, SideEntry<W> *st, u64 cap, u32 *bm,
                                                         int bits) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots || !(flags[i] & want)) return;
    typename KeyOf<W>::type key = table[i].key;
    u64 e = side_locate<W>(st, cap, key, true);
    if (e != NONE64) st[e].slot = i;
    const u32 b = hash32_key(key) >> (32 - bits);
    atomicOr(&bm[b >> 5], 1u << (b & 31));
}

template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_order_apply(const Slot<W> *__restrict__ table, TGeom tg, int k,
                                                           double threshold, const SideEntry<W> *__restrict__ ot,
                                                           u64 cap, u16 *__restrict__ kept, u32 *__restrict__ succ,
                                                           u8 *flags, int sharded, DevCounters *ctr) {
    typedef typename KeyOf<W>::type K;
    u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= cap) return;
    if (ot[e].key == empty_key<W>()) return;
    u64 i = ot[e].slot;
    u32 cnt[4] = { table[i].cnt[0], table[i].cnt[1], table[i].cnt[2], table[i].cnt[3] };
    u64 first[4] = { ot[e].first[0], ot[e].first[1], ot[e].first[2], ot[e].first[3] };
    u32 word = prune_successors(cnt, first, threshold);
    kept[i] = (u16)word;
    if (kept_count(word) == 1) {        // threshold < 1: the tie decides WHICH successor survives
        K sk = key_roll(table[i].key, (u32)kept_base(word, 0), key_mask<W>(k));
        u64 fi = table_find<W>(table, tg, sk);
        if (fi == NONE64 && !sharded) raise_fault(ctr, FAULT_SUCC_MISSING, (u32)i);
        succ[i] = (u32)fi;
        if (sharded) flags[i] = (u8)((flags[i] & ~F_EXIT) | (fi == NONE64 ? F_EXIT : 0u));   // only this thread writes slot i
    }
}

// ---------------------------------------------------------------------------------------------
// k_collect: append (tag, slot) of every slot whose flags hit `want` (starts / branch nodes).
// ---------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_collect(const Slot<W> *__restrict__ table, u64 n_slots,
                                                       const u8 *__restrict__ flags, u32 want,
                                                       u64 *__restrict__ out_tag, u32 *__restrict__ out_slot,
                                                       u64 *cursor) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = i < n_slots && (flags[i] & want);
    u32 b = __ballot_sync(0xFFFFFFFFu, hit);
    if (!b) return;
    u32 lane = threadIdx.x & 31;
    u64 base = 0;
    if (lane == 0) base = atomicAdd(cursor, (u64)__popc(b));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (hit) {
        u64 o = base + __popc(b & ((1u << lane) - 1));
        out_tag[o] = table[i].tag;
        out_slot[o] = (u32)i;
    }
}

// ---------------------------------------------------------------------------------------------
// k_tip_paths: candidate paths of pruningErrorContigFromBranch (debruijn.py:169-186) for every
// branch node, ignoring already_pull_out (the host applies that filter in branch order, which
// is the only sequential part).  Record = 6 words: len | branch-bitmask << 8, then <= 5 slots.
// mode 0 counts, mode 1 fills at offs[b].
// ---------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_tip_paths(const Slot<W> *__restrict__ table, TGeom tg, int k,
                                                         const u8 *__restrict__ flags, const u16 *__restrict__ kept,
                                                         const u32 *__restrict__ branch, u64 n_branch, int mode,
                                                         u32 *__restrict__ counts, const u64 *__restrict__ offs,
                                                         u32 *__restrict__ recs, DevCounters *ctr) {
    typedef typename KeyOf<W>::type K;
    u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_branch) return;
    const K mask = key_mask<W>(k);
    u32 vec[5], word[5], child[5];
    K keyv[5];
    u32 n = 0;
    vec[0] = branch[b]; word[0] = kept[vec[0]]; child[0] = 0; keyv[0] = table[vec[0]].key;
    int depth = 1;
    while (depth > 0) {
        int d = depth - 1;
        u32 ci = child[d];
        if ((int)ci >= kept_count(word[d])) { depth--; continue; }
        child[d] = ci + 1;
        if (depth == 5) continue;                                   // callee would see depth == 0
        K sk = key_roll(keyv[d], (u32)kept_base(word[d], (int)ci), mask);
        u64 fi = table_find<W>(table, tg, sk);
        if (fi == NONE64) { raise_fault(ctr, FAULT_TIP_MISSING, vec[d]); continue; }
        u32 nx = (u32)fi;
        vec[depth] = nx;
        u32 fl = flags[nx];
        if (fl & F_DEADEND) {                                       // debruijn.py:173
            if (mode == 1) {
                u32 *rec = recs + (offs[b] + n) * 6;
                u32 bm = 0;
                for (int j = 0; j <= depth; j++) {
                    rec[1 + j] = vec[j];
                    if (flags[vec[j]] & F_BRANCH) bm |= 1u << j;
                }
                rec[0] = (u32)(depth + 1) | (bm << 8);
            }
            n++;
            continue;
        }
        word[depth] = kept[nx]; child[depth] = 0; keyv[depth] = table[nx].key;
        depth++;
    }
    if (mode == 0) counts[b] = n;
}

__global__ void __launch_bounds__(DBG_BLOCK) k_mark_pulled(const u32 *__restrict__ pulled, u64 n, u8 *flags) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[pulled[i]] |= (u8)F_PULLED;
}

// ---------------------------------------------------------------------------------------------
// k_pull_reads: read r is pulled out iff any of its k-mers is a branch node (debruijn.py:274-278)
// ---------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_pull_reads(const u64 *__restrict__ packed, const u64 *__restrict__ off,
                                                          const u32 *__restrict__ tile_first, u64 n_bases, int k,
                                                          SideEntry<W> *st, u64 cap, const u32 *__restrict__ bm, int bits,
                                                          u8 *read_flags, const u32 *__restrict__ read_batch) {
    typedef typename KeyOf<W>::type K;
    for_windows<W, false>(packed, off, tile_first, n_bases, k,
        [&](K key, u32 r, u32, bool, u32) {
            if (read_batch) key = key_add_batch(key, __ldg(read_batch + r), k);
            if (!bitmap_test(bm, bits, hash32_key(key))) return;
            if (side_locate<W>(st, cap, key, false) != NONE64) read_flags[r] = 1;     // key is a branch node
        });
}

// ---------------------------------------------------------------------------------------------
// Contig walk as list ranking (debruijn.py:288-316 without the recursion).
//
// link[v] = SPLIT | dist << 32 | next, "dist hops from v lead to next".  Roots are the nodes at
// which the reference's DFS stops (debruijn.py:292, :304): pulled nodes, nodes without successor,
// nodes with > 1 kept successor; a root is link == (v, 0).  Every other node ("chain node") has
// exactly one kept successor.  Only the START nodes need an answer (one contig per start,
// debruijn.py:331-345), so instead of pointer-jumping all n nodes for log n rounds (O(n log n)
// sector traffic) the chain is cut at SPLITTERS -- every start plus a 1/2^WALK_SPLIT_LOG2 hash
// sample of the chain nodes:
//   k_link_init  writes the links, flags splitters, compacts them into a list
//   k_sublist    one thread per splitter walks to the next splitter or root (expected 2^LOG2 hops,
//                L2-resident 8-byte reads) and stores (that node, hops) in its own link
//   k_jump       pointer jumping over the splitter list only (n / 2^LOG2 + starts elements)
// The 64-bit (next, dist) pair is read and written as one word, so in k_jump any interleaving
// reads a valid earlier or later state of the same chain.  A start that never reaches a root sits
// on, or leads into, a cycle (debruijn.py:289: nothing is emitted): k_sublist finds splitter-free
// cycles with Brent's checkpointing and marks the splitter as its own target; cycles through
// splitters are what is left not pointing at a root after ceil(log2(m)) + 2 rounds.
// ---------------------------------------------------------------------------------------------
#ifndef WALK_SPLIT_LOG2
#define WALK_SPLIT_LOG2 5
#endif
#ifndef WALK_SUPER_LOG2
#define WALK_SUPER_LOG2 3             // measured (split, super) = (5,5) 0.466 / (5,3) 0.436 / (4,3) 0.445 / (5,2) 0.447 ms of walk at C2
#endif
#ifndef WALK_BLOCK_JUMP_CAP
#define WALK_BLOCK_JUMP_CAP 131072         // super-splitter list capacity up to which ONE cluster launch jumps it
#endif
#define LINK_SPLIT 0x8000000000000000ULL
#define LINK_DIST_MAX 0x7FFFFFFFu
DBG_HD u64 link_pack(u32 next, u32 dist) { return ((u64)dist << 32) | next; }
DBG_HD u32 link_next(u64 l) { return (u32)l; }
DBG_HD u32 link_dist(u64 l) { return (u32)(l >> 32) & LINK_DIST_MAX; }
DBG_HD bool link_is_root(u64 l, u32 v) { return l == (u64)v; }
DBG_HD u32 link_dist_add(u32 a, u32 b) { u64 s = (u64)a + b; return s > LINK_DIST_MAX ? LINK_DIST_MAX : (u32)s; }
DBG_HD u32 hash32(u32 x) { return mix32(x); }

// second sampling level: a splitter whose hash bits 5..9 are zero is a SUPER-splitter (1 in 32)
DBG_HD bool super_hash(u32 v) { return ((hash32(v) >> WALK_SPLIT_LOG2) & ((1u << WALK_SUPER_LOG2) - 1)) == 0; }

__global__ void __launch_bounds__(DBG_BLOCK) k_link_init(const u8 *__restrict__ flags, const u16 *__restrict__ kept,
                                                         const u32 *__restrict__ succ, u64 n_slots, u64 *link,
                                                         u32 *__restrict__ split_list, u64 split_cap, u64 *cursor,
                                                         u32 *__restrict__ super_list, u64 super_cap, u64 *super_cursor) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool split = false;
    if (i < n_slots) {
        u32 f = flags[i];
        // sharded table (dbg_shard.cuh): an exit is a root of the local ranking, an entry a splitter
        bool chain = (f & F_OCC) && !(f & (F_PULLED | F_EXIT)) && kept_count(kept[i]) == 1;
        split = chain && ((f & (F_START | F_ENTRY)) || (hash32((u32)i) & ((1u << WALK_SPLIT_LOG2) - 1)) == 0);
        link[i] = chain ? (link_pack(succ[i], 1) | (split ? LINK_SPLIT : 0)) : link_pack((u32)i, 0);
    }
    // list append, aggregated per block: one global atomic per block and list (one per warp on the same
    // word serialised in L2: 3 x 10^5 of them cost more than the kernel's memory traffic)
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

// walk from a chain node (exclusive; its own link leads to `from_next` over `first_dist` hops) to the first
// splitter or root; returns that node and the hop count, or NONE32 when the walk is caught in a splitter-free
// cycle.  Hops are weighted by the links' own distances: 1 on the node table, anything on the reduced
// list of a sharded walk (dbg_shard.cuh).
DBG_D u32 walk_to_splitter(const u64 *link, u32 from_next, u32 first_dist, u32 *hops) {
    u32 cur = from_next, d = first_dist;
    u32 check = cur, power = 1, lam = 0;              // Brent: checkpoint every power of two
    for (;;) {
        u64 l = __ldcg(&link[cur]);
        if ((l & LINK_SPLIT) || link_is_root(l, cur)) { *hops = d; return cur; }
        cur = link_next(l);
        d = link_dist_add(d, link_dist(l));
        if (cur == check) return NONE32;
        if (++lam == power) { check = cur; power <<= 1; lam = 0; }
    }
}

__global__ void __launch_bounds__(DBG_BLOCK) k_sublist(u64 *link, const u32 *__restrict__ split_list,
                                                       const u64 *__restrict__ n_split_dev, u64 split_cap) {
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 n_split = __ldg(n_split_dev);
    if (n_split > split_cap) n_split = split_cap;                           // the host reports the overflow
    if (j >= n_split) return;
    u32 v = split_list[j];
    const u64 l0 = __ldcg(&link[v]);
    u32 first = link_next(l0);
    u32 hops = 0;
    u32 t = first == v ? NONE32 : walk_to_splitter(link, first, link_dist(l0), &hops);     // first == v: self loop
    // only this thread writes link[v]; walkers that pass by only look at its SPLIT bit
    __stcg(&link[v], t == NONE32 ? (LINK_SPLIT | link_pack(v, LINK_DIST_MAX)) : (LINK_SPLIT | link_pack(t, hops)));
}

// walk over splitter-level links from splitter-or-root `cur` (inclusive) to the first super-splitter
// or root, adding the hop counts to *hops; NONE32 when caught in a cycle of plain splitters
DBG_D u32 walk_to_super(const u64 *link, u32 cur, u32 *hops) {
    u32 d = *hops;
    u32 check = cur, power = 1, lam = 0;              // Brent: checkpoint every power of two
    for (;;) {
        const u64 l = __ldcg(&link[cur]);
        if (link_is_root(l, cur) || super_hash(cur)) { *hops = d; return cur; }
        const u32 nx = link_next(l);
        if (nx == cur) return NONE32;                 // marked splitter-free cycle
        d = link_dist_add(d, link_dist(l));
        cur = nx;
        if (cur == check) return NONE32;
        if (++lam == power) { check = cur; power <<= 1; lam = 0; }
    }
}

// level 2: every super-splitter walks the splitter list to the next super-splitter or root
__global__ void __launch_bounds__(DBG_BLOCK) k_suplist(u64 *link, const u32 *__restrict__ super_list,
                                                       const u64 *__restrict__ n_super_dev, u64 super_cap) {
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 n_super = __ldg(n_super_dev);
    if (n_super > super_cap) n_super = super_cap;
    if (j >= n_super) return;
    const u32 v = super_list[j];
    const u64 l = __ldcg(&link[v]);
    const u32 first = link_next(l);
    u32 hops = link_dist(l);
    const u32 t = first == v ? NONE32 : walk_to_super(link, first, &hops);
    // only this thread writes link[v]; walkers stop AT a super-splitter (hash test) without reading through it
    __stcg(&link[v], t == NONE32 ? (LINK_SPLIT | link_pack(v, LINK_DIST_MAX)) : (LINK_SPLIT | link_pack(t, hops)));
}

__global__ void __launch_bounds__(DBG_BLOCK) k_jump(u64 *link, const u32 *__restrict__ split_list, u64 n_split,
                                                    u64 n_slots, DevCounters *ctr) {
    u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool changed = false;
    if (j < n_split) {
        u32 v = split_list[j];
        u64 l = __ldcg(&link[v]);
        u32 nx = link_next(l);
        if (nx != v) {                                  // nx == v: marked cycle
            if (nx >= n_slots) raise_fault(ctr, FAULT_JUMP_RANGE, v);
            else {
                u64 ln = __ldcg(&link[nx]);
                u32 nn = link_next(ln);
                if ((ln & LINK_SPLIT) && nn != nx) {    // else nx is a root (final) or a marked cycle
                    __stcg(&link[v], LINK_SPLIT | link_pack(nn, link_dist_add(link_dist(l), link_dist(ln))));
                    changed = true;
                }
            }
        }
    }
    // one store per warp: a store per thread to the same word serialises in L2
    if (__any_sync(0xFFFFFFFFu, changed) && (threadIdx.x & 31) == 0) ctr->changed = 1;
}

// level 3 in ONE launch: a single block pointer-jumps the super-splitter list (<= WALK_BLOCK_JUMP_CAP
// elements) to convergence, __syncthreads between rounds; leaves the round count in ctr->changed
__global__ void __launch_bounds__(1024) k_jump_block(u64 *link, const u32 *__restrict__ super_list,
                                                     const u64 *__restrict__ n_super_dev, u64 super_cap,
                                                     u64 n_slots, int max_rounds, DevCounters *ctr) {
    __shared__ u32 s_changed;
    u64 n = __ldg(n_super_dev);
    if (n > super_cap) n = super_cap;
    int rounds = 0;
    for (int r = 0; r < max_rounds; r++) {
        if (threadIdx.x == 0) s_changed = 0;
        __syncthreads();
        bool changed = false;
        for (u64 j = threadIdx.x; j < n; j += blockDim.x) {
            const u32 v = super_list[j];
            const u64 l = __ldcg(&link[v]);
            const u32 nx = link_next(l);
            if (nx == v) continue;                      // marked cycle
            if (nx >= n_slots) { raise_fault(ctr, FAULT_JUMP_RANGE, v); continue; }
            const u64 ln = __ldcg(&link[nx]);
            const u32 nn = link_next(ln);
            if ((ln & LINK_SPLIT) && nn != nx) {        // else nx is a root (final) or a marked cycle
                __stcg(&link[v], LINK_SPLIT | link_pack(nn, link_dist_add(link_dist(l), link_dist(ln))));
                changed = true;
            }
        }
        if (changed) s_changed = 1;
        __syncthreads();
        rounds = r + 1;
        if (!s_changed) break;
        __syncthreads();
    }
    if (threadIdx.x == 0) ctr->changed = (u32)rounds;
}

#ifndef DBG_EMU
// The same on a thread-block CLUSTER: 8 CTAs x 1024 threads share the list (one or two elements per
// thread), cluster.sync() between rounds.  One block alone spent 98 us on 5.7 K elements (13 rounds x 6
// elements per thread x 2 dependent loads); rounds are now two dependent loads + a cluster barrier.
// No early exit (a cluster-wide vote would cost as much as a round): it runs the ceil(log2 n) + 2 rounds.
#include <cooperative_groups.h>
#define WALK_CLUSTER 8
__global__ void __cluster_dims__(WALK_CLUSTER, 1, 1) __launch_bounds__(1024)
k_jump_cluster(u64 *link, const u32 *__restrict__ super_list, const u64 *__restrict__ n_super_dev, u64 super_cap,
               u64 n_slots, DevCounters *ctr) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    u64 n = __ldg(n_super_dev);
    if (n > super_cap) n = super_cap;
    int max_rounds = 2;
    for (u64 m = n; m > 1; m >>= 1) max_rounds++;
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x, nth = (u64)gridDim.x * blockDim.x;
    for (int r = 0; r < max_rounds; r++) {
        for (u64 j = tid; j < n; j += nth) {
            const u32 v = super_list[j];
            const u64 l = __ldcg(&link[v]);
            const u32 nx = link_next(l);
            if (nx == v) continue;                      // marked cycle
            if (nx >= n_slots) { raise_fault(ctr, FAULT_JUMP_RANGE, v); continue; }
            const u64 ln = __ldcg(&link[nx]);
            const u32 nn = link_next(ln);
            if ((ln & LINK_SPLIT) && nn != nx)          // else nx is a root (final) or a marked cycle
                __stcg(&link[v], LINK_SPLIT | link_pack(nn, link_dist_add(link_dist(l), link_dist(ln))));
        }
        __threadfence();
        cluster.sync();
    }
    if (tid == 0) ctr->changed = (u32)max_rounds;
}
#endif

// (terminal, hops) of an arbitrary node after the super-splitter list has converged:
// chain nodes -> splitter -> super-splitter -> root
DBG_D bool resolve_chain(const u64 *link, u32 v, u32 *terminal, u32 *hops) {
    u64 l = __ldcg(&link[v]);
    u32 d = 0, w = v;
    if (!link_is_root(l, v) && !(l & LINK_SPLIT)) {          // plain chain node: reach a splitter first
        u32 first = link_next(l);
        if (first == v) return false;
        w = walk_to_splitter(link, first, link_dist(l), &d);
        if (w == NONE32) return false;
        l = __ldcg(&link[w]);
    }
    if (link_is_root(l, w)) { *terminal = w; *hops = d; return true; }
    w = walk_to_super(link, w, &d);                          // w is a splitter
    if (w == NONE32) return false;
    l = __ldcg(&link[w]);
    if (link_is_root(l, w)) { *terminal = w; *hops = d; return true; }
    u32 t = link_next(l);
    if (!link_is_root(__ldcg(&link[t]), t)) return false;    // cycle
    *terminal = t; *hops = link_dist_add(d, link_dist(l));
    return true;
}

struct ContigRec {          // one per start, in start order; kind T_CYCLE / n_nodes 0 = nothing emitted.
    u32 start, n_nodes, terminal, kind;     // Same layout as dbg_contig_t (include/dbg_b200.h): the host hands
    u64 seg_begin, seg_end;                 // the device's array straight to the caller when nothing is dropped
};

__global__ void __launch_bounds__(DBG_BLOCK) k_emit(const u32 *__restrict__ starts, u64 n_starts,
                                                    const u64 *__restrict__ link, const u8 *__restrict__ flags,
                                                    ContigRec *__restrict__ out, u64 *n_dropped) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    bool dropped = false;
    if (i < n_starts) {
        u32 s = starts[i];
        u32 t = s, d = 0;
        ContigRec c;
        c.start = s;
        c.seg_begin = 0; c.seg_end = 0;
        if (!resolve_chain(link, s, &t, &d)) { c.terminal = t; c.kind = T_CYCLE; c.n_nodes = 0; }
        else {
            c.terminal = t;
            u32 f = flags[t];
            if (f & F_PULLED) { c.kind = T_PULLED; c.n_nodes = d; }          // path stops in front of t
            else { c.kind = (f & F_DEADEND) ? T_DEADEND : T_BRANCH; c.n_nodes = d + 1; }
        }
        dropped = c.n_nodes == 0;
        out[i] = c;
    }
    const u32 b = __ballot_sync(0xFFFFFFFFu, dropped);
    if (b && (threadIdx.x & 31) == 0) atomicAdd(n_dropped, (u64)__popc(b));
}

// junction table for the final=True enumeration: for every branch node and each kept successor,
// where that successor's chain ends.
struct JunctionSucc {
    u32 entry, n_hops, terminal, kind;      // kind as above; n_hops = distance entry -> terminal
};
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_junctions(const Slot<W> *__restrict__ table, TGeom tg, int k,
                                                         const u8 *__restrict__ flags, const u16 *__restrict__ kept,
                                                         const u64 *__restrict__ link, const u32 *__restrict__ branch,
                                                         u64 n_branch, JunctionSucc *__restrict__ out, DevCounters *ctr) {
    typedef typename KeyOf<W>::type K;
    u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_branch) return;
    u32 v = branch[b];
    u32 word = kept[v];
    K key = table[v].key;
    for (int j = 0; j < 4; j++) {
        JunctionSucc js;
        js.entry = NONE32; js.n_hops = 0; js.terminal = NONE32; js.kind = T_CYCLE;
        if (j < kept_count(word)) {
            K sk = key_roll(key, (u32)kept_base(word, j), key_mask<W>(k));
            u64 fi = table_find<W>(table, tg, sk);
            if (fi == NONE64) { raise_fault(ctr, FAULT_JUNCTION_MISSING, v); out[b * 4 + j] = js; continue; }
            u32 s = (u32)fi, t = s, d = 0;
            js.entry = s;
            if (!resolve_chain(link, s, &t, &d)) js.kind = T_CYCLE;
            else {
                u32 f = flags[t];
                js.n_hops = d; js.terminal = t;
                js.kind = (f & F_PULLED) ? T_PULLED : (f & F_DEADEND) ? T_DEADEND : T_BRANCH;
            }
        }
        out[b * 4 + j] = js;
    }
}

// ---------------------------------------------------------------------------------------------
// strings and scores (debruijn.py:298-300; II_assembleFromReads.py:14-18).  A contig is a chain of up
// to 10^5..10^8 nodes, and "first node's label, then the last base of every later node" walked node by
// node is that many DEPENDENT loads on one thread.  Instead the successor array is lifted once per
// build: up[j][v] = succ^(2^(j+1))(v), j = 0..LIFT_LEVELS-1 (pointer doubling), and a block of 2^LIFT_LEVELS threads takes
// one chain segment: thread t reaches node t with <= LIFT_LEVELS hops, then strides 2^LIFT_LEVELS nodes a hop, so
// consecutive threads write consecutive characters (coalesced) and the dependent chain is n / 256 long.
// ---------------------------------------------------------------------------------------------
#define LIFT_LEVELS 8
#define LIFT_BLOCK (1 << LIFT_LEVELS)
struct SegJob {
    u32 entry, n_nodes;
    u64 out;            // offset into the char buffer
    u32 first;          // 1 = first segment of its contig: write the whole entry label
    u32 pad;
};
struct ScoreJob {
    u64 seg_begin, seg_end;     // into the uploaded segment array
};

__global__ void __launch_bounds__(DBG_BLOCK) k_lift(const u32 *__restrict__ prev, u32 *__restrict__ next, u64 n_slots) {
    const u64 v = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_slots) return;
    const u32 a = prev[v];
    next[v] = (a == NONE32 || a >= n_slots) ? NONE32 : prev[a];     // a >= n_slots: a sharded table keeps remote ids there
}

// node t hops after `entry` (t < 2^LIFT_LEVELS); level 0 = succ, level j = up + (j-1) * n_slots
DBG_D u32 lift_to(const u32 *__restrict__ succ, const u32 *__restrict__ up, u64 n_slots, u32 entry, u32 t) {
    u32 v = entry;
    if (t & 1u) v = succ[v];
#pragma unroll
    for (int j = 1; j < LIFT_LEVELS; j++)
        if (((t >> j) & 1u) && v != NONE32) v = up[(u64)(j - 1) * n_slots + v];
    return v;
}

// last symbol / symbol i of a node label as a character: 2-bit DNA keys, or generic keys of lut[256] bits a symbol
// (lut = the generic engine's code -> byte table followed by its code width, dbg_generic.cuh GEN_LUT_BYTES)
template <int W> DBG_D char label_last_char(typename KeyOf<W>::type key, const u8 *__restrict__ lut) {
    if (lut) return (char)lut[(u32)key_lo64(key) & ((1u << lut[256]) - 1u)];
    return (char)((0x54474341u >> (8 * key_last_base(key))) & 0xFFu);
}
DBG_D u32 key_symbol_g(u64, int, int, int) { return 0; }             // 64-bit keys are DNA only
DBG_D u32 key_symbol_g(K128 key, int n, int j, int gb) {
    const int sh = gb * (n - 1 - j);
    const u32 mask = (1u << gb) - 1u;
    if (sh >= 64) return (u32)(key.hi >> (sh - 64)) & mask;
    u64 x = key.lo >> sh;
    if (sh + gb > 64) x |= key.hi << (64 - sh);
    return (u32)x & mask;
}
template <int W> DBG_D char label_char(typename KeyOf<W>::type key, int k, int i, const u8 *__restrict__ lut) {
    if (lut) return (char)lut[key_symbol_g(key, k, i, (int)lut[256])];
    return (char)((0x54474341u >> (8 * key_base(key, k, i))) & 0xFFu);
}

template <int W>
__global__ void __launch_bounds__(LIFT_BLOCK) k_materialize(const Slot<W> *__restrict__ table, u64 n_slots, int k,
                                                            const u32 *__restrict__ succ, const u32 *__restrict__ up,
                                                            const u8 *__restrict__ lut, const SegJob *__restrict__ jobs,
                                                            char *__restrict__ buf) {
    const SegJob job = jobs[blockIdx.x];
    const u32 n = job.n_nodes, t = threadIdx.x;
    if (n == 0) return;
    char *o = buf + job.out;
    u64 shift = 0;                                                   // node i writes its last character at o[shift + i]
    if (job.first) {
        const typename KeyOf<W>::type key = table[job.entry].key;
        for (int i = (int)t; i < k; i += LIFT_BLOCK) o[i] = label_char<W>(key, k, i, lut);
        shift = (u64)k - 1;                                          // node 0's last character is o[k - 1], written above
    }
    if (t >= n) return;
    u32 v = lift_to(succ, up, n_slots, job.entry, t);
    const u32 *top = up + (u64)(LIFT_LEVELS - 1) * n_slots;          // succ^(2^LIFT_LEVELS)
    for (u32 i = t; i < n && v != NONE32; i += LIFT_BLOCK) {
        if (i || !job.first) o[shift + i] = label_last_char<W>(table[v].key, lut);
        v = top[v];
    }
}

// edge of the generic-alphabet engine (dbg_generic.cuh): one distinct (k+1)-mer
struct GEdge {
    u32 src, dst;                          // node-table slots
    u32 cnt;                               // multiplicity of the (k+1)-mer
    u32 sym;                               // code of its last symbol
    u64 first;                             // tag of its first appearance
};

// getScore: sum of the counts of the edges along a contig.  Inside a segment every edge is the single kept
// successor of its source (DNA: cnt[kept base]; generic: the first edge of the node's run); between two segments
// (final=True: the path runs through a multi-successor node) it is looked up by the pair (last node, next entry).
template <int W>
DBG_D u32 edge_count_between(const Slot<W> *__restrict__ table, const GEdge *__restrict__ ge, const u32 *__restrict__ estart,
                             const u32 *__restrict__ edeg, u32 a, u32 b) {
    if (ge) {
        const u32 s = estart[a], m = edeg[a];
        for (u32 q = 0; q < m; q++) if (ge[s + q].dst == b) return ge[s + q].cnt;
        return 0;
    }
    return table[a].cnt[key_last_base(table[b].key)];
}

template <int W>
__global__ void __launch_bounds__(LIFT_BLOCK) k_scores(const Slot<W> *__restrict__ table, u64 n_slots,
                                                       const u16 *__restrict__ kept, const u32 *__restrict__ succ,
                                                       const u32 *__restrict__ up, const GEdge *__restrict__ ge,
                                                       const u32 *__restrict__ estart, const u32 *__restrict__ edeg,
                                                       const u32 *__restrict__ seg_entry, const u32 *__restrict__ seg_nodes,
                                                       const ScoreJob *__restrict__ jobs, u64 *__restrict__ scores) {
    __shared__ u64 s_sum[LIFT_BLOCK / 32];
    __shared__ u32 s_last;
    const ScoreJob job = jobs[blockIdx.x];
    const u32 t = threadIdx.x;
    const u32 *top = up + (u64)(LIFT_LEVELS - 1) * n_slots;
    u64 total = 0;
    u32 prev_last = NONE32;
    for (u64 s = job.seg_begin; s < job.seg_end; s++) {
        const u32 entry = seg_entry[s], n = seg_nodes[s];
        if (n == 0) continue;
        if (t == 0 && prev_last != NONE32) total += edge_count_between<W>(table, ge, estart, edeg, prev_last, entry);
        if (t < n) {
            u32 v = lift_to(succ, up, n_slots, entry, t);
            for (u32 i = t; i < n && v != NONE32; i += LIFT_BLOCK) {
                if (i + 1 < n) total += ge ? ge[estart[v]].cnt : table[v].cnt[kept_base(kept[v], 0)];
                else s_last = v;
                v = top[v];
            }
        }
        __syncthreads();
        prev_last = s_last;
        __syncthreads();
    }
    for (int d = 16; d > 0; d >>= 1) total += __shfl_down_sync(0xFFFFFFFFu, total, d);
    if ((t & 31) == 0) s_sum[t >> 5] = total;
    __syncthreads();
    if (t == 0) {
        u64 x = 0;
        for (int q = 0; q < LIFT_BLOCK / 32; q++) x += s_sum[q];
        scores[blockIdx.x] = x;
    }
}

template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_lookup_edges(const Slot<W> *__restrict__ table, TGeom tg, int k,
                                                            const u8 *__restrict__ q, const u32 *__restrict__ q_batch,
                                                            u64 n, u32 *__restrict__ out) {
    typedef typename KeyOf<W>::type K;
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u8 *s = q + i * (u64)(k + 1);
    K key = zero_key<W>();
    bool bad = false;
    u32 code = 0;
    for (int j = 0; j <= k; j++) {
        u32 c = s[j];
        code = ((c >> 1) & 3u) ^ ((c >> 2) & 1u);
        bad |= (c != ((0x54474341u >> (8 * code)) & 0xFFu));
        if (j < k) key = key_push(key, code);
    }
    u32 res = 0;
    if (!bad) {
        if (q_batch) key = key_add_batch(key, q_batch[i], k);
        u64 idx = table_find<W>(table, tg, key);
        if (idx != NONE64) res = table[idx].cnt[code];
    }
    out[i] = res;
}

// ---------------------------------------------------------------------------------------------
// Both strands (DBG_BOTH_STRANDS): the reads followed by their reverse complements, as a forward-strand assembler is
// usually fed -- an EXTENSION (north-star "canonical k-mers"; the reference has no reverse complement anywhere,
// SURVEY.md 9 delta 1).  The graph then holds every k-mer next to its reverse complement ("twin nodes"), which is the
// canonical graph unfolded: counts of e and rc(e) are equal, and the nodes with key <= rc(key) are the canonical ones.
// ---------------------------------------------------------------------------------------------
DBG_D u8 comp_base(u8 c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c; }
__global__ void __launch_bounds__(DBG_BLOCK) k_strand_reads(const u8 *__restrict__ src, const u64 *__restrict__ off, u64 n_reads,
                                                            u64 n_bases, u8 *__restrict__ dst, u64 *__restrict__ off2) {
    const u32 lane = threadIdx.x & 31;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 r = warp; r < n_reads; r += n_warps) {
        const u64 a = off[r], L = off[r + 1] - a;
        for (u64 j = lane; j < L; j += 32) {
            dst[a + j] = src[a + j];
            dst[n_bases + a + j] = comp_base(src[a + L - 1 - j]);
        }
        if (lane == 0) { off2[r] = a; off2[n_reads + r] = n_bases + a; }
    }
    if (warp == 0 && lane == 0) off2[2 * n_reads] = 2 * n_bases;
}
DBG_HD u64 rev2bit64(u64 x) {                               // the 32 two-bit groups of a word in reverse order
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
    return (x >> 32) | (x << 32);
}
DBG_HD u64 key_revcomp(u64 key, int k) { return rev2bit64(~key) >> (64 - 2 * k); }
DBG_HD K128 key_revcomp(K128 key, int k) {
    const u64 hi = rev2bit64(~key.lo), lo = rev2bit64(~key.hi);       // 128 bits reversed, left-aligned
    const int sh = 128 - 2 * k;                                       // 2 <= sh <= 126 (a handle with 128-bit keys may hold k <= 31 too)
    K128 r;
    if (sh >= 64) { r.hi = 0; r.lo = sh == 64 ? hi : hi >> (sh - 64); }
    else { r.hi = hi >> sh; r.lo = (lo >> sh) | (hi << (64 - sh)); }
    return r;
}
DBG_HD bool key_le(u64 a, u64 b) { return a <= b; }
DBG_HD bool key_le(K128 a, K128 b) { return a.hi < b.hi || (a.hi == b.hi && a.lo <= b.lo); }
// twin[j] = node id of the reverse complement of node ids[j] (NONE32 if the graph does not hold it), canonical[j] = key <= rc(key)
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_twin_nodes(const Slot<W> *__restrict__ table, TGeom tg, int k,
                                                          const u32 *__restrict__ ids, u64 n, u32 *__restrict__ twin,
                                                          u8 *__restrict__ canonical) {
    typedef typename KeyOf<W>::type K;
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const K key = load_key(&table[ids[j]]);
    const K rc = key_revcomp(key, k);
    const u64 fi = table_find<W>(table, tg, rc);
    twin[j] = fi == NONE64 ? NONE32 : (u32)fi;
    canonical[j] = key_le(key, rc) ? 1 : 0;
}

// TRUE in-degree / out-degree / balance of the (unpruned) graph: in = distinct predecessors = distinct (k+1)-mers that end
// in the node, out = distinct successors, balance = out - in.  An EXTRA output (north-star kernel 4, SURVEY.md 9
// delta 4): the reference keeps only the 0/1 first-seen `indegree` flag (debruijn.py:137-143) and no balance.
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_true_indegree(const Slot<W> *__restrict__ table, TGeom tg, int k,
                                                             u32 *__restrict__ indeg) {
    typedef typename KeyOf<W>::type K;
    const K mask = key_mask<W>(k);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < tg.n_slots; i += (u64)gridDim.x * blockDim.x) {
        const K key = load_key(&table[i]);
        if (key == empty_key<W>()) continue;
        for (u32 c = 0; c < 4; c++) {
            if (!table[i].cnt[c]) continue;
            const u64 fi = table_find<W>(table, tg, key_roll(key, c, mask));
            if (fi != NONE64) atomicAdd(&indeg[fi], 1u);
        }
    }
}
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_degree_gather(const Slot<W> *__restrict__ table, const u32 *__restrict__ indeg,
                                                             const u32 *__restrict__ ids, u64 n, u8 *__restrict__ o_in,
                                                             u8 *__restrict__ o_out, signed char *__restrict__ o_bal) {
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const u32 i = ids[j];
    u32 out = 0;
    for (u32 c = 0; c < 4; c++) out += table[i].cnt[c] != 0u;
    o_in[j] = (u8)indeg[i];
    o_out[j] = (u8)out;
    o_bal[j] = (signed char)((int)out - (int)indeg[i]);
}

template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_gather_keys(const Slot<W> *__restrict__ table,
                                                           const u32 *__restrict__ ids, u64 n,
                                                           u64 *__restrict__ lo, u64 *__restrict__ hi) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    typename KeyOf<W>::type key = table[ids[i]].key;
    lo[i] = key_lo64(key);
    hi[i] = key_hi64(key);
}

// export: the fields of the given slots (dbg_export_nodes: the occupied slots in first-occurrence order)
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_export_gather(const Slot<W> *__restrict__ table, const u8 *__restrict__ flags,
                                                             const u16 *__restrict__ kept, const u32 *__restrict__ ids, u64 n,
                                                             u64 *__restrict__ o_lo, u64 *__restrict__ o_hi,
                                                             u64 *__restrict__ o_tag, u32 *__restrict__ o_cnt,
                                                             u8 *__restrict__ o_flags, u16 *__restrict__ o_kept) {
    const u64 o = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    const u32 i = ids[o];
    typename KeyOf<W>::type key = table[i].key;
    o_lo[o] = key_lo64(key);
    o_hi[o] = key_hi64(key);
    o_tag[o] = table[i].tag;
    for (int c = 0; c < 4; c++) o_cnt[4 * o + c] = table[i].cnt[c];
    o_flags[o] = flags[i];
    o_kept[o] = kept[i];
}

// ---------------------------------------------------------------------------------------------
// Node records: one per occupied slot, W + 3 u64 words { key (W) | tag | cnt0,cnt1 | cnt2,cnt3 }.
// They are what ranks exchange in the multi-GPU build: the owner of a key is
// (hash32(key) * n_parts) >> 32, and merging records is associative (min tag, summed counters).
// ---------------------------------------------------------------------------------------------
template <int W> DBG_D u32 record_owner(typename KeyOf<W>::type key, u32 n_parts) {
    return (u32)(((u64)hash32_key(key) * n_parts) >> 32);
}
DBG_D void key_to_words(u64 key, u64 *o) { o[0] = key; }
DBG_D void key_to_words(K128 key, u64 *o) { o[0] = key.lo; o[1] = key.hi; }
template <int W> DBG_D typename KeyOf<W>::type key_from_words(const u64 *w);
template <> DBG_D u64 key_from_words<1>(const u64 *w) { return w[0]; }
template <> DBG_D K128 key_from_words<2>(const u64 *w) { K128 r; r.lo = w[0]; r.hi = w[1]; return r; }

// Both kernels aggregate per block in shared memory first: one global atomic per (block, owner)
// instead of one per record (millions of atomics on a handful of addresses serialise in L2).
#define DBG_MAX_PARTS 1024
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_rec_count(const Slot<W> *__restrict__ table, u64 n_slots, u32 n_parts,
                                                         u64 *counts) {
    __shared__ u32 s_cnt[DBG_MAX_PARTS];
    for (u32 p = threadIdx.x; p < n_parts; p += blockDim.x) s_cnt[p] = 0;
    __syncthreads();
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_slots) {
        typename KeyOf<W>::type key = table[i].key;
        if (key != empty_key<W>()) atomicAdd(&s_cnt[record_owner<W>(key, n_parts)], 1u);
    }
    __syncthreads();
    for (u32 p = threadIdx.x; p < n_parts; p += blockDim.x)
        if (s_cnt[p]) atomicAdd(&counts[p], (u64)s_cnt[p]);
}

template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_rec_scatter(const Slot<W> *__restrict__ table, u64 n_slots, u32 n_parts,
                                                           u64 *cursors, u64 *__restrict__ out) {
    __shared__ u32 s_cnt[DBG_MAX_PARTS];
    __shared__ u64 s_base[DBG_MAX_PARTS];
    for (u32 p = threadIdx.x; p < n_parts; p += blockDim.x) s_cnt[p] = 0;
    __syncthreads();
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    typename KeyOf<W>::type key = empty_key<W>();
    u32 owner = 0, local = 0;
    if (i < n_slots) key = table[i].key;
    const bool have = key != empty_key<W>();
    if (have) {
        owner = record_owner<W>(key, n_parts);
        local = atomicAdd(&s_cnt[owner], 1u);
    }
    __syncthreads();
    for (u32 p = threadIdx.x; p < n_parts; p += blockDim.x)
        if (s_cnt[p]) s_base[p] = atomicAdd(&cursors[p], (u64)s_cnt[p]);
    __syncthreads();
    if (have) {
        const u64 o = (s_base[owner] + local) * (W + 3);
        key_to_words(key, out + o);
        out[o + W] = table[i].tag;
        out[o + W + 1] = (u64)table[i].cnt[0] | ((u64)table[i].cnt[1] << 32);
        out[o + W + 2] = (u64)table[i].cnt[2] | ((u64)table[i].cnt[3] << 32);
    }
}

template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_rec_merge(Slot<W> *table, TGeom tg, const u64 *__restrict__ recs, u64 n,
                                                         DevCounters *ctr) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 *r = recs + i * (W + 3);
    u64 idx = table_upsert<W>(table, tg, key_from_words<W>(r));
    if (idx == NONE64) { ctr->overflow = 1; return; }
    Slot<W> *s = &table[idx];
    const u64 tag = r[W];
    if (tag < __ldcg(&s->tag)) atomicMin(&s->tag, tag);
    const u32 c[4] = {(u32)r[W + 1], (u32)(r[W + 1] >> 32), (u32)r[W + 2], (u32)(r[W + 2] >> 32)};
#pragma unroll
    for (int b = 0; b < 4; b++) if (c[b]) atomicAdd(&s->cnt[b], c[b]);
}

// order records: W + 4 words { key (W) | first[4] }
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_side_export(const SideEntry<W> *__restrict__ st, u64 cap, u64 *__restrict__ out,
                                                           u64 capacity, u64 *cursor) {
    u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    typename KeyOf<W>::type key = empty_key<W>();
    if (e < cap) key = st[e].key;
    const bool have = key != empty_key<W>();
    const u32 b = __ballot_sync(0xFFFFFFFFu, have);
    if (!b) return;
    const u32 lane = threadIdx.x & 31;
    u64 base = 0;
    if (lane == 0) base = atomicAdd(cursor, (u64)__popc(b));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (!have) return;
    const u64 o = base + __popc(b & ((1u << lane) - 1));
    if (o >= capacity) return;
    u64 *w = out + o * (W + 4);
    key_to_words(key, w);
#pragma unroll
    for (int c = 0; c < 4; c++) w[W + c] = st[e].first[c];
}

template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_side_merge(SideEntry<W> *st, u64 cap, const u64 *__restrict__ recs, u64 n) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 *r = recs + i * (W + 4);
    u64 e = side_locate<W>(st, cap, key_from_words<W>(r), false);
    if (e == NONE64) return;                    // not a multi-successor node of the merged graph
#pragma unroll
    for (int b = 0; b < 4; b++) if (r[W + b] < __ldcg(&st[e].first[b])) atomicMin(&st[e].first[b], r[W + b]);
}
