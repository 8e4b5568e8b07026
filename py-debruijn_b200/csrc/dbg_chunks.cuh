// dbg_chunks.cuh -- the window loop (debruijn.py:123-144) restated over SUPER-K-MERS ("chunks").
//
// Why.  A scattered table touch costs about one L1TEX/LSU cycle per lane on every SM whatever it
// hits, so a build that probes the node table once per window (k_insert, dbg_kernels.cuh) is bound
// by ~2.5 scattered accesses x 10^8 windows.  Reads that cover the same stretch of the genome
// repeat the same k-mers in the same order, so the windows are first grouped into content-defined
// runs and the RUNS are deduplicated:
//
//   chunk    a maximal run of consecutive node occurrences (debruijn.py:127) of one read that
//            share the position of their minimizer (leftmost smallest hashed m-mer among the
//            first w m-mers of the k-mer; w = 2^logw <= 16 nodes).  Boundaries depend on content
//            only, so two reads that cover the same stretch cut it at the same places and yield
//            byte-identical chunk strings: n nodes + their successor bases = n + k bases
//            (n + k - 1 for the last chunk of a read, whose last node has no successor).
//   phase 0  k_masks        per 32 stream positions: V (is a node occurrence), St (starts a chunk),
//                           the read index, and a HyperLogLog sketch of the chunk starts
//   phase 1  k_chunk_insert every chunk instance -> chunk table keyed by the WHOLE string
//                           (exact, no fingerprints): count += 1, tag = min (read, pos).  The
//                           first and last chunk of a read are cut by the read, not by content,
//                           and are unique: they go to an append-only list, not to the table.
//   phase 2  k_expand       every DISTINCT chunk (table + list) -> its n nodes -> node table:
//                           count[next base] += multiplicity, tag = min(chunk tag + offset).
//                           Node-table passes own a hash range each so their slots stay in L2.
//
// The result is the same node table the per-window build produces (sums and minima are
// associative), whatever the cut rule: correctness does not depend on where chunks are cut, only
// the amount of deduplication does.  configs[1] (150 bp x 30): 99.2 M windows -> 13.5 M chunk
// instances -> 0.6 M distinct interior chunks + 1.7 M boundary chunks -> 15.7 M node upserts.
#pragma once
#include "dbg_common.cuh"

#define CH_WMAX 16                         // nodes per chunk at most
#define CH_MMAX 16                         // minimizer length at most (an m-mer fits 32 bits)
#define CH_HALF 60                         // bases per 128-bit key half (120 bits + 8 bits of meta)

// ---------------------------------------------------------------------------------------------
// Chunk keys.  A chunk string of B <= 60 W bases is stored left-aligned in W halves of 128 bits:
//   half.hi = bases 0..31 of the half, half.lo = bases 32..59 in bits 63..8 | meta in bits 7..0
//   meta    = n << 1 | term    (n = nodes in the chunk, term = the last node has no successor)
// Bases beyond B are zero, so equal (string, n, term) <=> equal key.  meta is never 0xFF, so the
// all-ones half is free: EMPTY for half 0, NOT_READY for half 1 (k > 31: the 256-bit key cannot be
// written by one CAS; the winner of the CAS on half 0 publishes half 1 with a plain store and
// probes that see NOT_READY come back later).
// ---------------------------------------------------------------------------------------------
template <int W> struct ChunkKey { K128 h[W]; };

template <int W> struct ChunkSlot;
template <> struct __align__(32) ChunkSlot<1> {
    K128 h[1];
    u64 tag;
    u32 cnt;
    u32 pad;
    DBG_HD void zero_pad() { pad = 0; }
};
template <> struct __align__(64) ChunkSlot<2> {
    K128 h[2];
    u64 tag;
    u32 cnt;
    u32 pad;
    u64 pad2[2];
    DBG_HD void zero_pad() { pad = 0; pad2[0] = 0; pad2[1] = 0; }
};

DBG_HD K128 k128_ones() { K128 r; r.lo = ~0ULL; r.hi = ~0ULL; return r; }
DBG_HD u64 keep_top_bases(u64 word, int nb) {          // keep the first nb (0..32) bases of a word
    return nb >= 32 ? word : (nb <= 0 ? 0ULL : word & ~(~0ULL >> (2 * nb)));
}
DBG_HD u32 chunk_meta(u32 n, u32 term) { return (n << 1) | term; }
DBG_HD u32 meta_nodes(u32 meta) { return (meta >> 1) & 31u; }
DBG_HD u32 meta_term(u32 meta) { return meta & 1u; }

template <int W> DBG_HD u32 chunk_hash(const ChunkKey<W> &key) {
    u64 x = key.h[0].lo * 0x9E3779B97F4A7C15ULL + key.h[0].hi * 0xC2B2AE3D27D4EB4FULL;
    if (W == 2) x += key.h[1].lo * 0xD6E8FEB86659FD93ULL + key.h[1].hi * 0xA0761D6478BD642FULL;
    x ^= x >> 29;
    return (u32)((x * 0x9E3779B97F4A7C15ULL) >> 32);
}

// 64 bits of a staged 2-bit word array starting at base `rel`
DBG_D u64 staged_word(const u64 *words, u32 rel) {
    const u32 w = rel >> 5;
    const int s = 2 * (int)(rel & 31);
    const u64 w0 = words[w], w1 = words[w + 1];
    return s ? ((w0 << s) | (w1 >> (64 - s))) : w0;
}

// key of the chunk whose string is the B bases at `rel` of a staged word array
template <int W> DBG_D ChunkKey<W> chunk_key_from_stream(const u64 *words, u32 rel, int B, u32 meta) {
    ChunkKey<W> key;
#pragma unroll
    for (int h = 0; h < W; h++) {
        const int nb = B - CH_HALF * h;                                  // bases of this half (may be <= 0)
        u64 hi = 0, lo = 0;
        if (nb > 0) {
            hi = keep_top_bases(staged_word(words, rel + CH_HALF * h), nb);
            if (nb > 32) lo = keep_top_bases(staged_word(words, rel + CH_HALF * h + 32), nb - 32 > 28 ? 28 : nb - 32);
        }
        key.h[h].hi = hi;
        key.h[h].lo = (lo & ~0xFFULL) | meta;
    }
    return key;
}

// the chunk string as contiguous 2-bit words (same layout as the packed read stream), 2 W words
template <int W> DBG_HD void chunk_words(const K128 *h, u64 *out);
template <> DBG_HD void chunk_words<1>(const K128 *h, u64 *out) {
    out[0] = h[0].hi;
    out[1] = h[0].lo & ~0xFFULL;
}
template <> DBG_HD void chunk_words<2>(const K128 *h, u64 *out) {
    const u64 l0 = h[0].lo & ~0xFFULL, l1 = h[1].lo & ~0xFFULL;
    out[0] = h[0].hi;
    out[1] = l0 | (h[1].hi >> 56);
    out[2] = (h[1].hi << 8) | (l1 >> 56);
    out[3] = l1 << 8;
}

struct ChunkCounters {
    u64 n_first;        // entries created in the chunk table = distinct interior chunks
    u64 n_instances;    // chunk instances seen by k_chunk_insert
    u64 n_list;         // unused (the list is indexed by read: 2 r = first chunk, 2 r + 1 = last chunk)
    u32 chunk_overflow; // chunk table full
    u32 list_overflow;  // list capacity exceeded (cannot happen: capacity is 2 x n_first)
    u32 bad_chunk;      // a chunk longer than CH_WMAX (internal consistency)
    u32 pad;
};

DBG_D u32 bit_range(u32 lo, u32 hi) { return (~0u >> (31 - hi)) & (~0u << lo); }   // bits lo..hi, 0 <= lo <= hi <= 31

// ---------------------------------------------------------------------------------------------
// k_masks (phase 0).  One thread per span of 32 stream positions.
//   V  bit j: position p0 + j is a node occurrence (its read is longer than k, pos <= len - k)
//   St bit j: a chunk starts there: V and (first position of its read, or its minimizer is not at
//             the same place as the previous position's)
// Minimizer of the node at p: among the m-mers at p .. p + w - 1 (all inside the k-mer because
// m <= k - w + 1) the one with the smallest 18-bit multiplicative hash, leftmost on ties.  The 48
// hashed m-mers a span needs (positions p0 - 1 .. p0 + 46) come from the 64 bases at p0 - 1, held
// in two registers; the sliding minimum over w = 2^logw is logw rounds of a[i] = min(a[i],
// a[i + 2^r]) on a register array with compile-time indices.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DBG_BLOCK) k_masks(const u64 *__restrict__ packed, const u64 *__restrict__ off,
                                                     u64 n_reads, const u32 *__restrict__ tile_first, u64 n_bases,
                                                     u64 s_begin, u64 n_spans, int k, int logw, int m, u32 sample_mask,
                                                     u32 *__restrict__ st_mask,
                                                     u32 *__restrict__ v_mask, u32 *__restrict__ span_read, u32 *regs) {
    const u64 s = s_begin + (u64)blockIdx.x * blockDim.x + threadIdx.x;   // spans [s_begin, n_spans)
    if (s >= n_spans) return;
    const u64 p0 = s * 32;
    u32 V = 0, S = 0, r0 = 0;
    if (p0 < n_bases) {
        const u64 tile = p0 / DBG_TILE;
        u32 rlo = __ldg(tile_first + tile), rhi = __ldg(tile_first + tile + 1);
        while (rlo < rhi) {
            const u32 mid = rlo + ((rhi - rlo + 1) >> 1);
            if (__ldg(off + mid) <= p0) rlo = mid; else rhi = mid - 1;
        }
        u64 r = rlo;
        r0 = rlo;
        u64 rs = __ldg(off + r), re = __ldg(off + r + 1);
        const u64 pend = p0 + 32;
        for (;;) {
            if (re - rs > (u64)k) {                                      // node occurrences: rs .. re - k
                const u64 lo = rs > p0 ? rs : p0;
                u64 hi = re - (u64)k;
                if (hi >= pend) hi = pend - 1;
                if (lo <= hi) V |= bit_range((u32)(lo - p0), (u32)(hi - p0));
                if (rs >= p0 && rs < pend) S |= 1u << (u32)(rs - p0);
            }
            if (re >= pend || r + 1 >= n_reads) break;
            r++; rs = re; re = __ldg(off + r + 1);
        }
    }
    u32 St = 0;
    if (V) {
        const u64 wm = s ? __ldg(packed + s - 1) : 0ULL, w0 = __ldg(packed + s), w1 = __ldg(packed + s + 1);
        const u64 A = (wm << 62) | (w0 >> 2);                            // bases p0 - 1 .. p0 + 30
        const u64 B = (w0 << 62) | (w1 >> 2);                            // bases p0 + 31 .. p0 + 62
        // the 128 bits as four 32-bit words; m-mer i = the 32 bits at bit 2 i, shifted down to 2 m bits.
        // Order of the candidates: 26-bit multiplicative hash, ties to the leftmost (the index rides in the low 6 bits)
        const u32 X[5] = {(u32)(A >> 32), (u32)A, (u32)(B >> 32), (u32)B, 0u};
        const int down = 32 - 2 * m;
        u32 a[48];
#pragma unroll
        for (int i = 0; i < 48; i++) {
            const int o = 2 * i;
            const u32 t = (o & 31) ? __funnelshift_l(X[o / 32 + 1], X[o / 32], o & 31) : X[o / 32];
            const u32 x = t >> down;
            a[i] = ((x * 0x9E3779B1u) | 63u) - (u32)(63 - i);              // 26-bit hash << 6 | i  (3 instructions)
        }
        {
            if (logw > 0) {
#pragma unroll
                for (int i = 0; i < 47; i++) a[i] = a[i] < a[i + 1] ? a[i] : a[i + 1];
            }
            if (logw > 1) {
#pragma unroll
                for (int i = 0; i < 45; i++) a[i] = a[i] < a[i + 2] ? a[i] : a[i + 2];
            }
            if (logw > 2) {
#pragma unroll
                for (int i = 0; i < 41; i++) a[i] = a[i] < a[i + 4] ? a[i] : a[i + 4];
            }
            if (logw > 3) {
#pragma unroll
                for (int i = 0; i < 33; i++) a[i] = a[i] < a[i + 8] ? a[i] : a[i + 8];
            }
        }
        u32 N = 0, Smp = 0;
#pragma unroll
        for (int j = 0; j < 32; j++) N |= (a[j + 1] != a[j] ? 1u : 0u) << j;
        if (regs) {                                                      // no sketch when the table is sized from the handle's hint
#pragma unroll
            for (int j = 0; j < 32; j++) Smp |= ((a[j + 1] & sample_mask) == 0u ? 1u : 0u) << j;   // distinct sampling by minimizer hash
        }
        St = V & (S | N);
        // sketch of the (sampled) chunk starts that are not read starts: the first min(32, k + 1) bases
        u32 todo = St & ~S & Smp;
        const int nbp = k + 1 < 32 ? k + 1 : 32;
        while (todo) {
            const int j = __ffs((int)todo) - 1;
            todo &= todo - 1;
            const int sh = 2 * (j + 1);
            const u64 t0 = keep_top_bases(sh == 64 ? B : ((A << sh) | (B >> (64 - sh))), nbp);
            const u32 hsh = (u32)(fmix64(t0) >> 32);
            const u32 idx = hsh >> (32 - HLL_P);
            const u32 rank = (u32)__clz((int)((hsh << HLL_P) | (1u << (HLL_P - 1)))) + 1;
            if (regs[idx] < rank) atomicMax(&regs[idx], rank);
        }
    }
    st_mask[s] = St;
    v_mask[s] = V;
    span_read[s] = r0;
}

template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_chunk_table_init(ChunkSlot<W> *table, u64 n_slots) {
    ChunkSlot<W> s;
#pragma unroll
    for (int h = 0; h < W; h++) s.h[h] = k128_ones();
    s.tag = ~0ULL;
    s.cnt = 0;
    s.zero_pad();
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (u64)gridDim.x * blockDim.x) table[i] = s;   // grid-stride
}

// ---------------------------------------------------------------------------------------------
// k_chunk_insert (phase 1).  Each warp works through units of 32 spans (1024 stream positions):
//   stage   the unit's 2-bit words (+ overhang for the longest chunk string) in shared memory
//   select  the unit's chunk instances are dealt to the lanes 32 at a time through prefix sums of
//           popc(St) (no per-lane queueing loop); the end of a chunk is the next bit of St | ~V, in
//           its own span or the next one
//   probe   one chunk per lane: read index and position (span_read + off[]), the key from the
//           staged words, then ONE probe of the chunk table; a probe that finds another key does
//           not loop, it re-queues (key, tag, probes) in a retry ring
// First / last chunks of a read go to the list at slot 2 r / 2 r + 1 (r = read index): no atomics, and the
// list is in read order (a shared append cursor was measured at 1 ms: 4 x 10^5 same-address atomics).
// ---------------------------------------------------------------------------------------------
#define CH_UNIT_WORDS (DBG_UNIT / 32 + 6)
#define CH_RCAP 64
#define CH_MAX_PROBES 1024u
template <int W> struct ChunkInsertShared {
    u64 words[DBG_WARPS][CH_UNIT_WORDS];
    u32 c_read[DBG_WARPS][32];                  // per span: read holding its first position,
    u32 c_pos[DBG_WARPS][32];                   //   position of the span start in that read,
    u16 c_end[DBG_WARPS][32];                   //   distance to that read's end (clamped),
    u16 c_end2[DBG_WARPS][32];                  //   distance to the end of the next non-empty read (clamped)
    u8 c_skip[DBG_WARPS][32];                   //   and that read's index minus c_read (0 = unknown: ask off[])
    ChunkKey<W> rkey[DBG_WARPS][CH_RCAP];
    u64 rtag[DBG_WARPS][CH_RCAP];
    u16 rprobes[DBG_WARPS][CH_RCAP];
};

DBG_D K128 load_k128(const K128 *p) {
    const ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2 *>(p));
    K128 r; r.lo = v.x; r.hi = v.y;
    return r;
}
DBG_D void store_k128(K128 *p, K128 v) {
    __stcg(reinterpret_cast<ulonglong2 *>(p), make_ulonglong2(v.lo, v.hi));
}

// one probe of one slot: 0 = counted, 3 = counted and this probe created the entry, 1 = another key lives
// here (go on to the next slot), 2 = the slot's second half is not published yet (come back to the same slot)
template <int W>
DBG_D int chunk_probe(ChunkSlot<W> *s, const ChunkKey<W> &key, u64 tag) {
    const K128 c0 = load_k128(&s->h[0]);
    u64 seen = __ldcg(&s->tag);
    if (c0 != key.h[0]) {
        if (c0 != k128_ones()) return 1;
        const K128 old = cas_key(&s->h[0], k128_ones(), key.h[0]);
        if (old == k128_ones()) {                                        // this probe created the entry
            if (W == 2) store_k128(&s->h[W - 1], key.h[W - 1]);
            atomicMin(&s->tag, tag);
            atomicAdd(&s->cnt, 1u);
            return 3;
        }
        if (old != key.h[0]) return 1;
        seen = __ldcg(&s->tag);
    }
    if (W == 2) {
        const K128 c1 = load_k128(&s->h[W - 1]);
        if (c1 == k128_ones()) return 2;
        if (c1 != key.h[W - 1]) return 1;
    }
    if (tag < seen) atomicMin(&s->tag, tag);
    atomicAdd(&s->cnt, 1u);
    return 0;
}

// the chunk table grows while reads are still arriving (streamed host build): every entry of the old
// table is moved into the new one.  Keys are unique, so a probe only looks for an EMPTY slot.
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_chunk_rehash(const ChunkSlot<W> *__restrict__ old_table, u64 n_old,
                                                            ChunkSlot<W> *table, u32 n_slots, ChunkCounters *cc) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_old) return;
    ChunkKey<W> key;
#pragma unroll
    for (int hh = 0; hh < W; hh++) key.h[hh] = old_table[i].h[hh];
    if (key.h[0] == k128_ones()) return;
    u64 idx = slot_home(chunk_hash<W>(key), n_slots);
    for (u32 probes = 0; probes < n_slots; probes++) {
        if (load_k128(&table[idx].h[0]) == k128_ones() &&
            cas_key(&table[idx].h[0], k128_ones(), key.h[0]) == k128_ones()) {
            if (W == 2) table[idx].h[W - 1] = key.h[W - 1];
            table[idx].tag = old_table[i].tag;
            table[idx].cnt = old_table[i].cnt;
            return;
        }
        if (++idx == n_slots) idx = 0;
    }
    cc->chunk_overflow = 1;
}

// ---------------------------------------------------------------------------------------------
// Probe-length statistics (dbg_probe_stats; north-star "probe length" evidence, reported by bench.py): both tables
// probe linearly from a home slot, so an entry's probe length = 1 + the cyclic distance from its home to the slot it
// lives in = the number of slots a lookup of that key touches.  hist[j] counts lengths j + 1 (the last bucket takes
// everything longer); acc[0] = entries, acc[1] = sum of lengths, acc[2] = longest.
// ---------------------------------------------------------------------------------------------
#define PROBE_HIST 64
DBG_D void probe_stats_tally(u32 *sh_hist, bool have, u64 len, u64 *acc) {
    if (have) {
        atomicAdd(&sh_hist[len <= PROBE_HIST ? (u32)len - 1u : PROBE_HIST - 1u], 1u);
        atomicAdd(&acc[1], len);
        atomicMax(&acc[2], len);
    }
}
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_probe_stats_nodes(const Slot<W> *__restrict__ table, TGeom g, u64 *hist, u64 *acc) {
    __shared__ u32 sh_hist[PROBE_HIST];
    for (u32 j = threadIdx.x; j < PROBE_HIST; j += blockDim.x) sh_hist[j] = 0;
    __syncthreads();
    typedef typename KeyOf<W>::type K;
    const K EMPTY = empty_key<W>();
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < g.n_slots; i += (u64)gridDim.x * blockDim.x) {
        const K key = load_key(&table[i]);
        const bool have = key != EMPTY;
        u64 len = 0;
        if (have) {
            const u64 home = node_home(g, key);
            len = (i >= home ? i - home : i + g.n_slots - home) + 1;
        }
        probe_stats_tally(sh_hist, have, len, acc);
    }
    __syncthreads();
    for (u32 j = threadIdx.x; j < PROBE_HIST; j += blockDim.x)
        if (sh_hist[j]) { atomicAdd(&hist[j], (u64)sh_hist[j]); atomicAdd(&acc[0], (u64)sh_hist[j]); }
}
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_probe_stats_chunks(const ChunkSlot<W> *__restrict__ table, u64 n_slots, u64 *hist,
                                                                 u64 *acc) {
    __shared__ u32 sh_hist[PROBE_HIST];
    for (u32 j = threadIdx.x; j < PROBE_HIST; j += blockDim.x) sh_hist[j] = 0;
    __syncthreads();
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (u64)gridDim.x * blockDim.x) {
        ChunkKey<W> key;
#pragma unroll
        for (int hh = 0; hh < W; hh++) key.h[hh] = load_k128(&table[i].h[hh]);
        const bool have = key.h[0] != k128_ones();
        u64 len = 0;
        if (have) {
            const u64 home = slot_home(chunk_hash<W>(key), n_slots);
            len = (i >= home ? i - home : i + n_slots - home) + 1;
        }
        probe_stats_tally(sh_hist, have, len, acc);
    }
    __syncthreads();
    for (u32 j = threadIdx.x; j < PROBE_HIST; j += blockDim.x)
        if (sh_hist[j]) { atomicAdd(&hist[j], (u64)sh_hist[j]); atomicAdd(&acc[0], (u64)sh_hist[j]); }
}

#ifndef DBG_CHUNK_MIN_BLOCKS
#define DBG_CHUNK_MIN_BLOCKS 4
#endif
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK, DBG_CHUNK_MIN_BLOCKS)
k_chunk_insert(const u64 *__restrict__ packed, u64 n_words_padded, const u64 *__restrict__ off, u64 n_reads,
               const u32 *__restrict__ st_mask, const u32 *__restrict__ v_mask, const u32 *__restrict__ span_read,
               u64 n_spans, u64 unit_base, u64 n_units, u64 units_per_block, int k, ChunkSlot<W> *table, u32 n_slots,
               ChunkSlot<W> *list, u64 list_cap, u64 read_idx_base, const u32 *__restrict__ read_batch,
               ChunkCounters *cc) {
    __shared__ ChunkInsertShared<W> sh;
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 lt = (1u << lane) - 1;
    u32 head2 = 0, n2 = 0;                                               // warp-uniform state of the retry ring
    u32 flags = 0;                                                       // 1 table overflow, 2 list overflow, 4 bad chunk
    u64 ub = 0, n_inst = 0;
    u32 n_new = 0;                                                       // entries this lane created = distinct interior chunks

    auto probe = [&](bool act, const ChunkKey<W> &key, u64 tag, u32 probes) {
        int res = 0;
        if (act) {
            u64 idx = slot_home(chunk_hash<W>(key), n_slots) + probes;
            if (idx >= n_slots) idx -= n_slots;
            res = chunk_probe<W>(&table[idx], key, tag);
            if (res == 3) { n_new++; res = 0; }
            if (res == 1) {
                probes++;
                if (probes >= CH_MAX_PROBES || probes >= n_slots) { flags |= 1u; res = 0; cc->chunk_overflow = 1; }
            }
        }
        const bool again = res != 0;
        const u32 mk = __ballot_sync(0xFFFFFFFFu, again);
        if (again) {
            const u32 e = (head2 + n2 + __popc(mk & lt)) & (CH_RCAP - 1);
            sh.rkey[w][e] = key; sh.rtag[w][e] = tag; sh.rprobes[w][e] = (u16)probes;
        }
        n2 += __popc(mk);
    };
    auto drain2 = [&]() {                                                // up to 32 retries
        __syncwarp();
        const u32 take = n2 < 32u ? n2 : 32u;
        const u32 e = (head2 + lane) & (CH_RCAP - 1);
        const bool act = lane < take;
        const ChunkKey<W> key = sh.rkey[w][e];
        const u64 tag = sh.rtag[w][e];
        const u32 probes = sh.rprobes[w][e];
        __syncwarp();
        head2 = (head2 + take) & (CH_RCAP - 1);
        n2 -= take;
        probe(act, key, tag, probes);
    };
    auto fresh = [&](bool act, u32 rel, u32 n, u32 term) {               // up to 32 fresh chunks, one per lane
        ChunkKey<W> key;
        u64 tag = 0;
        bool boundary = false;
        if (act) {
            const u32 sp = rel >> 5, j = rel & 31;
            u64 r = sh.c_read[w][sp];
            u32 pos = sh.c_pos[w][sp] + j;
            const u32 end = sh.c_end[w][sp];
            u32 left = end - j;                                          // bases from the chunk start to its read's end (clamped)
            if (j >= end) {                                              // a read ends inside the span
                const u32 end2 = sh.c_end2[w][sp], skip = sh.c_skip[w][sp];
                if (skip && j < end2) { r += skip; pos = j - end; left = end2 - j; }   // ... and the chunk is in the next one
                else {                                                   // reads shorter than a span: ask off[]
                    const u64 c = ub + rel;
                    u64 re = __ldg(off + r + 1);
                    while (re <= c) { r++; re = __ldg(off + r + 1); }    // also skips empty reads
                    pos = (u32)(c - __ldg(off + r));
                    left = (u32)(re - c < (u64)0xFFFF ? re - c : (u64)0xFFFF);
                }
            }
            // the last node has a successor only if its read goes on for one more base (debruijn.py:126).  The
            // masks say so too whenever a non-node position separates two reads' nodes, i.e. for k >= 2; with
            // k = 1 every base is a node and the next read's first base must not be taken for a successor.
            if (left < n + (u32)k) term = 1u;
            tag = ((read_idx_base + r) << 32) | pos;
            boundary = pos == 0 || term;
            key = chunk_key_from_stream<W>(sh.words[w], rel, (int)n + k - (int)term, chunk_meta(n, term));
            if (read_batch) key.h[W - 1].lo |= (u64)__ldg(read_batch + r) << 8;     // batch id: spare bits past the string
        }
        if (boundary) {                                                  // first / last chunk of a read: its own list slot
            const u64 o = 2 * (tag >> 32) - 2 * read_idx_base + ((u32)tag == 0 ? 0u : 1u);   // 2 x local read index (+ 1: last chunk)
            if (o < list_cap) {
                ChunkSlot<W> rec;
#pragma unroll
                for (int hh = 0; hh < W; hh++) rec.h[hh] = key.h[hh];
                rec.tag = tag; rec.cnt = 1;
                rec.zero_pad();
                list[o] = rec;
            } else flags |= 2u;
        }
        probe(act && !boundary, key, tag, 0);
    };

    const u64 unit_begin = unit_base + (u64)blockIdx.x * units_per_block;       // units [unit_base, n_units)
    u64 unit_end = unit_begin + units_per_block;
    if (unit_end > n_units) unit_end = n_units;
    for (u64 u = unit_begin + w; u < unit_end; u += DBG_WARPS) {
        ub = u * DBG_UNIT;
        if ((u & 15) == 0 && __any_sync(0xFFFFFFFFu, __ldcg(&cc->chunk_overflow) != 0u)) break;   // table full: the host rebuilds it larger
        __syncwarp();
        for (u32 i = lane; i < CH_UNIT_WORDS; i += 32) {
            const u64 wi = ub / 32 + i;
            sh.words[w][i] = wi < n_words_padded ? __ldg(packed + wi) : 0;
        }
        const u64 s = u * 32 + lane;
        u32 st = 0, E = ~0u, En = ~0u, v = 0, vn = 0;
        if (s < n_spans) { st = __ldg(st_mask + s); v = __ldg(v_mask + s); E = st | ~v; }
        if (s + 1 < n_spans) { vn = __ldg(v_mask + s + 1); En = __ldg(st_mask + s + 1) | ~vn; }
        n_inst += __popc(st);
        {                                                                // read bookkeeping of this lane's span
            u32 r0 = 0, pos0 = 0, end = 0, end2 = 0, skip = 0;
            if (st) {
                const u64 p0 = s * 32;
                r0 = __ldg(span_read + s);
                const u64 rs = __ldg(off + r0), re = __ldg(off + r0 + 1);
                pos0 = (u32)(p0 - rs);
                end = (u32)(re - p0 < (u64)0xFFFF ? re - p0 : (u64)0xFFFF);
                if (end < 32u) {                                         // the span runs into the next read(s)
                    u64 r2 = (u64)r0 + 1, re2 = re;
                    while (r2 < n_reads && skip < 200u) {
                        re2 = __ldg(off + r2 + 1);
                        skip++;
                        if (re2 > re) break;                             // first non-empty read after r0
                        r2++;
                    }
                    if (re2 > re && skip < 200u) end2 = (u32)(re2 - p0 < (u64)0xFFFF ? re2 - p0 : (u64)0xFFFF);
                    else skip = 0;
                }
            }
            sh.c_read[w][lane] = r0; sh.c_pos[w][lane] = pos0; sh.c_end[w][lane] = (u16)end;
            sh.c_end2[w][lane] = (u16)end2; sh.c_skip[w][lane] = (u8)skip;
        }
        __syncwarp();
        // ---- hand the unit's chunk instances to the lanes, 32 at a time.  Instance g belongs to the span (lane) s with
        // excl[s] <= g < incl[s] (prefix sums of popc(St), found by a 5-step search over shuffles) and starts at the
        // (g - excl[s])-th set bit of that span's St; its end is the next bit of St | ~V, in that span or the next.
        // (Each lane used to walk its own span's bits and queue them: a loop at 1/4 lane occupancy that was a quarter
        // of this kernel's instructions.)
        {
            const u32 c = __popc(st);
            u32 incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u32 x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= (u32)d) incl += x;
            }
            const u32 total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            const u32 excl = incl - c;
            for (u32 base = 0; base < total; base += 32u) {
                const u32 g = base + lane;
                const bool act = g < total;
                u32 lo = 0;                                              // first lane whose inclusive sum exceeds g
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const u32 x = __shfl_sync(0xFFFFFFFFu, incl, (int)((lo + step - 1) & 31u));
                    if (x <= g) lo += (u32)step;
                }
                const int src = (int)(lo & 31u);
                const u32 ex_s = __shfl_sync(0xFFFFFFFFu, excl, src);
                const u32 st_s = __shfl_sync(0xFFFFFFFFu, st, src);
                const u32 E_s = __shfl_sync(0xFFFFFFFFu, E, src), En_s = __shfl_sync(0xFFFFFFFFu, En, src);
                const u32 v_s = __shfl_sync(0xFFFFFFFFu, v, src), vn_s = __shfl_sync(0xFFFFFFFFu, vn, src);
                u32 b = 0, n = 1, term = 1;
                if (act) {
                    b = nth_set_bit(st_s, g - ex_s);
                    const u32 rest = E_s & ~((2u << b) - 1u);            // bits above b
                    u32 e = 64;                                          // e = 64: no end found (flagged below)
                    if (rest) { e = (u32)__ffs((int)rest) - 1; term = ((v_s >> e) & 1u) ^ 1u; }
                    else if (En_s) { const u32 en = (u32)__ffs((int)En_s) - 1; e = 32 + en; term = ((vn_s >> en) & 1u) ^ 1u; }
                    n = e - b;
                    if (n > CH_WMAX || n == 0) { flags |= 4u; n = 1; term = 1; }
                }
                fresh(act, (u32)src * 32u + b, n, term);
                while (n2 >= 32u) drain2();
            }
        }
    }
    while (n2) drain2();
    for (int d = 16; d > 0; d >>= 1) n_inst += __shfl_down_sync(0xFFFFFFFFu, n_inst, d);
    if (lane == 0 && n_inst) atomicAdd(&cc->n_instances, n_inst);
    for (int d = 16; d > 0; d >>= 1) n_new += __shfl_down_sync(0xFFFFFFFFu, n_new, d);
    if (lane == 0 && n_new) atomicAdd(&cc->n_first, (u64)n_new);
    if (flags & 1u) cc->chunk_overflow = 1;
    if (flags & 2u) cc->list_overflow = 1;
    if (flags & 4u) cc->bad_chunk = 1;
}

// ---------------------------------------------------------------------------------------------
// k_expand (phase 2).  Records = the chunk table's slots followed by the boundary list.  Each lane
// takes one record and keeps its string in registers.
//   MODE 0  rolls the node key along the string (one shift per node, as the reference's window loop
//           slides, debruijn.py:127-128), feeds a HyperLogLog sketch of the node keys (sizes the node
//           table) and leaves ONE BYTE per (record, node) = top 8 bits of the node's slot hash
//           (255 = no node), 16 bytes per record
//   MODE 2  MODE 1 without the hash bytes of a sketch pass: the lane hashes its record's nodes itself (the
//           build that knows its table sizes from the previous one on the handle skips MODE 0 altogether)
//   MODE 1  node-table pass: owns the hash bytes [lo8, lo8 + span8], whose home slots form a table
//           region small enough to stay L2-resident.  A lane compares its record's 16 bytes at once
//           (SIMD byte compares) and only rebuilds the keys of the nodes in range; those are queued in
//           a ring and probed 32 at a time -- one probe each, a miss re-queues (as k_insert does) --
//           so the table traffic runs with full warps although few nodes of a record are in range.
// ---------------------------------------------------------------------------------------------
#define EX_RCAP 128
#ifndef EX_ILP
#define EX_ILP 2                           // entries per lane per drain
#endif
template <int W> struct ExpandShared {
    typename KeyOf<W>::type rkey[DBG_WARPS][EX_RCAP];
    u64 rtag[DBG_WARPS][EX_RCAP];
    u32 rcnt[DBG_WARPS][EX_RCAP];
    u16 rmeta[DBG_WARPS][EX_RCAP];            // probes so far << 3 | has_next << 2 | next base
    u32 rhome[DBG_WARPS][EX_RCAP];            // home slot
};

// k bases at base j (< 32) of a chunk string held in registers (2 W words), right-aligned
template <int W> DBG_D typename KeyOf<W>::type reg_key(const u64 *cw, u32 j, int k);
template <> DBG_D u64 reg_key<1>(const u64 *cw, u32 j, int k) {
    const int s = 2 * (int)j;
    const u64 top = s ? ((cw[0] << s) | (cw[1] >> (64 - s))) : cw[0];
    return top >> (64 - 2 * k);
}
template <> DBG_D K128 reg_key<2>(const u64 *cw, u32 j, int k) {
    const int s = 2 * (int)j;
    const u64 hi = s ? ((cw[0] << s) | (cw[1] >> (64 - s))) : cw[0];
    const u64 lo = s ? ((cw[1] << s) | (cw[2] >> (64 - s))) : cw[1];
    const int sh = 128 - 2 * k;                                          // 2 <= sh <= 126
    K128 r;
    if (sh >= 64) { r.lo = sh == 64 ? hi : hi >> (sh - 64); r.hi = 0; }
    else { r.lo = (lo >> sh) | (hi << (64 - sh)); r.hi = hi >> sh; }
    return r;
}
// base q of a chunk string held in registers
template <int W> DBG_D u32 string_base(const u64 *cw, u32 q) {
    u64 word = cw[0];
    if (q >= 32) word = cw[1];
    if (W == 2 && q >= 64) word = cw[2];
    return (u32)(word >> (62 - 2 * (q & 31))) & 3u;
}
// bit b = byte b of x lies in [lo, hi] (lo4 / hi4 = the bound in every byte)
DBG_D u32 bytes_in_range(u32 x, u32 lo4, u32 hi4) {
    const u32 t = __vcmpgeu4(x, lo4) & __vcmpleu4(x, hi4) & 0x01010101u;
    return (t * 0x10204080u) >> 28;
}

#ifndef DBG_EXPAND_MIN_BLOCKS
#define DBG_EXPAND_MIN_BLOCKS 4
#endif
#ifndef EX_NP
#define EX_NP 2                            // MODE 3: slots of the probe sequence read per round
#endif

template <int W, int MODE>
__global__ void __launch_bounds__(DBG_BLOCK, DBG_EXPAND_MIN_BLOCKS)
k_expand(const ChunkSlot<W> *__restrict__ recs_a, u64 n_a, const ChunkSlot<W> *__restrict__ recs_b, u64 n_b, int k,
         uint4 *__restrict__ hbytes, Slot<W> *table, u32 n_slots_, u32 lo8, u32 span8, u32 *regs, int batch_bits,
         DevCounters *ctr, const u64 *__restrict__ order, const u32 *__restrict__ n_order, TGeom tg) {
    typedef typename KeyOf<W>::type K;
    __shared__ ExpandShared<W> sh;
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 lt = (1u << lane) - 1;
    const K mask = key_mask<W>(k);
    const u32 lo4 = lo8 * 0x01010101u, hi4 = (lo8 + span8) * 0x01010101u;
    u32 head = 0, nq = 0;
    bool overflow = false;
    const u64 n_slots = MODE == 3 ? tg.n_slots : (u64)n_slots_;

    // up to 32 * EX_ILP entries at a time: every lane issues the slot loads of all its entries before it
    // looks at the first one, then the CAS of every entry that met an EMPTY slot before it looks at the first
    // result, so a round costs two memory round trips whatever EX_ILP is.  With the table laid out by region
    // (MODE 3) the NEXT slot of the probe sequence is read in the same round trip (same or neighbouring
    // sector).  A probe that finds other keys does not loop: the entry goes to the retry ring (shared memory)
    // and comes back in a later drain.
    constexpr int NP = MODE == 3 ? EX_NP : 1;                                // slots looked at per round
    auto probe_entries = [&](const bool (&act)[EX_ILP], K (&key)[EX_ILP], u64 (&tag)[EX_ILP], u32 (&cnt)[EX_ILP],
                             u32 (&meta)[EX_ILP], u32 (&home)[EX_ILP]) {
        Slot<W> *s[EX_ILP]; K cur[EX_ILP][NP]; u64 seen[EX_ILP][NP];
        u32 first[EX_ILP];
#pragma unroll
        for (int i = 0; i < EX_ILP; i++) {
            u64 idx = (u64)home[i] + (meta[i] >> 3);
            if (idx >= n_slots) idx %= n_slots;                          // (measured: a subtract loop here is 50 us SLOWER at C2)
            if (!act[i]) idx = 0;
            first[i] = (u32)idx;
#pragma unroll
            for (int q = 0; q < NP; q++) {
                u64 iq = idx + q;
                if (iq >= n_slots) iq -= n_slots;
                cur[i][q] = load_key_tag(&table[iq], &seen[i][q]);
            }
        }
        int which[EX_ILP];                                               // slot of the round that is, or may become, this key's
        bool cas[EX_ILP];
        K old[EX_ILP];
#pragma unroll
        for (int i = 0; i < EX_ILP; i++) {
            which[i] = -1; cas[i] = false;
#pragma unroll
            for (int q = NP - 1; q >= 0; q--) {                          // the first slot that holds the key or is EMPTY,
                bool earlier_other = true;                               // provided every slot before it holds another key
#pragma unroll
                for (int p = 0; p < q; p++) earlier_other = earlier_other && cur[i][p] != key[i] && cur[i][p] != empty_key<W>();
                if (earlier_other && (cur[i][q] == key[i] || cur[i][q] == empty_key<W>())) { which[i] = q; cas[i] = cur[i][q] != key[i]; }
            }
            if (!act[i]) { which[i] = -1; cas[i] = false; }
            u64 it = (u64)first[i] + (which[i] > 0 ? (u32)which[i] : 0u);
            if (it >= n_slots) it -= n_slots;
            s[i] = &table[it];
            old[i] = empty_key<W>();
            if (cas[i]) old[i] = cas_key(&s[i]->key, empty_key<W>(), key[i]);
        }
#pragma unroll
        for (int i = 0; i < EX_ILP; i++) {
            bool again = false;
            if (act[i]) {
                bool hit = which[i] >= 0;
                u64 sn = which[i] > 0 ? seen[i][NP - 1] : seen[i][0];
                if (cas[i]) {
                    if (old[i] == empty_key<W>()) sn = ~0ULL;                     // this probe created the node
                    else if (old[i] == key[i]) sn = __ldcg(&s[i]->tag);
                    else hit = false;                                            // somebody else's key got there first
                }
                if (hit) {
                    if (tag[i] < sn) atomicMin(&s[i]->tag, tag[i]);                  // first occurrence
                    if (meta[i] & 4u) atomicAdd(&s[i]->cnt[meta[i] & 3u], cnt[i]);   // debruijn.py:132/136, x multiplicity
                } else {
                    meta[i] += 8u * (which[i] >= 0 ? (u32)which[i] + 1u : (u32)NP);  // on past the slots that hold other keys
                    if ((meta[i] >> 3) >= CH_MAX_PROBES) { overflow = true; ctr->overflow = 1; } else again = true;
                }
            }
            const u32 mk = __ballot_sync(0xFFFFFFFFu, again);
            if (again) {
                const u32 o = (head + nq + __popc(mk & lt)) & (EX_RCAP - 1);
                sh.rkey[w][o] = key[i]; sh.rtag[w][o] = tag[i]; sh.rcnt[w][o] = cnt[i]; sh.rmeta[w][o] = (u16)meta[i];
                sh.rhome[w][o] = home[i];
            }
            nq += __popc(mk);
        }
    };
    auto drain = [&]() {                                                 // retries out of the ring
        __syncwarp();
        const u32 cap = 32u * EX_ILP;
        const u32 take = nq < cap ? nq : cap;
        bool act[EX_ILP]; K key[EX_ILP]; u64 tag[EX_ILP]; u32 cnt[EX_ILP]; u32 meta[EX_ILP]; u32 home[EX_ILP];
#pragma unroll
        for (int i = 0; i < EX_ILP; i++) {
            const u32 e = (head + 32 * i + lane) & (EX_RCAP - 1);
            act[i] = 32 * i + lane < take;
            key[i] = sh.rkey[w][e]; tag[i] = sh.rtag[w][e]; cnt[i] = sh.rcnt[w][e]; meta[i] = sh.rmeta[w][e];
            home[i] = sh.rhome[w][e];
        }
        __syncwarp();
        head = (head + take) & (EX_RCAP - 1);
        nq -= take;
        probe_entries(act, key, tag, cnt, meta, home);
    };

    // MODE 3: the records in region order
    const u64 n_total = MODE == 3 ? (u64)__ldg(n_order) : n_a + n_b;
    const u64 n_batches = (n_total + 31) / 32;
    const u64 warps_total = (u64)gridDim.x * DBG_WARPS;
    for (u64 b = (u64)blockIdx.x * DBG_WARPS + w; b < n_batches; b += warps_total) {
        // the table is full: the host rebuilds it larger (looked at now and then; the flag is set at once)
        if (MODE != 0 && ((b / warps_total) & 7) == 0 && __any_sync(0xFFFFFFFFu, __ldcg(&ctr->overflow) != 0u)) break;
        const u64 idx = b * 32 + lane;
        u32 n = 0, B = 0, cnt = 0, m16 = 0, batch = 0, vh = 0;
        u64 tag = 0;
        u64 cw[2 * W];
#pragma unroll
        for (int i = 0; i < 2 * W; i++) cw[i] = 0;
        if (idx < n_total) {
            u64 oe = 0;
            if (MODE == 3) oe = __ldg(order + idx);
            u64 ridx = idx;
            if (MODE == 3) { ridx = (u32)oe; vh = (u32)(oe >> 32); }
            const ChunkSlot<W> *rec = ridx < n_a ? recs_a + ridx : recs_b + (ridx - n_a);
            K128 h[W];
#pragma unroll
            for (int hh = 0; hh < W; hh++) h[hh] = load_k128(&rec->h[hh]);
            if (h[0] != k128_ones()) {
                const u32 meta = (u32)(h[0].lo & 0xFFu);
                n = meta_nodes(meta);
                B = n + (u32)k - meta_term(meta);
                chunk_words<W>(h, cw);
                if (batch_bits) batch = (u32)(h[W - 1].lo >> 8) & (u32)((1ull << batch_bits) - 1);
                if (MODE == 1) {
                    const uint4 hb = __ldg(hbytes + idx);
                    m16 = bytes_in_range(hb.x, lo4, hi4) | (bytes_in_range(hb.y, lo4, hi4) << 4) |
                          (bytes_in_range(hb.z, lo4, hi4) << 8) | (bytes_in_range(hb.w, lo4, hi4) << 12);
                }
                if ((MODE == 2 && span8 >= 254u) || MODE == 3) m16 = (1u << n) - 1u;   // one pass owns every node
                if (MODE == 2 || MODE == 3 || m16) {
                    const ulonglong2 tc = __ldcg(reinterpret_cast<const ulonglong2 *>(&rec->tag));   // tag | cnt, pad
                    tag = tc.x;
                    cnt = (u32)tc.y;
                }
            }
        }
        if (MODE == 2 && span8 < 254u) {
            // no hash bytes were left by a sketch pass (the table was sized from the previous build on this handle):
            // roll the node key along the string and pick the nodes of this pass by their hash byte
            K key = key_add_batch(reg_key<W>(cw, 0, k), batch, k);
            u32 nmax = n;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const u32 x = __shfl_xor_sync(0xFFFFFFFFu, nmax, d);
                nmax = x > nmax ? x : nmax;
            }
#pragma unroll
            for (int j = 0; j < CH_WMAX; j++) {
                if ((u32)j >= nmax) break;
                const u32 hsh = hash32_key(key);
                const u32 byte = (hsh >> 24) > 254u ? 254u : (hsh >> 24);
                if ((u32)j < n && byte - lo8 <= span8) m16 |= 1u << j;
                key = key_roll(key, string_base<W>(cw, (u32)(j + k)), mask);
            }
        }
        if (MODE == 0) {                                                  // (plain hash layout only)
            K key = key_add_batch(reg_key<W>(cw, 0, k), batch, k);
            u32 nmax = n;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const u32 x = __shfl_xor_sync(0xFFFFFFFFu, nmax, d);
                nmax = x > nmax ? x : nmax;
            }
            u32 hb[4] = {~0u, ~0u, ~0u, ~0u};
#pragma unroll
            for (int j = 0; j < CH_WMAX; j++) {
                if ((u32)j >= nmax) break;
                const u32 nb = string_base<W>(cw, (u32)(j + k));          // garbage past the string: unused then
                const u32 hsh = hash32_key(key);
                if ((u32)j < n) {
                    const u32 ri = hsh >> (32 - HLL_P);
                    const u32 rank = (u32)__clz((int)((hsh << HLL_P) | (1u << (HLL_P - 1)))) + 1;
                    if (regs[ri] < rank) atomicMax(&regs[ri], rank);
                    const u32 byte = (hsh >> 24) > 254u ? 254u : (hsh >> 24);
                    hb[j >> 2] = (hb[j >> 2] & ~(0xFFu << (8 * (j & 3)))) | (byte << (8 * (j & 3)));
                }
                key = key_roll(key, nb, mask);
            }
            if (idx < n_total) hbytes[idx] = make_uint4(hb[0], hb[1], hb[2], hb[3]);
        } else {
            // ---- hand the selected (record, node) pairs to the lanes, 32 * EX_ILP at a time.  No lane walks its own
            // record's nodes one by one (that loop ran at 1/4 lane occupancy and was half of this kernel's
            // instructions): pair g belongs to the lane s with excl[s] <= g < incl[s] (prefix sums of the per-record
            // counts, found by a 5-step search over shuffles) and is the (g - excl[s])-th set bit of that lane's mask;
            // the record's words, tag and multiplicity come over by shuffle and the node key is cut in registers.
            const u32 c = __popc(m16);
            u32 incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u32 x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= (u32)d) incl += x;
            }
            const u32 total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            const u32 excl = incl - c;
            const u32 packed = m16 | (B << 16);
            for (u32 base = 0; base < total; base += 32u * EX_ILP) {
                bool act[EX_ILP]; K key[EX_ILP]; u64 tg_[EX_ILP]; u32 cn[EX_ILP]; u32 meta[EX_ILP]; u32 home[EX_ILP];
#pragma unroll
                for (int i = 0; i < EX_ILP; i++) {
                    const u32 g = base + 32u * i + lane;
                    act[i] = g < total;
                    u32 lo = 0;                                          // first lane whose inclusive sum exceeds g
#pragma unroll
                    for (int step = 16; step > 0; step >>= 1) {
                        const u32 v = __shfl_sync(0xFFFFFFFFu, incl, (int)((lo + step - 1) & 31u));
                        if (v <= g) lo += (u32)step;
                    }
                    const int src = (int)(lo & 31u);
                    const u32 ex_s = __shfl_sync(0xFFFFFFFFu, excl, src);
                    const u32 pm = __shfl_sync(0xFFFFFFFFu, packed, src);
                    u64 cws[2 * W];
#pragma unroll
                    for (int q = 0; q < 2 * W; q++) cws[q] = __shfl_sync(0xFFFFFFFFu, cw[q], src);
                    const u64 tag_s = __shfl_sync(0xFFFFFFFFu, tag, src);
                    const u32 cnt_s = __shfl_sync(0xFFFFFFFFu, cnt, src);
                    const u32 batch_s = batch_bits ? __shfl_sync(0xFFFFFFFFu, batch, src) : 0u;
                    const u32 vh_s = MODE == 3 ? __shfl_sync(0xFFFFFFFFu, vh, src) : 0u;
                    const u32 j = act[i] ? nth_set_bit(pm & 0xFFFFu, g - ex_s) : 0u;
                    const u32 q = j + (u32)k;
                    const bool has_next = q < (pm >> 16);
                    const u32 nb = string_base<W>(cws, q);
                    key[i] = key_add_batch(reg_key<W>(cws, j, k), batch_s, k);
                    tg_[i] = tag_s + j;
                    cn[i] = cnt_s;
                    meta[i] = (has_next ? 4u : 0u) | (has_next ? nb : 0u);
                    home[i] = MODE == 3 ? (u32)region_home(tg, vh_s, hash32_key(key[i])) : (u32)slot_home(hash32_key(key[i]), n_slots);
                }
                probe_entries(act, key, tg_, cn, meta, home);
                while (nq >= 32u * EX_ILP) drain();
            }
        }
    }
    if (MODE != 0) {
        while (nq) drain();
        if (overflow) ctr->overflow = 1;
    }
}

// ---------------------------------------------------------------------------------------------
// Region order of the distinct chunks (node table laid out by minimizer, TGeom in dbg_common.cuh).
//   k_rec_vhash    one thread per record slot (chunk table + boundary list): vhash = node_vhash of the record's
//                  first node = of all its nodes; histogram over virtual buckets = top vb_bits of vhash (a
//                  bucketing that does not depend on the node table's size, which is not known yet when the
//                  sketch runs).  SKETCH also rolls the node key along the string and feeds the HyperLogLog
//                  sketch of the node keys that sizes the node table.
//   (exclusive scan of the histogram -> offs[]; offs[2^vb_bits] = number of records)
//   k_rec_order    order[offs[b] + --hist[b]] = vhash << 32 | record index: the records in bucket order, which
//                  is region order for ANY region count (region = vhash * n_regions >> 32 is monotone in vhash).
//                  Any order is correct; this one keeps k_expand's table traffic in L2.
// ---------------------------------------------------------------------------------------------
template <int W, bool SKETCH>
__global__ void __launch_bounds__(DBG_BLOCK) k_rec_vhash(const ChunkSlot<W> *__restrict__ recs_a, u64 n_a,
                                                         const ChunkSlot<W> *__restrict__ recs_b, u64 n_b, int k, int w,
                                                         int m, int batch_bits, int vb_bits, u32 *__restrict__ rec_vhash,
                                                         u32 *vhist, u32 *regs) {
    typedef typename KeyOf<W>::type K;
    const u64 n_total = n_a + n_b;
    const K mask = key_mask<W>(k);
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < n_total; idx += (u64)gridDim.x * blockDim.x) {
        const ChunkSlot<W> *rec = idx < n_a ? recs_a + idx : recs_b + (idx - n_a);
        K128 h[W];
#pragma unroll
        for (int hh = 0; hh < W; hh++) h[hh] = load_k128(&rec->h[hh]);
        u32 vh = NONE32;
        if (h[0] != k128_ones()) {
            u64 cw[2 * W];
            chunk_words<W>(h, cw);
            const u32 batch = batch_bits ? (u32)(h[W - 1].lo >> 8) & (u32)((1ull << batch_bits) - 1) : 0u;
            K key = key_add_batch(reg_key<W>(cw, 0, k), batch, k);
            vh = node_vhash(key, k, w, m);
            if (vh == NONE32) vh = NONE32 - 1;                           // NONE32 marks an empty record slot
            atomicAdd(&vhist[vh >> (32 - vb_bits)], 1u);
            if (SKETCH) {
                const u32 n = meta_nodes((u32)(h[0].lo & 0xFFu));
                for (u32 j = 0; j < n; j++) {
                    const u32 hsh = hash32_key(key);
                    const u32 ri = hsh >> (32 - HLL_P);
                    const u32 rank = (u32)__clz((int)((hsh << HLL_P) | (1u << (HLL_P - 1)))) + 1;
                    if (regs[ri] < rank) atomicMax(&regs[ri], rank);
                    key = key_roll(key, string_base<W>(cw, j + (u32)k), mask);   // garbage past the string: unused then
                }
            }
        }
        rec_vhash[idx] = vh;
    }
}

__global__ void __launch_bounds__(DBG_BLOCK) k_rec_order(const u32 *__restrict__ rec_vhash, u64 n_total, int vb_bits,
                                                         u32 *vhist, const u32 *__restrict__ offs, u64 *__restrict__ order) {
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < n_total; idx += (u64)gridDim.x * blockDim.x) {
        const u32 vh = __ldg(rec_vhash + idx);
        if (vh == NONE32) continue;
        const u32 b = vh >> (32 - vb_bits);
        const u32 o = __ldg(offs + b) + (atomicSub(&vhist[b], 1u) - 1u);  // the bucket fills from its end; its count returns to 0
        order[o] = ((u64)vh << 32) | idx;
    }
}

// ---------------------------------------------------------------------------------------------
// k_order_chunks: first appearance of every (node, successor base) edge of the multi-successor nodes
// (Counter.most_common tie order, debruijn.py:159-164) from the DISTINCT chunks instead of a second
// pass over all windows.  A window at offset j of a chunk instance tagged (read, pos) has tag
// (read, pos + j), so over all instances of one chunk the earliest window at offset j is chunk.tag + j
// (the chunk's tag is the minimum over its instances); the minimum over the chunks that hold the edge
// is its first appearance.  Same side set + bitmap as k_order_scan.  Measured at C3 (1 % errors, 0.75 M
// multi-successor nodes): 1.22 ms instead of 2.47.  (The same idea for pull_out_read -- mark the chunks
// that hold a branch node, then look every read's chunk instances up -- measured SLOWER than the
// re-scan, 1.63 vs 1.43 ms, and was dropped.)
// ---------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_order_chunks(const ChunkSlot<W> *__restrict__ recs_a, u64 n_a,
                                                            const ChunkSlot<W> *__restrict__ recs_b, u64 n_b, int k,
                                                            int batch_bits, SideEntry<W> *st, u64 cap,
                                                            const u32 *__restrict__ bm, int bits) {
    typedef typename KeyOf<W>::type K;
    const u64 n_total = n_a + n_b;
    const K mask = key_mask<W>(k);
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < n_total; idx += (u64)gridDim.x * blockDim.x) {
        const ChunkSlot<W> *rec = idx < n_a ? recs_a + idx : recs_b + (idx - n_a);
        K128 h[W];
#pragma unroll
        for (int hh = 0; hh < W; hh++) h[hh] = load_k128(&rec->h[hh]);
        if (h[0] == k128_ones()) continue;
        const u32 meta = (u32)(h[0].lo & 0xFFu);
        const u32 n = meta_nodes(meta), B = n + (u32)k - meta_term(meta);
        u64 cw[2 * W];
        chunk_words<W>(h, cw);
        const u32 batch = batch_bits ? (u32)(h[W - 1].lo >> 8) & (u32)((1ull << batch_bits) - 1) : 0u;
        const u64 tag = __ldcg(&rec->tag);
        K key = key_add_batch(reg_key<W>(cw, 0, k), batch, k);
        for (u32 j = 0; j < n; j++) {
            const u32 q = j + (u32)k;
            if (q >= B) break;                                           // the last node of a read: no edge
            const u32 nb = string_base<W>(cw, q);
            if (bitmap_test(bm, bits, hash32_key(key))) {
                const u64 e = side_locate<W>(st, cap, key, false);
                if (e != NONE64 && tag + j < __ldcg(&st[e].first[nb])) atomicMin(&st[e].first[nb], tag + j);
            }
            key = key_roll(key, nb, mask);
        }
    }
}

// min_count (extension, see k_mc_filter in dbg_kernels.cuh): first-occurrence tags of the nodes that surviving
// windows touch, from the distinct chunks.  The counters in the table are already filtered (0 = the window's
// (k+1)-mer was too rare).
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_mc_retag(const ChunkSlot<W> *__restrict__ recs_a, u64 n_a,
                                                        const ChunkSlot<W> *__restrict__ recs_b, u64 n_b, int k,
                                                        int batch_bits, Slot<W> *table, TGeom tg, DevCounters *ctr) {
    typedef typename KeyOf<W>::type K;
    const u64 n_total = n_a + n_b;
    const K mask = key_mask<W>(k);
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < n_total; idx += (u64)gridDim.x * blockDim.x) {
        const ChunkSlot<W> *rec = idx < n_a ? recs_a + idx : recs_b + (idx - n_a);
        K128 h[W];
#pragma unroll
        for (int hh = 0; hh < W; hh++) h[hh] = load_k128(&rec->h[hh]);
        if (h[0] == k128_ones()) continue;
        const u32 meta = (u32)(h[0].lo & 0xFFu);
        const u32 n = meta_nodes(meta), B = n + (u32)k - meta_term(meta);
        u64 cw[2 * W];
        chunk_words<W>(h, cw);
        const u32 batch = batch_bits ? (u32)(h[W - 1].lo >> 8) & (u32)((1ull << batch_bits) - 1) : 0u;
        const u64 tag = __ldcg(&rec->tag);
        K key = key_add_batch(reg_key<W>(cw, 0, k), batch, k);
        bool touched = false;                                            // the window before this node survived
        for (u32 j = 0; j < n; j++) {
            const u32 q = j + (u32)k;
            const bool has_next = q < B;                                 // (the last node of a read: no window)
            const u32 nb = has_next ? string_base<W>(cw, q) : 0u;
            const u64 s = table_find<W>(table, tg, key);
            if (s == NONE64) { raise_fault(ctr, FAULT_SUCC_MISSING, (u32)idx); break; }
            const bool alive = has_next && __ldcg(&table[s].cnt[nb]) != 0u;
            if ((alive || touched) && tag + j < __ldcg(&table[s].tag)) atomicMin(&table[s].tag, tag + j);
            touched = alive;
            if (!has_next) break;
            key = key_roll(key, nb, mask);
        }
        if (touched) {                                                   // the successor of the chunk's last node: another chunk's node
            const u64 s = table_find<W>(table, tg, key);
            if (s == NONE64) raise_fault(ctr, FAULT_SUCC_MISSING, (u32)idx);
            else if (tag + n < __ldcg(&table[s].tag)) atomicMin(&table[s].tag, tag + n);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Chunk records: the unit of the multi-GPU exchange when it happens BEFORE any node table exists.
// A rank's window loop leaves its distinct chunks {string | min tag | multiplicity} (chunk table + the
// boundary list); merging them is associative (min tag, summed multiplicity).  The owner of a chunk is a
// function of its MINIMIZER -- the m-mer all its nodes share, and which every occurrence of any of its
// nodes has wherever it turns up -- so all contributions to one node meet on one rank, which expands
// only what it owns.  The owner function must pick exactly the minimizer k_masks picked: same hash,
// leftmost on ties.
// ---------------------------------------------------------------------------------------------
template <int W>
DBG_D u32 chunk_owner(const u64 *cw, int w, int m, u32 batch, u32 n_parts) {
    u32 best = 0xFFFFFFFFu, best_x = 0;
    for (int i = 0; i < w; i++) {                              // the first node's w candidate m-mers
        const int s = 2 * i;                                   // i < 16: inside the first two words
        const u64 t = s ? ((cw[0] << s) | (cw[1] >> (64 - s))) : cw[0];
        const u32 x = (u32)(t >> 32) >> (32 - 2 * m);
        const u32 g = ((x * 0x9E3779B1u) | 63u) - (u32)(63 - i);          // exactly k_masks' order: hash, then leftmost
        if (g < best) { best = g; best_x = x; }
    }
    const u32 hsh = hash32_key((u64)best_x | ((u64)batch << 32));
    return (u32)(((u64)hsh * n_parts) >> 32);
}

template <int W>
DBG_D bool chunk_record_load(const ChunkSlot<W> *recs_a, u64 n_a, const ChunkSlot<W> *recs_b, u64 idx, K128 *h) {
    const ChunkSlot<W> *rec = idx < n_a ? recs_a + idx : recs_b + (idx - n_a);
#pragma unroll
    for (int hh = 0; hh < W; hh++) h[hh] = load_k128(&rec->h[hh]);
    return h[0] != k128_ones();
}

template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_chunk_rec_count(const ChunkSlot<W> *__restrict__ recs_a, u64 n_a,
                                                               const ChunkSlot<W> *__restrict__ recs_b, u64 n_b,
                                                               int w, int m, int batch_bits, u32 n_parts, u64 *counts) {
    __shared__ u32 s_cnt[DBG_MAX_PARTS];
    for (u32 p = threadIdx.x; p < n_parts; p += blockDim.x) s_cnt[p] = 0;
    __syncthreads();
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    K128 h[W];
    if (idx < n_a + n_b && chunk_record_load<W>(recs_a, n_a, recs_b, idx, h)) {
        u64 cw[2 * W];
        chunk_words<W>(h, cw);
        const u32 batch = batch_bits ? (u32)(h[W - 1].lo >> 8) & (u32)((1ull << batch_bits) - 1) : 0u;
        atomicAdd(&s_cnt[chunk_owner<W>(cw, w, m, batch, n_parts)], 1u);
    }
    __syncthreads();
    for (u32 p = threadIdx.x; p < n_parts; p += blockDim.x)
        if (s_cnt[p]) atomicAdd(&counts[p], (u64)s_cnt[p]);
}

template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_chunk_rec_scatter(const ChunkSlot<W> *__restrict__ recs_a, u64 n_a,
                                                                 const ChunkSlot<W> *__restrict__ recs_b, u64 n_b,
                                                                 int w, int m, int batch_bits, u32 n_parts,
                                                                 u64 *cursors, ChunkSlot<W> *__restrict__ out) {
    __shared__ u32 s_cnt[DBG_MAX_PARTS];
    __shared__ u64 s_base[DBG_MAX_PARTS];
    for (u32 p = threadIdx.x; p < n_parts; p += blockDim.x) s_cnt[p] = 0;
    __syncthreads();
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    K128 h[W];
    u32 owner = 0, local = 0;
    const bool have = idx < n_a + n_b && chunk_record_load<W>(recs_a, n_a, recs_b, idx, h);
    if (have) {
        u64 cw[2 * W];
        chunk_words<W>(h, cw);
        const u32 batch = batch_bits ? (u32)(h[W - 1].lo >> 8) & (u32)((1ull << batch_bits) - 1) : 0u;
        owner = chunk_owner<W>(cw, w, m, batch, n_parts);
        local = atomicAdd(&s_cnt[owner], 1u);
    }
    __syncthreads();
    for (u32 p = threadIdx.x; p < n_parts; p += blockDim.x)
        if (s_cnt[p]) s_base[p] = atomicAdd(&cursors[p], (u64)s_cnt[p]);
    __syncthreads();
    if (have) {
        const ChunkSlot<W> *rec = idx < n_a ? recs_a + idx : recs_b + (idx - n_a);
        ChunkSlot<W> o;
#pragma unroll
        for (int hh = 0; hh < W; hh++) o.h[hh] = h[hh];
        o.tag = __ldcg(&rec->tag);
        o.cnt = __ldcg(&rec->cnt);
        o.zero_pad();
        out[s_base[owner] + local] = o;
    }
}

// merge received chunk records into this rank's chunk table: multiplicities add, tags take the minimum
template <int W>
__global__ void __launch_bounds__(DBG_BLOCK) k_chunk_rec_merge(ChunkSlot<W> *table, u32 n_slots,
                                                               const ChunkSlot<W> *__restrict__ recs, u64 n,
                                                               ChunkCounters *cc) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ChunkKey<W> key;
#pragma unroll
    for (int hh = 0; hh < W; hh++) key.h[hh] = load_k128(&recs[i].h[hh]);
    const u64 tag = __ldcg(&recs[i].tag);
    const u32 cnt = __ldcg(&recs[i].cnt);
    u64 idx = slot_home(chunk_hash<W>(key), n_slots);
    for (u32 probes = 0; probes < CH_MAX_PROBES && probes < n_slots;) {
        ChunkSlot<W> *s = &table[idx];
        const K128 c0 = load_k128(&s->h[0]);
        bool mine = c0 == key.h[0];
        if (!mine && c0 == k128_ones()) {
            const K128 old = cas_key(&s->h[0], k128_ones(), key.h[0]);
            if (old == k128_ones()) { if (W == 2) store_k128(&s->h[W - 1], key.h[W - 1]); mine = true; }
            else mine = old == key.h[0];
        }
        if (mine && W == 2) {
            const K128 c1 = load_k128(&s->h[W - 1]);
            if (c1 == k128_ones()) continue;                              // being published: look again
            mine = c1 == key.h[W - 1];
        }
        if (mine) {
            atomicMin(&s->tag, tag);
            atomicAdd(&s->cnt, cnt);
            return;
        }
        probes++;
        if (++idx == n_slots) idx = 0;
    }
    cc->chunk_overflow = 1;
}
