// dbg_common.cuh -- key types, hashing, table slots and probe routines shared by every kernel.
//
// Vocabulary follows the reference (debruijn.py): a *node* is a k-mer (debruijn.py:127), an
// *edge* a (k+1)-mer (debruijn.py:217), a *window* one iteration of debruijn.py:126-144.
//
// Table layout (HBM).  One open-addressing table keyed on the NODE (2k <= 126 bits, so the
// all-ones pattern is free for EMPTY; the (k+1)-mer would fill the word -- SURVEY.md section 7):
//   k <= 31: 32-byte slot = one DRAM sector  { key 8 | first-occurrence tag 8 | 4 x u32 count }
//   k <= 63: 48-byte slot                    { key 16 | tag 8 | pad 8 | 4 x u32 count }
// count[c] is edge_count_table[node + "ACGT"[c]] (debruijn.py:213-222); outdegree is the number
// of non-zero counts; the reference's 0/1 indegree is (tag.pos != 0).
#pragma once
#include <stdint.h>

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned short u16;
typedef unsigned char u8;
#define DBG_HD __host__ __device__ __forceinline__
#define DBG_D __device__ __forceinline__

// slot flag bits (device side; low bits match DBG_NODE_* in include/dbg_b200.h)
#define F_OCC     0x01u
#define F_START   0x02u
#define F_BRANCH  0x04u
#define F_DEADEND 0x08u
#define F_PULLED  0x10u
#define F_MULTI   0x20u
#define F_EXIT    0x40u   // sharded table: the single kept successor lives on another rank (dbg_shard.cuh)
#define F_ENTRY   0x80u   // sharded table: this node is the kept successor of another rank's exit

#define NONE32 0xFFFFFFFFu
#define NONE64 0xFFFFFFFFFFFFFFFFULL

// 128-bit keys are two explicit 64-bit words with hand-written shifts.  (nvcc 12.9 miscompiles
// rolling `unsigned __int128` shifts inside the unrolled window loop for sm_100a -- bits carried
// from the low into the high word were lost on the B200 -- so no __int128 arithmetic is used in
// device code.)
struct __align__(16) K128 {
    u64 lo, hi;
};
DBG_HD bool operator==(K128 a, K128 b) { return ((a.lo ^ b.lo) | (a.hi ^ b.hi)) == 0; }
DBG_HD bool operator!=(K128 a, K128 b) { return ((a.lo ^ b.lo) | (a.hi ^ b.hi)) != 0; }

template <int W> struct KeyOf;
template <> struct KeyOf<1> { typedef u64 type; };
template <> struct KeyOf<2> { typedef K128 type; };

template <int W> struct Slot;
template <> struct __align__(32) Slot<1> {
    u64 key;
    u64 tag;
    u32 cnt[4];
};
template <> struct __align__(16) Slot<2> {
    K128 key;
    u64 tag;
    u64 pad;
    u32 cnt[4];
};

template <int W> DBG_HD typename KeyOf<W>::type empty_key();
template <> DBG_HD u64 empty_key<1>() { return ~0ULL; }
template <> DBG_HD K128 empty_key<2>() { K128 r; r.lo = ~0ULL; r.hi = ~0ULL; return r; }

// mask of the low 2k bits: 2k <= 62 (W=1) or <= 126 (W=2), never the full key, so the all-ones
// pattern is free for EMPTY
template <int W> DBG_HD typename KeyOf<W>::type key_mask(int k);
template <> DBG_HD u64 key_mask<1>(int k) { return (1ULL << (2 * k)) - 1; }
template <> DBG_HD K128 key_mask<2>(int k) {
    K128 r;
    const int b = 2 * k;
    if (b >= 64) { r.lo = ~0ULL; r.hi = (1ULL << (b - 64)) - 1; }
    else { r.lo = (1ULL << b) - 1; r.hi = 0; }
    return r;
}

// successor key: drop the first base, append `nb` (debruijn.py:128: read[i+1:i+k+1]).  Bits above
// the 2k key bits are kept as they are: a batched build (many independent read sets in one table)
// stores the batch id there, so a node and its successor stay in the same batch.
DBG_HD u64 key_roll(u64 key, u32 nb, u64 mask) { return (((key << 2) | (u64)nb) & mask) | (key & ~mask); }
DBG_HD K128 key_roll(K128 key, u32 nb, K128 mask) {
    K128 r;
    r.hi = (((key.hi << 2) | (key.lo >> 62)) & mask.hi) | (key.hi & ~mask.hi);
    r.lo = (((key.lo << 2) | (u64)nb) & mask.lo) | (key.lo & ~mask.lo);
    return r;
}
// batch id in the bits above the k-mer (0 = not batched: a no-op)
DBG_HD u64 key_add_batch(u64 key, u32 batch, int k) { return key | ((u64)batch << (2 * k)); }
DBG_HD K128 key_add_batch(K128 key, u32 batch, int k) {
    const int sh = 2 * k;                                    // 2 .. 126
    if (sh >= 64) key.hi |= (u64)batch << (sh - 64);
    else { key.lo |= (u64)batch << sh; if (sh > 32) key.hi |= (u64)batch >> (64 - sh); }
    return key;
}
// append without masking (building a key base by base)
DBG_HD u64 key_push(u64 key, u32 nb) { return (key << 2) | (u64)nb; }
DBG_HD K128 key_push(K128 key, u32 nb) {
    K128 r;
    r.hi = (key.hi << 2) | (key.lo >> 62);
    r.lo = (key.lo << 2) | (u64)nb;
    return r;
}
template <int W> DBG_HD typename KeyOf<W>::type zero_key();
template <> DBG_HD u64 zero_key<1>() { return 0; }
template <> DBG_HD K128 zero_key<2>() { K128 r; r.lo = 0; r.hi = 0; return r; }

// pieces for "shift the k-mer right by one base and put base pb in front" (predecessor key)
DBG_HD u64 key_and_mask(u64 key, u64 mask) { return key & mask; }
DBG_HD K128 key_and_mask(K128 key, K128 mask) { K128 r; r.lo = key.lo & mask.lo; r.hi = key.hi & mask.hi; return r; }
DBG_HD u64 key_andnot_mask(u64 key, u64 mask) { return key & ~mask; }
DBG_HD K128 key_andnot_mask(K128 key, K128 mask) { K128 r; r.lo = key.lo & ~mask.lo; r.hi = key.hi & ~mask.hi; return r; }
DBG_HD u64 key_or(u64 a, u64 b) { return a | b; }
DBG_HD K128 key_or(K128 a, K128 b) { K128 r; r.lo = a.lo | b.lo; r.hi = a.hi | b.hi; return r; }
DBG_HD u64 key_shr2_put(u64 low, u32 pb, int k) { return (low >> 2) | ((u64)pb << (2 * (k - 1))); }
DBG_HD K128 key_shr2_put(K128 low, u32 pb, int k) {
    K128 r;
    r.lo = (low.lo >> 2) | (low.hi << 62);
    r.hi = low.hi >> 2;
    const int sh = 2 * (k - 1);
    if (sh >= 64) r.hi |= (u64)pb << (sh - 64); else r.lo |= (u64)pb << sh;
    return r;
}
DBG_HD u32 key_last_base(u64 key) { return (u32)(key & 3); }
DBG_HD u32 key_last_base(K128 key) { return (u32)(key.lo & 3); }
// base j (0 = first) of a k-mer key
DBG_HD u32 key_base(u64 key, int k, int j) { return (u32)(key >> (2 * (k - 1 - j))) & 3u; }
DBG_HD u32 key_base(K128 key, int k, int j) {
    const int sh = 2 * (k - 1 - j);
    return (u32)(sh >= 64 ? key.hi >> (sh - 64) : key.lo >> sh) & 3u;
}
DBG_HD u64 key_lo64(u64 key) { return key; }
DBG_HD u64 key_lo64(K128 key) { return key.lo; }
DBG_HD u64 key_hi64(u64) { return 0; }
DBG_HD u64 key_hi64(K128 key) { return key.hi; }

DBG_HD u64 fmix64(u64 x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}
DBG_HD u64 hash_key(u64 k) { return fmix64(k + 0x9e3779b97f4a7c15ULL); }
DBG_HD u64 hash_key(K128 k) { return fmix64(k.lo ^ fmix64(k.hi + 0x9e3779b97f4a7c15ULL)); }

// slot hash: multiply-shift, top 32 bits of a 64-bit product (3 IMADs for 64-bit keys); the home
// slot is (h * n_slots) >> 32, one IMAD.WIDE, so n_slots must stay below 2^32 (node ids are 32 bit)
DBG_HD u32 hash32_key(u64 k) { return (u32)((k * 0x9E3779B97F4A7C15ULL) >> 32); }
DBG_HD u32 hash32_key(K128 k) { return (u32)((k.lo * 0x9E3779B97F4A7C15ULL + k.hi * 0xC2B2AE3D27D4EB4FULL) >> 32); }
DBG_HD u64 slot_home(u32 h, u64 n_slots) { return ((u64)h * n_slots) >> 32; }

DBG_HD u64 mulhi64(u64 a, u64 b) {
#if defined(__CUDA_ARCH__) || defined(DBG_EMU)
    return __umul64hi(a, b);
#else
    return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}

// ---------------------------------------------------------------------------------------------
// packed read stream: base i lives in word i/32 at bits [62 - 2*(i%32), 63 - 2*(i%32)]
// (first base most significant, so a k-mer's integer value is its lexicographic rank).
// ---------------------------------------------------------------------------------------------
DBG_D u32 base_at(const u64 *__restrict__ packed, u64 i) {
    return (u32)(__ldg(packed + (i >> 5)) >> (62 - 2 * (i & 31))) & 3u;
}

// n <= 16 bases starting at base index p, right-aligned in 32 bits (first base most significant)
DBG_D u32 extract_bases(const u64 *__restrict__ packed, u64 p, int n) {
    u64 w = p >> 5;
    int s = 2 * (int)(p & 31);
    u64 w0 = __ldg(packed + w), w1 = __ldg(packed + w + 1);
    u64 top = s ? ((w0 << s) | (w1 >> (64 - s))) : w0;
    return (u32)(top >> (64 - 2 * n));
}

// k bases starting at base index p, as a right-aligned integer
template <int W> DBG_D typename KeyOf<W>::type extract_key(const u64 *__restrict__ packed, u64 p, int k);

template <> DBG_D u64 extract_key<1>(const u64 *__restrict__ packed, u64 p, int k) {
    u64 w = p >> 5;
    int s = 2 * (int)(p & 31);
    u64 w0 = __ldg(packed + w), w1 = __ldg(packed + w + 1);
    u64 top = s ? ((w0 << s) | (w1 >> (64 - s))) : w0;       // 64 bits starting at base p
    return top >> (64 - 2 * k);
}
template <> DBG_D K128 extract_key<2>(const u64 *__restrict__ packed, u64 p, int k) {
    u64 w = p >> 5;
    int s = 2 * (int)(p & 31);
    u64 w0 = __ldg(packed + w), w1 = __ldg(packed + w + 1), w2 = __ldg(packed + w + 2);
    u64 hi = s ? ((w0 << s) | (w1 >> (64 - s))) : w0;        // 128 bits starting at base p
    u64 lo = s ? ((w1 << s) | (w2 >> (64 - s))) : w1;
    const int sh = 128 - 2 * k;                              // right-align: 2 <= sh <= 126
    K128 r;
    if (sh >= 64) { r.lo = sh == 64 ? hi : hi >> (sh - 64); r.hi = 0; }
    else { r.lo = (lo >> sh) | (hi << (64 - sh)); r.hi = hi >> sh; }
    return r;
}

// ---------------------------------------------------------------------------------------------
// probes.  Linear probing, slot index = mulhi(hash, n_slots) so n_slots need not be a power of 2.
// Keys are written once (CAS from EMPTY) and never change, so a stale read can only say EMPTY
// for an occupied slot, which the CAS then corrects.
// ---------------------------------------------------------------------------------------------
DBG_D u64 load_key(const Slot<1> *s) { return __ldcg(&s->key); }
DBG_D K128 load_key(const Slot<2> *s) {
    ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2 *>(&s->key));
    K128 r; r.lo = v.x; r.hi = v.y;
    return r;
}

// compare-and-swap of a whole key; returns the previous value
DBG_D u64 cas_key(u64 *addr, u64 expect, u64 val) { return atomicCAS(addr, expect, val); }
DBG_D K128 cas_key(K128 *addr, K128 expect, K128 val) {
    K128 old;
#ifdef DBG_EMU
    emu::chaos_point();
    old = *addr;
    if (old == expect) *addr = val;
#else
    asm volatile("{\n\t.reg .b128 c, v, d;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 v, {%4, %5};\n\t"
                 "atom.global.cas.b128 d, [%6], c, v;\n\tmov.b128 {%0, %1}, d;\n\t}"
                 : "=l"(old.lo), "=l"(old.hi)
                 : "l"(expect.lo), "l"(expect.hi), "l"(val.lo), "l"(val.hi), "l"(addr)
                 : "memory");
#endif
    return old;
}

// key and first-occurrence tag of a slot in ONE 16-byte L2 request (k <= 31), or key then tag
DBG_D u64 load_key_tag(const Slot<1> *s, u64 *tag) {
    ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2 *>(s));
    *tag = v.y;
    return v.x;
}
DBG_D K128 load_key_tag(const Slot<2> *s, u64 *tag) {
    K128 r = load_key(s);
    *tag = __ldcg(&s->tag);
    return r;
}

// find-or-insert starting at the key's home slot; also returns the tag the slot held when it was
// read (~0 for a slot this call created).  NONE64 when the table is full.
template <int W>
DBG_D u64 table_upsert_tag(Slot<W> *table, u64 n_slots, u64 home, typename KeyOf<W>::type key, u64 *tag_seen) {
    typedef typename KeyOf<W>::type K;
    const K EMPTY = empty_key<W>();
    u64 idx = home;
    for (u64 probes = 0; probes < n_slots; ++probes) {
        u64 t;
        K cur = load_key_tag(&table[idx], &t);
        if (cur == key) { *tag_seen = t; return idx; }
        if (cur == EMPTY) {
            K old = cas_key(&table[idx].key, EMPTY, key);
            if (old == EMPTY) { *tag_seen = ~0ULL; return idx; }
            if (old == key) { *tag_seen = __ldcg(&table[idx].tag); return idx; }
        }
        if (++idx == n_slots) idx = 0;
    }
    return NONE64;
}

// find-or-insert; returns slot index or NONE64 when the table is full
template <int W>
DBG_D u64 table_upsert(Slot<W> *table, u64 n_slots, typename KeyOf<W>::type key) {
    typedef typename KeyOf<W>::type K;
    const K EMPTY = empty_key<W>();
    u64 idx = slot_home(hash32_key(key), n_slots);
    for (u64 probes = 0; probes < n_slots; ++probes) {
        K cur = load_key(&table[idx]);
        if (cur == key) return idx;
        if (cur == EMPTY) {
            K old = cas_key(&table[idx].key, EMPTY, key);
            if (old == EMPTY || old == key) return idx;
        }
        if (++idx == n_slots) idx = 0;
    }
    return NONE64;
}

template <int W>
DBG_D u64 table_find(const Slot<W> *table, u64 n_slots, typename KeyOf<W>::type key) {
    typedef typename KeyOf<W>::type K;
    const K EMPTY = empty_key<W>();
    u64 idx = slot_home(hash32_key(key), n_slots);
    for (u32 probes = (u32)n_slots; probes; --probes) {          // n_slots < 2^32; a torn (overflowed) table may be full
        K cur = load_key(&table[idx]);
        if (cur == key) return idx;
        if (cur == EMPTY) return NONE64;
        if (++idx == n_slots) idx = 0;
    }
    return NONE64;
}

// ---------------------------------------------------------------------------------------------
// Node table laid out by MINIMIZER ("regions").  The table is n_regions x region_slots slots; a node's
// home slot is  region(minimizer of the node) * region_slots + hash(node) mod region_slots,  then linear
// probing as before (over the whole table, so a crowded region spills into the next one and nothing else
// changes).  Every occurrence of a node has the same minimizer, and all nodes of a chunk share theirs, so
// (1) a chunk's node upserts land inside one region (k_expand walks the chunks in region order: the table
// is built front to back through L2 instead of by random DRAM sectors), and (2) a node's successor is in
// the same region unless the minimizer changes between them (1 node in ~8): k_finalize's probes stay local.
// The minimizer is exactly the one k_masks cuts chunks by: smallest 26-bit multiplicative hash among the
// node's first w m-mers, leftmost on ties.  n_regions == 1 is the plain hash layout (short minimizers do
// not spread over regions; the generic-alphabet engine).
// ---------------------------------------------------------------------------------------------
struct TGeom {
    u64 n_slots;
    u32 n_regions;       // >= 1
    u32 region_slots;    // n_slots == n_regions * region_slots when n_regions > 1
    int k, w, m;         // chunk geometry: w candidate m-mers per node
};
DBG_HD TGeom plain_geom(u64 n_slots) {
    TGeom g; g.n_slots = n_slots; g.n_regions = 1; g.region_slots = (u32)n_slots; g.k = 0; g.w = 0; g.m = 0;
    return g;
}
DBG_HD u32 mix32(u32 x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// the 2 m bits of the m-mer at offset i of a k-mer key (m <= 16)
DBG_HD u32 key_mmer(u64 key, int k, int m, int i) {
    return (u32)(key >> (2 * (k - m - i))) & (u32)((1ULL << (2 * m)) - 1);
}
DBG_HD u32 key_mmer(K128 key, int k, int m, int i) {
    const int sh = 2 * (k - m - i);
    const u64 x = sh >= 64 ? key.hi >> (sh - 64) : (sh ? (key.lo >> sh) | (key.hi << (64 - sh)) : key.lo);
    return (u32)x & (u32)((1ULL << (2 * m)) - 1);
}
// batch id of a batched build: the key bits above the k-mer
DBG_HD u32 key_batch(u64 key, int k) { return (u32)(key >> (2 * k)); }
DBG_HD u32 key_batch(K128 key, int k) {
    const int sh = 2 * k;
    return (u32)(sh >= 64 ? key.hi >> (sh - 64) : (key.lo >> sh) | (key.hi << (64 - sh)));
}
// order of the minimizer candidates: 26-bit multiplicative hash, ties to the leftmost (k_masks, chunk_owner)
DBG_HD u32 mmer_rank(u32 x, int i) { return ((x * 0x9E3779B1u) | 63u) - (u32)(63 - i); }
DBG_HD u32 mmer_owner_hash(u32 x, u32 batch) { return hash32_key((u64)x | ((u64)batch << 32)); }

// 32-bit "where in the table" hash of a node: a function of its minimizer (and batch) only; remixed so
// that it is independent of the multi-GPU owner function, which uses mmer_owner_hash directly
template <typename K> DBG_HD u32 node_vhash(K key, int k, int w, int m) {
    u32 best = 0xFFFFFFFFu, best_x = 0;
    for (int i = 0; i < w; i++) {
        const u32 x = key_mmer(key, k, m, i);
        const u32 g = mmer_rank(x, i);
        if (g < best) { best = g; best_x = x; }
    }
    return mix32(mmer_owner_hash(best_x, key_batch(key, k)) ^ 0x5bd1e995u);
}
DBG_HD u64 region_home(const TGeom &g, u32 vhash, u32 h) {
    return (((u64)vhash * g.n_regions) >> 32) * g.region_slots + (((u64)h * g.region_slots) >> 32);
}
template <typename K> DBG_HD u64 node_home(const TGeom &g, K key) {
    const u32 h = hash32_key(key);
    if (g.n_regions <= 1) return slot_home(h, g.n_slots);
    return region_home(g, node_vhash(key, g.k, g.w, g.m), h);
}

template <int W>
DBG_D u64 table_find(const Slot<W> *table, const TGeom &g, typename KeyOf<W>::type key) {
    typedef typename KeyOf<W>::type K;
    const K EMPTY = empty_key<W>();
    u64 idx = node_home(g, key);
    for (u32 probes = (u32)g.n_slots; probes; --probes) {        // n_slots < 2^32; a torn (overflowed) table may be full
        K cur = load_key(&table[idx]);
        if (cur == key) return idx;
        if (cur == EMPTY) return NONE64;
        if (++idx == g.n_slots) idx = 0;
    }
    return NONE64;
}
template <int W>
DBG_D u64 table_upsert(Slot<W> *table, const TGeom &g, typename KeyOf<W>::type key) {
    typedef typename KeyOf<W>::type K;
    const K EMPTY = empty_key<W>();
    u64 idx = node_home(g, key);
    for (u64 probes = 0; probes < g.n_slots; ++probes) {
        K cur = load_key(&table[idx]);
        if (cur == key) return idx;
        if (cur == EMPTY) {
            K old = cas_key(&table[idx].key, EMPTY, key);
            if (old == EMPTY || old == key) return idx;
        }
        if (++idx == g.n_slots) idx = 0;
    }
    return NONE64;
}

// position of the r-th (0-based) set bit of m; m must have more than r bits set
DBG_D u32 nth_set_bit(u32 m, u32 r) {
    u32 pos = 0, c;
    c = __popc(m & 0xFFFFu); if (r >= c) { r -= c; pos += 16; m >>= 16; }
    c = __popc(m & 0xFFu);   if (r >= c) { r -= c; pos += 8;  m >>= 8; }
    c = __popc(m & 0xFu);    if (r >= c) { r -= c; pos += 4;  m >>= 4; }
    c = __popc(m & 0x3u);    if (r >= c) { r -= c; pos += 2;  m >>= 2; }
    if (r >= (m & 1u)) pos += 1;
    return pos;
}


// ---------------------------------------------------------------------------------------------
// pruningEdges (debruijn.py:150-166) for one node.  cnt[c] = multiplicity of successor base c,
// first[c] = window sequence tag of its first appearance (ties in count are broken by it, as
// Counter.most_common does; pass NULL to break ties by base code).  Returns the `kept` word:
// bits 0-7 = kept bases in E order, 2 bits each; bits 8-10 = number kept; bits 12-14 = number
// of distinct successors (the reference's outdegree).
// ---------------------------------------------------------------------------------------------
DBG_HD u32 prune_successors(const u32 cnt[4], const u64 *first, double threshold) {
    // almost every node has at most one distinct successor: no sort, no indexed array (which would live in
    // local memory: that loop was a third of k_finalize's instructions)
    const u32 nz = (cnt[0] ? 1u : 0u) | (cnt[1] ? 2u : 0u) | (cnt[2] ? 4u : 0u) | (cnt[3] ? 8u : 0u);
    if ((nz & (nz - 1u)) == 0u) {
        if (nz == 0u) return 0u;
        const u32 c = (nz >> 1) - (nz >> 3);                        // 1, 2, 4, 8 -> 0, 1, 2, 3
        return (1u << 12) | (1u << 8) | c;
    }
    int order[4], m = 0;
    for (int c = 0; c < 4; c++) if (cnt[c]) order[m++] = c;
    for (int a = 1; a < m; a++) {                  // insertion sort: count desc, then first asc
        int x = order[a], b = a - 1;
        while (b >= 0) {
            int y = order[b];
            bool y_before_x = cnt[y] > cnt[x] ||
                              (cnt[y] == cnt[x] && (first ? first[y] < first[x] : y < x));
            if (y_before_x) break;
            order[b + 1] = y; b--;
        }
        order[b + 1] = x;
    }
    u32 word = (u32)m << 12, kept = 0;
    if (m >= 1) {
        word |= (u32)order[0];
        kept = 1;
        if (m > 1) {
            double lim = (double)cnt[order[0]] / threshold;          // debruijn.py:163
            for (int a = 1; a < m; a++)
                if ((double)cnt[order[a]] >= lim) { word |= (u32)order[a] << (2 * kept); kept++; }
        }
    }
    return word | (kept << 8);
}
DBG_HD int kept_count(u32 word) { return (int)((word >> 8) & 7); }
DBG_HD int kept_base(u32 word, int j) { return (int)((word >> (2 * j)) & 3); }
DBG_HD int distinct_count(u32 word) { return (int)((word >> 12) & 7); }
