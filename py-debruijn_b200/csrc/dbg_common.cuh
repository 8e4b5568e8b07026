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
typedef unsigned __int128 u128;

#define DBG_HD __host__ __device__ __forceinline__
#define DBG_D __device__ __forceinline__

// slot flag bits (device side; low bits match DBG_NODE_* in include/dbg_b200.h)
#define F_OCC     0x01u
#define F_START   0x02u
#define F_BRANCH  0x04u
#define F_DEADEND 0x08u
#define F_PULLED  0x10u
#define F_MULTI   0x20u

#define NONE32 0xFFFFFFFFu

template <int W> struct KeyOf;
template <> struct KeyOf<1> { typedef u64 type; };
template <> struct KeyOf<2> { typedef u128 type; };

template <int W> struct Slot;
template <> struct __align__(32) Slot<1> {
    u64 key;
    u64 tag;
    u32 cnt[4];
};
template <> struct __align__(16) Slot<2> {
    u128 key;
    u64 tag;
    u64 pad;
    u32 cnt[4];
};

template <int W> DBG_HD typename KeyOf<W>::type empty_key() { return ~(typename KeyOf<W>::type)0; }

template <int W> DBG_HD typename KeyOf<W>::type key_mask(int k) {
    typedef typename KeyOf<W>::type K;
    return (((K)1) << (2 * k)) - 1;          // 2k <= 62 (W=1) or <= 126 (W=2): never the full word
}

DBG_HD u64 fmix64(u64 x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}
DBG_HD u64 hash_key(u64 k) { return fmix64(k + 0x9e3779b97f4a7c15ULL); }
DBG_HD u64 hash_key(u128 k) { return fmix64((u64)k ^ fmix64((u64)(k >> 64) + 0x9e3779b97f4a7c15ULL)); }

DBG_HD u64 mulhi64(u64 a, u64 b) {
#if defined(__CUDA_ARCH__) || defined(DBG_EMU)
    return __umul64hi(a, b);
#else
    return (u64)(((u128)a * b) >> 64);
#endif
}

// ---------------------------------------------------------------------------------------------
// packed read stream: base i lives in word i/32 at bits [62 - 2*(i%32), 63 - 2*(i%32)]
// (first base most significant, so a k-mer's integer value is its lexicographic rank).
// ---------------------------------------------------------------------------------------------
DBG_D u32 base_at(const u64 *__restrict__ packed, u64 i) {
    return (u32)(__ldg(packed + (i >> 5)) >> (62 - 2 * (i & 31))) & 3u;
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
template <> DBG_D u128 extract_key<2>(const u64 *__restrict__ packed, u64 p, int k) {
    u64 w = p >> 5;
    int s = 2 * (int)(p & 31);
    u64 w0 = __ldg(packed + w), w1 = __ldg(packed + w + 1), w2 = __ldg(packed + w + 2);
    u64 hi = s ? ((w0 << s) | (w1 >> (64 - s))) : w0;
    u64 lo = s ? ((w1 << s) | (w2 >> (64 - s))) : w1;
    u128 top = ((u128)hi << 64) | lo;                        // 128 bits starting at base p
    return top >> (128 - 2 * k);
}

// ---------------------------------------------------------------------------------------------
// probes.  Linear probing, slot index = mulhi(hash, n_slots) so n_slots need not be a power of 2.
// Keys are written once (CAS from EMPTY) and never change, so a stale read can only say EMPTY
// for an occupied slot, which the CAS then corrects.
// ---------------------------------------------------------------------------------------------
DBG_D u64 load_key(const Slot<1> *s) { return __ldcg(&s->key); }
DBG_D u128 load_key(const Slot<2> *s) {
    ulonglong2 v = __ldcg(reinterpret_cast<const ulonglong2 *>(&s->key));
    return ((u128)v.y << 64) | v.x;
}

// find-or-insert; returns slot index or NONE64 when the table is full
#define NONE64 0xFFFFFFFFFFFFFFFFULL
template <int W>
DBG_D u64 table_upsert(Slot<W> *table, u64 n_slots, typename KeyOf<W>::type key) {
    typedef typename KeyOf<W>::type K;
    const K EMPTY = empty_key<W>();
    u64 idx = mulhi64(hash_key(key), n_slots);
    for (u64 probes = 0; probes < n_slots; ++probes) {
        K cur = load_key(&table[idx]);
        if (cur == key) return idx;
        if (cur == EMPTY) {
            K old = atomicCAS(&table[idx].key, EMPTY, key);
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
    u64 idx = mulhi64(hash_key(key), n_slots);
    for (u64 probes = 0; probes < n_slots; ++probes) {
        K cur = load_key(&table[idx]);
        if (cur == key) return idx;
        if (cur == EMPTY) return NONE64;
        if (++idx == n_slots) idx = 0;
    }
    return NONE64;
}

// ---------------------------------------------------------------------------------------------
// pruningEdges (debruijn.py:150-166) for one node.  cnt[c] = multiplicity of successor base c,
// first[c] = window sequence tag of its first appearance (ties in count are broken by it, as
// Counter.most_common does; pass NULL to break ties by base code).  Returns the `kept` word:
// bits 0-7 = kept bases in E order, 2 bits each; bits 8-10 = number kept; bits 12-14 = number
// of distinct successors (the reference's outdegree).
// ---------------------------------------------------------------------------------------------
DBG_HD u32 prune_successors(const u32 cnt[4], const u64 *first, double threshold) {
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
