// dbg_support.cuh -- findSupportReadScore of IV_sortOutputs.py:10-15:
//
//     score(contig) = sum of score_table[read] over the table's reads (dict order) with `read in contig`
//
// O(reads x contigs) substring tests in the reference.  Here: the reads go into a hash set keyed by a
// polynomial hash of the whole read; every contig gets prefix hashes, so the hash of any substring is
// two multiply-adds; every (contig position, distinct read length) looks its substring up and a hit is
// verified byte by byte (exact, no fingerprints).  The (contig, read) hits are sorted, made unique (a
// read that occurs twice in a contig counts once, as `in` does) and summed PER CONTIG IN READ ORDER in
// float64 -- the reference's left-to-right `score +=`, so the sums are bit-identical for float scores too.
#pragma once
#include "dbg_kernels.cuh"

#define SUP_BASE 0x100000001B3ULL
#define SUP_EMPTY 0xFFFFFFFFu

struct SupSlot { u64 hash; u32 read; u32 len; };

DBG_HD u64 sup_step(u64 h, u32 byte) { return h * SUP_BASE + (u64)(byte + 1u); }

__global__ void __launch_bounds__(DBG_BLOCK) k_sup_reads(const u8 *__restrict__ bytes, const u64 *__restrict__ off, u64 n,
                                                         SupSlot *table, u64 cap) {
    const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const u64 a = off[r], b = off[r + 1];
    u64 h = 0;
    for (u64 i = a; i < b; i++) h = sup_step(h, bytes[i]);
    u64 idx = mulhi64(h * 0x9E3779B97F4A7C15ULL, cap);
    for (u64 probes = 0; probes < cap; probes++) {                       // every read takes its own slot (cap >= 2 n)
        if (atomicCAS(&table[idx].read, SUP_EMPTY, (u32)r) == SUP_EMPTY) {
            table[idx].hash = h; table[idx].len = (u32)(b - a);
            return;
        }
        if (++idx == cap) idx = 0;
    }
}

// prefix hashes of contig c at pre[off[c] + c .. off[c + 1] + c]  (len + 1 values)
__global__ void __launch_bounds__(DBG_BLOCK) k_sup_prefix(const u8 *__restrict__ bytes, const u64 *__restrict__ off, u64 n,
                                                          u64 *__restrict__ pre) {
    const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const u64 a = off[c], b = off[c + 1];
    u64 h = 0;
    pre[a + c] = 0;
    for (u64 i = a; i < b; i++) { h = sup_step(h, bytes[i]); pre[i + c + 1] = h; }
}

// one thread per contig byte; mode 0 counts the hits, mode 1 writes (contig << 32 | read)
__global__ void __launch_bounds__(DBG_BLOCK) k_sup_match(const u8 *__restrict__ cbytes, const u64 *__restrict__ coff, u64 n_contigs,
                                                         const u64 *__restrict__ pre, const u8 *__restrict__ rbytes,
                                                         const u64 *__restrict__ roff, const SupSlot *__restrict__ table, u64 cap,
                                                         const u32 *__restrict__ lens, const u64 *__restrict__ pows, u32 n_lens,
                                                         int mode, u64 *cursor, u64 *__restrict__ hits, u64 hit_cap) {
    const u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 total = coff[n_contigs];
    if (g >= total) return;
    u64 lo = 0, hi = n_contigs - 1;                                      // contig holding byte g
    while (lo < hi) {
        const u64 mid = (lo + hi + 1) >> 1;
        if (coff[mid] <= g) lo = mid; else hi = mid - 1;
    }
    const u64 c = lo, p = g - coff[c], clen = coff[c + 1] - coff[c];
    const u64 *H = pre + coff[c] + c;
    for (u32 li = 0; li < n_lens; li++) {
        const u64 len = lens[li];
        if (p + len > clen) break;                                       // lens ascend
        const u64 h = H[p + len] - H[p] * pows[li];
        u64 idx = mulhi64(h * 0x9E3779B97F4A7C15ULL, cap);
        for (u64 probes = 0; probes < cap; probes++) {
            const u32 r = table[idx].read;
            if (r == SUP_EMPTY) break;
            if (table[idx].hash == h && table[idx].len == (u32)len) {
                const u8 *a = rbytes + roff[r], *b = cbytes + g;
                bool same = true;
                for (u64 i = 0; i < len && same; i++) same = a[i] == b[i];
                if (same) {
                    const u64 o = atomicAdd(cursor, 1ULL);
                    if (mode == 1 && o < hit_cap) hits[o] = (c << 32) | r;
                }
            }
            if (++idx == cap) idx = 0;
        }
    }
}

// hits sorted by (contig, read): per contig, the distinct reads in read order, summed left to right
__global__ void __launch_bounds__(DBG_BLOCK) k_sup_sum(const u64 *__restrict__ hits, u64 n_hits, const double *__restrict__ score,
                                                       u64 n_contigs, double *__restrict__ out) {
    const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_contigs) return;
    u64 lo = 0, hi = n_hits;                                             // first hit with contig >= c
    while (lo < hi) {
        const u64 mid = (lo + hi) >> 1;
        if ((hits[mid] >> 32) < c) lo = mid + 1; else hi = mid;
    }
    double s = 0.0;
    u64 last = ~0ULL;
    for (u64 i = lo; i < n_hits && (hits[i] >> 32) == c; i++) {
        if (hits[i] == last) continue;
        last = hits[i];
        s += score[(u32)hits[i]];
    }
    out[c] = s;
}
