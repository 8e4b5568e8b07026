// dbg_generic.cuh -- the same hot path for ANY single-byte alphabet (up to 255 symbols, codes of 1 .. 8 bits).
//
// The reference treats reads as opaque strings (debruijn.py works on any characters) and its real inputs
// are de-novo PEPTIDES (II_assembleFromReads.py:43-47, :55), so the drop-in must take them.  The 2-bit
// engine (dbg_chunks.cuh / dbg_kernels.cuh) is built around four successor counters per node; with 20+
// letters a node can have 20+ distinct successors, so this engine keeps the graph edge-centric:
//
//   node table   Slot<2>: key = k symbols x b bits, b = ceil(log2(symbols)) chosen per handle (peptides: 5 bits,
//                k <= 24; ACGT + N: 3 bits, k <= 40; 8 bits: k <= 14), tag = first occurrence
//   edge table   Slot<2>: key = the (k+1)-mer (b (k + 1) <= 125 bits), tag = first appearance of the edge
//                (Counter.most_common tie order, debruijn.py:159-164), cnt[0] = multiplicity
//                = edge_count_table[(k+1)-mer] (debruijn.py:213-222)
//   edge array   occupied edge slots compacted and sorted by (source node, count desc, first appearance):
//                every node's successors are one contiguous run, already in the reference's E order, and
//                pruningEdges (debruijn.py:150-166) keeps a PREFIX of it
//
// Everything downstream of that (lists, tip pruning, pull-out reads, list ranking, contig descriptors,
// strings, getScore) is the same code as the 2-bit engine or a CSR twin of it.  Inputs of this shape are
// small (10^3 .. 10^6 reads of 10 .. 40 residues), so the table build is one thread per window with plain
// upserts -- no super-k-mers, no partition passes.
#pragma once
#include "dbg_kernels.cuh"

#define GEN_MAX_SYMBOLS 255                // code 0xFF = "not a symbol of this alphabet"
#define GEN_KEY_BITS 125                   // b (k + 1) <= 125: the edge key keeps the all-ones pattern free
#define GEN_LUT_BYTES (256 + 256 + 1)      // byte -> code | code -> byte | b  (device copy of the handle's alphabet)

DBG_HD int gen_bits_for(int n_symbols) {   // smallest code width that tells n_symbols apart
    int b = 1;
    while ((1 << b) < n_symbols) b++;
    return b;
}
DBG_HD int gen_max_k(int gb) { const int m = GEN_KEY_BITS / gb - 1; return m < 63 ? m : 63; }   // (63: what the contig / label buffers of the ABI are sized for)

DBG_HD K128 gkey_push(K128 key, u32 code, int gb) {
    K128 r;
    r.hi = (key.hi << gb) | (key.lo >> (64 - gb));
    r.lo = (key.lo << gb) | (u64)code;
    return r;
}
DBG_HD K128 gkey_mask(int n, int gb) {     // low gb n bits (<= 125)
    K128 r;
    const int b = gb * n;
    if (b >= 64) { r.lo = ~0ULL; r.hi = b == 64 ? 0ULL : ((1ULL << (b - 64)) - 1); }
    else { r.lo = (1ULL << b) - 1; r.hi = 0; }
    return r;
}
DBG_HD K128 gkey_and(K128 a, K128 m) { K128 r; r.lo = a.lo & m.lo; r.hi = a.hi & m.hi; return r; }
DBG_HD K128 gkey_drop_last(K128 key, int gb) {     // (k+1)-mer -> its first k symbols
    K128 r;
    r.lo = (key.lo >> gb) | (key.hi << (64 - gb));
    r.hi = key.hi >> gb;
    return r;
}

// bytes -> symbol codes through the handle's table (0xFF = not a symbol of this alphabet)
__global__ void __launch_bounds__(DBG_BLOCK) k_g_encode(const u8 *__restrict__ bases, u64 n_bases,
                                                        const u8 *__restrict__ lut, u8 *__restrict__ codes,
                                                        DevCounters *ctr) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_bases) return;
    const u8 c = lut[bases[i]];
    codes[i] = c;
    if (c == 0xFFu) ctr->bad_alphabet = 1;
}

// read holding stream position p: last r with off[r] <= p (empty reads are skipped by the strict test below)
DBG_D u64 g_read_of(const u64 *__restrict__ off, u64 n_reads, u64 p) {
    u64 lo = 0, hi = n_reads - 1;
    while (lo < hi) {
        const u64 mid = (lo + hi + 1) >> 1;
        if (__ldg(off + mid) <= p) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// the window loop, debruijn.py:123-144: one thread per stream position
__global__ void __launch_bounds__(DBG_BLOCK) k_g_insert(const u8 *__restrict__ codes, const u64 *__restrict__ off,
                                                        u64 n_reads, u64 n_bases, int k, int gb, Slot<2> *nodes,
                                                        u64 n_nslots, Slot<2> *edges, u64 n_eslots, u64 read_idx_base,
                                                        DevCounters *ctr) {
    const u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n_bases) {
        const u64 r = g_read_of(off, n_reads, p);
        const u64 rs = __ldg(off + r), re = __ldg(off + r + 1);
        const u64 len = re - rs, pos = p - rs;
        if (len > (u64)k && pos <= len - (u64)k) {                      // a node occurrence (debruijn.py:126-127)
            K128 key = zero_key<2>();
            for (int j = 0; j < k; j++) key = gkey_push(key, codes[p + j], gb);
            const u64 tag = ((read_idx_base + r) << 32) | pos;
            u64 seen;
            const u64 idx = table_upsert_tag<2>(nodes, n_nslots, slot_home(hash32_key(key), n_nslots), key, &seen);
            if (idx == NONE64) ctr->overflow = 1;
            else if (tag < seen) atomicMin(&nodes[idx].tag, tag);
            if (pos < len - (u64)k) {                                    // ... with a successor (debruijn.py:128-136)
                const K128 ekey = gkey_push(key, codes[p + k], gb);
                const u64 e = table_upsert_tag<2>(edges, n_eslots, slot_home(hash32_key(ekey), n_eslots), ekey, &seen);
                if (e == NONE64) ctr->overflow = 1;
                else {
                    if (tag < seen) atomicMin(&edges[e].tag, tag);
                    atomicAdd(&edges[e].cnt[0], 1u);
                }
            }
        }
    }
}

// occupied edge slots -> edge array (unordered)
__global__ void __launch_bounds__(DBG_BLOCK) k_g_edges(const Slot<2> *__restrict__ etable, u64 n_eslots,
                                                       const Slot<2> *__restrict__ nodes, u64 n_nslots, int k, int gb,
                                                       GEdge *__restrict__ out, u64 *cursor, DevCounters *ctr) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    K128 ekey = empty_key<2>();
    if (i < n_eslots) ekey = etable[i].key;
    const bool have = ekey != empty_key<2>();
    const u32 b = __ballot_sync(0xFFFFFFFFu, have);
    if (!b) return;
    const u32 lane = threadIdx.x & 31;
    u64 base = 0;
    if (lane == 0) base = atomicAdd(cursor, (u64)__popc(b));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (!have) return;
    GEdge e;
    const u64 s = table_find<2>(nodes, n_nslots, gkey_drop_last(ekey, gb));
    const u64 d = table_find<2>(nodes, n_nslots, gkey_and(ekey, gkey_mask(k, gb)));
    if (s == NONE64 || d == NONE64) raise_fault(ctr, FAULT_SUCC_MISSING, (u32)i);
    e.src = (u32)s; e.dst = (u32)d;
    e.cnt = etable[i].cnt[0];
    e.sym = (u32)ekey.lo & ((1u << gb) - 1u);
    e.first = etable[i].tag;
    out[base + __popc(b & ((1u << lane) - 1))] = e;
}

// sort keys: pass 1 by first appearance, pass 2 (stable) by (source, count descending)
__global__ void __launch_bounds__(DBG_BLOCK) k_g_sortkeys1(const GEdge *__restrict__ e, u64 n, u64 *__restrict__ key,
                                                           u32 *__restrict__ idx) {
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) { key[j] = e[j].first; idx[j] = (u32)j; }
}
__global__ void __launch_bounds__(DBG_BLOCK) k_g_sortkeys2(const GEdge *__restrict__ e, const u32 *__restrict__ idx,
                                                           u64 n, u64 *__restrict__ key) {
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) { const GEdge &x = e[idx[j]]; key[j] = ((u64)x.src << 32) | (u64)(~x.cnt); }
}
__global__ void __launch_bounds__(DBG_BLOCK) k_g_gather(const GEdge *__restrict__ e, const u32 *__restrict__ idx, u64 n,
                                                        GEdge *__restrict__ out) {
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) out[j] = e[idx[j]];
}

// CSR over the sorted edge array: first edge and number of edges of every source slot
__global__ void __launch_bounds__(DBG_BLOCK) k_g_csr(const GEdge *__restrict__ e, u64 n, u32 *__restrict__ estart,
                                                     u32 *edeg) {
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const u32 s = e[j].src;
    if (j == 0 || e[j - 1].src != s) estart[s] = (u32)j;
    atomicAdd(&edeg[s], 1u);
}

// degrees, pruningEdges, branch flags, the single surviving successor (debruijn.py:129-166, 230-236)
__global__ void __launch_bounds__(DBG_BLOCK) k_g_finalize(const Slot<2> *__restrict__ nodes, u64 n_slots,
                                                          const GEdge *__restrict__ e, const u32 *__restrict__ estart,
                                                          const u32 *__restrict__ edeg, double threshold,
                                                          u8 *__restrict__ flags, u16 *__restrict__ kept,
                                                          u32 *__restrict__ succ, u32 *__restrict__ enk, DevCounters *ctr) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    u32 f = 0, word = 0, sx = NONE32, nk = 0;
    if (nodes[i].key != empty_key<2>()) {
        const u32 m = edeg[i], s = estart[i];
        if (m >= 1) {
            nk = 1;
            if (m > 1) {
                const double lim = (double)e[s].cnt / threshold;                 // debruijn.py:163
                while (nk < m && (double)e[s + nk].cnt >= lim) nk++;             // counts descend: the kept ones are a prefix
            }
        }
        f = F_OCC | ((u32)nodes[i].tag == 0 ? F_START : 0) | (nk > 1 ? F_BRANCH : 0) | (m == 0 ? F_DEADEND : 0) |
            (m > 1 ? F_MULTI : 0);
        if (nk == 1) sx = e[s].dst;
        word = ((nk < 7 ? nk : 7u) << 8) | ((m < 7 ? m : 7u) << 12);            // what the list-ranking kernels look at
        atomicAdd(&ctr->n_nodes, 1ull);
        if (m) atomicAdd(&ctr->n_edges, (u64)m);
        if (f & F_START) atomicAdd(&ctr->n_starts, 1ull);
        if (nk > 1) atomicAdd(&ctr->n_branch, 1ull);
        if (m > 1) atomicAdd(&ctr->n_multi, 1ull);
    }
    flags[i] = (u8)f;
    kept[i] = (u16)word;
    succ[i] = sx;
    enk[i] = nk;
}

// candidate paths of pruningErrorContigFromBranch (debruijn.py:169-186); same records as k_tip_paths
__global__ void __launch_bounds__(DBG_BLOCK) k_g_tip_paths(const u8 *__restrict__ flags, const u32 *__restrict__ enk,
                                                           const u32 *__restrict__ estart, const GEdge *__restrict__ e,
                                                           const u32 *__restrict__ branch, u64 n_branch, int mode,
                                                           u32 *__restrict__ counts, const u64 *__restrict__ offs,
                                                           u32 *__restrict__ recs) {
    const u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_branch) return;
    u32 vec[5], child[5];
    u32 n = 0;
    vec[0] = branch[b]; child[0] = 0;
    int depth = 1;
    while (depth > 0) {
        const int d = depth - 1;
        const u32 ci = child[d];
        if (ci >= enk[vec[d]]) { depth--; continue; }
        child[d] = ci + 1;
        if (depth == 5) continue;                                   // callee would see depth == 0
        const u32 nx = e[estart[vec[d]] + ci].dst;
        vec[depth] = nx;
        if (flags[nx] & F_DEADEND) {                                // debruijn.py:173
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
        child[depth] = 0;
        depth++;
    }
    if (mode == 0) counts[b] = n;
}

// pull_out_read (debruijn.py:274-278): a read is pulled iff one of its k-mers is a branch node
__global__ void __launch_bounds__(DBG_BLOCK) k_g_pull_reads(const u8 *__restrict__ codes, const u64 *__restrict__ off,
                                                            u64 n_reads, u64 n_bases, int k, int gb,
                                                            const Slot<2> *__restrict__ nodes, u64 n_nslots,
                                                            const u8 *__restrict__ flags, u8 *read_flags) {
    const u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_bases) return;
    const u64 r = g_read_of(off, n_reads, p);
    const u64 rs = __ldg(off + r), re = __ldg(off + r + 1);
    if (re - rs < (u64)k || p - rs > re - rs - (u64)k) return;
    K128 key = zero_key<2>();
    for (int j = 0; j < k; j++) key = gkey_push(key, codes[p + j], gb);
    const u64 idx = table_find<2>(nodes, n_nslots, key);
    if (idx != NONE64 && (flags[idx] & F_BRANCH)) read_flags[r] = 1;
}

__global__ void __launch_bounds__(DBG_BLOCK) k_g_gather_u32(const u32 *__restrict__ src, const u32 *__restrict__ ids, u64 n,
                                                            u32 *__restrict__ out) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[ids[i]];
}

// junction table of the final=True enumeration: where every kept successor of every branch node ends
__global__ void __launch_bounds__(DBG_BLOCK) k_g_junctions(const u8 *__restrict__ flags, const u32 *__restrict__ enk,
                                                           const u32 *__restrict__ estart, const GEdge *__restrict__ e,
                                                           const u64 *__restrict__ link, const u32 *__restrict__ branch,
                                                           u64 n_branch, const u64 *__restrict__ joff,
                                                           JunctionSucc *__restrict__ out) {
    const u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_branch) return;
    const u32 v = branch[b];
    for (u32 j = 0; j < enk[v]; j++) {
        JunctionSucc js;
        const u32 s = e[estart[v] + j].dst;
        u32 t = s, d = 0;
        js.entry = s; js.n_hops = 0; js.terminal = NONE32; js.kind = T_CYCLE;
        if (resolve_chain(link, s, &t, &d)) {
            const u32 f = flags[t];
            js.n_hops = d; js.terminal = t;
            js.kind = (f & F_PULLED) ? T_PULLED : (f & F_DEADEND) ? T_DEADEND : T_BRANCH;
        }
        out[joff[b] + j] = js;
    }
}

// edge_count_table[(k+1)-mer] for ASCII queries laid back to back; 0 = absent
__global__ void __launch_bounds__(DBG_BLOCK) k_g_lookup_edges(const Slot<2> *__restrict__ etable, u64 n_eslots, int k,
                                                              int gb, const u8 *__restrict__ lut, const u8 *__restrict__ q, u64 n,
                                                              u32 *__restrict__ out) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u8 *s = q + i * (u64)(k + 1);
    K128 key = zero_key<2>();
    bool bad = false;
    for (int j = 0; j <= k; j++) {
        const u8 c = lut[s[j]];
        bad |= c == 0xFFu;
        key = gkey_push(key, bad ? 0u : (u32)c, gb);
    }
    u32 res = 0;
    if (!bad) {
        const u64 idx = table_find<2>(etable, n_eslots, key);
        if (idx != NONE64) res = etable[idx].cnt[0];
    }
    out[i] = res;
}

// the sorted edge array as columns (dbg_export_edges)
__global__ void __launch_bounds__(DBG_BLOCK) k_g_export_edges(const GEdge *__restrict__ e, u64 n, const u32 *__restrict__ estart,
                                                              const u32 *__restrict__ enk, const u8 *__restrict__ sym_of_code,
                                                              u32 *__restrict__ src, u32 *__restrict__ dst, u8 *__restrict__ sym,
                                                              u32 *__restrict__ cnt, u8 *__restrict__ kept) {
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const GEdge x = e[j];
    src[j] = x.src; dst[j] = x.dst; cnt[j] = x.cnt;
    sym[j] = sym_of_code[x.sym];
    kept[j] = (u8)((u32)j - estart[x.src] < enk[x.src] ? 1 : 0);
}
