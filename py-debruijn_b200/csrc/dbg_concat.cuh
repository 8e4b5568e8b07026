// dbg_concat.cuh -- the contig concatenation of III_concatenateOutpus.py:57-137 and the substring dedup of
// :139-142.
//
//     for k in 20 .. 6:   repeat up to 100000 times:
//         current = random.choice(the contigs that are not in nonconcatenatable)            (:62-65)
//         every other contig longer than k whose last k letters are current's first k (or whose first k
//         are current's last k), exactly or after exchanging two adjacent letters, becomes the merged
//         string (:70-131); if any did, the first contig equal to current leaves the list (:134), otherwise
//         current joins nonconcatenatable (:136)
//
// The loop is sequential by construction: every draw depends on the list the previous iteration left.  What
// is parallel is the inside of one iteration -- the scan over all contigs, the copies of the merged strings,
// the membership tests -- and what is expensive in the reference is everything around it (a Python scan of
// ~40 string operations per contig and iteration, `item not in nonconcatenatable` over two lists).  So the
// WHOLE loop runs inside ONE persistent thread block per k (k_concat_run): contigs live in an append-only
// byte arena in HBM (slot -> offset, length, hash; strings are immutable, a merge writes a new one), the
// Mersenne Twister of Python's `random` runs on the device (state handed in and out, so the draws are the
// reference's draws after `random.seed(s)`), the list membership is a flag per slot kept exact by comparing
// strings (hash first, then bytes), and nothing goes back to the host between iterations.  The kernel stops
// early only when an iteration found a swap-tolerant match (:85-131): those are decided by
// findSupportReadScore of two candidate strings (dbg_support_scores), so the host scores them, hands back one
// choice per match and the kernel resumes in the middle of that iteration.
#pragma once
#include "dbg_kernels.cuh"

#define CC_ALIVE 1u
#define CC_NONCAT 2u
#define CC_KMAX 32
#define CC_MAX_THREADS 1024
enum { CC_RUNNING = 0, CC_DONE = 1, CC_NEED_CHOICE = 2, CC_NEED_ARENA = 3 };
enum { CC_NONE = 0, CC_HEAD_TAIL = 1, CC_TAIL_HEAD = 2, CC_SWAP_HEAD_TAIL = 3, CC_SWAP_TAIL_HEAD = 4 };

struct ConcatState {
    u32 status, k;
    u32 n_slots, cur;              // cur .. arena_need: the iteration in flight when the kernel stopped early
    u32 n_match, first_equal;
    u64 arena_need;
    u64 arena_top, arena_cap;
    u32 n_list, list_cap;
    u64 iter, max_iter;
    u32 n_alive, max_len;
    u32 mt_pos, pad;
    u32 mt[624];                   // MT19937 state of Python's random.Random (getstate()[1][:624], [624] = mt_pos)
};

DBG_HD u64 cc_mix(u32 byte, u64 pos) { return fmix64(((pos << 8) | (u64)byte) ^ 0x9E3779B97F4A7C15ULL); }

DBG_D bool cc_equal(const u8 *a, const u8 *b, u64 n) {
    for (u64 i = 0; i < n; i++)
        if (a[i] != b[i]) return false;
    return true;
}

// a = k letters of current, b = k letters of the contig: 1 = equal, 2 = equal after exchanging two adjacent
// letters of a (the loops at :85-90 / :108-113 find the same j: the first mismatch), 0 = neither
DBG_D int cc_match(const u8 *a, const u8 *b, int k) {
    int j = 0;
    while (j < k && a[j] == b[j]) j++;
    if (j == k) return 1;
    if (j + 1 >= k || a[j] != b[j + 1] || a[j + 1] != b[j]) return 0;
    for (int i = j + 2; i < k; i++)
        if (a[i] != b[i]) return 0;
    return 2;
}

__global__ void __launch_bounds__(DBG_BLOCK) k_concat_init(const u8 *__restrict__ arena, const u64 *__restrict__ off,
                                                           u32 *__restrict__ len, u64 *__restrict__ hash, u8 *__restrict__ flag, u64 n) {
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 a = off[i], L = off[i + 1] - a;
    u64 h = 0;
    for (u64 j = 0; j < L; j++) h += cc_mix(arena[a + j], j);
    len[i] = (u32)L;
    hash[i] = h;
    flag[i] = CC_ALIVE;
}

// ---- MT19937 as CPython's _randommodule.c runs it (thread 0 only; state in shared memory) ----
DBG_D u32 cc_genrand(u32 *mt, u32 *pos) {
    if (*pos >= 624u) {
        for (int i = 0; i < 624; i++) {
            const u32 y = (mt[i] & 0x80000000u) | (mt[i + 1 == 624 ? 0 : i + 1] & 0x7FFFFFFFu);
            mt[i] = mt[i + 397 < 624 ? i + 397 : i + 397 - 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
        }
        *pos = 0;
    }
    u32 y = mt[(*pos)++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9D2C5680u;
    y ^= (y << 15) & 0xEFC60000u;
    y ^= y >> 18;
    return y;
}
// random.Random._randbelow_with_getrandbits(n), n >= 1: k = n.bit_length(); r = getrandbits(k) until r < n
DBG_D u32 cc_randbelow(u32 *mt, u32 *pos, u32 n) {
    const int bits = 32 - __clz((int)n);
    u32 r;
    do r = cc_genrand(mt, pos) >> (32 - bits); while (r >= n);
    return r;
}

// One block.  resume = 0: start of a k (flags from the lengths, empty nonconcatenatable list, :59);
// resume = 1: continue the iteration the state describes (after CC_NEED_CHOICE / CC_NEED_ARENA).
__global__ void __launch_bounds__(CC_MAX_THREADS) k_concat_run(ConcatState *S, u8 *arena, u64 *off, u32 *len, u64 *hash, u8 *flag,
                                                               u8 *code, u32 *mslot, const u8 *choice, u64 *l_off, u32 *l_len,
                                                               u64 *l_hash, int resume) {
    __shared__ u32 s_mt[624];
    __shared__ u32 s_pos, s_total, s_r, s_cur, s_nmatch, s_swap, s_first, s_status, s_alive, s_maxlen;
    __shared__ u64 s_need, s_iter;
    __shared__ u32 s_cnt[CC_MAX_THREADS];
    __shared__ u8 s_head[CC_KMAX], s_tail[CC_KMAX];
    const u32 tid = threadIdx.x, T = blockDim.x;
    const u32 n = S->n_slots;
    const int k = (int)S->k;
    const u64 max_iter = S->max_iter;
    for (u32 i = tid; i < 624; i += T) s_mt[i] = S->mt[i];
    if (tid == 0) { s_pos = S->mt_pos; s_status = CC_RUNNING; s_iter = resume ? S->iter : 0; }
    if (!resume) {
        for (u32 i = tid; i < n; i += T)
            if (flag[i] & CC_ALIVE) flag[i] = (u8)(CC_ALIVE | (len[i] <= (u32)k ? CC_NONCAT : 0u));
        if (tid == 0) S->n_list = 0;
    }
    __syncthreads();
    const u32 per = (n + T - 1) / T;
    const u32 lo = tid * per < n ? tid * per : n, hi = lo + per < n ? lo + per : n;
    bool pending = resume != 0;
    for (;;) {
        u32 cur, n_match, first_equal;
        u64 need;
        if (!pending) {
            if (s_iter >= max_iter) { if (tid == 0) s_status = CC_DONE; break; }
            // ---- concatednatable (:62) and the draw (:65) ----
            u32 c = 0;
            for (u32 i = lo; i < hi; i++) c += (flag[i] & (CC_ALIVE | CC_NONCAT)) == CC_ALIVE;
            s_cnt[tid] = c;
            __syncthreads();
            if (tid == 0) {
                u32 total = 0;
                for (u32 t = 0; t < T; t++) { const u32 x = s_cnt[t]; s_cnt[t] = total; total += x; }
                s_total = total;
                if (total) s_r = cc_randbelow(s_mt, &s_pos, total);
                s_nmatch = 0; s_swap = 0; s_first = 0xFFFFFFFFu; s_need = 0;
            }
            __syncthreads();
            if (s_total == 0) { if (tid == 0) s_status = CC_DONE; break; }
            const u32 r = s_r;
            if (c && s_cnt[tid] <= r && r < s_cnt[tid] + c) {
                u32 left = r - s_cnt[tid];
                for (u32 i = lo; i < hi; i++)
                    if ((flag[i] & (CC_ALIVE | CC_NONCAT)) == CC_ALIVE && left-- == 0) { s_cur = i; break; }
            }
            __syncthreads();
            cur = s_cur;
            const u64 coff = off[cur];
            const u32 clen = len[cur];
            const u64 chash = hash[cur];
            if (tid < (u32)k) { s_head[tid] = arena[coff + tid]; s_tail[tid] = arena[coff + clen - k + tid]; }
            __syncthreads();
            // ---- the scan over all contigs (:66-131) ----
            for (u32 i = tid; i < n; i += T) {
                u32 cd = CC_NONE;
                const u32 L = len[i];
                if ((flag[i] & CC_ALIVE) && L > (u32)k) {
                    const u8 *p = arena + off[i];
                    if (L == clen && hash[i] == chash && cc_equal(p, arena + coff, clen)) {
                        atomicMin(&s_first, i);                                     // contig == current (:70)
                    } else {
                        const int ht = cc_match(s_head, p + L - k, k);              // current's head on the contig's tail
                        if (ht == 1) cd = CC_HEAD_TAIL;                            // :76
                        else {
                            const int th = cc_match(s_tail, p, k);                  // current's tail on the contig's head
                            if (th == 1) cd = CC_TAIL_HEAD;                        // :80
                            else if (ht == 2) cd = CC_SWAP_HEAD_TAIL;              // :85-106
                            else if (th == 2) cd = CC_SWAP_TAIL_HEAD;              // :108-131
                        }
                    }
                }
                code[i] = (u8)cd;
                if (cd) {
                    mslot[atomicAdd(&s_nmatch, 1u)] = i;
                    atomicAdd(&s_need, (u64)(L + clen - (u32)k));
                    if (cd >= CC_SWAP_HEAD_TAIL) s_swap = 1;
                }
            }
            __syncthreads();
            n_match = s_nmatch; first_equal = s_first; need = s_need;
            if (s_swap) {                                                           // the host decides result_1 / result_2
                if (tid == 0) {
                    S->cur = cur; S->n_match = n_match; S->first_equal = first_equal; S->arena_need = need;
                    s_status = CC_NEED_CHOICE;
                }
                break;
            }
        } else {
            cur = S->cur; n_match = S->n_match; first_equal = S->first_equal; need = S->arena_need;
            pending = false;
        }
        const u64 coff = off[cur];
        const u32 clen = len[cur];
        if (n_match) {
            // ---- the merged strings (:77, :81, :91-100, :114-123) ----
            if (S->arena_top + need > S->arena_cap) {
                if (tid == 0) {
                    S->cur = cur; S->n_match = n_match; S->first_equal = first_equal; S->arena_need = need;
                    s_status = CC_NEED_ARENA;
                }
                break;
            }
            __syncthreads();                                                        // everyone has read arena_top
            const u32 lane = tid & 31u, warp = tid >> 5, n_warps = (T + 31u) >> 5;
            const u32 n_list = S->n_list;
            for (u32 m = warp; m < n_match; m += n_warps) {
                const u32 i = mslot[m];
                const u32 cd = code[i], L = len[i];
                const u8 *ps = arena + off[i], *pc = arena + coff;
                const bool second = cd >= CC_SWAP_HEAD_TAIL && choice[i] == 2;
                const u8 *A, *B;
                u32 nA, nB;
                if (cd == CC_HEAD_TAIL || cd == CC_SWAP_HEAD_TAIL) {
                    if (!second) { A = ps; nA = L; B = pc + k; nB = clen - k; }      // contig + current[k:]
                    else { A = ps; nA = L - k; B = pc; nB = clen; }                  // contig[:len-k] + current
                } else {
                    if (!second) { A = pc; nA = clen; B = ps + k; nB = L - k; }      // current + contig[k:]
                    else { A = pc; nA = clen - k; B = ps; nB = L; }                  // current[:len-k] + contig
                }
                const u32 nl = nA + nB;
                u64 dst = 0;
                if (lane == 0) dst = atomicAdd(&S->arena_top, (u64)nl);
                dst = __shfl_sync(0xFFFFFFFFu, dst, 0);
                u64 hs = 0;
                for (u32 j = lane; j < nl; j += 32) {
                    const u8 b = j < nA ? A[j] : B[j - nA];
                    arena[dst + j] = b;
                    hs += cc_mix(b, j);
                }
                for (int d = 16; d > 0; d >>= 1) hs += __shfl_down_sync(0xFFFFFFFFu, hs, d);
                hs = __shfl_sync(0xFFFFFFFFu, hs, 0);
                __syncwarp();
                int in_list = 0;                                                    // is the new string in nonconcatenatable?
                for (u32 e = lane; e < n_list; e += 32)
                    if (l_len[e] == nl && l_hash[e] == hs && cc_equal(arena + l_off[e], arena + dst, nl)) in_list = 1;
                in_list = __any_sync(0xFFFFFFFFu, in_list);
                if (lane == 0) {
                    off[i] = dst; len[i] = nl; hash[i] = hs;
                    flag[i] = (u8)(CC_ALIVE | (in_list ? CC_NONCAT : 0u));
                }
            }
            __syncthreads();
            if (tid == 0) flag[first_equal] = 0;                                    // contigs.remove(current) (:134)
        } else {
            // ---- nonconcatenatable.append(current) (:136) ----
            const u64 chash = hash[cur];
            if (tid == 0) {
                const u32 e = S->n_list;
                l_off[e] = coff; l_len[e] = clen; l_hash[e] = chash;
                S->n_list = e + 1;
            }
            for (u32 i = tid; i < n; i += T)
                if ((flag[i] & CC_ALIVE) && len[i] == clen && hash[i] == chash && cc_equal(arena + off[i], arena + coff, clen))
                    flag[i] |= CC_NONCAT;
        }
        if (tid == 0) s_iter++;
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0) { s_alive = 0; s_maxlen = 0; }
    __syncthreads();
    u32 a = 0, ml = 0;
    for (u32 i = tid; i < n; i += T)
        if (flag[i] & CC_ALIVE) { a++; ml = len[i] > ml ? len[i] : ml; }
    atomicAdd(&s_alive, a);
    atomicMax(&s_maxlen, ml);
    for (u32 i = tid; i < 624; i += T) S->mt[i] = s_mt[i];
    __syncthreads();
    if (tid == 0) {
        S->mt_pos = s_pos; S->status = s_status; S->iter = s_iter;
        S->n_alive = s_alive; S->max_len = s_maxlen;
    }
}

// selected strings, back to back (dst_off[j] .. dst_off[j + 1]): one block per string
__global__ void __launch_bounds__(DBG_BLOCK) k_concat_gather(const u8 *__restrict__ arena, const u64 *__restrict__ off,
                                                             const u32 *__restrict__ len, const u32 *__restrict__ slots,
                                                             const u64 *__restrict__ dst_off, u8 *__restrict__ dst) {
    const u32 i = slots[blockIdx.x];
    const u8 *p = arena + off[i];
    u8 *q = dst + dst_off[blockIdx.x];
    for (u32 j = threadIdx.x; j < len[i]; j += blockDim.x) q[j] = p[j];
}

// III_concatenateOutpus.py:139-142.  A contig is kept iff it is not a substring of an earlier KEPT one, which is
// the same as: not a substring of ANY earlier contig (a dropped one is inside a kept one that is earlier still),
// so every contig is decided independently: one block per contig j, threads over the positions of contig i < j.
__global__ void __launch_bounds__(DBG_BLOCK) k_concat_dedup(const u8 *__restrict__ arena, const u64 *__restrict__ off,
                                                            const u32 *__restrict__ len, const u8 *__restrict__ flag, u32 n,
                                                            u8 *__restrict__ keep) {
    __shared__ u32 s_found;
    const u32 j = blockIdx.x;
    if (!(flag[j] & CC_ALIVE)) { if (threadIdx.x == 0) keep[j] = 0; return; }
    if (threadIdx.x == 0) s_found = 0;
    __syncthreads();
    const u8 *q = arena + off[j];
    const u32 Lj = len[j];
    for (u32 i = 0; i < j; i++) {
        if (!(flag[i] & CC_ALIVE) || len[i] < Lj) continue;
        const u8 *p = arena + off[i];
        const u32 np = len[i] - Lj + 1;
        for (u32 s = threadIdx.x; s < np; s += blockDim.x)
            if (cc_equal(p + s, q, Lj)) s_found = 1;
        __syncthreads();
        if (s_found) break;
        __syncthreads();
    }
    if (threadIdx.x == 0) keep[j] = s_found ? 0 : 1;
}
