// dbg_concat_host.cuh -- host side of the dbg_concat_* calls (kernels and the design: dbg_concat.cuh).
// Included at the end of dbg_b200.cu (uses its handle, fail / CK / TRY / reserve / LAUNCH).
#pragma once

struct ConcatSession {
    DevBuf arena, off, len, hash, flag, code, mslot, choice, l_off, l_len, l_hash, state, gbytes, gslots, goff, keep;
    uint64_t n = 0;
    ConcatState st;                               // host mirror of the device state
    bool pending = false;                         // the kernel stopped inside an iteration (CC_NEED_CHOICE)
    std::vector<uint32_t> pend_slots;
    std::vector<uint8_t> pend_codes;
};

static void concat_free(dbg_t *h) {
    ConcatSession *c = h->cc;
    if (!c) return;
    DevBuf *all[] = {&c->arena, &c->off, &c->len, &c->hash, &c->flag, &c->code, &c->mslot, &c->choice, &c->l_off, &c->l_len,
                     &c->l_hash, &c->state, &c->gbytes, &c->gslots, &c->goff, &c->keep};
    for (DevBuf *b : all) if (b->p) cudaFree(b->p);
    delete c;
    h->cc = nullptr;
}

#define NEED_CONCAT() do { if (!h) return DBG_ERR_ARG; if (!h->cc) return fail(h, DBG_ERR_STATE, "no concatenation session: call dbg_concat_begin first"); } while (0)

DBG_API int dbg_concat_begin(dbg_t *h, const uint8_t *contigs, const uint64_t *contig_off, uint64_t n_contigs) {
    if (!h) return DBG_ERR_ARG;
    if (!contig_off || n_contigs == 0) return fail(h, DBG_ERR_ARG, "dbg_concat_begin needs at least one contig");
    if (n_contigs >= 0x7FFFFFF0ull) return fail(h, DBG_ERR_OVERFLOW, "more than 2^31 contigs");
    CK(cudaSetDevice(h->device));
    concat_free(h);
    const uint64_t total = contig_off[n_contigs];
    if (total && !contigs) return fail(h, DBG_ERR_ARG, "NULL argument");
    for (uint64_t i = 0; i < n_contigs; i++)
        if (contig_off[i + 1] < contig_off[i] || contig_off[i + 1] - contig_off[i] > 0x7FFFFFFFull)
            return fail(h, DBG_ERR_ARG, "contig offsets must ascend (a contig holds at most 2^31-1 bytes)");
    ConcatSession *c = h->cc = new ConcatSession();
    c->n = n_contigs;
    memset(&c->st, 0, sizeof c->st);
    uint64_t cap = 4 * total + (1ull << 20);
    if (const char *e = getenv("DBG_B200_CC_ARENA")) cap = total + (uint64_t)atoll(e);   // tests: force the arena to grow
    TRY(reserve(h, c->arena, cap));
    TRY(reserve(h, c->off, (n_contigs + 1) * 8));
    TRY(reserve(h, c->len, n_contigs * 4));
    TRY(reserve(h, c->hash, n_contigs * 8));
    TRY(reserve(h, c->flag, n_contigs));
    TRY(reserve(h, c->code, n_contigs));
    TRY(reserve(h, c->mslot, n_contigs * 4));
    TRY(reserve(h, c->choice, n_contigs));
    TRY(reserve(h, c->state, sizeof(ConcatState)));
    if (total) CK(cudaMemcpyAsync(c->arena.p, contigs, total, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(c->off.p, contig_off, (n_contigs + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(c->choice.p, 0, n_contigs, h->stream));
    LAUNCH(k_concat_init, blocks_for(n_contigs), DBG_BLOCK, h->stream, ptr<const u8>(c->arena), ptr<const u64>(c->off),
           ptr<u32>(c->len), ptr<u64>(c->hash), ptr<u8>(c->flag), n_contigs);
    CKK();
    CK(cudaStreamSynchronize(h->stream));
    c->st.n_slots = (u32)n_contigs;
    c->st.arena_top = total;
    c->st.arena_cap = c->arena.cap;
    c->st.n_alive = (u32)n_contigs;
    return DBG_OK;
}

static int concat_threads(uint64_t n) {
#ifdef DBG_EMU
    (void)n;
    int t = 64;
#else
    int t = n <= 256 ? 256 : (n <= 512 ? 512 : CC_MAX_THREADS);
#endif
    if (const char *e = getenv("DBG_B200_CC_THREADS")) {
        const int v = atoi(e);
        if (v >= 32 && v <= CC_MAX_THREADS && v % 32 == 0) t = v;
    }
    return t;
}

DBG_API int dbg_concat_run(dbg_t *h, int k, uint64_t max_iter, uint32_t *mt_state, int resume, dbg_concat_status_t *out) {
    NEED_CONCAT();
    ConcatSession *c = h->cc;
    if (!mt_state || !out) return fail(h, DBG_ERR_ARG, "NULL argument");
    if (k < 1 || k > CC_KMAX) return fail(h, DBG_ERR_ARG, "dbg_concat_run: k = %d is outside 1..%d", k, CC_KMAX);
    if (mt_state[624] > 624) return fail(h, DBG_ERR_ARG, "not a Mersenne Twister state: position %u", mt_state[624]);
    if (resume && !c->pending) return fail(h, DBG_ERR_STATE, "nothing to resume");
    if (!resume && c->pending) return fail(h, DBG_ERR_STATE, "an iteration is waiting for dbg_concat_choose + resume");
    if (resume && ((int)c->st.k != k || c->st.max_iter != max_iter)) return fail(h, DBG_ERR_ARG, "resume with a different k or max_iter");
    if (max_iter >= 0xFFFFFFF0ull) return fail(h, DBG_ERR_ARG, "max_iter too large");
    CK(cudaSetDevice(h->device));
    if (!resume) {                                  // nonconcatenatable grows by at most one string per iteration
        TRY(reserve(h, c->l_off, (max_iter + 1) * 8));
        TRY(reserve(h, c->l_len, (max_iter + 1) * 4));
        TRY(reserve(h, c->l_hash, (max_iter + 1) * 8));
        c->st.k = (u32)k;
        c->st.max_iter = max_iter;
        c->st.list_cap = (u32)(max_iter + 1);
    }
    memcpy(c->st.mt, mt_state, 624 * 4);
    c->st.mt_pos = mt_state[624];
    const int threads = concat_threads(c->n);
    int res = resume ? 1 : 0;
    for (;;) {
        c->st.status = CC_RUNNING;
        CK(cudaMemcpyAsync(c->state.p, &c->st, sizeof(ConcatState), cudaMemcpyHostToDevice, h->stream));
        LAUNCH(k_concat_run, 1, threads, h->stream, ptr<ConcatState>(c->state), ptr<u8>(c->arena), ptr<u64>(c->off), ptr<u32>(c->len),
               ptr<u64>(c->hash), ptr<u8>(c->flag), ptr<u8>(c->code), ptr<u32>(c->mslot), ptr<const u8>(c->choice),
               ptr<u64>(c->l_off), ptr<u32>(c->l_len), ptr<u64>(c->l_hash), res);
        CKK();
        CK(cudaMemcpyAsync(&c->st, c->state.p, sizeof(ConcatState), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (c->st.status != CC_NEED_ARENA) break;
        // the merged strings of the iteration in flight do not fit: a larger arena, then the same iteration again
        const uint64_t want = std::max(2 * c->st.arena_cap, c->st.arena_top + c->st.arena_need + c->st.arena_cap / 2);
        DevBuf bigger;
        TRY(reserve(h, bigger, want));
        CK(cudaMemcpyAsync(bigger.p, c->arena.p, c->st.arena_top, cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        cudaFree(c->arena.p);
        c->arena = bigger;
        c->st.arena_cap = bigger.cap;
        res = 1;
    }
    memcpy(mt_state, c->st.mt, 624 * 4);
    mt_state[624] = c->st.mt_pos;
    c->pending = false;
    c->pend_slots.clear();
    c->pend_codes.clear();
    if (c->st.status == CC_NEED_CHOICE) {
        std::vector<uint8_t> code(c->n);
        CK(cudaMemcpyAsync(code.data(), c->code.p, c->n, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (uint64_t i = 0; i < c->n; i++)
            if (code[i]) { c->pend_slots.push_back((uint32_t)i); c->pend_codes.push_back(code[i]); }
        c->pending = true;
    } else if (c->st.status != CC_DONE) {
        return fail(h, DBG_ERR_STATE, "dbg_concat_run: the kernel left status %u", c->st.status);
    }
    out->status = c->st.status == CC_DONE ? DBG_CONCAT_DONE : DBG_CONCAT_NEED_CHOICE;
    out->current = c->st.cur;
    out->iterations = c->st.iter;
    out->n_alive = c->st.n_alive;
    out->max_len = c->st.max_len;
    out->n_pending = (uint32_t)c->pend_slots.size();
    out->arena_bytes = c->st.arena_top;
    return DBG_OK;
}

DBG_API int dbg_concat_pending(dbg_t *h, uint32_t *slots, uint8_t *codes) {
    NEED_CONCAT();
    ConcatSession *c = h->cc;
    if (!c->pending) return fail(h, DBG_ERR_STATE, "no iteration is waiting");
    if (!slots || !codes) return fail(h, DBG_ERR_ARG, "NULL argument");
    memcpy(slots, c->pend_slots.data(), c->pend_slots.size() * 4);
    memcpy(codes, c->pend_codes.data(), c->pend_codes.size());
    return DBG_OK;
}

DBG_API int dbg_concat_choose(dbg_t *h, const uint8_t *choice) {
    NEED_CONCAT();
    ConcatSession *c = h->cc;
    if (!c->pending) return fail(h, DBG_ERR_STATE, "no iteration is waiting");
    if (!choice) return fail(h, DBG_ERR_ARG, "NULL argument");
    std::vector<uint8_t> all(c->n, 0);
    for (size_t j = 0; j < c->pend_slots.size(); j++) {
        if (c->pend_codes[j] >= CC_SWAP_HEAD_TAIL && choice[j] != 1 && choice[j] != 2)
            return fail(h, DBG_ERR_ARG, "choice %zu must be 1 (result_1) or 2 (result_2)", j);
        all[c->pend_slots[j]] = choice[j];
    }
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(c->choice.p, all.data(), c->n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DBG_OK;
}

DBG_API int dbg_concat_lengths(dbg_t *h, uint32_t *len, uint8_t *alive) {
    NEED_CONCAT();
    ConcatSession *c = h->cc;
    if (!len || !alive) return fail(h, DBG_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(len, c->len.p, c->n * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(alive, c->flag.p, c->n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (uint64_t i = 0; i < c->n; i++) alive[i] &= CC_ALIVE;
    return DBG_OK;
}

DBG_API int dbg_concat_strings(dbg_t *h, const uint32_t *slots, uint64_t n, uint8_t *bytes, const uint64_t *off) {
    NEED_CONCAT();
    ConcatSession *c = h->cc;
    if (n == 0) return DBG_OK;
    if (!slots || !off || (off[n] && !bytes)) return fail(h, DBG_ERR_ARG, "NULL argument");
    for (uint64_t j = 0; j < n; j++)
        if (slots[j] >= c->n) return fail(h, DBG_ERR_ARG, "slot out of range");
    CK(cudaSetDevice(h->device));
    std::vector<uint32_t> len(c->n);
    CK(cudaMemcpyAsync(len.data(), c->len.p, c->n * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (uint64_t j = 0; j < n; j++)
        if (off[j + 1] - off[j] != len[slots[j]]) return fail(h, DBG_ERR_ARG, "offsets do not match the lengths (dbg_concat_lengths)");
    TRY(reserve(h, c->gslots, n * 4));
    TRY(reserve(h, c->goff, (n + 1) * 8));
    TRY(reserve(h, c->gbytes, off[n] - off[0] + 16));
    std::vector<uint64_t> rel(n + 1);
    for (uint64_t j = 0; j <= n; j++) rel[j] = off[j] - off[0];
    CK(cudaMemcpyAsync(c->gslots.p, slots, n * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(c->goff.p, rel.data(), (n + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    LAUNCH(k_concat_gather, n, DBG_BLOCK, h->stream, ptr<const u8>(c->arena), ptr<const u64>(c->off), ptr<const u32>(c->len),
           ptr<const u32>(c->gslots), ptr<const u64>(c->goff), ptr<u8>(c->gbytes));
    CKK();
    if (rel[n]) CK(cudaMemcpyAsync(bytes + off[0], c->gbytes.p, rel[n], cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DBG_OK;
}

DBG_API int dbg_concat_dedup(dbg_t *h, uint8_t *keep) {
    NEED_CONCAT();
    ConcatSession *c = h->cc;
    if (!keep) return fail(h, DBG_ERR_ARG, "NULL argument");
    if (c->pending) return fail(h, DBG_ERR_STATE, "an iteration is waiting for dbg_concat_choose + resume");
    CK(cudaSetDevice(h->device));
    TRY(reserve(h, c->keep, c->n));
    LAUNCH(k_concat_dedup, c->n, DBG_BLOCK, h->stream, ptr<const u8>(c->arena), ptr<const u64>(c->off), ptr<const u32>(c->len),
           ptr<const u8>(c->flag), (u32)c->n, ptr<u8>(c->keep));
    CKK();
    CK(cudaMemcpyAsync(keep, c->keep.p, c->n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DBG_OK;
}

DBG_API int dbg_concat_end(dbg_t *h) {
    if (!h) return DBG_ERR_ARG;
    cudaSetDevice(h->device);
    concat_free(h);
    return DBG_OK;
}
