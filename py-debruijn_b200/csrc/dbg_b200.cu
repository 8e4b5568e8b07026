// dbg_b200.cu -- host side of the C ABI declared in include/dbg_b200.h.
//
// One handle = one graph on one device.  dbg_build* runs the construct_graph stages
// (debruijn.py:206-285), dbg_walk the output_contigs stage (debruijn.py:326-347); every stage is a
// kernel in dbg_kernels.cuh.  The only work done on the host is what the reference defines
// sequentially and what is tiny: the branch-order conflict resolution of tip pruning
// (debruijn.py:241-254, over the candidate paths the device enumerated) and the path enumeration
// of final=True through surviving multi-successor nodes (debruijn.py:314, over the junction table
// the device ranked).  There is no CPU implementation of any stage behind this ABI: without a
// CUDA device dbg_create fails.
//
// tests/emu builds this same file with g++ and tests/emu/cuda_emu.h (-DDBG_EMU) so that the CPU
// test-suite can run the host logic and kernel bodies; that library is test infrastructure and is
// never loaded by the product.
#include <math.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <unordered_set>
#include <utility>
#include <vector>

#ifdef DBG_EMU
#define LAUNCH(kernel, grid, block, stream, ...) \
    (h->st.n_launches++, emu::launch(dim3((unsigned)(grid)), dim3((unsigned)(block)), 0, [&]() { kernel(__VA_ARGS__); }))
#define LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) LAUNCH(kernel, grid, block, stream, __VA_ARGS__)
#else
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#define LAUNCH(kernel, grid, block, stream, ...) \
    (h->st.n_launches++, kernel<<<(unsigned)(grid), (unsigned)(block), 0, stream>>>(__VA_ARGS__))
#define LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) \
    (h->st.n_launches++, kernel<<<(unsigned)(grid), (unsigned)(block), (smem), stream>>>(__VA_ARGS__))
#endif

#include "../../include/dbg_b200.h"
#include "dbg_kernels.cuh"
#include "dbg_chunks.cuh"
#include "dbg_shard.cuh"
#include "dbg_generic.cuh"
#include "dbg_support.cuh"
#include "dbg_concat.cuh"

#define DBG_ABI_VERSION 1
#define DBG_RETRY_COLD 1                      // internal: a build sized from the handle's hint did not fit; redo it cold
#define MAX_ENUM_CONTIGS (1ull << 26)
#define MAX_ENUM_STEPS (1ull << 32)

// -------------------------------------------------------------------------------------------------
// handle
// -------------------------------------------------------------------------------------------------
struct DevBuf {                 // grow-only device buffer: repeated builds on one handle reuse HBM
    void *p = nullptr;
    size_t cap = 0;
};

enum { EV_MAX = 64, EV_WALK = 48, MAX_SLICES = 32 };
#define CTL_BYTES 65536
#define CTL_SKETCH 0                       // HLL_M x u32
#define CTL_CC (HLL_M * 4)                 // ChunkCounters
#define CTL_CTR (HLL_M * 4 + 1024)         // DevCounters
#define SLICE_BYTES (16ull << 20)         // host reads are uploaded in slices of this size

struct dbg_handle {
    int device = 0;
    int k = 0;
    int W = 1;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    bool ext_stream = false;                  // the caller put the handle on ITS stream (dbg_set_stream): device buffers a call
                                              //   leaves for the caller are ordered by that stream, no host wait at the end of the call
    const uint8_t *host_bases = nullptr;      // dbg_build*: the reads still sit in host memory and are copied in
                                              //   slices on copy_stream while k_pack / k_masks work on earlier ones
    cudaEvent_t slice_ev[MAX_SLICES + 1];
    int geom_logw = 0, geom_m = 1;
    uint32_t sample_mask = 0;
    std::string err;
    bool built = false, walked = false;
    uint32_t build_flags = 0;
    double threshold = 3.0;
    uint64_t n_reads = 0, n_bases = 0, n_words = 0, n_tiles = 0, n_slots = 0, read_idx_base = 0;
    float pending_h2d = 0;
    bool table_ready = false, order_active = false;
    uint64_t order_cap = 0;
    int order_bits = 0;
    const uint64_t *d_off = nullptr;          // offsets in use (caller's for build_device, ours for build)
    const uint32_t *d_read_batch = nullptr;   // batched build: batch id per read (device), else NULL
    int batch_bits = 0;                       // bits the batch id takes above the 2k key bits
    uint32_t n_batches = 0;
    DevBuf b_bases, b_off, b_packed, b_tile, b_table, b_flags, b_kept, b_succ, b_link, b_ctr, b_hll,
           b_order, b_tag, b_tag2, b_slot, b_slot2, b_sorttmp, b_branch, b_starts, b_pulled, b_readflags,
           b_recs, b_cnt32, b_off64, b_contig, b_junction, b_split, b_h8, b_bitmap, b_stmask, b_vmask,
           b_spanread, b_cc, b_ctable, b_clist, b_super, b_readbatch, b_ctable2, b_scratch[6],
           b_exitslot, b_reqslot, b_rvslot, b_rvord, b_link2, b_scontig, b_owneroff;
    DevCounters ctr;
    dbg_stats_t st;
    std::vector<uint32_t> branch, pulled;     // reference order
    std::vector<uint8_t> read_flags;
    bool read_flags_valid = false;
    std::vector<dbg_contig_t> contigs;        // slow path (dropped starts, final=True enumeration)
    const dbg_contig_t *cview = nullptr;      // what dbg_contigs / materialize / scores read: `contigs`, or the
    uint64_t n_cview = 0;                     //   pinned buffer the device's records landed in (fast path)
    unsigned char *ctl = nullptr;             // pinned 64 KB for the small control read-backs (sketch registers,
                                              //   counters): a pageable destination would make every such copy a host sync
    void *pin = nullptr;                      // pinned host staging, grow-only
    size_t pin_cap = 0;
    std::vector<dbg_segment_t> segments;
    cudaEvent_t ev[EV_MAX];
    int n_ev = 0;                             // build stages use events [0, EV_WALK), dbg_walk [EV_WALK, EV_MAX)
    int n_ev_walk = EV_WALK;
    struct TimeRec { float *dst; int a, b; bool add; };
    std::vector<TimeRec> tm;                  // stage times still to be read from their events (after a sync)
    bool expand_pending = false;              // the node table was built without looking at its overflow flag yet
    bool chunks_merged = false;               // the chunk table holds records merged from other ranks (dbg_chunk_merge_records)
    bool chunks_valid = false;                // the chunk table + list of the last build still describe the local reads
    bool tags_local = false;                  // every tag in the table comes from this handle's own reads
    bool chunks_streamed = false;             // phase 1 already ran, slice by slice, while the reads were arriving
    uint64_t n_cslots = 0, n_clist = 0, ex_grid = 1;
    // generic-alphabet engine (dbg_generic.cuh): generic, sym_bits = 1 .. 8 by the number of symbols; the DNA engine has 2
    bool generic = false;
    int sym_bits = 2, n_symbols = 4;
    uint8_t sym_of_code[256 + 1] = {'A', 'C', 'G', 'T'}, code_of_byte[256];   // [256] = sym_bits (the device copy's layout)
    uint64_t n_eslots = 0, n_earr = 0;
    // Table sizes remembered from the last successful build on this handle (like the grow-only buffers): a build of
    // a similar read set sizes its chunk and node table from them at once -- no sketch round trips, no sketch pass,
    // tables initialised on a side stream under the packing -- and CHECKS afterwards (overflow flags, read with the
    // counters graph_finalize fetches anyway); a build they do not fit is redone the cold way.
    struct { bool valid = false; uint64_t n_bases = 0, chunks = 0, nodes = 0; uint32_t n_batches = 0; } hint;
    bool hinted = false, preinit = false;
    bool have_hbytes = false;                 // a sketch pass (k_expand MODE 0) left hash bytes for the chunks now in the tables
    // node table laid out by minimizer (TGeom, dbg_common.cuh): region_slots > 0.  region_mode is decided per build
    // (minimizer long enough to spread over regions); have_vhash / have_order: the per-record region hashes +
    // histogram / the record order of the chunks now in the tables are valid
    bool region_mode = false, have_vhash = false, have_order = false;
    uint32_t region_slots = 0;
    uint32_t min_count = 1;                   // EXTENSION (dbg_set_min_count): windows whose (k+1)-mer is rarer are skipped
    int expand_evs = 3;                       // events the last expand_node_table marked
    int vb_bits = 0;
    DevBuf b_recvh, b_vhist, b_voffs, b_recorder;
    uint64_t last_est_chunks = 0;
    uint64_t shard_hint_nodes = 0;            // nodes of the last shard this handle expanded (dbg_chunk_expand)
    bool merge_check_pending = false;         // dbg_chunk_merge_records' overflow flag is looked at by graph_finalize
    cudaStream_t init_stream = nullptr;
    cudaEvent_t ev_init[3];                   // [0] build start, [1] chunk table + boundary list cleared, [2] node table cleared
    DevBuf b_up;                              // lifted successor arrays (strings / scores), built on demand
    bool up_ready = false;
    DevBuf b_codes, b_etable, b_earr, b_earr2, b_estart, b_edeg, b_enk, b_lut;
    // sharded node table (dbg_shard.cuh): this handle holds shard `shard_rank` of `shard_n` (0 / 1 = not sharded)
    uint32_t shard_rank = 0, shard_n = 0;
    uint64_t n_exit = 0, n_req = 0, n_rv = 0;
    std::vector<uint64_t> exit_off;           // exits grouped by the owner of their successor: n_ranks + 1 offsets
    std::vector<dbg_shard_contig_t> scontigs;
    const dbg_shard_contig_t *sview = nullptr;  // what dbg_shard_contigs reads: the pinned records, or `scontigs` when some were dropped
    uint64_t n_sview = 0;
    int shard_stage = 0;                      // 0 none, 1 exits listed, 2 entries marked, 3 ranked locally, 4 reduced records out, 5 walked
    DevBuf b_bases2, b_off2;                  // DBG_BOTH_STRANDS: the reads followed by their reverse complements
    bool avoid_regions = false;               // a build on this handle overflowed its regions (skewed minimizers): plain layout from then on
    struct ConcatSession *cc = nullptr;       // dbg_concat_* (III_concatenateOutpus.py:57-142; dbg_concat_host.cuh)
};
static void concat_free(dbg_t *h);

static std::string g_create_error;
static TGeom geom_of(const dbg_t *h);

static int fail(dbg_t *h, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(h, e_ == cudaErrorMemoryAllocation ? DBG_ERR_NOMEM : DBG_ERR_CUDA,        \
                        "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));       \
    } while (0)
#define CKK() CK(cudaGetLastError())
#define TRY(expr) do { int r_ = (expr); if (r_ != DBG_OK) return r_; } while (0)

static int reserve(dbg_t *h, DevBuf &b, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return DBG_OK;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 16 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        b.p = nullptr;
        return fail(h, DBG_ERR_NOMEM, "device allocation of %zu bytes failed: %s", want, cudaGetErrorString(e));
    }
    b.cap = want;
    return DBG_OK;
}
template <typename T> static T *ptr(DevBuf &b) { return (T *)b.p; }

static int reserve_pinned(dbg_t *h, size_t bytes) {
    if (h->pin_cap >= bytes) return DBG_OK;
#ifdef DBG_EMU
    if (h->pin) free(h->pin);
    h->pin = malloc(bytes + bytes / 8 + 4096);
    if (!h->pin) return fail(h, DBG_ERR_NOMEM, "host allocation failed");
#else
    if (h->pin) { cudaFreeHost(h->pin); h->pin = nullptr; h->pin_cap = 0; }
    if (cudaMallocHost(&h->pin, bytes + bytes / 8 + 4096) != cudaSuccess) {
        h->pin = nullptr;
        return fail(h, DBG_ERR_NOMEM, "pinned host allocation of %zu bytes failed", bytes);
    }
#endif
    h->pin_cap = bytes + bytes / 8 + 4096;
    return DBG_OK;
}

// end of a call whose only result is a DEVICE buffer the caller will use (or feed to a collective)
#define SYNC_FOR_CALLER() do { if (!h->ext_stream) CK(cudaStreamSynchronize(h->stream)); } while (0)

static inline uint64_t cdiv(uint64_t a, uint64_t b) { return (a + b - 1) / b; }
static inline uint64_t blocks_for(uint64_t n) { return n ? cdiv(n, DBG_BLOCK) : 1; }

static int ev_mark(dbg_t *h) {                       // returns the event index
    int i = h->n_ev++;
    cudaEventRecord(h->ev[i], h->stream);
    return i;
}
static float ev_ms(dbg_t *h, int a, int b) {
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[a], h->ev[b]);
    return ms;
}
static int ev_mark_walk(dbg_t *h) {
    int i = h->n_ev_walk++;
    cudaEventRecord(h->ev[i], h->stream);
    return i;
}
// Stage times are read from their events later, when the stream is idle anyway (dbg_stats, or the next
// sync the data path needs): a stage never waits for the GPU just to time itself.
static void tm_set(dbg_t *h, float *dst, int a, int b) { h->tm.push_back(dbg_handle::TimeRec{dst, a, b, false}); }
static void tm_add(dbg_t *h, float *dst, int a, int b) { h->tm.push_back(dbg_handle::TimeRec{dst, a, b, true}); }
static void tm_flush(dbg_t *h) {                      // the stream must be idle
    for (const dbg_handle::TimeRec &r : h->tm) {
        const float ms = ev_ms(h, r.a, r.b);
        if (r.add) *r.dst += ms; else *r.dst = ms;
    }
    h->tm.clear();
}

// -------------------------------------------------------------------------------------------------
// sort (tag, slot) pairs by tag: the reference's dict order is first-occurrence order
// -------------------------------------------------------------------------------------------------
static int sort_by_tag(dbg_t *h, uint64_t n, u32 *d_sorted_slots_out, int begin_bit) {
    // tags are (global read index << 32 | pos): the read index needs only so many bits
    int end_bit = 64;
    if (h->tags_local) {                                // not for a table merged from other ranks' records
        end_bit = 33;
        while (end_bit < 64 && ((h->read_idx_base + h->n_reads) >> (end_bit - 32)) != 0) end_bit++;
    }
    // input pairs are in b_tag / b_slot; sorted slots land in d_sorted_slots_out
    if (n == 0) return DBG_OK;
#ifdef DBG_EMU
    std::vector<std::pair<u64, u32>> v(n);
    u64 *t = ptr<u64>(h->b_tag);
    u32 *s = ptr<u32>(h->b_slot);
    for (uint64_t i = 0; i < n; i++) v[i] = std::make_pair(t[i], s[i]);
    std::sort(v.begin(), v.end());
    for (uint64_t i = 0; i < n; i++) d_sorted_slots_out[i] = v[i].second;
#else
    TRY(reserve(h, h->b_tag2, n * 8));
    size_t tmp = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, ptr<u64>(h->b_tag), ptr<u64>(h->b_tag2), ptr<u32>(h->b_slot),
                                       d_sorted_slots_out, (int64_t)n, begin_bit, end_bit, h->stream));
    TRY(reserve(h, h->b_sorttmp, tmp));
    h->st.n_launches++;                                   // cub launches several kernels; counted as one
    CK(cub::DeviceRadixSort::SortPairs(h->b_sorttmp.p, tmp, ptr<u64>(h->b_tag), ptr<u64>(h->b_tag2),
                                       ptr<u32>(h->b_slot), d_sorted_slots_out, (int64_t)n, begin_bit, end_bit, h->stream));
#endif
    return DBG_OK;
}

// -------------------------------------------------------------------------------------------------
// templated stage drivers
// -------------------------------------------------------------------------------------------------
static double hll_estimate(const std::vector<u32> &regs) {
    const double m = (double)HLL_M;
    double sum = 0;
    int zeros = 0;
    for (u32 r : regs) { sum += ldexp(1.0, -(int)r); zeros += (r == 0); }
    double alpha = 0.7213 / (1.0 + 1.079 / m);
    double e = alpha * m * m / sum;
    if (e <= 2.5 * m && zeros) e = m * log(m / (double)zeros);      // small-range correction
    return e;
}

static int fetch_counters(dbg_t *h) {
    void *dst = h->ctl ? (void *)(h->ctl + CTL_CTR) : (void *)&h->ctr;
    CK(cudaMemcpyAsync(dst, h->b_ctr.p, sizeof(DevCounters), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->ctl) memcpy(&h->ctr, dst, sizeof(DevCounters));
    return DBG_OK;
}

static int check_fault(dbg_t *h, const char *where) {
    TRY(fetch_counters(h));
    if (h->ctr.fault && !getenv("DBG_B200_IGNORE_FAULT"))
        return fail(h, DBG_ERR_STATE, "internal consistency fault %u (info %u) in %s", h->ctr.fault,
                    h->ctr.fault_info, where);
    return DBG_OK;
}

template <int W>
static int collect_sorted(dbg_t *h, u32 want, uint64_t expect, DevBuf &out, u64 *cursor) {
    // slots whose flags hit `want`, sorted by first-occurrence tag
    TRY(reserve(h, out, (expect + 1) * 4));
    if (expect == 0) return DBG_OK;
    TRY(reserve(h, h->b_tag, expect * 8));
    TRY(reserve(h, h->b_slot, expect * 4));
    CK(cudaMemsetAsync(cursor, 0, 8, h->stream));
    LAUNCH(k_collect<W>, blocks_for(h->n_slots), DBG_BLOCK, h->stream, ptr<const Slot<W>>(h->b_table), h->n_slots,
           ptr<const u8>(h->b_flags), want, ptr<u64>(h->b_tag), ptr<u32>(h->b_slot), cursor);
    CKK();
    // a start's first occurrence is at position 0 of its read: only the read-index half of the tag orders starts
    // (not under min_count: a node whose incoming window was filtered away starts a contig at any position)
    return sort_by_tag(h, expect, ptr<u32>(out), want == F_START && h->min_count <= 1 ? 32 : 0);
}

// side set (dbg_kernels.cuh) of the slots whose flags hit `want`: entries in b_order, bitmap in b_bitmap
template <int W>
static int side_set(dbg_t *h, u32 want, uint64_t n_members, uint64_t *cap_out, int *bits_out) {
    const uint64_t cap = n_members * 2 + 32;
    int bits = 10;                                              // bitmap of >= 32 bits per member, <= 2^28 bits
    while (bits < 28 && (1ull << bits) < n_members * 32) bits++;
    TRY(reserve(h, h->b_order, cap * sizeof(SideEntry<W>)));
    TRY(reserve(h, h->b_bitmap, (1ull << bits) / 8));
    CK(cudaMemsetAsync(h->b_order.p, 0xFF, cap * sizeof(SideEntry<W>), h->stream));
    CK(cudaMemsetAsync(h->b_bitmap.p, 0, (1ull << bits) / 8, h->stream));
    LAUNCH(k_side_fill<W>, blocks_for(h->n_slots), DBG_BLOCK, h->stream, ptr<const Slot<W>>(h->b_table),
           ptr<const u8>(h->b_flags), h->n_slots, want, ptr<SideEntry<W>>(h->b_order), cap, ptr<u32>(h->b_bitmap), bits);
    CKK();
    *cap_out = cap;
    *bits_out = bits;
    return DBG_OK;
}

template <int W>
static int tip_prune(dbg_t *h) {
    // debruijn.py:241-254.  Device: every candidate path of every branch node (k_tip_paths).
    // Host: walk the branch nodes in reference order and drop candidates that run into a node an
    // earlier branch already pulled (debruijn.py:184) -- the one order-dependent step.
    const uint64_t nb = h->branch.size();
    h->pulled.clear();
    if (nb == 0) return DBG_OK;
    TRY(reserve(h, h->b_cnt32, nb * 4));
    TRY(reserve(h, h->b_off64, (nb + 1) * 8));
    if (h->generic)
        LAUNCH(k_g_tip_paths, blocks_for(nb), DBG_BLOCK, h->stream, ptr<const u8>(h->b_flags), ptr<const u32>(h->b_enk),
               ptr<const u32>(h->b_estart), ptr<const GEdge>(h->b_earr), ptr<const u32>(h->b_branch), nb, 0,
               ptr<u32>(h->b_cnt32), (const u64 *)nullptr, (u32 *)nullptr);
    else
        LAUNCH(k_tip_paths<W>, blocks_for(nb), DBG_BLOCK, h->stream, ptr<const Slot<W>>(h->b_table), geom_of(h), h->k,
               ptr<const u8>(h->b_flags), ptr<const u16>(h->b_kept), ptr<const u32>(h->b_branch), nb, 0,
               ptr<u32>(h->b_cnt32), (const u64 *)nullptr, (u32 *)nullptr, ptr<DevCounters>(h->b_ctr));
    CKK();
    std::vector<u32> counts(nb);
    CK(cudaMemcpyAsync(counts.data(), h->b_cnt32.p, nb * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    TRY(check_fault(h, "k_tip_paths"));
    std::vector<u64> offs(nb + 1, 0);
    for (uint64_t i = 0; i < nb; i++) offs[i + 1] = offs[i] + counts[i];
    const uint64_t total = offs[nb];
    if (total == 0) return DBG_OK;
    TRY(reserve(h, h->b_recs, total * 6 * 4));
    CK(cudaMemcpyAsync(h->b_off64.p, offs.data(), (nb + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    if (h->generic)
        LAUNCH(k_g_tip_paths, blocks_for(nb), DBG_BLOCK, h->stream, ptr<const u8>(h->b_flags), ptr<const u32>(h->b_enk),
               ptr<const u32>(h->b_estart), ptr<const GEdge>(h->b_earr), ptr<const u32>(h->b_branch), nb, 1,
               ptr<u32>(h->b_cnt32), ptr<const u64>(h->b_off64), ptr<u32>(h->b_recs));
    else
        LAUNCH(k_tip_paths<W>, blocks_for(nb), DBG_BLOCK, h->stream, ptr<const Slot<W>>(h->b_table), geom_of(h), h->k,
               ptr<const u8>(h->b_flags), ptr<const u16>(h->b_kept), ptr<const u32>(h->b_branch), nb, 1,
               ptr<u32>(h->b_cnt32), ptr<const u64>(h->b_off64), ptr<u32>(h->b_recs), ptr<DevCounters>(h->b_ctr));
    CKK();
    std::vector<u32> recs(total * 6);
    CK(cudaMemcpyAsync(recs.data(), h->b_recs.p, total * 6 * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    std::unordered_set<u32> pulled_set;
    std::vector<u32> fresh;
    for (uint64_t b = 0; b < nb; b++) {
        fresh.clear();                                  // already_pull_out is fixed while b's DFS runs
        for (uint64_t r = offs[b]; r < offs[b + 1]; r++) {
            const u32 *rec = &recs[r * 6];
            const int len = (int)(rec[0] & 0xFF);
            const u32 bm = rec[0] >> 8;
            bool blocked = false;                       // runs into a node an EARLIER branch pulled (:184)
            for (int j = 1; j < len && !blocked; j++) blocked = pulled_set.count(rec[1 + j]) != 0;
            if (blocked) continue;
            for (int j = 0; j < len; j++) {
                const u32 v = rec[1 + j];
                if ((bm >> j) & 1) continue;                           // `item not in branch_kmer`
                if (std::find(fresh.begin(), fresh.end(), v) == fresh.end()) fresh.push_back(v);
            }
        }
        for (u32 v : fresh) { pulled_set.insert(v); h->pulled.push_back(v); }
    }
    const uint64_t np = h->pulled.size();
    if (np) {
        TRY(reserve(h, h->b_pulled, np * 4));
        CK(cudaMemcpyAsync(h->b_pulled.p, h->pulled.data(), np * 4, cudaMemcpyHostToDevice, h->stream));
        LAUNCH(k_mark_pulled, blocks_for(np), DBG_BLOCK, h->stream, ptr<const u32>(h->b_pulled), np,
               ptr<u8>(h->b_flags));
        CKK();
        CK(cudaStreamSynchronize(h->stream));          // h->pulled is host memory of this handle; keep it simple
    }
    return DBG_OK;
}

// ---- chunk table build (dbg_chunks.cuh): masks -> chunk table + boundary list -> node table ----
struct ChunkGeom { int w, logw, m; };
static ChunkGeom chunk_geom(int k) {
    int t = k - 4;                                  // leave the minimizer at least ~5 bases where k allows it
    if (t < 1) t = 1;
    if (t > CH_WMAX) t = CH_WMAX;
    ChunkGeom g;
    g.logw = 0;
    while ((2 << g.logw) <= t) g.logw++;
    g.w = 1 << g.logw;                              // nodes per chunk at most (power of two <= 16, <= k)
    g.m = k - g.w + 1;                              // every candidate m-mer lies inside the k-mer
    if (g.m > CH_MMAX) g.m = CH_MMAX;
    return g;
}

static uint32_t pass_count(dbg_t *h, uint64_t table_bytes) {
    // Partition passes: each pass owns a hash range, so the table region it touches is bounded.  A pass
    // re-reads every distinct chunk (32 + 16 bytes), so few large regions beat many L2-sized ones:
    // measured on B200, 312 MB table: 1 / 2 / 3 / 5 passes = 0.61 / 0.50 / 0.53 / 0.61 ms;
    // 2.1 GB table: 5 / 6 / 9 / 10 passes = 3.41 / 3.05 / 2.70 / 2.81 ms.
    uint64_t region = 160ull << 20;
    if (table_bytes / 9 > region) region = table_bytes / 9;
    if (region > 512ull << 20) region = 512ull << 20;
    if (const char *e = getenv("DBG_B200_L2_REGION_BYTES")) {         // tuning / tests: force more passes
        const unsigned long long v = strtoull(e, nullptr, 10);
        if (v >= 64) region = v;
    }
    uint64_t n = cdiv(table_bytes, region);
    if (n < 1) n = 1;
    if (n > 255) n = 255;
    (void)h;
    return (uint32_t)n;
}

// ---- node table laid out by minimizer: geometry, record order ----
static bool want_regions(const dbg_t *h) {
    // the region of a node is a hash of its minimizer (+ batch id): short minimizers take too few values to spread
    const ChunkGeom g = chunk_geom(h->k);
    bool on = 2 * g.m + h->batch_bits >= 24 && !h->generic && !h->avoid_regions;
    if (const char *e = getenv("DBG_B200_REGIONS")) on = atoi(e) != 0 && !h->generic;    // tests / tuning: 0 = plain hash layout, 1 = force
    return on;
}
#define VB_BITS_MAX 21
static int vb_bits_for(uint64_t n_record_slots) {             // virtual buckets: >= record slots / 16, 2^10 .. 2^21
    int b = 10;
    while (b < VB_BITS_MAX && (1ull << b) < n_record_slots / 16) b++;
    if (const char *e = getenv("DBG_B200_VB_BITS")) {              // tests: a few buckets, nearly every record spills
        const int v = atoi(e);
        if (v >= 1 && v <= VB_BITS_MAX) b = v;
    }
    return b;
}
// node-table size for `want` slots under the current layout: whole regions, at most 2^vb_bits / 4 of them
static uint64_t fit_slots(dbg_t *h, uint64_t want) {
    h->region_slots = 0;
    if (!h->region_mode) return want;
    uint64_t R = 2048;
    if (const char *e = getenv("DBG_B200_REGION_SLOTS")) {        // tests: many small regions
        const unsigned long long v = strtoull(e, nullptr, 10);
        if (v >= 2) R = v;
    }
    const uint64_t max_regions = (1ull << (h->vb_bits && !getenv("DBG_B200_VB_BITS") ? h->vb_bits : VB_BITS_MAX)) / 4;
    if (cdiv(want, R) > max_regions) R = cdiv(want, max_regions);
    if (R >= 0xFFFFFFFFull) return want;
    h->region_slots = (uint32_t)R;
    return cdiv(want, R) * R;
}
static TGeom geom_of(const dbg_t *h) {
    TGeom g = plain_geom(h->n_slots);
    if (h->region_slots && h->n_slots >= h->region_slots && h->n_slots % h->region_slots == 0) {
        const ChunkGeom cg = chunk_geom(h->k);
        g.n_regions = (u32)(h->n_slots / h->region_slots);
        g.region_slots = h->region_slots;
        g.k = h->k; g.w = cg.w; g.m = cg.m;
    }
    return g;
}

// per record slot: region hash + histogram over virtual buckets (+ the node-key sketch that sizes the node table)
template <int W>
static int launch_rec_vhash(dbg_t *h, bool sketch) {
    typedef ChunkSlot<W> CS;
    const uint64_t n_rec = h->n_cslots + h->n_clist;
    if (n_rec >= 0xFFFFFFF0ull) return fail(h, DBG_ERR_OVERFLOW, "more than 2^32 chunk record slots");
    const ChunkGeom cg = chunk_geom(h->k);
    h->vb_bits = vb_bits_for(n_rec);
    TRY(reserve(h, h->b_recvh, (n_rec + 32) * 4));
    TRY(reserve(h, h->b_vhist, ((1ull << h->vb_bits) + 2) * 4));
    CK(cudaMemsetAsync(h->b_vhist.p, 0, ((1ull << h->vb_bits) + 2) * 4, h->stream));
    uint64_t grid = blocks_for(n_rec);
    if (grid > 148 * 16) grid = 148 * 16;
    if (sketch) {
        CK(cudaMemsetAsync(h->b_hll.p, 0, HLL_M * 4, h->stream));
        LAUNCH((k_rec_vhash<W, true>), grid, DBG_BLOCK, h->stream, ptr<const CS>(h->b_ctable), h->n_cslots,
               ptr<const CS>(h->b_clist), h->n_clist, h->k, cg.w, cg.m, h->batch_bits, h->vb_bits, ptr<u32>(h->b_recvh),
               ptr<u32>(h->b_vhist), ptr<u32>(h->b_hll));
    } else {
        LAUNCH((k_rec_vhash<W, false>), grid, DBG_BLOCK, h->stream, ptr<const CS>(h->b_ctable), h->n_cslots,
               ptr<const CS>(h->b_clist), h->n_clist, h->k, cg.w, cg.m, h->batch_bits, h->vb_bits, ptr<u32>(h->b_recvh),
               ptr<u32>(h->b_vhist), (u32 *)nullptr);
    }
    CKK();
    h->have_vhash = true;
    h->have_order = false;
    return DBG_OK;
}

// the records in region order: exclusive scan of the histogram, then one scatter
static int launch_rec_order(dbg_t *h) {
    const uint64_t n_rec = h->n_cslots + h->n_clist;
    const uint64_t nv = (1ull << h->vb_bits) + 1;                  // + 1: the total lands behind the last bucket
    TRY(reserve(h, h->b_voffs, (nv + 1) * 4));
    TRY(reserve(h, h->b_recorder, (n_rec + 32) * 8));
#ifdef DBG_EMU
    {
        const u32 *hist = ptr<const u32>(h->b_vhist);
        u32 *offs = ptr<u32>(h->b_voffs);
        u32 acc = 0;
        for (uint64_t i = 0; i < nv; i++) { offs[i] = acc; acc += hist[i]; }
    }
#else
    {
        size_t tmp = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, ptr<const u32>(h->b_vhist), ptr<u32>(h->b_voffs), (int)nv, h->stream));
        TRY(reserve(h, h->b_sorttmp, tmp));
        h->st.n_launches++;
        CK(cub::DeviceScan::ExclusiveSum(h->b_sorttmp.p, tmp, ptr<const u32>(h->b_vhist), ptr<u32>(h->b_voffs), (int)nv, h->stream));
    }
#endif
    uint64_t grid = blocks_for(n_rec);
    if (grid > 148 * 16) grid = 148 * 16;
    LAUNCH(k_rec_order, grid, DBG_BLOCK, h->stream, ptr<const u32>(h->b_recvh), n_rec, h->vb_bits, ptr<u32>(h->b_vhist),
           ptr<const u32>(h->b_voffs), ptr<u64>(h->b_recorder));
    CKK();
    h->have_order = true;
    return DBG_OK;
}

// sketch of the node keys of the distinct chunks (sizes the node table); leaves what the node-table pass(es) want
// with it: hash bytes per (record, node) for the plain layout, region hashes + histogram for the region layout
template <int W>
static int launch_node_sketch(dbg_t *h, uint64_t ex_grid) {
    typedef Slot<W> S;
    typedef ChunkSlot<W> CS;
    if (h->region_mode) return launch_rec_vhash<W>(h, true);
    TRY(reserve(h, h->b_h8, (h->n_cslots + h->n_clist + 32) * 16));      // one hash byte per (record, node)
    CK(cudaMemsetAsync(h->b_hll.p, 0, HLL_M * 4, h->stream));
    LAUNCH((k_expand<W, 0>), ex_grid, DBG_BLOCK, h->stream, ptr<const CS>(h->b_ctable), h->n_cslots,
           ptr<const CS>(h->b_clist), h->n_clist, h->k, ptr<uint4>(h->b_h8), (S *)nullptr, 0u, 0u, 0u,
           ptr<u32>(h->b_hll), h->batch_bits, ptr<DevCounters>(h->b_ctr), (const u64 *)nullptr, (const u32 *)nullptr,
           plain_geom(0));
    CKK();
    h->have_hbytes = true;
    return DBG_OK;
}

// phase 1 over the units [u0, u1) of the read stream
template <int W>
static int launch_chunk_insert(dbg_t *h, const u64 *d_off, uint64_t u0, uint64_t u1) {
    typedef ChunkSlot<W> CS;
    if (u1 <= u0) return DBG_OK;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    const uint64_t n = u1 - u0;
    uint64_t grid = (uint64_t)sms * DBG_CHUNK_MIN_BLOCKS;       // persistent: measured 0.39 ms vs 0.43 (x16) / 0.47 (x32)
    if (const char *e = getenv("DBG_B200_INSERT_GRID")) {            // tuning / tests
        const unsigned long long v = strtoull(e, nullptr, 10);
        if (v >= 1) grid = v;
    }
    if (grid > cdiv(n, DBG_WARPS)) grid = cdiv(n, DBG_WARPS);
    const uint64_t units_per_block = cdiv(n, grid);
    grid = cdiv(n, units_per_block);
    LAUNCH(k_chunk_insert<W>, grid, DBG_BLOCK, h->stream, ptr<const u64>(h->b_packed), h->n_words + 4, d_off,
           h->n_reads, ptr<const u32>(h->b_stmask), ptr<const u32>(h->b_vmask), ptr<const u32>(h->b_spanread),
           h->n_words, u0, u1, units_per_block, h->k, ptr<CS>(h->b_ctable), (u32)h->n_cslots, ptr<CS>(h->b_clist),
           2 * h->n_reads + 2, h->read_idx_base, (const u32 *)h->d_read_batch, ptr<ChunkCounters>(h->b_cc));
    CKK();
    return DBG_OK;
}

// streamed host build: make sure the chunk table can take what the sketch has seen so far
template <int W>
static int chunk_table_fit(dbg_t *h, uint64_t est_chunks_so_far) {
    typedef ChunkSlot<W> CS;
    const uint64_t need = (uint64_t)((double)est_chunks_so_far * 1.7) + 4096;
    if (h->n_cslots == 0) {                                          // first slice: leave room for what later slices add
        uint64_t n = need + need / 3;
        if (const char *e = getenv("DBG_B200_CHUNK_SLOTS")) {        // tests: force the growth / overflow paths
            const unsigned long long v = strtoull(e, nullptr, 10);
            if (v >= 1) n = v;
        }
        if (n >= 0xFFFFFFF0ull) return fail(h, DBG_ERR_OVERFLOW, "chunk table would exceed 2^32 slots");
        TRY(reserve(h, h->b_ctable, n * sizeof(CS)));
        LAUNCH(k_chunk_table_init<W>, blocks_for(n), DBG_BLOCK, h->stream, ptr<CS>(h->b_ctable), n);
        CKK();
        h->n_cslots = n;
    } else if (need > h->n_cslots) {                                 // grow: move every entry into a larger table
        const uint64_t n = need + need / 2;
        if (n >= 0xFFFFFFF0ull) return fail(h, DBG_ERR_OVERFLOW, "chunk table would exceed 2^32 slots");
        TRY(reserve(h, h->b_ctable2, n * sizeof(CS)));
        LAUNCH(k_chunk_table_init<W>, blocks_for(n), DBG_BLOCK, h->stream, ptr<CS>(h->b_ctable2), n);
        CKK();
        LAUNCH(k_chunk_rehash<W>, blocks_for(h->n_cslots), DBG_BLOCK, h->stream, ptr<const CS>(h->b_ctable), h->n_cslots,
               ptr<CS>(h->b_ctable2), (u32)n, ptr<ChunkCounters>(h->b_cc));
        CKK();
        std::swap(h->b_ctable, h->b_ctable2);
        h->n_cslots = n;
    }
    return DBG_OK;
}

// node table from the distinct chunks (phase 2b): init + the k_expand passes at h->n_slots; no host sync.
template <int W>
static int expand_node_table(dbg_t *h) {
    typedef Slot<W> S;
    typedef ChunkSlot<W> CS;
    dbg_stats_t &st = h->st;
    DevCounters *d_ctr = ptr<DevCounters>(h->b_ctr);
    const uint64_t n_slots = h->n_slots;
    if (n_slots >= 0xFFFFFFF0ull) return fail(h, DBG_ERR_OVERFLOW, "node table would exceed 2^32 slots");
    TRY(reserve(h, h->b_table, n_slots * sizeof(S)));
    const int a = ev_mark(h);
    if (!h->preinit) {
        LAUNCH(k_table_init<W>, blocks_for(n_slots), DBG_BLOCK, h->stream, ptr<S>(h->b_table), n_slots);
        CKK();
    } else CK(cudaStreamWaitEvent(h->stream, h->ev_init[2], 0));        // cleared on the side stream
    const bool self_hash = !h->have_hbytes;                          // no sketch pass ran: no hash bytes
    h->preinit = false;                                              // a rebuild (overflow) initialises its own table
    const int b = ev_mark(h);
    if (st.n_windows && geom_of(h).n_regions > 1) {
        // table laid out by minimizer: ONE pass over the records in region order, the table is built front to back
        if (!h->have_vhash) TRY(launch_rec_vhash<W>(h, false));
        if (!h->have_order) TRY(launch_rec_order(h));
        const int b2 = ev_mark(h);
        tm_add(h, &st.ms_region_order, b, b2);
        st.n_insert_passes = 1;
        LAUNCH((k_expand<W, 3>), h->ex_grid, DBG_BLOCK, h->stream, ptr<const CS>(h->b_ctable), h->n_cslots,
               ptr<const CS>(h->b_clist), h->n_clist, h->k, (uint4 *)nullptr, ptr<S>(h->b_table), 0u, 0u, 254u,
               (u32 *)nullptr, h->batch_bits, d_ctr, ptr<const u64>(h->b_recorder),
               ptr<const u32>(h->b_voffs) + (1ull << h->vb_bits), geom_of(h));
        CKK();
    } else if (st.n_windows) {
        uint32_t n_parts = pass_count(h, n_slots * sizeof(S));
        if (n_parts > 255) n_parts = 255;
        st.n_insert_passes = n_parts;
        for (uint32_t part = 0; part < n_parts; part++) {
            // hash bytes 0..254 (255 = no node) split evenly over the passes
            const uint32_t lo8 = part * 255u / n_parts;
            const uint32_t hi8 = (part + 1) * 255u / n_parts - 1;
            if (self_hash)
                LAUNCH((k_expand<W, 2>), h->ex_grid, DBG_BLOCK, h->stream, ptr<const CS>(h->b_ctable), h->n_cslots,
                       ptr<const CS>(h->b_clist), h->n_clist, h->k, (uint4 *)nullptr, ptr<S>(h->b_table), (u32)n_slots,
                       lo8, n_parts == 1 ? 254u : hi8 - lo8, (u32 *)nullptr, h->batch_bits, d_ctr, (const u64 *)nullptr,
                       (const u32 *)nullptr, plain_geom(n_slots));
            else
                LAUNCH((k_expand<W, 1>), h->ex_grid, DBG_BLOCK, h->stream, ptr<const CS>(h->b_ctable), h->n_cslots,
                       ptr<const CS>(h->b_clist), h->n_clist, h->k, ptr<uint4>(h->b_h8), ptr<S>(h->b_table), (u32)n_slots,
                       lo8, hi8 - lo8, (u32 *)nullptr, h->batch_bits, d_ctr, (const u64 *)nullptr, (const u32 *)nullptr,
                       plain_geom(n_slots));
            CKK();
        }
    }
    const int c = ev_mark(h);
    tm_add(h, &st.ms_table_init, a, b);
    tm_add(h, &st.ms_insert, b, c);
    tm_add(h, &st.ms_expand, b, c);
    st.table_slots = n_slots;
    st.table_bytes = n_slots * sizeof(S);
    st.n_regions = geom_of(h).n_regions;
    h->expand_evs = h->n_ev - a;
    return DBG_OK;
}

// look at the overflow flag of the last expand_node_table (counters fetched = stream idle); rebuild larger
// until it fits.  `fetched` = h->ctr is already current.
template <int W>
static int settle_node_table(dbg_t *h, bool fetched) {
    DevCounters *d_ctr = ptr<DevCounters>(h->b_ctr);
    const uint64_t n_slots0 = h->n_slots;
    for (int attempt = 0;; attempt++) {
        if (!fetched) TRY(fetch_counters(h));
        fetched = false;
        tm_flush(h);
        if (!h->ctr.overflow) break;
        if (attempt >= 8) return fail(h, DBG_ERR_NOMEM, "node table overflowed repeatedly");
        if (attempt == 1 && geom_of(h).n_regions > 1) {
            // Twice the slots did not help: the nodes are not short of slots, they crowd a few REGIONS -- reads that pile
            // up on a few loci (amplicons, a high-copy repeat) put thousands of k-mer variants behind the same
            // minimizers, and a larger table only adds regions nobody maps to.  Fall back to the plain hash layout
            // (partition passes) for this build and for the following builds on the handle.
            h->region_mode = false;
            h->avoid_regions = true;
            h->region_slots = 0;
            h->have_hbytes = false;
            h->n_slots = n_slots0;
        }
        h->n_slots *= 2;                                               // estimate was low: rebuild larger
        CK(cudaMemsetAsync(&d_ctr->overflow, 0, 4, h->stream));
        if (h->n_ev >= h->expand_evs) h->n_ev -= h->expand_evs;                                // (the table may have been expanded by an earlier call)
        TRY(expand_node_table<W>(h));
    }
    h->expand_pending = false;
    return DBG_OK;
}

template <int W> static int node_table_from_chunks(dbg_t *h, uint64_t est);

template <int W>
static int build_chunks(dbg_t *h, const u64 *d_off, int e0, int e1) {
    typedef Slot<W> S;
    typedef ChunkSlot<W> CS;
    const int k = h->k;
    dbg_stats_t &st = h->st;
    DevCounters *d_ctr = ptr<DevCounters>(h->b_ctr);
    const ChunkGeom geom = chunk_geom(k);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);

    // ---- phase 0 (span masks + sketch of the chunk starts) ran with the packing, slice by slice ----
    const uint64_t n_spans = h->n_words;
    const u32 sample_mask = h->sample_mask;
    ChunkCounters *d_cc = ptr<ChunkCounters>(h->b_cc);
    ChunkCounters cc;
    memset(&cc, 0, sizeof cc);
    uint64_t est_chunks = 0;
    if (h->hinted) {
        // Tables were sized from the previous build on this handle and initialised on the side stream: no host
        // round trip here.  Sketch registers and chunk counters land in pinned memory and are looked at, with
        // the graph counters, by graph_finalize -- which redoes the build the cold way if they say "did not fit".
        const int e2h = ev_mark(h);
        if (!h->chunks_streamed) {
            CK(cudaStreamWaitEvent(h->stream, h->ev_init[1], 0));
            TRY(launch_chunk_insert<W>(h, d_off, 0, cdiv(h->n_bases, DBG_UNIT)));
        }
        h->chunks_streamed = false;
        const int ch = ev_mark(h);
        CK(cudaMemcpyAsync(h->ctl + CTL_CC, d_cc, sizeof cc, cudaMemcpyDeviceToHost, h->stream));
        h->n_clist = 2 * h->n_reads + 2;
        uint64_t g = cdiv(cdiv(h->n_cslots + h->n_clist, 32), DBG_WARPS);
        uint64_t per_sm = DBG_EXPAND_MIN_BLOCKS;
        if (const char *e = getenv("DBG_B200_EXPAND_GRID")) {
            const unsigned long long v = strtoull(e, nullptr, 10);
            if (v >= 1) per_sm = v;
        }
        if (g > (uint64_t)sms * per_sm) g = (uint64_t)sms * per_sm;
        h->ex_grid = g < 1 ? 1 : g;
        st.n_windows = 1;                                            // the real count arrives with graph_finalize's counters
        st.n_chunk_list = h->n_clist;
        st.chunk_slots = h->n_cslots;
        tm_set(h, &st.ms_pack, e0, e1);
        tm_set(h, &st.ms_estimate, e1, e2h);
        tm_set(h, &st.ms_chunk_insert, e2h, ch);
        tm_set(h, &st.ms_insert, e2h, ch);
        if (h->build_flags & DBG_CHUNKS_ONLY) {                      // multi-GPU: the chunks leave for their owners
            TRY(fetch_counters(h));                                   // ONE round trip: counters of packing + phase 1
            tm_flush(h);
            memcpy(&cc, h->ctl + CTL_CC, sizeof cc);
            if (h->ctr.bad_alphabet) return fail(h, DBG_ERR_ALPHABET, "a read holds a byte other than A/C/G/T");
            if (h->ctr.too_long) return fail(h, DBG_ERR_OVERFLOW, "a read is longer than 2^32-1 bases");
            if (cc.chunk_overflow || cc.bad_chunk || cc.list_overflow) return DBG_RETRY_COLD;
            st.n_windows = h->ctr.n_windows;
            st.n_chunk_instances = cc.n_instances;
            h->hint.chunks = cc.n_first;
            h->hint.n_bases = h->n_bases;
            h->hinted = false;
            h->chunks_valid = true;
            const int e_endc = ev_mark(h);
            tm_set(h, &st.ms_build_total, e0, e_endc);
            return DBG_OK;
        }
        if (!h->region_mode && pass_count(h, h->n_slots * sizeof(S)) > 2) {
            // many node-table passes: one sweep that leaves a hash byte per (record, node) is cheaper than hashing every
            // node again in every pass (measured at C3, 10 passes: 2.8 ms with the bytes, 4.3 ms without)
            TRY(launch_node_sketch<W>(h, h->ex_grid));
        }
        TRY(expand_node_table<W>(h));
        h->expand_pending = true;
        const int e_end = ev_mark(h);
        tm_set(h, &st.ms_build_total, e0, e_end);
        h->table_ready = true;
        h->chunks_valid = true;
        return DBG_OK;
    }
    if (h->n_bases) {
        std::vector<u32> regs(HLL_M);
        void *r_dst = h->ctl ? (void *)(h->ctl + CTL_SKETCH) : (void *)regs.data();
        void *c_dst = h->ctl ? (void *)(h->ctl + CTL_CC) : (void *)&cc;
        CK(cudaMemcpyAsync(r_dst, h->b_hll.p, HLL_M * 4, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(c_dst, d_cc, sizeof cc, cudaMemcpyDeviceToHost, h->stream));
        TRY(fetch_counters(h));
        if (h->ctl) { memcpy(regs.data(), r_dst, HLL_M * 4); memcpy(&cc, c_dst, sizeof cc); }
        est_chunks = (uint64_t)hll_estimate(regs) * ((sample_mask >> 6) + 1);
    } else {
        TRY(fetch_counters(h));
    }
    if (h->ctr.bad_alphabet) return fail(h, DBG_ERR_ALPHABET, "a read holds a byte other than A/C/G/T");
    if (h->ctr.too_long) return fail(h, DBG_ERR_OVERFLOW, "a read is longer than 2^32-1 bases");
    st.n_windows = h->ctr.n_windows;
    const int e2 = ev_mark(h);

    // ---- phase 1: chunk instances -> chunk table (interior) / list (first and last chunk of a read) ----
    // list of boundary chunks: slot 2 r = first chunk of read r, 2 r + 1 = its last chunk (EMPTY where a
    // read has none: reads of length <= k, or one chunk that is both)
    const uint64_t list_cap = 2 * h->n_reads + 2;                     // reserved and cleared by build_impl
    // the sketch hashes the 32 bases at every chunk start that is not a read start; the last chunk of a
    // read starts at a content-defined cut like the interior chunk other reads have there, so the
    // sketch counts (about) the distinct INTERIOR chunks; it misses variants that differ only past
    // base 32, which the slack and the overflow retry absorb
    const uint64_t est_interior = est_chunks;
    uint64_t n_cslots = (uint64_t)((double)est_interior * 1.7) + 4096;
    if (const char *e = getenv("DBG_B200_CHUNK_SLOTS")) {             // tests: force the overflow / retry path
        const unsigned long long v = strtoull(e, nullptr, 10);
        if (v >= 1) n_cslots = v;
    }
    if (h->chunks_streamed) n_cslots = h->n_cslots;                   // phase 1 ran while the reads were arriving
    float ms_init = 0, ms_insert = 0;
    const uint64_t n_list = list_cap;
    uint64_t est = 0;
    std::vector<u32> sketch(HLL_M, 0);
    uint64_t ex_grid = 1;
    for (int attempt = 0;; attempt++) {
        if (n_cslots >= 0xFFFFFFF0ull) return fail(h, DBG_ERR_OVERFLOW, "chunk table would exceed 2^32 slots");
        TRY(reserve(h, h->b_ctable, n_cslots * sizeof(CS)));
        TRY(reserve(h, h->b_h8, (n_cslots + n_list + 32) * 16));      // one hash byte per (record, node)
        const int a = ev_mark(h);
        if (!h->chunks_streamed) {
            LAUNCH(k_chunk_table_init<W>, blocks_for(n_cslots), DBG_BLOCK, h->stream, ptr<CS>(h->b_ctable), n_cslots);
            CKK();
        }
        const int b = ev_mark(h);
        if (st.n_windows && !h->chunks_streamed) {
            h->n_cslots = n_cslots;
            TRY(launch_chunk_insert<W>(h, d_off, 0, cdiv(h->n_bases, DBG_UNIT)));
        }
        h->chunks_streamed = false;                                  // a retry redoes phase 1 in one go
        const int c = ev_mark(h);
        void *c_dst = h->ctl ? (void *)(h->ctl + CTL_CC) : (void *)&cc;
        void *s_dst = h->ctl ? (void *)(h->ctl + CTL_SKETCH) : (void *)sketch.data();
        CK(cudaMemcpyAsync(c_dst, d_cc, sizeof cc, cudaMemcpyDeviceToHost, h->stream));
        // ---- phase 2a right behind it (no host round trip in between): sketch of the node keys of the
        // distinct chunks -> node table size; wasted only when the chunk table turns out to have overflowed
        ex_grid = cdiv(cdiv(n_cslots + n_list, 32), DBG_WARPS);
        uint64_t ex_per_sm = DBG_EXPAND_MIN_BLOCKS;                   // persistent: one resident wave (0.48 vs 0.51 ms at x8)
        if (const char *e = getenv("DBG_B200_EXPAND_GRID")) {         // tuning: blocks per SM
            const unsigned long long v = strtoull(e, nullptr, 10);
            if (v >= 1) ex_per_sm = v;
        }
        if (ex_grid > (uint64_t)sms * ex_per_sm) ex_grid = (uint64_t)sms * ex_per_sm;
        if (ex_grid < 1) ex_grid = 1;
        if (st.n_windows && !(h->build_flags & DBG_CHUNKS_ONLY)) {      // (chunks that get exchanged are expanded by their owners)
            h->n_cslots = n_cslots;
            h->n_clist = n_list;
            TRY(launch_node_sketch<W>(h, ex_grid));
            CK(cudaMemcpyAsync(s_dst, h->b_hll.p, HLL_M * 4, cudaMemcpyDeviceToHost, h->stream));
        }
        const int d = ev_mark(h);
        CK(cudaStreamSynchronize(h->stream));
        if (h->ctl) { memcpy(&cc, c_dst, sizeof cc); memcpy(sketch.data(), s_dst, HLL_M * 4); }
        ms_init += ev_ms(h, a, b);
        ms_insert += ev_ms(h, b, d);
        st.ms_chunk_insert += ev_ms(h, b, c);
        st.ms_expand_sketch += ev_ms(h, c, d);
        if (cc.bad_chunk) return fail(h, DBG_ERR_STATE, "internal consistency fault: a chunk without an end");
        if (cc.list_overflow) return fail(h, DBG_ERR_STATE, "internal consistency fault: boundary list overflow");
        if (!cc.chunk_overflow) break;
        if (attempt >= 8) return fail(h, DBG_ERR_NOMEM, "chunk table overflowed repeatedly");
        n_cslots = n_cslots * 2 + 4096;                                // sketch was low: rebuild larger
        CK(cudaMemsetAsync(&d_cc->n_instances, 0, sizeof(ChunkCounters) - offsetof(ChunkCounters, n_instances),
                           h->stream));
        h->n_ev -= 4;
    }
    if (st.n_windows) est = (uint64_t)hll_estimate(sketch);
    st.hll_estimate = est;
    h->last_est_chunks = est_interior;
    h->n_cslots = n_cslots;
    h->n_clist = n_list;
    h->ex_grid = ex_grid;
    st.n_chunk_instances = cc.n_instances;
    st.n_chunk_list = n_list;
    st.chunk_slots = n_cslots;
    st.ms_table_init = ms_init;
    st.ms_insert = ms_insert;
    tm_set(h, &st.ms_pack, e0, e1);
    tm_set(h, &st.ms_estimate, e1, e2);
    if (h->build_flags & DBG_CHUNKS_ONLY) {                          // multi-GPU: the chunks are exchanged before any node table
        h->hint.valid = true;                                        // sizes for the next build of this kind on the handle
        h->hint.n_bases = h->n_bases;
        h->hint.chunks = est_interior;
        h->hint.n_batches = 0;
        const int e_end = ev_mark(h);
        tm_set(h, &st.ms_build_total, e0, e_end);
        h->chunks_valid = true;
        return DBG_OK;
    }
    TRY(node_table_from_chunks<W>(h, est));
    const int e_end = ev_mark(h);
    tm_set(h, &st.ms_build_total, e0, e_end);
    h->table_ready = true;
    h->chunks_valid = true;
    return DBG_OK;
}

// ---- phase 2b: node table (kernel 3): every node of every distinct chunk, x multiplicity ----
template <int W>
static int node_table_from_chunks(dbg_t *h, uint64_t est) {
    dbg_stats_t &st = h->st;
    const uint64_t max_nodes = st.n_windows + h->n_reads;             // every window + last k-mer per read
    uint64_t want = (uint64_t)((double)est * 1.06) + 64;              // HLL std error at 2^12 registers: 1.6 %
    if (want > max_nodes && !h->chunks_merged) want = max_nodes;
    double load = 0.5;                                                // load factor: measured 0.5 / 0.6 / 0.7 = 0.51 / 0.53 / 0.60 ms of passes
    if (const char *e = getenv("DBG_B200_TABLE_LOAD")) {              // tuning: percent
        const double v = atof(e) / 100.0;
        if (v >= 0.05 && v <= 0.95) load = v;
    }
    h->n_slots = (uint64_t)((double)want / load) + 1024;
    if (const char *e = getenv("DBG_B200_NODE_SLOTS")) {              // tests: force the overflow / rebuild path
        const unsigned long long v = strtoull(e, nullptr, 10);
        if (v >= 1) h->n_slots = v;
    }
    h->n_slots = fit_slots(h, h->n_slots);
    TRY(expand_node_table<W>(h));
    if ((h->build_flags & DBG_TABLE_ONLY) || h->chunks_merged) TRY(settle_node_table<W>(h, false));   // nobody downstream looks at the flag
    else h->expand_pending = true;                                   // graph_finalize looks at it with its own counters
    return DBG_OK;
}

template <int W>
static int build_impl(dbg_t *h, const u8 *d_bases, const u64 *d_off) {
    typedef Slot<W> S;
    const int k = h->k;
    dbg_stats_t &st = h->st;
    DevCounters *d_ctr = ptr<DevCounters>(h->b_ctr);
    const int e0 = ev_mark(h);

    // ---- read packing (kernel 1) + span masks (phase 0 of the chunk build), slice by slice ----
    h->n_words = cdiv(h->n_bases, 32);
    TRY(reserve(h, h->b_packed, (h->n_words + 4) * 8));
    CK(cudaMemsetAsync(ptr<u64>(h->b_packed) + h->n_words, 0, 4 * 8, h->stream));
    h->n_tiles = cdiv(h->n_bases, DBG_TILE);
    TRY(reserve(h, h->b_tile, (h->n_tiles + 2) * 4));
    const uint64_t n_spans = h->n_words;
    {
        const ChunkGeom geom = chunk_geom(k);
        h->geom_logw = geom.logw;
        h->geom_m = geom.m;
        // the sketch looks at 1 chunk start in 4, chosen by two bits of the minimizer's hash (distinct sampling:
        // a function of the chunk's content); with short minimizers the hash has too few values for that
        h->sample_mask = geom.m >= 8 ? (3u << 6) : 0u;
        TRY(reserve(h, h->b_stmask, (n_spans + 2) * 4));
        TRY(reserve(h, h->b_vmask, (n_spans + 2) * 4));
        TRY(reserve(h, h->b_spanread, (n_spans + 2) * 4));
        TRY(reserve(h, h->b_hll, HLL_M * 4));
        TRY(reserve(h, h->b_cc, sizeof(ChunkCounters)));
        CK(cudaMemsetAsync(h->b_hll.p, 0, HLL_M * 4, h->stream));
        CK(cudaMemsetAsync(h->b_cc.p, 0, sizeof(ChunkCounters), h->stream));
        // list of boundary chunks: slot 2 r = first chunk of read r, 2 r + 1 = its last chunk (EMPTY where a
        // read has none: reads of length <= k, or one chunk that is both)
        const uint64_t list_cap = 2 * h->n_reads + 2;
        TRY(reserve(h, h->b_clist, list_cap * sizeof(ChunkSlot<W>)));
        bool list_cleared = false;
        h->n_cslots = 0;
        h->chunks_streamed = false;
        if (h->hinted) {
            typedef ChunkSlot<W> CS;
            const uint64_t n_cs = (uint64_t)((double)h->hint.chunks * 1.7 * 1.15) + 4096;
            double load = 0.5;
            if (const char *e = getenv("DBG_B200_TABLE_LOAD")) {
                const double v = atof(e) / 100.0;
                if (v >= 0.05 && v <= 0.95) load = v;
            }
            const bool chunks_only = (h->build_flags & DBG_CHUNKS_ONLY) != 0;
            h->vb_bits = 0;                                              // (not known yet: fit_slots assumes the largest)
            const uint64_t n_ns = chunks_only ? 16 : fit_slots(h, (uint64_t)((double)h->hint.nodes * 1.06 * 1.03 / load) + 1024);
            if (n_cs >= 0xFFFFFFF0ull || n_ns >= 0xFFFFFFF0ull) h->hinted = false;
            else {
                TRY(reserve(h, h->b_ctable, n_cs * sizeof(CS)));
                if (!chunks_only) TRY(reserve(h, h->b_table, n_ns * sizeof(S)));
                h->n_cslots = n_cs;
                if (!chunks_only) h->n_slots = n_ns;
                // both tables are cleared on the side stream while the reads are packed and masked
                CK(cudaEventRecord(h->ev_init[0], h->stream));
                CK(cudaStreamWaitEvent(h->init_stream, h->ev_init[0], 0));
                CK(cudaMemsetAsync(h->b_clist.p, 0xFF, list_cap * sizeof(ChunkSlot<W>), h->init_stream));   // (k_chunk_insert waits for ev_init[1])
                list_cleared = true;
                int sms_ = 148;
                cudaDeviceGetAttribute(&sms_, cudaDevAttrMultiProcessorCount, h->device);
                uint64_t bg = (uint64_t)sms_;                             // background grid: one block an SM (measured 1 / 2 / 8 / 64:
                                                                          //   C2 step 2.070 / 2.089 / 2.122 / 2.125 ms)
                if (const char *e = getenv("DBG_B200_BG_BLOCKS")) {       // tuning: blocks per SM of the background clears
                    const unsigned long long v = strtoull(e, nullptr, 10);
                    if (v >= 1) bg = (uint64_t)sms_ * v;
                }
                LAUNCH(k_chunk_table_init<W>, std::min<uint64_t>(bg, blocks_for(n_cs)), DBG_BLOCK, h->init_stream, ptr<CS>(h->b_ctable), n_cs);
                CKK();
                CK(cudaEventRecord(h->ev_init[1], h->init_stream));
                if (!chunks_only) {
                    // the node table is not needed before k_expand: its 0.3 GB of writes land under k_chunk_insert (which
                    // lives in L2) instead of competing with k_pack for HBM
                    LAUNCH(k_table_init<W>, std::min<uint64_t>(bg, blocks_for(n_ns)), DBG_BLOCK, h->init_stream, ptr<S>(h->b_table), n_ns);
                    CKK();
                    CK(cudaEventRecord(h->ev_init[2], h->init_stream));
                }
                h->preinit = !chunks_only;
            }
        }
        if (!list_cleared) CK(cudaMemsetAsync(h->b_clist.p, 0xFF, list_cap * sizeof(ChunkSlot<W>), h->stream));
    }
    if (h->n_reads) {
        uint64_t rs_grid = blocks_for(h->n_reads);
        if (rs_grid > 148 * 4) rs_grid = 148 * 4;
        LAUNCH(k_read_stats, rs_grid, DBG_BLOCK, h->stream, d_off, h->n_reads, k, d_ctr);
        CKK();
    }
    if (h->n_bases) {
        LAUNCH(k_tile_first, blocks_for(h->n_tiles + 1), DBG_BLOCK, h->stream, d_off, h->n_reads, h->n_bases,
               h->n_tiles, ptr<u32>(h->b_tile));
        CKK();
        // Reads that still sit in host memory (dbg_build*) come over in slices on the copy stream; slice i
        // is packed, and the spans whose 2-word lookahead it completes are masked, while slice i + 1 is
        // on the wire.  Reads already in HBM are one slice.
        uint64_t slice_bytes = SLICE_BYTES;
        if (const char *e = getenv("DBG_B200_SLICE_BYTES")) {             // tests: many small slices
            const unsigned long long v = strtoull(e, nullptr, 10);
            if (v >= 1) slice_bytes = v;
        }
        uint64_t n_slices = h->host_bases ? cdiv(h->n_bases, slice_bytes) : 1;
        if (n_slices > MAX_SLICES) n_slices = MAX_SLICES;
        const uint64_t slice = cdiv(cdiv(h->n_bases, n_slices), DBG_TILE) * DBG_TILE;    // whole tiles
        if (h->host_bases) {
            CK(cudaEventRecord(h->slice_ev[MAX_SLICES], h->stream));                 // the buffers are ours from here on
            CK(cudaStreamWaitEvent(h->copy_stream, h->slice_ev[MAX_SLICES], 0));
            for (uint64_t i = 0; i < n_slices; i++) {
                const uint64_t b0 = i * slice, b1 = b0 + slice < h->n_bases ? b0 + slice : h->n_bases;
                if (b0 >= b1) { n_slices = i; break; }
                CK(cudaMemcpyAsync((u8 *)d_bases + b0, h->host_bases + b0, b1 - b0, cudaMemcpyHostToDevice, h->copy_stream));
                CK(cudaEventRecord(h->slice_ev[i], h->copy_stream));
            }
        }
        // with the reads arriving in several slices, phase 1 of the chunk build (k_chunk_insert) also runs slice
        // by slice, hidden behind the copies; the chunk table is sized from the running sketch and grows
        // (k_chunk_rehash) when later slices bring more distinct chunks than the first ones promised
        const bool stream_insert = h->host_bases && n_slices > 1;
        std::vector<u32> regs(stream_insert ? HLL_M : 0);
        uint64_t s_done = 0, u_done = 0;
        for (uint64_t i = 0; i < n_slices; i++) {
            const uint64_t b0 = i * slice, b1 = b0 + slice < h->n_bases ? b0 + slice : h->n_bases;
            if (b0 >= b1) break;
            const bool last = b1 == h->n_bases;
            if (h->host_bases) CK(cudaStreamWaitEvent(h->stream, h->slice_ev[i], 0));
            const uint64_t w0 = b0 / 32, w1 = last ? h->n_words : b1 / 32;
            LAUNCH(k_pack, blocks_for(w1 - w0), DBG_BLOCK, h->stream, d_bases, h->n_bases, ptr<u64>(h->b_packed),
                   w0, w1, (int)(((uintptr_t)d_bases & 15) == 0), d_ctr);
            CKK();
            {
                const uint64_t s1 = last ? n_spans : (w1 > 2 ? w1 - 2 : 0);           // k_masks reads words s - 1 .. s + 1
                if (s1 > s_done) {
                    LAUNCH(k_masks, blocks_for(s1 - s_done), DBG_BLOCK, h->stream, ptr<const u64>(h->b_packed), d_off,
                           h->n_reads, ptr<const u32>(h->b_tile), h->n_bases, s_done, s1, k, h->geom_logw, h->geom_m,
                           h->sample_mask, ptr<u32>(h->b_stmask), ptr<u32>(h->b_vmask), ptr<u32>(h->b_spanread),
                           h->hinted ? (u32 *)nullptr : ptr<u32>(h->b_hll));
                    CKK();
                    s_done = s1;
                }
                if (stream_insert) {
                    // units whose masks (32 spans + 1) and string overhang (3 words) are complete
                    const uint64_t u1 = last ? cdiv(h->n_bases, DBG_UNIT) : (s_done >= 36 ? (s_done - 4) / 32 : 0);
                    if (u1 > u_done && h->hinted) {
                        if (!h->chunks_streamed) CK(cudaStreamWaitEvent(h->stream, h->ev_init[1], 0));
                        TRY(launch_chunk_insert<W>(h, d_off, u_done, u1));
                        u_done = u1;
                        h->chunks_streamed = true;
                    } else if (u1 > u_done) {
                        void *r_dst = h->ctl ? (void *)(h->ctl + CTL_SKETCH) : (void *)regs.data();
                        CK(cudaMemcpyAsync(r_dst, h->b_hll.p, HLL_M * 4, cudaMemcpyDeviceToHost, h->stream));
                        CK(cudaStreamSynchronize(h->stream));              // the next slices are already on the wire
                        if (h->ctl) memcpy(regs.data(), r_dst, HLL_M * 4);
                        const uint64_t est = (uint64_t)hll_estimate(regs) * ((h->sample_mask >> 6) + 1);
                        TRY(chunk_table_fit<W>(h, est));
                        TRY(launch_chunk_insert<W>(h, d_off, u_done, u1));
                        u_done = u1;
                        h->chunks_streamed = true;
                    }
                }
            }
        }
        if (h->host_bases) {                                                          // H2D time: first to last slice
            CK(cudaStreamSynchronize(h->copy_stream));
            float ms = 0;
            cudaEventElapsedTime(&ms, h->slice_ev[MAX_SLICES], h->slice_ev[n_slices ? n_slices - 1 : 0]);
            st.ms_h2d = ms;
        }
    }
    const int e1 = ev_mark(h);

    return build_chunks<W>(h, d_off, e0, e1);
}

// ---- min_count (extension): the graph of the windows whose (k+1)-mer occurs >= min_count times, made from the
// unfiltered table (whose counters are those totals); kernels k_mc_* in dbg_kernels.cuh ----
template <int W>
static int min_count_stage(dbg_t *h) {
    typedef Slot<W> S;
    if (h->expand_pending) TRY(settle_node_table<W>(h, false));        // the counters must be complete
    if (!h->n_bases || !h->st.n_windows) return DBG_OK;
    uint64_t grid = blocks_for(h->n_slots);
    if (grid > 148 * 16) grid = 148 * 16;
    LAUNCH(k_mc_filter<W>, grid, DBG_BLOCK, h->stream, ptr<S>(h->b_table), h->n_slots, h->min_count);
    CKK();
    if (!h->chunks_valid || h->chunks_merged) return fail(h, DBG_ERR_STATE, "min_count needs the chunks of this build");
    uint64_t rgrid = blocks_for(h->n_cslots + h->n_clist);
    if (rgrid > 148 * 16) rgrid = 148 * 16;
    LAUNCH(k_mc_retag<W>, rgrid, DBG_BLOCK, h->stream, ptr<const ChunkSlot<W>>(h->b_ctable), h->n_cslots,
           ptr<const ChunkSlot<W>>(h->b_clist), h->n_clist, h->k, h->batch_bits, ptr<S>(h->b_table), geom_of(h),
           ptr<DevCounters>(h->b_ctr));
    CKK();
    return DBG_OK;
}

// ---- graph stage A: degrees, pruningEdges, branch flags, successor links (kernel 4) ----
template <int W>
static int graph_finalize(dbg_t *h) {
    typedef Slot<W> S;
    dbg_stats_t &st = h->st;
    DevCounters *d_ctr = ptr<DevCounters>(h->b_ctr);
    const int e3 = ev_mark(h);
    for (;;) {
        const uint64_t n_slots = h->n_slots;
        CK(cudaMemsetAsync(&d_ctr->n_nodes, 0, 6 * 8, h->stream));      // n_nodes .. n_starts, n_list_a
        TRY(reserve(h, h->b_link, n_slots * 8));                        // (the walk's links; until then the far list of k_finalize)
        TRY(reserve(h, h->b_flags, n_slots));
        TRY(reserve(h, h->b_kept, n_slots * 2));
        TRY(reserve(h, h->b_succ, n_slots * 4));
        uint64_t fin_grid = cdiv(cdiv(n_slots, FIN_SLOTS), DBG_WARPS);
        if (fin_grid > 148 * 8) fin_grid = 148 * 8;
        if (fin_grid < 1) fin_grid = 1;
        LAUNCH(k_finalize<W>, fin_grid, DBG_BLOCK, h->stream, ptr<const S>(h->b_table), geom_of(h), h->k,
               h->threshold, ptr<u8>(h->b_flags), ptr<u16>(h->b_kept), ptr<u32>(h->b_succ), h->shard_n > 1 ? 1 : 0,
               ptr<u32>(h->b_link), d_ctr);
        CKK();
        if (geom_of(h).n_regions > 1) {
            uint64_t far_grid = 148 * 16;                               // (measured 4 / 8 / 16 / 32 blocks an SM: finalize 0.278 / 0.268 / 0.265 / 0.266 ms at C2)
            if (const char *e = getenv("DBG_B200_FAR_GRID")) {           // tuning: blocks per SM
                const unsigned long long v = strtoull(e, nullptr, 10);
                if (v >= 1) far_grid = 148 * v;
            }
            if (far_grid > blocks_for(n_slots)) far_grid = blocks_for(n_slots);   // (the far list is shorter than the table)
            LAUNCH(k_finalize_far<W>, far_grid, DBG_BLOCK, h->stream, ptr<const S>(h->b_table), geom_of(h), h->k,
                   ptr<const u16>(h->b_kept), ptr<const u32>(h->b_link), ptr<u8>(h->b_flags), ptr<u32>(h->b_succ),
                   h->shard_n > 1 ? 1 : 0, d_ctr);
            CKK();
        }
        if (h->min_count > 1 && h->n_bases) {                           // nodes created as the SOURCE of a surviving window
            LAUNCH(k_mc_starts<W>, fin_grid, DBG_BLOCK, h->stream, ptr<const u64>(h->b_packed), (const u64 *)h->d_off,
                   ptr<const S>(h->b_table), geom_of(h), h->k, h->read_idx_base, ptr<u8>(h->b_flags), d_ctr);
            CKK();
        }
        TRY(fetch_counters(h));
        tm_flush(h);
        if (h->hinted) {                                               // what the cold path checks between its phases
            ChunkCounters cc;
            memcpy(&cc, h->ctl + CTL_CC, sizeof cc);
            if (h->ctr.bad_alphabet) return fail(h, DBG_ERR_ALPHABET, "a read holds a byte other than A/C/G/T");
            if (h->ctr.too_long) return fail(h, DBG_ERR_OVERFLOW, "a read is longer than 2^32-1 bases");
            if (cc.chunk_overflow || cc.bad_chunk || cc.list_overflow) return DBG_RETRY_COLD;   // the hint did not fit
            h->last_est_chunks = cc.n_first;                          // exact: entries the chunk table's CAS winners created
            st.n_windows = h->ctr.n_windows;
            st.n_chunk_instances = cc.n_instances;
            h->hinted = false;                                         // validated; a node-table rebuild below runs cold
        }
        if (h->merge_check_pending) {
            h->merge_check_pending = false;
            ChunkCounters cc;
            memcpy(&cc, h->ctl + CTL_CC, sizeof cc);
            if (cc.chunk_overflow) return fail(h, DBG_ERR_NOMEM, "chunk table overflowed while merging records");
        }
        if (!(h->expand_pending && h->ctr.overflow)) break;
        // the node table the chunk build left had overflowed (its flag is looked at here, with this
        // stage's own counters, to save a host round trip on the common path): rebuild it larger
        CK(cudaMemsetAsync(&d_ctr->fault, 0, 8, h->stream));            // whatever k_finalize met in the torn table
        TRY(settle_node_table<W>(h, true));
    }
    h->expand_pending = false;
    st.n_nodes = h->ctr.n_nodes;
    st.n_edges = h->ctr.n_edges;
    st.n_multi = h->ctr.n_multi;
    st.n_branch = h->ctr.n_branch;
    st.n_starts = h->ctr.n_starts;
    if (h->ctr.fault && !getenv("DBG_B200_IGNORE_FAULT"))
        return fail(h, DBG_ERR_STATE, "internal consistency fault %u (info %u) in k_finalize", h->ctr.fault,
                    h->ctr.fault_info);
    if (st.n_multi && h->threshold == 0.0)
        return fail(h, DBG_ERR_ZERODIV, "threshold == 0 with a multi-successor node (debruijn.py:163)");
    const int e4 = ev_mark(h);
    tm_set(h, &st.ms_finalize, e3, e4);
    tm_add(h, &st.ms_build_total, e3, e4);
    h->order_active = false;
    if (h->chunks_merged) h->shard_hint_nodes = st.n_nodes;          // size of the next shard this handle expands
    if (h->chunks_valid && !h->chunks_merged && h->shard_n < 2) {      // sizes for the next build on this handle
        h->hint.valid = true;
        h->hint.n_bases = h->n_bases;
        h->hint.chunks = h->last_est_chunks;
        h->hint.nodes = st.n_nodes;
        h->hint.n_batches = h->n_batches;
    }
    return DBG_OK;
}

// ---- graph stage B: Counter.most_common tie order (debruijn.py:159-164), scan of the LOCAL reads ----
template <int W>
static int graph_order_scan(dbg_t *h) {
    dbg_stats_t &st = h->st;
    const int e4 = ev_mark(h);
    if ((h->build_flags & DBG_EXACT_ORDER) && st.n_multi) {
        TRY(side_set<W>(h, F_MULTI, st.n_multi, &h->order_cap, &h->order_bits));
        h->order_active = true;
        if ((h->n_bases || h->chunks_merged) && h->chunks_valid) {
            // from the distinct chunks of the local reads (2 - 20 x fewer than windows)
            typedef ChunkSlot<W> CS;
            uint64_t g = blocks_for(h->n_cslots + h->n_clist);
            if (g > 148 * 32) g = 148 * 32;
            LAUNCH(k_order_chunks<W>, g, DBG_BLOCK, h->stream, ptr<const CS>(h->b_ctable), h->n_cslots,
                   ptr<const CS>(h->b_clist), h->n_clist, h->k, h->batch_bits, ptr<SideEntry<W>>(h->b_order),
                   h->order_cap, ptr<const u32>(h->b_bitmap), h->order_bits);
            CKK();
        }
    }
    const int e5 = ev_mark(h);
    tm_set(h, &st.ms_order, e4, e5);
    tm_add(h, &st.ms_build_total, e4, e5);
    return DBG_OK;
}

// ---- graph stage C: apply the order, branch / start lists, tip pruning, pull-out reads ----
template <int W>
static int graph_complete(dbg_t *h) {
    typedef Slot<W> S;
    dbg_stats_t &st = h->st;
    const uint64_t n_slots = h->n_slots;
    const int k = h->k;
    DevCounters *d_ctr = ptr<DevCounters>(h->b_ctr);
    const int e4 = ev_mark(h);
    if (h->order_active) {
        LAUNCH(k_order_apply<W>, blocks_for(h->order_cap), DBG_BLOCK, h->stream, ptr<const S>(h->b_table), geom_of(h), k,
               h->threshold, ptr<const SideEntry<W>>(h->b_order), h->order_cap, ptr<u16>(h->b_kept),
               ptr<u32>(h->b_succ), ptr<u8>(h->b_flags), h->shard_n > 1 ? 1 : 0, d_ctr);
        CKK();
        TRY(check_fault(h, "k_order_*"));
        h->order_active = false;
    }
    const int e5 = ev_mark(h);

    // ---- branch_kmer and starts, in first-occurrence order ----
    TRY(collect_sorted<W>(h, F_BRANCH, st.n_branch, h->b_branch, &d_ctr->n_list_a));
    TRY(collect_sorted<W>(h, F_START, st.n_starts, h->b_starts, &d_ctr->n_list_b));
    h->branch.resize(st.n_branch);
    if (st.n_branch) {
        CK(cudaMemcpyAsync(h->branch.data(), h->b_branch.p, st.n_branch * 4, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    const int e6 = ev_mark(h);
    if (h->shard_n > 1 && st.n_branch)
        return fail(h, DBG_ERR_STATE, "a sharded graph with branch nodes needs the whole table on one rank for tip pruning "
                    "(debruijn.py:241-254): gather the shards (dbg_export_records / dbg_merge_records) instead");
    h->shard_stage = 0;

    // ---- tip pruning (debruijn.py:241-254) ----
    TRY(tip_prune<W>(h));
    st.n_pulled = h->pulled.size();
    const int e7 = ev_mark(h);

    // ---- pull_out_read (debruijn.py:274-283), over the LOCAL reads ----
    h->read_flags_valid = false;
    st.n_pull_reads = 0;
    const bool want_reads = !(h->build_flags & (DBG_FINAL | DBG_NO_PULL_READS));
    if (want_reads) {
        h->read_flags.assign(h->n_reads, 0);
        if (st.n_branch && h->n_reads && h->n_bases) {
            TRY(reserve(h, h->b_readflags, h->n_reads));
            CK(cudaMemsetAsync(h->b_readflags.p, 0, h->n_reads, h->stream));
            uint64_t cap; int bits;
            TRY(side_set<W>(h, F_BRANCH, st.n_branch, &cap, &bits));
            LAUNCH(k_pull_reads<W>, h->n_tiles, DBG_BLOCK, h->stream, ptr<const u64>(h->b_packed), (const u64 *)h->d_off,
                   ptr<const u32>(h->b_tile), h->n_bases, k, ptr<SideEntry<W>>(h->b_order), cap,
                   ptr<const u32>(h->b_bitmap), bits, ptr<u8>(h->b_readflags), (const u32 *)h->d_read_batch);
            CKK();
            CK(cudaMemcpyAsync(h->read_flags.data(), h->b_readflags.p, h->n_reads, cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            uint64_t n = 0;
            for (uint8_t f : h->read_flags) n += f;
            st.n_pull_reads = n;
        }
        h->read_flags_valid = true;
    }
    const int e8 = ev_mark(h);
    tm_add(h, &st.ms_order, e4, e5);
    tm_set(h, &st.ms_lists, e5, e6);
    tm_set(h, &st.ms_tips, e6, e7);
    tm_set(h, &st.ms_pull_reads, e7, e8);
    tm_add(h, &st.ms_build_total, e4, e8);
    h->built = true;
    return DBG_OK;
}

// ---- node records: the unit of the multi-GPU exchange (one per occupied slot) ----
template <int W>
static int export_counts_impl(dbg_t *h, uint32_t n_parts, uint64_t *counts) {
    TRY(reserve(h, h->b_off64, (uint64_t)n_parts * 8 * 2));
    CK(cudaMemsetAsync(h->b_off64.p, 0, (uint64_t)n_parts * 8 * 2, h->stream));
    LAUNCH(k_rec_count<W>, blocks_for(h->n_slots), DBG_BLOCK, h->stream, ptr<const Slot<W>>(h->b_table), h->n_slots,
           n_parts, ptr<u64>(h->b_off64));
    CKK();
    CK(cudaMemcpyAsync(counts, h->b_off64.p, (uint64_t)n_parts * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DBG_OK;
}

template <int W>
static int export_records_impl(dbg_t *h, uint32_t n_parts, const uint64_t *counts, uint64_t *d_out) {
    std::vector<u64> cursors(n_parts);
    u64 acc = 0;
    for (uint32_t p = 0; p < n_parts; p++) { cursors[p] = acc; acc += counts[p]; }
    u64 *d_cur = ptr<u64>(h->b_off64) + n_parts;
    CK(cudaMemcpyAsync(d_cur, cursors.data(), (uint64_t)n_parts * 8, cudaMemcpyHostToDevice, h->stream));
    LAUNCH(k_rec_scatter<W>, blocks_for(h->n_slots), DBG_BLOCK, h->stream, ptr<const Slot<W>>(h->b_table), h->n_slots,
           n_parts, d_cur, (u64 *)d_out);
    CKK();
    SYNC_FOR_CALLER();                                    // (pageable H2D sources are staged before cudaMemcpyAsync returns)
    return DBG_OK;
}

// ---- chunk records: the unit of the multi-GPU exchange when it happens before any node table ----
template <int W>
static int chunk_export_counts_impl(dbg_t *h, uint32_t n_parts, uint64_t *counts) {
    typedef ChunkSlot<W> CS;
    const ChunkGeom geom = chunk_geom(h->k);
    TRY(reserve(h, h->b_off64, (uint64_t)n_parts * 8 * 2));
    CK(cudaMemsetAsync(h->b_off64.p, 0, (uint64_t)n_parts * 8 * 2, h->stream));
    LAUNCH(k_chunk_rec_count<W>, blocks_for(h->n_cslots + h->n_clist), DBG_BLOCK, h->stream, ptr<const CS>(h->b_ctable),
           h->n_cslots, ptr<const CS>(h->b_clist), h->n_clist, geom.w, geom.m, h->batch_bits, n_parts, ptr<u64>(h->b_off64));
    CKK();
    CK(cudaMemcpyAsync(counts, h->b_off64.p, (uint64_t)n_parts * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DBG_OK;
}

template <int W>
static int chunk_export_records_impl(dbg_t *h, uint32_t n_parts, const uint64_t *counts, uint64_t *d_out) {
    typedef ChunkSlot<W> CS;
    const ChunkGeom geom = chunk_geom(h->k);
    std::vector<u64> cursors(n_parts);
    u64 acc = 0;
    for (uint32_t p = 0; p < n_parts; p++) { cursors[p] = acc; acc += counts[p]; }
    u64 *d_cur = ptr<u64>(h->b_off64) + n_parts;
    CK(cudaMemcpyAsync(d_cur, cursors.data(), (uint64_t)n_parts * 8, cudaMemcpyHostToDevice, h->stream));
    LAUNCH(k_chunk_rec_scatter<W>, blocks_for(h->n_cslots + h->n_clist), DBG_BLOCK, h->stream, ptr<const CS>(h->b_ctable),
           h->n_cslots, ptr<const CS>(h->b_clist), h->n_clist, geom.w, geom.m, h->batch_bits, n_parts, d_cur, (CS *)d_out);
    CKK();
    SYNC_FOR_CALLER();
    return DBG_OK;
}

template <int W>
static int chunk_table_reset_impl(dbg_t *h, uint64_t expected_chunks) {
    typedef ChunkSlot<W> CS;
    const uint64_t n = (uint64_t)((double)expected_chunks * 1.7) + 4096;      // received records are an upper bound
    if (n >= 0xFFFFFFF0ull) return fail(h, DBG_ERR_OVERFLOW, "chunk table would exceed 2^32 slots");
    TRY(reserve(h, h->b_ctable, n * sizeof(CS)));
    TRY(reserve(h, h->b_cc, sizeof(ChunkCounters)));
    LAUNCH(k_chunk_table_init<W>, blocks_for(n), DBG_BLOCK, h->stream, ptr<CS>(h->b_ctable), n);
    CKK();
    CK(cudaMemsetAsync(h->b_cc.p, 0, sizeof(ChunkCounters), h->stream));
    h->n_cslots = n;
    h->n_clist = 0;                                                   // boundary chunks arrive as records like the others
    h->region_mode = want_regions(h);
    h->chunks_merged = true;
    h->have_hbytes = false; h->have_vhash = false; h->have_order = false;
    h->chunks_valid = true;
    h->tags_local = false;
    h->expand_pending = false;
    h->table_ready = false;
    h->built = h->walked = false;
    return DBG_OK;
}

template <int W>
static int chunk_merge_records_impl(dbg_t *h, const uint64_t *d_records, uint64_t n) {
    typedef ChunkSlot<W> CS;
    if (n == 0) return DBG_OK;
    LAUNCH(k_chunk_rec_merge<W>, blocks_for(n), DBG_BLOCK, h->stream, ptr<CS>(h->b_ctable), (u32)h->n_cslots,
           (const CS *)d_records, n, ptr<ChunkCounters>(h->b_cc));
    CKK();
    return DBG_OK;                                        // the overflow flag is looked at by dbg_chunk_expand
}

// phase 2 on the merged chunk table: this rank's shard of the node table
template <int W>
static int chunk_expand_impl(dbg_t *h) {
    typedef Slot<W> S;
    typedef ChunkSlot<W> CS;
    DevCounters *d_ctr = ptr<DevCounters>(h->b_ctr);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    if (!h->tm.empty()) { CK(cudaStreamSynchronize(h->stream)); tm_flush(h); }
    h->n_ev = 0;
    uint64_t ex_grid = cdiv(cdiv(h->n_cslots + h->n_clist, 32), DBG_WARPS);
    if (ex_grid > (uint64_t)sms * DBG_EXPAND_MIN_BLOCKS) ex_grid = (uint64_t)sms * DBG_EXPAND_MIN_BLOCKS;
    if (ex_grid < 1) ex_grid = 1;
    h->ex_grid = ex_grid;
    std::vector<u32> sketch(HLL_M, 0);
    TRY(reserve(h, h->b_hll, HLL_M * 4));
    CK(cudaMemsetAsync(&d_ctr->overflow, 0, 4, h->stream));
    if (h->shard_hint_nodes && h->ctl && !getenv("DBG_B200_NO_HINT") && !getenv("DBG_B200_NODE_SLOTS")) {
        // the shard this handle expanded last time tells the size of this one: no sketch pass, no host round trip
        // here; the overflow flags (chunk merge, node table) are looked at by graph_finalize with its own counters
        double load = 0.5;
        if (const char *e = getenv("DBG_B200_TABLE_LOAD")) {
            const double v = atof(e) / 100.0;
            if (v >= 0.05 && v <= 0.95) load = v;
        }
        h->vb_bits = 0;
        h->n_slots = fit_slots(h, (uint64_t)((double)h->shard_hint_nodes * 1.06 * 1.03 / load) + 1024);
        h->have_hbytes = false; h->have_vhash = false; h->have_order = false;
        if (!h->region_mode && pass_count(h, h->n_slots * sizeof(S)) > 2) TRY(launch_node_sketch<W>(h, ex_grid));
        CK(cudaMemcpyAsync(h->ctl + CTL_CC, h->b_cc.p, sizeof(ChunkCounters), cudaMemcpyDeviceToHost, h->stream));
        h->merge_check_pending = true;
        const uint64_t lw = h->st.n_windows;
        h->st.n_windows = 1;
        const int rc2 = expand_node_table<W>(h);
        h->st.n_windows = lw;
        TRY(rc2);
        h->expand_pending = true;
        h->table_ready = true;
        return DBG_OK;
    }
    TRY(launch_node_sketch<W>(h, ex_grid));
    CK(cudaMemcpyAsync(sketch.data(), h->b_hll.p, HLL_M * 4, cudaMemcpyDeviceToHost, h->stream));
    ChunkCounters cc;
    CK(cudaMemcpyAsync(&cc, h->b_cc.p, sizeof cc, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (cc.chunk_overflow) return fail(h, DBG_ERR_NOMEM, "chunk table overflowed while merging records");
    const uint64_t local_windows = h->st.n_windows;
    h->st.n_windows = 1;                                              // expand_node_table skips its passes for an empty read set,
    const int rc = node_table_from_chunks<W>(h, (uint64_t)hll_estimate(sketch));   // but the chunks here may all be other ranks'
    h->st.n_windows = local_windows;
    TRY(rc);
    h->table_ready = true;
    return DBG_OK;
}

template <int W>
static int table_reset_impl(dbg_t *h, uint64_t expected_nodes) {
    typedef Slot<W> S;
    h->region_mode = want_regions(h);
    h->vb_bits = 0;
    const uint64_t n_slots = fit_slots(h, expected_nodes * 2 + 1024);
    if (n_slots >= 0xFFFFFFF0ull) return fail(h, DBG_ERR_OVERFLOW, "node table would exceed 2^32 slots");
    h->n_slots = n_slots;
    TRY(reserve(h, h->b_table, n_slots * sizeof(S)));
    LAUNCH(k_table_init<W>, blocks_for(n_slots), DBG_BLOCK, h->stream, ptr<S>(h->b_table), n_slots);
    CKK();
    DevCounters *d_ctr = ptr<DevCounters>(h->b_ctr);
    CK(cudaMemsetAsync(&d_ctr->overflow, 0, 4, h->stream));
    h->st.table_slots = n_slots;
    h->st.table_bytes = n_slots * sizeof(S);
    h->table_ready = true;
    h->tags_local = false;                              // records of other ranks' reads are merged in
    h->expand_pending = false;
    h->built = h->walked = false;
    return DBG_OK;
}

template <int W>
static int merge_records_impl(dbg_t *h, const uint64_t *d_records, uint64_t n) {
    if (n == 0) return DBG_OK;
    LAUNCH(k_rec_merge<W>, blocks_for(n), DBG_BLOCK, h->stream, ptr<Slot<W>>(h->b_table), geom_of(h),
           (const u64 *)d_records, n, ptr<DevCounters>(h->b_ctr));
    CKK();
    TRY(fetch_counters(h));
    if (h->ctr.overflow) return fail(h, DBG_ERR_NOMEM, "node table overflowed while merging records");
    return DBG_OK;
}

template <int W>
static int order_export_impl(dbg_t *h, uint64_t *d_out, uint64_t capacity, uint64_t *n_out) {
    *n_out = 0;
    if (!h->order_active) return DBG_OK;
    DevCounters *d_ctr = ptr<DevCounters>(h->b_ctr);
    CK(cudaMemsetAsync(&d_ctr->n_list_a, 0, 8, h->stream));
    LAUNCH(k_side_export<W>, blocks_for(h->order_cap), DBG_BLOCK, h->stream, ptr<const SideEntry<W>>(h->b_order),
           h->order_cap, (u64 *)d_out, capacity, &d_ctr->n_list_a);
    CKK();
    TRY(fetch_counters(h));
    *n_out = h->ctr.n_list_a;
    if (h->ctr.n_list_a > capacity) return fail(h, DBG_ERR_ARG, "order export buffer too small");
    return DBG_OK;
}

template <int W>
static int order_merge_impl(dbg_t *h, const uint64_t *d_records, uint64_t n) {
    if (n == 0 || !h->order_active) return DBG_OK;
    LAUNCH(k_side_merge<W>, blocks_for(n), DBG_BLOCK, h->stream, ptr<SideEntry<W>>(h->b_order), h->order_cap,
           (const u64 *)d_records, n);
    CKK();
    CK(cudaStreamSynchronize(h->stream));
    return DBG_OK;
}

// -------------------------------------------------------------------------------------------------
// generic-alphabet engine (dbg_generic.cuh): construct_graph for reads over any single-byte alphabet
// -------------------------------------------------------------------------------------------------
static int gen_sort_edges(dbg_t *h, uint64_t n) {
    // edge array -> sorted by (source slot, count descending, first appearance): b_earr holds the result
    if (n == 0) return DBG_OK;
#ifdef DBG_EMU
    GEdge *e = ptr<GEdge>(h->b_earr);
    std::stable_sort(e, e + n, [](const GEdge &a, const GEdge &b) {
        if (a.src != b.src) return a.src < b.src;
        if (a.cnt != b.cnt) return a.cnt > b.cnt;
        return a.first < b.first;
    });
#else
    TRY(reserve(h, h->b_tag, n * 8));
    TRY(reserve(h, h->b_tag2, n * 8));
    TRY(reserve(h, h->b_slot, n * 4));
    TRY(reserve(h, h->b_slot2, n * 4));
    TRY(reserve(h, h->b_earr2, n * sizeof(GEdge)));
    size_t tmp = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, ptr<u64>(h->b_tag), ptr<u64>(h->b_tag2), ptr<u32>(h->b_slot),
                                       ptr<u32>(h->b_slot2), (int64_t)n, 0, 64, h->stream));
    TRY(reserve(h, h->b_sorttmp, tmp));
    // LSD: stable sort by the minor key (first appearance), then by the major one (source, ~count)
    LAUNCH(k_g_sortkeys1, blocks_for(n), DBG_BLOCK, h->stream, ptr<const GEdge>(h->b_earr), n, ptr<u64>(h->b_tag), ptr<u32>(h->b_slot));
    CKK();
    h->st.n_launches++;
    CK(cub::DeviceRadixSort::SortPairs(h->b_sorttmp.p, tmp, ptr<u64>(h->b_tag), ptr<u64>(h->b_tag2), ptr<u32>(h->b_slot),
                                       ptr<u32>(h->b_slot2), (int64_t)n, 0, 64, h->stream));
    LAUNCH(k_g_sortkeys2, blocks_for(n), DBG_BLOCK, h->stream, ptr<const GEdge>(h->b_earr), ptr<const u32>(h->b_slot2), n,
           ptr<u64>(h->b_tag));
    CKK();
    h->st.n_launches++;
    CK(cub::DeviceRadixSort::SortPairs(h->b_sorttmp.p, tmp, ptr<u64>(h->b_tag), ptr<u64>(h->b_tag2), ptr<u32>(h->b_slot2),
                                       ptr<u32>(h->b_slot), (int64_t)n, 0, 64, h->stream));
    LAUNCH(k_g_gather, blocks_for(n), DBG_BLOCK, h->stream, ptr<const GEdge>(h->b_earr), ptr<const u32>(h->b_slot), n,
           ptr<GEdge>(h->b_earr2));
    CKK();
    std::swap(h->b_earr, h->b_earr2);
#endif
    return DBG_OK;
}

static int gen_build(dbg_t *h, const u8 *d_bases, const u64 *d_off) {
    typedef Slot<2> S;
    const int k = h->k;
    dbg_stats_t &st = h->st;
    DevCounters *d_ctr = ptr<DevCounters>(h->b_ctr);
    const int e0 = ev_mark(h);
    if (h->host_bases && h->n_bases)
        CK(cudaMemcpyAsync((u8 *)d_bases, h->host_bases, h->n_bases, cudaMemcpyHostToDevice, h->stream));
    // ---- symbols -> codes; window count ----
    TRY(reserve(h, h->b_lut, GEN_LUT_BYTES));
    CK(cudaMemcpyAsync(h->b_lut.p, h->code_of_byte, 256, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(ptr<u8>(h->b_lut) + 256, h->sym_of_code, 257, cudaMemcpyHostToDevice, h->stream));
    TRY(reserve(h, h->b_codes, h->n_bases + 64));
    if (h->n_bases) {
        LAUNCH(k_g_encode, blocks_for(h->n_bases), DBG_BLOCK, h->stream, d_bases, h->n_bases, ptr<const u8>(h->b_lut),
               ptr<u8>(h->b_codes), d_ctr);
        CKK();
    }
    if (h->n_reads) {
        uint64_t rs_grid = blocks_for(h->n_reads);
        if (rs_grid > 148 * 4) rs_grid = 148 * 4;
        LAUNCH(k_read_stats, rs_grid, DBG_BLOCK, h->stream, d_off, h->n_reads, k, d_ctr);
        CKK();
    }
    TRY(fetch_counters(h));
    if (h->ctr.bad_alphabet) return fail(h, DBG_ERR_ALPHABET, "a read holds a byte that is not a symbol of this handle's alphabet");
    if (h->ctr.too_long) return fail(h, DBG_ERR_OVERFLOW, "a read is longer than 2^32-1 symbols");
    st.n_windows = h->ctr.n_windows;
    const int e1 = ev_mark(h);
    // ---- node table + edge table (debruijn.py:123-144, 213-222) ----
    const uint64_t n_nslots = 2 * (st.n_windows + h->n_reads) + 64, n_eslots = 2 * st.n_windows + 64;
    if (n_nslots >= 0xFFFFFFF0ull) return fail(h, DBG_ERR_OVERFLOW, "node table would exceed 2^32 slots");
    h->n_slots = n_nslots;
    h->n_eslots = n_eslots;
    TRY(reserve(h, h->b_table, n_nslots * sizeof(S)));
    TRY(reserve(h, h->b_etable, n_eslots * sizeof(S)));
    LAUNCH(k_table_init<2>, blocks_for(n_nslots), DBG_BLOCK, h->stream, ptr<S>(h->b_table), n_nslots);
    CKK();
    LAUNCH(k_table_init<2>, blocks_for(n_eslots), DBG_BLOCK, h->stream, ptr<S>(h->b_etable), n_eslots);
    CKK();
    if (st.n_windows) {
        LAUNCH(k_g_insert, blocks_for(h->n_bases), DBG_BLOCK, h->stream, ptr<const u8>(h->b_codes), d_off, h->n_reads,
               h->n_bases, k, h->sym_bits, ptr<S>(h->b_table), n_nslots, ptr<S>(h->b_etable), n_eslots, h->read_idx_base, d_ctr);
        CKK();
    }
    // ---- edge array in E order, CSR ----
    TRY(reserve(h, h->b_earr, (st.n_windows + 1) * sizeof(GEdge)));
    CK(cudaMemsetAsync(&d_ctr->n_list_a, 0, 8, h->stream));
    LAUNCH(k_g_edges, blocks_for(n_eslots), DBG_BLOCK, h->stream, ptr<const S>(h->b_etable), n_eslots, ptr<const S>(h->b_table),
           n_nslots, k, h->sym_bits, ptr<GEdge>(h->b_earr), &d_ctr->n_list_a, d_ctr);
    CKK();
    TRY(check_fault(h, "the generic table build"));
    if (h->ctr.overflow) return fail(h, DBG_ERR_STATE, "internal consistency fault: a table sized by its upper bound overflowed");
    const uint64_t n_e = h->ctr.n_list_a;
    h->n_earr = n_e;
    const int e2 = ev_mark(h);
    TRY(gen_sort_edges(h, n_e));
    TRY(reserve(h, h->b_estart, n_nslots * 4));
    TRY(reserve(h, h->b_edeg, n_nslots * 4));
    TRY(reserve(h, h->b_enk, n_nslots * 4));
    CK(cudaMemsetAsync(h->b_edeg.p, 0, n_nslots * 4, h->stream));
    CK(cudaMemsetAsync(h->b_estart.p, 0, n_nslots * 4, h->stream));
    if (n_e) {
        LAUNCH(k_g_csr, blocks_for(n_e), DBG_BLOCK, h->stream, ptr<const GEdge>(h->b_earr), n_e, ptr<u32>(h->b_estart),
               ptr<u32>(h->b_edeg));
        CKK();
    }
    // ---- degrees / pruningEdges / branch flags ----
    CK(cudaMemsetAsync(&d_ctr->n_nodes, 0, 5 * 8, h->stream));
    TRY(reserve(h, h->b_flags, n_nslots));
    TRY(reserve(h, h->b_kept, n_nslots * 2));
    TRY(reserve(h, h->b_succ, n_nslots * 4));
    LAUNCH(k_g_finalize, blocks_for(n_nslots), DBG_BLOCK, h->stream, ptr<const S>(h->b_table), n_nslots,
           ptr<const GEdge>(h->b_earr), ptr<const u32>(h->b_estart), ptr<const u32>(h->b_edeg), h->threshold,
           ptr<u8>(h->b_flags), ptr<u16>(h->b_kept), ptr<u32>(h->b_succ), ptr<u32>(h->b_enk), d_ctr);
    CKK();
    TRY(fetch_counters(h));
    st.n_nodes = h->ctr.n_nodes;
    st.n_edges = h->ctr.n_edges;
    st.n_multi = h->ctr.n_multi;
    st.n_branch = h->ctr.n_branch;
    st.n_starts = h->ctr.n_starts;
    st.table_slots = n_nslots;
    st.table_bytes = (n_nslots + n_eslots) * sizeof(S);
    st.hll_estimate = 0;
    if (st.n_multi && h->threshold == 0.0)
        return fail(h, DBG_ERR_ZERODIV, "threshold == 0 with a multi-successor node (debruijn.py:163)");
    const int e3 = ev_mark(h);
    tm_set(h, &st.ms_pack, e0, e1);
    tm_set(h, &st.ms_insert, e1, e2);
    tm_set(h, &st.ms_finalize, e2, e3);
    h->table_ready = true;
    h->tags_local = true;
    // ---- lists, tip pruning, pull-out reads ----
    TRY(collect_sorted<2>(h, F_BRANCH, st.n_branch, h->b_branch, &d_ctr->n_list_a));
    TRY(collect_sorted<2>(h, F_START, st.n_starts, h->b_starts, &d_ctr->n_list_b));
    h->branch.resize(st.n_branch);
    if (st.n_branch) {
        CK(cudaMemcpyAsync(h->branch.data(), h->b_branch.p, st.n_branch * 4, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    const int e4 = ev_mark(h);
    TRY(tip_prune<2>(h));
    st.n_pulled = h->pulled.size();
    const int e5 = ev_mark(h);
    h->read_flags_valid = false;
    st.n_pull_reads = 0;
    if (!(h->build_flags & (DBG_FINAL | DBG_NO_PULL_READS))) {
        h->read_flags.assign(h->n_reads, 0);
        if (st.n_branch && h->n_reads && h->n_bases) {
            TRY(reserve(h, h->b_readflags, h->n_reads));
            CK(cudaMemsetAsync(h->b_readflags.p, 0, h->n_reads, h->stream));
            LAUNCH(k_g_pull_reads, blocks_for(h->n_bases), DBG_BLOCK, h->stream, ptr<const u8>(h->b_codes), d_off, h->n_reads,
                   h->n_bases, k, h->sym_bits, ptr<const S>(h->b_table), n_nslots, ptr<const u8>(h->b_flags), ptr<u8>(h->b_readflags));
            CKK();
            CK(cudaMemcpyAsync(h->read_flags.data(), h->b_readflags.p, h->n_reads, cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            uint64_t n = 0;
            for (uint8_t f : h->read_flags) n += f;
            st.n_pull_reads = n;
        }
        h->read_flags_valid = true;
    }
    const int e6 = ev_mark(h);
    tm_set(h, &st.ms_lists, e3, e4);
    tm_set(h, &st.ms_tips, e4, e5);
    tm_set(h, &st.ms_pull_reads, e5, e6);
    tm_set(h, &st.ms_build_total, e0, e6);
    h->built = true;
    return DBG_OK;
}

// -------------------------------------------------------------------------------------------------
// walk
// -------------------------------------------------------------------------------------------------
struct EnumFrame {
    u32 junction_index;     // index into the branch list
    int next_child;
};

template <int W>
static int walk_impl(dbg_t *h, int stop_at_branch) {
    typedef Slot<W> S;
    dbg_stats_t &st = h->st;
    const uint64_t n_slots = h->n_slots;
    DevCounters *d_ctr = ptr<DevCounters>(h->b_ctr);
    h->contigs.clear();
    h->segments.clear();
    h->cview = nullptr; h->n_cview = 0;
    st.n_contigs = 0;
    st.n_cycle_starts = 0;
    st.jump_rounds = 0;
    h->n_ev_walk = EV_WALK;
    const int e0 = ev_mark_walk(h);

    // ---- list ranking over the pruned successor array: chain nodes -> splitters -> super-splitters ----
    TRY(reserve(h, h->b_link, n_slots * 8));
    const uint64_t split_cap = (n_slots >> WALK_SPLIT_LOG2) * 3 / 2 + st.n_starts + 4096;
    const uint64_t super_cap = (split_cap >> WALK_SUPER_LOG2) * 3 / 2 + 4096;
    TRY(reserve(h, h->b_split, split_cap * 4));
    TRY(reserve(h, h->b_super, super_cap * 4));
    CK(cudaMemsetAsync(&d_ctr->n_list_a, 0, 16, h->stream));             // n_list_a (splitters), n_list_b (supers)
    LAUNCH(k_link_init, blocks_for(n_slots), DBG_BLOCK, h->stream, ptr<const u8>(h->b_flags),
           ptr<const u16>(h->b_kept), ptr<const u32>(h->b_succ), n_slots, ptr<u64>(h->b_link), ptr<u32>(h->b_split),
           split_cap, &d_ctr->n_list_a, ptr<u32>(h->b_super), super_cap, &d_ctr->n_list_b);
    CKK();
    // the list lengths stay on the device: the kernels below read them there (no host round trip)
    LAUNCH(k_sublist, blocks_for(split_cap), DBG_BLOCK, h->stream, ptr<u64>(h->b_link), ptr<const u32>(h->b_split),
           (const u64 *)&d_ctr->n_list_a, split_cap);
    CKK();
    LAUNCH(k_suplist, blocks_for(super_cap), DBG_BLOCK, h->stream, ptr<u64>(h->b_link), ptr<const u32>(h->b_super),
           (const u64 *)&d_ctr->n_list_b, super_cap);
    CKK();
    bool deferred_check = false;
    {
        int max_rounds = 2;
        for (uint64_t n = super_cap; n > 1; n >>= 1) max_rounds++;
        if (super_cap <= WALK_BLOCK_JUMP_CAP) {                      // the whole list in one launch
#ifdef DBG_EMU
            LAUNCH(k_jump_block, 1, 1024, h->stream, ptr<u64>(h->b_link), ptr<const u32>(h->b_super),
                   (const u64 *)&d_ctr->n_list_b, super_cap, n_slots, max_rounds, d_ctr);
#else
            LAUNCH(k_jump_cluster, WALK_CLUSTER, 1024, h->stream, ptr<u64>(h->b_link), ptr<const u32>(h->b_super),
                   (const u64 *)&d_ctr->n_list_b, super_cap, n_slots, d_ctr);
#endif
            CKK();
            deferred_check = true;                                   // looked at with k_emit's counters below
        } else {
            TRY(fetch_counters(h));
            const uint64_t n_super = h->ctr.n_list_b < super_cap ? h->ctr.n_list_b : super_cap;
            CK(cudaMemsetAsync(&d_ctr->changed, 0, 4, h->stream));
            for (int r = 0; r < max_rounds && n_super; r++) {        // the list is small: look every 8 rounds
                LAUNCH(k_jump, blocks_for(n_super), DBG_BLOCK, h->stream, ptr<u64>(h->b_link),
                       ptr<const u32>(h->b_super), n_super, n_slots, d_ctr);
                CKK();
                st.jump_rounds++;
                if ((r & 7) == 7 || r == max_rounds - 1) {
                    TRY(check_fault(h, "k_jump"));
                    if (!h->ctr.changed) break;
                    CK(cudaMemsetAsync(&d_ctr->changed, 0, 4, h->stream));
                }
            }
            TRY(check_fault(h, "k_jump"));
        }
    }
    const int e1 = ev_mark_walk(h);

    // ---- one record per start (debruijn.py:331-345), already in dbg_contig_t layout ----
    const uint64_t ns = st.n_starts;
    const dbg_contig_t *recs = nullptr;
    uint64_t n_dropped = 0;
    if (ns) {
        TRY(reserve(h, h->b_contig, ns * sizeof(dbg_contig_t)));
        TRY(reserve_pinned(h, ns * sizeof(dbg_contig_t)));
        CK(cudaMemsetAsync(&d_ctr->n_dropped, 0, 8, h->stream));
        LAUNCH(k_emit, blocks_for(ns), DBG_BLOCK, h->stream, ptr<const u32>(h->b_starts), ns,
               ptr<const u64>(h->b_link), ptr<const u8>(h->b_flags), ptr<ContigRec>(h->b_contig), &d_ctr->n_dropped);
        CKK();
        CK(cudaMemcpyAsync(h->pin, h->b_contig.p, ns * sizeof(dbg_contig_t), cudaMemcpyDeviceToHost, h->stream));
        recs = (const dbg_contig_t *)h->pin;
    }
    TRY(check_fault(h, "the list ranking kernels"));                    // ONE host round trip for the whole walk
    if (deferred_check) st.jump_rounds = h->ctr.changed;
    if (h->ctr.n_list_a > split_cap || h->ctr.n_list_b > super_cap)
        return fail(h, DBG_ERR_STATE, "splitter list overflow (%llu > %llu or %llu > %llu)",
                    (unsigned long long)h->ctr.n_list_a, (unsigned long long)split_cap,
                    (unsigned long long)h->ctr.n_list_b, (unsigned long long)super_cap);
    if (ns) n_dropped = h->ctr.n_dropped;                              // starts on cycles or pulled themselves
    const uint64_t nb = h->branch.size();
    const bool enumerate = !stop_at_branch && nb > 0;
    std::vector<JunctionSucc> junc;
    std::vector<u64> joff;                            // kept successors of branch i: junc[joff[i] .. joff[i + 1])
    std::vector<std::pair<u32, u32>> jindex;           // (slot, index in branch list), sorted by slot
    if (enumerate) {
        joff.assign(nb + 1, 0);
        if (h->generic) {                 // any number of kept successors: their counts first
            std::vector<u32> nk(nb);
            TRY(reserve(h, h->b_cnt32, nb * 4));
            LAUNCH(k_g_gather_u32, blocks_for(nb), DBG_BLOCK, h->stream, ptr<const u32>(h->b_enk), ptr<const u32>(h->b_branch),
                   nb, ptr<u32>(h->b_cnt32));
            CKK();
            CK(cudaMemcpyAsync(nk.data(), h->b_cnt32.p, nb * 4, cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            for (uint64_t i = 0; i < nb; i++) joff[i + 1] = joff[i] + nk[i];
            TRY(reserve(h, h->b_off64, (nb + 1) * 8));
            CK(cudaMemcpyAsync(h->b_off64.p, joff.data(), (nb + 1) * 8, cudaMemcpyHostToDevice, h->stream));
            TRY(reserve(h, h->b_junction, (joff[nb] + 1) * sizeof(JunctionSucc)));
            LAUNCH(k_g_junctions, blocks_for(nb), DBG_BLOCK, h->stream, ptr<const u8>(h->b_flags), ptr<const u32>(h->b_enk),
                   ptr<const u32>(h->b_estart), ptr<const GEdge>(h->b_earr), ptr<const u64>(h->b_link),
                   ptr<const u32>(h->b_branch), nb, ptr<const u64>(h->b_off64), ptr<JunctionSucc>(h->b_junction));
            CKK();
        } else {
            for (uint64_t i = 0; i < nb; i++) joff[i + 1] = joff[i] + 4;          // 4 slots a node, unused ones say NONE32
            TRY(reserve(h, h->b_junction, nb * 4 * sizeof(JunctionSucc)));
            LAUNCH(k_junctions<W>, blocks_for(nb), DBG_BLOCK, h->stream, ptr<const S>(h->b_table), geom_of(h), h->k,
                   ptr<const u8>(h->b_flags), ptr<const u16>(h->b_kept), ptr<const u64>(h->b_link),
                   ptr<const u32>(h->b_branch), nb, ptr<JunctionSucc>(h->b_junction), d_ctr);
            CKK();
        }
        TRY(check_fault(h, "k_junctions"));
        junc.resize(joff[nb]);
        if (joff[nb]) CK(cudaMemcpyAsync(junc.data(), h->b_junction.p, joff[nb] * sizeof(JunctionSucc), cudaMemcpyDeviceToHost,
                                         h->stream));
        CK(cudaStreamSynchronize(h->stream));
        jindex.resize(nb);
        for (uint64_t i = 0; i < nb; i++) jindex[i] = std::make_pair(h->branch[i], (u32)i);
        std::sort(jindex.begin(), jindex.end());
    }
    auto junction_of = [&](u32 slot) -> u32 {
        auto it = std::lower_bound(jindex.begin(), jindex.end(), std::make_pair(slot, (u32)0));
        return it->second;
    };

    uint64_t enum_steps = 0;
    const bool fast = !enumerate && n_dropped == 0;                    // every start emits exactly its record
    for (uint64_t i = 0; i < ns && !fast; i++) {
        const dbg_contig_t &c = recs[i];
        if (c.kind == T_CYCLE) { st.n_cycle_starts++; continue; }      // debruijn.py:289
        if (!(enumerate && c.kind == T_BRANCH)) {
            if (c.n_nodes == 0) continue;                              // the start itself is pulled (:293-295)
            dbg_contig_t d;
            d.start = c.start; d.n_nodes = c.n_nodes; d.terminal = c.terminal; d.kind = c.kind;
            d.seg_begin = d.seg_end = 0;
            h->contigs.push_back(d);
            continue;
        }
        // final=True: the walk continues through multi-successor nodes (debruijn.py:314), one
        // contig per distinct root-to-stop path, paths that revisit a node emit nothing (:289).
        const uint64_t first_contig = h->contigs.size();
        std::vector<dbg_segment_t> segs;                               // current path as chain segments
        std::vector<u32> on_path;                                      // junction indices on the path
        std::vector<EnumFrame> stack;
        uint64_t total_nodes = c.n_nodes;
        segs.push_back(dbg_segment_t{c.start, c.n_nodes});
        u32 j0 = junction_of(c.terminal);
        on_path.push_back(j0);
        stack.push_back(EnumFrame{j0, 0});
        std::vector<uint64_t> node_stack;                              // total_nodes before each pushed segment
        while (!stack.empty()) {
            if (++enum_steps > MAX_ENUM_STEPS)
                return fail(h, DBG_ERR_OVERFLOW, "final=True path enumeration exceeded %llu steps (it is "
                            "exponential in the reference too, debruijn.py:314)", (unsigned long long)MAX_ENUM_STEPS);
            EnumFrame &fr = stack.back();
            const u32 jv = fr.junction_index;
            if ((u64)fr.next_child >= joff[jv + 1] - joff[jv] || junc[joff[jv] + fr.next_child].entry == NONE32) {
                stack.pop_back();
                on_path.pop_back();
                if (!node_stack.empty()) { total_nodes = node_stack.back(); node_stack.pop_back(); }
                segs.pop_back();
                continue;
            }
            const JunctionSucc js = junc[joff[jv] + fr.next_child];
            fr.next_child++;
            if (js.kind == T_CYCLE) continue;
            if (js.kind == T_BRANCH) {
                const u32 jn = junction_of(js.terminal);
                if (std::find(on_path.begin(), on_path.end(), jn) != on_path.end()) continue;   // revisit
                node_stack.push_back(total_nodes);
                segs.push_back(dbg_segment_t{js.entry, js.n_hops + 1});
                total_nodes += js.n_hops + 1;
                on_path.push_back(jn);
                stack.push_back(EnumFrame{jn, 0});
                continue;
            }
            // the path stops: at a dead end (terminal included) or in front of a pulled node
            const u32 add = js.kind == T_PULLED ? js.n_hops : js.n_hops + 1;
            dbg_contig_t d;
            d.start = c.start; d.terminal = js.terminal; d.kind = js.kind;
            d.n_nodes = (u32)(total_nodes + add);
            // per-start dedup of identical paths (debruijn.py:297): only paths that stop directly in
            // front of pulled successors of the same junction can coincide
            bool dup = false;
            if (add == 0) {
                for (uint64_t q = first_contig; q < h->contigs.size() && !dup; q++) {
                    const dbg_contig_t &o = h->contigs[q];
                    if (o.n_nodes != d.n_nodes || o.seg_end - o.seg_begin != segs.size()) continue;
                    bool same = true;
                    for (size_t s = 0; s < segs.size() && same; s++)
                        same = h->segments[o.seg_begin + s].entry == segs[s].entry &&
                               h->segments[o.seg_begin + s].n_nodes == segs[s].n_nodes;
                    dup = same;
                }
            }
            if (dup) continue;
            d.seg_begin = h->segments.size();
            for (const dbg_segment_t &s : segs) h->segments.push_back(s);
            if (add) h->segments.push_back(dbg_segment_t{js.entry, add});
            d.seg_end = h->segments.size();
            h->contigs.push_back(d);
            if (h->contigs.size() > MAX_ENUM_CONTIGS)
                return fail(h, DBG_ERR_OVERFLOW, "final=True path enumeration exceeded %llu contigs",
                            (unsigned long long)MAX_ENUM_CONTIGS);
        }
    }
    if (fast) { h->cview = recs; h->n_cview = ns; }
    else { h->cview = h->contigs.data(); h->n_cview = h->contigs.size(); }
    st.n_contigs = h->n_cview;
    const int e2 = ev_mark_walk(h);
    tm_set(h, &st.ms_walk_rank, e0, e1);
    tm_set(h, &st.ms_walk_emit, e1, e2);
    tm_set(h, &st.ms_walk_total, e0, e2);
    return DBG_OK;
}

// -------------------------------------------------------------------------------------------------
// sharded node table: owner-computes contig walk (dbg_shard.cuh)
// -------------------------------------------------------------------------------------------------
template <int W>
static int shard_exit_counts_impl(dbg_t *h, uint64_t *counts) {
    const uint32_t P = h->shard_n;
    const ChunkGeom geom = chunk_geom(h->k);
    TRY(reserve(h, h->b_off64, (uint64_t)P * 8 * 2));
    CK(cudaMemsetAsync(h->b_off64.p, 0, (uint64_t)P * 8 * 2, h->stream));
    LAUNCH(k_exit_count<W>, blocks_for(h->n_slots), DBG_BLOCK, h->stream, ptr<const Slot<W>>(h->b_table), h->n_slots,
           ptr<const u8>(h->b_flags), ptr<const u16>(h->b_kept), h->k, geom.w, geom.m, P, ptr<u64>(h->b_off64));
    CKK();
    CK(cudaMemcpyAsync(counts, h->b_off64.p, (uint64_t)P * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->exit_off.assign(P + 1, 0);
    for (uint32_t p = 0; p < P; p++) h->exit_off[p + 1] = h->exit_off[p] + counts[p];
    h->n_exit = h->exit_off[P];
    h->shard_stage = 1;
    return DBG_OK;
}

template <int W>
static int shard_exit_keys_impl(dbg_t *h, uint64_t *d_keys_out) {
    const uint32_t P = h->shard_n;
    const ChunkGeom geom = chunk_geom(h->k);
    TRY(reserve(h, h->b_exitslot, (h->n_exit + 1) * 4));
    u64 *d_cur = ptr<u64>(h->b_off64) + P;
    CK(cudaMemcpyAsync(d_cur, h->exit_off.data(), (uint64_t)P * 8, cudaMemcpyHostToDevice, h->stream));
    if (h->n_exit) {
        LAUNCH(k_exit_scatter<W>, blocks_for(h->n_slots), DBG_BLOCK, h->stream, ptr<const Slot<W>>(h->b_table), h->n_slots,
               ptr<const u8>(h->b_flags), ptr<const u16>(h->b_kept), h->k, geom.w, geom.m, P, d_cur,
               ptr<u32>(h->b_exitslot), (u64 *)d_keys_out);
        CKK();
    }
    SYNC_FOR_CALLER();
    return DBG_OK;
}

template <int W>
static int shard_mark_entries_impl(dbg_t *h, const uint64_t *d_keys, uint64_t n) {
    TRY(reserve(h, h->b_reqslot, (n + 1) * 4));
    h->n_req = n;
    if (n) {
        LAUNCH(k_mark_entries<W>, blocks_for(n), DBG_BLOCK, h->stream, ptr<const Slot<W>>(h->b_table), geom_of(h),
               (const u64 *)d_keys, n, ptr<u8>(h->b_flags), ptr<u32>(h->b_reqslot), ptr<DevCounters>(h->b_ctr));
        CKK();
    }
    h->shard_stage = 2;
    return DBG_OK;
}

// list ranking of chain nodes -> splitters -> super-splitters on link[] (shared by the local and the reduced ranking)
static int rank_links(dbg_t *h, u64 *d_link, uint64_t n_ids, uint64_t split_cap, uint64_t super_cap) {
    DevCounters *d_ctr = ptr<DevCounters>(h->b_ctr);
    LAUNCH(k_sublist, blocks_for(split_cap), DBG_BLOCK, h->stream, d_link, ptr<const u32>(h->b_split),
           (const u64 *)&d_ctr->n_list_a, split_cap);
    CKK();
    LAUNCH(k_suplist, blocks_for(super_cap), DBG_BLOCK, h->stream, d_link, ptr<const u32>(h->b_super),
           (const u64 *)&d_ctr->n_list_b, super_cap);
    CKK();
    int max_rounds = 2;
    for (uint64_t n = super_cap; n > 1; n >>= 1) max_rounds++;
    if (super_cap <= WALK_BLOCK_JUMP_CAP) {
#ifdef DBG_EMU
        LAUNCH(k_jump_block, 1, 1024, h->stream, d_link, ptr<const u32>(h->b_super), (const u64 *)&d_ctr->n_list_b,
               super_cap, n_ids, max_rounds, d_ctr);
#else
        LAUNCH(k_jump_cluster, WALK_CLUSTER, 1024, h->stream, d_link, ptr<const u32>(h->b_super),
               (const u64 *)&d_ctr->n_list_b, super_cap, n_ids, d_ctr);
#endif
        CKK();
    } else {
        TRY(fetch_counters(h));
        const uint64_t n_super = h->ctr.n_list_b < super_cap ? h->ctr.n_list_b : super_cap;
        CK(cudaMemsetAsync(&d_ctr->changed, 0, 4, h->stream));
        for (int r = 0; r < max_rounds && n_super; r++) {
            LAUNCH(k_jump, blocks_for(n_super), DBG_BLOCK, h->stream, d_link, ptr<const u32>(h->b_super), n_super, n_ids, d_ctr);
            CKK();
            if ((r & 7) == 7 || r == max_rounds - 1) {
                TRY(check_fault(h, "k_jump"));
                if (!h->ctr.changed) break;
                CK(cudaMemsetAsync(&d_ctr->changed, 0, 4, h->stream));
            }
        }
    }
    return DBG_OK;
}

template <int W>
static int shard_rank_local_impl(dbg_t *h, uint64_t *n_reduced) {
    dbg_stats_t &st = h->st;
    const uint64_t n_slots = h->n_slots;
    DevCounters *d_ctr = ptr<DevCounters>(h->b_ctr);
    h->n_ev_walk = EV_WALK;
    const int e0 = ev_mark_walk(h);
    TRY(reserve(h, h->b_link, n_slots * 8));
    // splitters: starts, entries (at most one per request received) and the hash sample
    const uint64_t split_cap = (n_slots >> WALK_SPLIT_LOG2) * 3 / 2 + st.n_starts + h->n_req + 4096;
    const uint64_t super_cap = (split_cap >> WALK_SUPER_LOG2) * 3 / 2 + 4096;
    TRY(reserve(h, h->b_split, split_cap * 4));
    TRY(reserve(h, h->b_super, super_cap * 4));
    CK(cudaMemsetAsync(&d_ctr->n_list_a, 0, 16, h->stream));
    LAUNCH(k_link_init, blocks_for(n_slots), DBG_BLOCK, h->stream, ptr<const u8>(h->b_flags),
           ptr<const u16>(h->b_kept), ptr<const u32>(h->b_succ), n_slots, ptr<u64>(h->b_link), ptr<u32>(h->b_split),
           split_cap, &d_ctr->n_list_a, ptr<u32>(h->b_super), super_cap, &d_ctr->n_list_b);
    CKK();
    TRY(rank_links(h, ptr<u64>(h->b_link), n_slots, split_cap, super_cap));
    // reduced vertices of this shard: entries + starts
    const uint64_t rv_cap = st.n_starts + h->n_req + 1;
    TRY(reserve(h, h->b_rvslot, rv_cap * 4));
    TRY(reserve(h, h->b_rvord, n_slots * 4));
    CK(cudaMemsetAsync(&d_ctr->n_dropped, 0, 8, h->stream));
    LAUNCH(k_rv_collect, blocks_for(n_slots), DBG_BLOCK, h->stream, ptr<const u8>(h->b_flags), n_slots,
           ptr<u32>(h->b_rvslot), ptr<u32>(h->b_rvord), &d_ctr->n_dropped);
    CKK();
    TRY(check_fault(h, "the local list ranking of the shard"));
    if (h->ctr.n_list_a > split_cap || h->ctr.n_list_b > super_cap)
        return fail(h, DBG_ERR_STATE, "splitter list overflow in the sharded walk");
    h->n_rv = h->ctr.n_dropped;
    if (h->n_rv > rv_cap) return fail(h, DBG_ERR_STATE, "reduced vertex list overflow");
    *n_reduced = h->n_rv;
    const int e1 = ev_mark_walk(h);
    tm_set(h, &st.ms_walk_rank, e0, e1);
    h->shard_stage = 3;
    return DBG_OK;
}

static int shard_entry_ordinals_impl(dbg_t *h, uint32_t *d_out) {
    if (h->n_req) {
        LAUNCH(k_entry_ordinals, blocks_for(h->n_req), DBG_BLOCK, h->stream, ptr<const u32>(h->b_reqslot), h->n_req,
               ptr<const u32>(h->b_rvord), (u32 *)d_out);
        CKK();
    }
    SYNC_FOR_CALLER();
    return DBG_OK;
}

static int shard_reduced_records_impl(dbg_t *h, const uint32_t *d_replies, const uint64_t *rid_base, uint64_t *d_out) {
    const uint32_t P = h->shard_n;
    TRY(reserve(h, h->b_owneroff, (uint64_t)(2 * P + 2) * 8));
    u64 *d_off = ptr<u64>(h->b_owneroff), *d_base = d_off + P + 1;
    CK(cudaMemcpyAsync(d_off, h->exit_off.data(), (uint64_t)(P + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_base, rid_base, (uint64_t)P * 8, cudaMemcpyHostToDevice, h->stream));
    if (h->n_exit) {
        LAUNCH(k_exit_store, blocks_for(h->n_exit), DBG_BLOCK, h->stream, ptr<const u32>(h->b_exitslot), h->n_exit,
               (const u64 *)d_off, P, (const u64 *)d_base, (const u32 *)d_replies, ptr<u32>(h->b_succ));
        CKK();
    }
    if (h->n_rv) {
        LAUNCH(k_reduced_build, blocks_for(h->n_rv), DBG_BLOCK, h->stream, ptr<const u32>(h->b_rvslot), h->n_rv,
               ptr<const u64>(h->b_link), ptr<const u8>(h->b_flags), ptr<const u16>(h->b_kept), ptr<const u32>(h->b_succ),
               h->shard_rank, (u64 *)d_out);
        CKK();
    }
    SYNC_FOR_CALLER();                                    // (rid_base: pageable H2D sources are staged before the call returns)
    h->shard_stage = 4;
    return DBG_OK;
}

template <int W>
static int shard_walk_impl(dbg_t *h, const uint64_t *d_all, uint64_t n_total, uint64_t my_base, uint64_t *n_contigs) {
    dbg_stats_t &st = h->st;
    DevCounters *d_ctr = ptr<DevCounters>(h->b_ctr);
    const int e0 = ev_mark_walk(h);
    h->scontigs.clear();
    h->sview = nullptr; h->n_sview = 0;
    *n_contigs = 0;
    if (n_total >= 0xFFFFFFF0ull) return fail(h, DBG_ERR_OVERFLOW, "the reduced list exceeds 2^32 vertices");
    TRY(reserve(h, h->b_link2, (n_total + 1) * 8));
    const uint64_t split_cap = (n_total >> WALK_SPLIT_LOG2) * 3 / 2 + h->n_rv + 4096;
    const uint64_t super_cap = (split_cap >> WALK_SUPER_LOG2) * 3 / 2 + 4096;
    TRY(reserve(h, h->b_split, split_cap * 4));
    TRY(reserve(h, h->b_super, super_cap * 4));
    CK(cudaMemsetAsync(&d_ctr->n_list_a, 0, 16, h->stream));
    LAUNCH(k_rlink_init, blocks_for(n_total), DBG_BLOCK, h->stream, (const u64 *)d_all, n_total, my_base, h->n_rv,
           ptr<u64>(h->b_link2), ptr<u32>(h->b_split), split_cap, &d_ctr->n_list_a, ptr<u32>(h->b_super), super_cap,
           &d_ctr->n_list_b, d_ctr);
    CKK();
    TRY(rank_links(h, ptr<u64>(h->b_link2), n_total, split_cap, super_cap));
    const uint64_t ns = st.n_starts;
    if (ns) {
        TRY(reserve(h, h->b_scontig, ns * sizeof(ShardContig)));
        CK(cudaMemsetAsync(&d_ctr->n_dropped, 0, 8, h->stream));
        LAUNCH(k_emit_reduced<W>, blocks_for(ns), DBG_BLOCK, h->stream, ptr<const Slot<W>>(h->b_table),
               ptr<const u32>(h->b_starts), ns, ptr<const u32>(h->b_rvord), my_base, h->shard_rank,
               ptr<const u64>(h->b_link2), (const u64 *)d_all, ptr<ShardContig>(h->b_scontig), &d_ctr->n_dropped);
        CKK();
    }
    if (ns) {
        TRY(reserve_pinned(h, ns * sizeof(ShardContig)));
        CK(cudaMemcpyAsync(h->pin, h->b_scontig.p, ns * sizeof(ShardContig), cudaMemcpyDeviceToHost, h->stream));
    }
    TRY(check_fault(h, "the ranking of the reduced list"));            // one host round trip: counters + records
    if (h->ctr.n_list_a > split_cap || h->ctr.n_list_b > super_cap)
        return fail(h, DBG_ERR_STATE, "splitter list overflow in the reduced walk");
    const uint64_t n_drop = ns ? h->ctr.n_dropped : 0;
    const dbg_shard_contig_t *recs = (const dbg_shard_contig_t *)h->pin;
    if (n_drop == 0) { h->sview = recs; h->n_sview = ns; }
    else {                                                              // starts on cycles / pulled starts emit nothing
        h->scontigs.reserve(ns - n_drop);
        for (uint64_t i = 0; i < ns; i++) if (recs[i].n_nodes) h->scontigs.push_back(recs[i]);
        h->sview = h->scontigs.data(); h->n_sview = h->scontigs.size();
    }
    const uint64_t nc = h->n_sview;
    st.n_contigs = nc;
    st.n_cycle_starts = 0;
    *n_contigs = nc;
    const int e1 = ev_mark_walk(h);
    tm_set(h, &st.ms_walk_emit, e0, e1);
    h->shard_stage = 5;
    return DBG_OK;
}

// -------------------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------------------
#define DISPATCH(fn, ...) (h->W == 1 ? fn<1>(__VA_ARGS__) : fn<2>(__VA_ARGS__))

#define DBG_API extern "C" __attribute__((visibility("default")))

DBG_API int dbg_version(void) { return DBG_ABI_VERSION; }

DBG_API const char *dbg_last_error(const dbg_t *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

DBG_API int dbg_create(dbg_t **out, int device, int k) { return dbg_create_ex(out, device, k, 0); }

DBG_API int dbg_create_ex(dbg_t **out, int device, int k, int key_words) {
    dbg_t *h = nullptr;
    if (!out) return fail(nullptr, DBG_ERR_ARG, "dbg_create: out is NULL");
    *out = nullptr;
    if (k < 1 || k > 63) return fail(nullptr, DBG_ERR_ARG, "k = %d out of range (1..63)", k);
    if (key_words < 0 || key_words > 2 || (key_words == 1 && k > 31))
        return fail(nullptr, DBG_ERR_ARG, "key_words = %d is not possible for k = %d", key_words, k);
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(nullptr, DBG_ERR_CUDA, "no CUDA device: %s (this library has no CPU path)",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n_dev) return fail(nullptr, DBG_ERR_ARG, "device %d out of range (0..%d)", device, n_dev - 1);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, DBG_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    h = new dbg_handle();
    h->device = device;
    h->k = k;
    h->W = key_words ? key_words : (k <= 31 ? 1 : 2);
    memset(&h->st, 0, sizeof h->st);
    memset(&h->ctr, 0, sizeof h->ctr);
    h->st.key_words = (uint32_t)h->W;
    e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete h; return fail(nullptr, DBG_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    h->stream = h->own_stream;
    for (int i = 0; i < EV_MAX; i++) cudaEventCreate(&h->ev[i]);
    for (int i = 0; i <= MAX_SLICES; i++) cudaEventCreate(&h->slice_ev[i]);
    cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
#ifdef DBG_EMU
    cudaStreamCreateWithFlags(&h->init_stream, cudaStreamNonBlocking);
#else
    {   // background work (table clears) at the LOWEST priority: blocks of the main stream's kernels go first
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&h->init_stream, cudaStreamNonBlocking, lo) != cudaSuccess)
            cudaStreamCreateWithFlags(&h->init_stream, cudaStreamNonBlocking);
    }
#endif
    for (int i = 0; i < 3; i++) cudaEventCreate(&h->ev_init[i]);
    if (cudaMallocHost(&h->ctl, CTL_BYTES) != cudaSuccess) h->ctl = nullptr;
    if (reserve(h, h->b_ctr, sizeof(DevCounters)) != DBG_OK) {
        g_create_error = h->err;
        dbg_destroy(h);
        return DBG_ERR_NOMEM;
    }
    *out = h;
    return DBG_OK;
}

DBG_API int dbg_create_alpha(dbg_t **out, int device, int k, const uint8_t *symbols, int n_symbols) {
    if (!out) return fail(nullptr, DBG_ERR_ARG, "dbg_create_alpha: out is NULL");
    *out = nullptr;
    if (!symbols || n_symbols < 1 || n_symbols > GEN_MAX_SYMBOLS)
        return fail(nullptr, DBG_ERR_ARG, "an alphabet has 1 .. %d symbols (got %d)", GEN_MAX_SYMBOLS, n_symbols);
    const int gb = gen_bits_for(n_symbols);
    if (k < 1 || k > gen_max_k(gb))
        return fail(nullptr, DBG_ERR_ARG, "k = %d out of range for an alphabet of %d symbols (%d-bit codes: 1..%d)", k, n_symbols,
                    gb, gen_max_k(gb));
    uint8_t code_of[256];
    memset(code_of, 0xFF, sizeof code_of);
    for (int i = 0; i < n_symbols; i++) {
        if (code_of[symbols[i]] != 0xFF) return fail(nullptr, DBG_ERR_ARG, "symbol %d appears twice in the alphabet", (int)symbols[i]);
        code_of[symbols[i]] = (uint8_t)i;
    }
    const int rc = dbg_create_ex(out, device, 1, 2);                   // 128-bit keys; k is set below (gb bits a symbol)
    if (rc != DBG_OK) return rc;
    dbg_t *h = *out;
    h->k = k;
    h->generic = true;
    h->sym_bits = gb;
    h->n_symbols = n_symbols;
    memset(h->sym_of_code, '?', sizeof h->sym_of_code);
    memcpy(h->sym_of_code, symbols, (size_t)n_symbols);
    h->sym_of_code[256] = (uint8_t)gb;
    memcpy(h->code_of_byte, code_of, sizeof code_of);
    return DBG_OK;
}

DBG_API void dbg_destroy(dbg_t *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    DevBuf *all[] = {&h->b_bases, &h->b_off, &h->b_packed, &h->b_tile, &h->b_table, &h->b_flags, &h->b_kept,
                     &h->b_succ, &h->b_link, &h->b_ctr, &h->b_hll, &h->b_order, &h->b_tag, &h->b_tag2, &h->b_slot,
                     &h->b_slot2, &h->b_sorttmp, &h->b_branch, &h->b_starts, &h->b_pulled, &h->b_readflags,
                     &h->b_recs, &h->b_cnt32, &h->b_off64, &h->b_contig, &h->b_junction, &h->b_split, &h->b_h8, &h->b_bitmap,
                     &h->b_stmask, &h->b_vmask, &h->b_spanread, &h->b_cc, &h->b_ctable, &h->b_clist, &h->b_super, &h->b_readbatch, &h->b_ctable2, &h->b_scratch[0],
                     &h->b_scratch[1], &h->b_scratch[2], &h->b_scratch[3], &h->b_scratch[4], &h->b_scratch[5],
                     &h->b_up, &h->b_codes, &h->b_etable, &h->b_earr, &h->b_earr2, &h->b_estart, &h->b_edeg, &h->b_enk, &h->b_lut,
                     &h->b_exitslot, &h->b_reqslot, &h->b_rvslot, &h->b_rvord, &h->b_link2, &h->b_scontig, &h->b_owneroff};
    for (DevBuf *b : all) if (b->p) cudaFree(b->p);
    if (h->b_bases2.p) cudaFree(h->b_bases2.p);
    if (h->b_off2.p) cudaFree(h->b_off2.p);
    concat_free(h);
    for (int i = 0; i < EV_MAX; i++) cudaEventDestroy(h->ev[i]);
    for (int i = 0; i <= MAX_SLICES; i++) cudaEventDestroy(h->slice_ev[i]);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->init_stream) cudaStreamDestroy(h->init_stream);
    for (int i = 0; i < 3; i++) cudaEventDestroy(h->ev_init[i]);
#ifdef DBG_EMU
    if (h->pin) free(h->pin);
#else
    if (h->pin) cudaFreeHost(h->pin);
#endif
    if (h->ctl) cudaFreeHost(h->ctl);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

// findSupportReadScore (IV_sortOutputs.py:10-15) for many contigs at once; kernels in dbg_support.cuh
DBG_API int dbg_support_scores(dbg_t *h, const uint8_t *reads, const uint64_t *read_off, uint64_t n_reads,
                               const double *read_score, const uint8_t *contigs, const uint64_t *contig_off,
                               uint64_t n_contigs, double *out_scores) {
    if (!h || !read_off || !contig_off || (n_contigs && !out_scores)) return h ? fail(h, DBG_ERR_ARG, "null argument") : DBG_ERR_ARG;
    cudaSetDevice(h->device);
    if (n_contigs == 0) return DBG_OK;
    const uint64_t rbytes = read_off[n_reads], cbytes = contig_off[n_contigs];
    if (n_reads >= 0xFFFFFFF0ull || n_contigs >= 0xFFFFFFF0ull) return fail(h, DBG_ERR_OVERFLOW, "more than 2^32 reads or contigs");
    // distinct non-empty read lengths, ascending, with BASE^len
    std::vector<u32> lens;
    uint64_t n_empty = 0;
    for (uint64_t r = 0; r < n_reads; r++) {
        const uint64_t l = read_off[r + 1] - read_off[r];
        if (l > 0xFFFFFFFFull) return fail(h, DBG_ERR_OVERFLOW, "a read is longer than 2^32-1");
        if (l == 0) n_empty++; else lens.push_back((u32)l);
    }
    std::sort(lens.begin(), lens.end());
    lens.erase(std::unique(lens.begin(), lens.end()), lens.end());
    std::vector<u64> pows(lens.size());
    for (size_t i = 0; i < lens.size(); i++) { u64 p = 1; for (u32 j = 0; j < lens[i]; j++) p *= SUP_BASE; pows[i] = p; }
    double empty_sum = 0.0;                                              // '' in contig is always True (handled in read order below)
    (void)empty_sum;
    const uint64_t cap = 2 * n_reads + 16;
    DevBuf &b_r = h->b_scratch[0], &b_c = h->b_scratch[1], &b_off = h->b_scratch[2], &b_tab = h->b_scratch[3],
           &b_pre = h->b_scratch[4], &b_misc = h->b_scratch[5];
    TRY(reserve(h, b_r, rbytes + 16));
    TRY(reserve(h, b_c, cbytes + 16));
    TRY(reserve(h, b_off, (n_reads + n_contigs + 2) * 8));
    TRY(reserve(h, b_tab, cap * sizeof(SupSlot)));
    TRY(reserve(h, b_pre, (cbytes + n_contigs + 1) * 8));
    // misc: lens (u32, padded to 8) | pows | scores | out | cursor
    const size_t o_lens = 0, o_pows = ((lens.size() * 4 + 7) / 8) * 8, o_sc = o_pows + pows.size() * 8,
                 o_out = o_sc + n_reads * 8, o_cur = o_out + n_contigs * 8;
    TRY(reserve(h, b_misc, o_cur + 16));
    char *misc = (char *)b_misc.p;
    u64 *d_roff = ptr<u64>(b_off), *d_coff = ptr<u64>(b_off) + n_reads + 1;
    if (rbytes) CK(cudaMemcpyAsync(b_r.p, reads, rbytes, cudaMemcpyHostToDevice, h->stream));
    if (cbytes) CK(cudaMemcpyAsync(b_c.p, contigs, cbytes, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_roff, read_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_coff, contig_off, (n_contigs + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    if (!lens.empty()) {
        CK(cudaMemcpyAsync(misc + o_lens, lens.data(), lens.size() * 4, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(misc + o_pows, pows.data(), pows.size() * 8, cudaMemcpyHostToDevice, h->stream));
    }
    if (n_reads) CK(cudaMemcpyAsync(misc + o_sc, read_score, n_reads * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(b_tab.p, 0xFF, cap * sizeof(SupSlot), h->stream));
    CK(cudaMemsetAsync(misc + o_cur, 0, 8, h->stream));
    if (n_reads) {
        LAUNCH(k_sup_reads, blocks_for(n_reads), DBG_BLOCK, h->stream, ptr<const u8>(b_r), (const u64 *)d_roff, n_reads,
               ptr<SupSlot>(b_tab), cap);
        CKK();
    }
    LAUNCH(k_sup_prefix, blocks_for(n_contigs), DBG_BLOCK, h->stream, ptr<const u8>(b_c), (const u64 *)d_coff, n_contigs,
           ptr<u64>(b_pre));
    CKK();
    uint64_t n_hits = 0;
    DevBuf &b_hits = h->b_tag, &b_hits2 = h->b_tag2;
    if (cbytes && !lens.empty()) {
        LAUNCH(k_sup_match, blocks_for(cbytes), DBG_BLOCK, h->stream, ptr<const u8>(b_c), (const u64 *)d_coff, n_contigs,
               ptr<const u64>(b_pre), ptr<const u8>(b_r), (const u64 *)d_roff, ptr<const SupSlot>(b_tab), cap,
               (const u32 *)(misc + o_lens), (const u64 *)(misc + o_pows), (u32)lens.size(), 0, (u64 *)(misc + o_cur),
               (u64 *)nullptr, (u64)0);
        CKK();
        CK(cudaMemcpyAsync(&n_hits, misc + o_cur, 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    // reads of length 0 are in every contig (`'' in contig`): one hit per (contig, empty read), made on the host side
    // of the list so that they take their place in read order
    std::vector<u64> empty_hits;
    if (n_empty)
        for (uint64_t c = 0; c < n_contigs; c++)
            for (uint64_t r = 0; r < n_reads; r++)
                if (read_off[r + 1] == read_off[r]) empty_hits.push_back((c << 32) | r);
    const uint64_t n_all = n_hits + empty_hits.size();
    TRY(reserve(h, b_hits, (n_all + 1) * 8));
    TRY(reserve(h, b_hits2, (n_all + 1) * 8));
    if (n_hits) {
        CK(cudaMemsetAsync(misc + o_cur, 0, 8, h->stream));
        LAUNCH(k_sup_match, blocks_for(cbytes), DBG_BLOCK, h->stream, ptr<const u8>(b_c), (const u64 *)d_coff, n_contigs,
               ptr<const u64>(b_pre), ptr<const u8>(b_r), (const u64 *)d_roff, ptr<const SupSlot>(b_tab), cap,
               (const u32 *)(misc + o_lens), (const u64 *)(misc + o_pows), (u32)lens.size(), 1, (u64 *)(misc + o_cur),
               ptr<u64>(b_hits), n_hits);
        CKK();
    }
    if (!empty_hits.empty())
        CK(cudaMemcpyAsync(ptr<u64>(b_hits) + n_hits, empty_hits.data(), empty_hits.size() * 8, cudaMemcpyHostToDevice, h->stream));
    const u64 *d_sorted = ptr<u64>(b_hits);
    if (n_all > 1) {
#ifdef DBG_EMU
        std::sort(ptr<u64>(b_hits), ptr<u64>(b_hits) + n_all);
#else
        size_t tmp = 0;
        CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp, ptr<u64>(b_hits), ptr<u64>(b_hits2), (int64_t)n_all, 0, 64, h->stream));
        TRY(reserve(h, h->b_sorttmp, tmp));
        h->st.n_launches++;
        CK(cub::DeviceRadixSort::SortKeys(h->b_sorttmp.p, tmp, ptr<u64>(b_hits), ptr<u64>(b_hits2), (int64_t)n_all, 0, 64, h->stream));
        d_sorted = ptr<u64>(b_hits2);
#endif
    }
    LAUNCH(k_sup_sum, blocks_for(n_contigs), DBG_BLOCK, h->stream, d_sorted, n_all, (const double *)(misc + o_sc), n_contigs,
           (double *)(misc + o_out));
    CKK();
    CK(cudaMemcpyAsync(out_scores, misc + o_out, n_contigs * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DBG_OK;
}

DBG_API int dbg_set_min_count(dbg_t *h, uint32_t min_count) {
    if (!h) return DBG_ERR_ARG;
    if (h->generic && min_count > 1) return fail(h, DBG_ERR_ARG, "min_count exists for the DNA engine only");
    h->min_count = min_count < 1 ? 1 : min_count;
    return DBG_OK;
}

DBG_API int dbg_set_stream(dbg_t *h, void *cuda_stream) {
    if (!h) return DBG_ERR_ARG;
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    h->ext_stream = cuda_stream != nullptr;
    return DBG_OK;
}

static int build_device_common(dbg_t *h, const uint8_t *d_bases, const uint64_t *d_read_off, uint64_t n_reads,
                               uint64_t n_bases, uint64_t read_idx_base, double threshold, uint32_t flags,
                               const uint32_t *d_read_batch, uint32_t n_batches) {
    if (!h) return DBG_ERR_ARG;
    if ((n_bases && !d_bases) || !d_read_off) return fail(h, DBG_ERR_ARG, "dbg_build_device: NULL input");
    h->d_read_batch = nullptr;
    h->batch_bits = 0;
    h->n_batches = 0;
    if (d_read_batch) {
        // batch ids live in the key bits above the k-mer; the all-ones key stays free for EMPTY
        int bb = 0;
        while (bb < 32 && ((1ull << bb) - 1) < (uint64_t)n_batches) bb++;
        if (((1ull << bb) - 1) < (uint64_t)n_batches || 2 * h->k + bb > 64 * h->W)
            return fail(h, DBG_ERR_ARG, "%u batches need %d key bits above the %d bits of a %d-mer: more than the %d-bit key has",
                        n_batches, bb, 2 * h->k, h->k, 64 * h->W);
        if (flags & (DBG_TABLE_ONLY | DBG_CHUNKS_ONLY)) return fail(h, DBG_ERR_ARG, "a batched build cannot be DBG_TABLE_ONLY / DBG_CHUNKS_ONLY");
        h->d_read_batch = d_read_batch;
        h->batch_bits = bb;
        h->n_batches = n_batches;
    }
    if (n_reads + read_idx_base >= 0xFFFFFFFFull)
        return fail(h, DBG_ERR_OVERFLOW, "read index exceeds 32 bits");
    if (threshold != threshold) return fail(h, DBG_ERR_ARG, "threshold is NaN");
    CK(cudaSetDevice(h->device));
    if (!h->tm.empty()) { CK(cudaStreamSynchronize(h->stream)); tm_flush(h); }   // a previous build's stage times
    h->expand_pending = false;
    h->tags_local = true;
    h->chunks_valid = false;
    h->chunks_merged = false;
    h->built = h->walked = false;
    h->up_ready = false;
    h->n_ev = 0;
    h->n_reads = n_reads;
    h->n_bases = n_bases;
    h->read_idx_base = read_idx_base;
    h->threshold = threshold;
    h->build_flags = flags;
    h->d_off = d_read_off;
    memset(&h->st, 0, sizeof h->st);
    h->st.ms_h2d = h->pending_h2d;
    h->pending_h2d = 0;
    h->st.key_words = (uint32_t)h->W;
    h->st.n_reads = n_reads;
    h->st.n_bases = n_bases;
    h->contigs.clear();
    h->segments.clear();
    CK(cudaMemsetAsync(h->b_ctr.p, 0, sizeof(DevCounters), h->stream));
    h->table_ready = false;
    if (h->generic) {
        if (d_read_batch || (flags & (DBG_TABLE_ONLY | DBG_CHUNKS_ONLY)))
            return fail(h, DBG_ERR_ARG, "batched and multi-GPU builds exist for the DNA engine only");
        return gen_build(h, (const u8 *)d_bases, (const u64 *)d_read_off);
    }
    // sizes remembered from the last build fit a read set of about the same size (checked afterwards, see dbg_handle::hint)
    h->hinted = h->hint.valid && h->ctl && n_bases > 0 && !(flags & DBG_TABLE_ONLY) && h->hint.chunks > 0 &&
                ((flags & DBG_CHUNKS_ONLY) || h->hint.nodes > 0) &&
                h->hint.n_batches == h->n_batches && n_bases >= h->hint.n_bases / 2 && n_bases <= h->hint.n_bases + h->hint.n_bases / 2 &&
                !getenv("DBG_B200_NO_HINT") && !getenv("DBG_B200_CHUNK_SLOTS") && !getenv("DBG_B200_NODE_SLOTS");
    if (h->min_count > 1) {
        if (flags & (DBG_TABLE_ONLY | DBG_CHUNKS_ONLY)) return fail(h, DBG_ERR_ARG, "min_count > 1 needs the whole graph stage on one handle");
        h->hinted = false;                                              // (the filter stage wants a settled table first)
    }
    h->preinit = false;
    h->have_hbytes = false; h->have_vhash = false; h->have_order = false;
    h->region_mode = want_regions(h);
    h->region_slots = 0;
    int rc = DISPATCH(build_impl, h, (const u8 *)d_bases, (const u64 *)d_read_off);
    if (rc != DBG_OK && rc != DBG_RETRY_COLD) return rc;
    const bool stop_early = (flags & (DBG_TABLE_ONLY | DBG_CHUNKS_ONLY)) != 0;
    if (rc == DBG_OK && stop_early) return DBG_OK;
    if (rc == DBG_OK && h->min_count > 1) rc = DISPATCH(min_count_stage, h);
    if (rc == DBG_OK) rc = DISPATCH(graph_finalize, h);
    if (rc == DBG_RETRY_COLD) {                                       // the remembered sizes did not fit this read set
        h->hint.valid = false;
        h->hinted = false;
        h->n_ev = 0;
        h->tm.clear();
        memset(&h->st, 0, sizeof h->st);
        h->st.key_words = (uint32_t)h->W;
        h->st.n_reads = n_reads;
        h->st.n_bases = n_bases;
        CK(cudaMemsetAsync(h->b_ctr.p, 0, sizeof(DevCounters), h->stream));
        h->table_ready = false;
        h->expand_pending = false;
        h->have_hbytes = false; h->have_vhash = false; h->have_order = false;
        if (h->host_bases) CK(cudaStreamSynchronize(h->copy_stream));   // the reads are all in HBM by now: one slice
        const uint8_t *hb = h->host_bases;
        h->host_bases = nullptr;
        rc = DISPATCH(build_impl, h, (const u8 *)d_bases, (const u64 *)d_read_off);
        h->host_bases = hb;
        if (rc == DBG_OK && stop_early) return DBG_OK;
        if (rc == DBG_OK) rc = DISPATCH(graph_finalize, h);
    }
    TRY(rc);
    TRY(DISPATCH(graph_order_scan, h));
    TRY(DISPATCH(graph_complete, h));
    return DBG_OK;
}

static int build_both_strands(dbg_t *h, const uint8_t *d_bases, const uint64_t *d_read_off, uint64_t n_reads, uint64_t n_bases,
                              uint64_t read_idx_base, double threshold, uint32_t flags);

DBG_API int dbg_build_device(dbg_t *h, const uint8_t *d_bases, const uint64_t *d_read_off, uint64_t n_reads,
                     uint64_t n_bases, uint64_t read_idx_base, double threshold, uint32_t flags) {
    if (h && (flags & DBG_BOTH_STRANDS)) {
        if ((n_bases && !d_bases) || !d_read_off) return fail(h, DBG_ERR_ARG, "dbg_build_device: NULL input");
        CK(cudaSetDevice(h->device));
        return build_both_strands(h, d_bases, d_read_off, n_reads, n_bases, read_idx_base, threshold, flags);
    }
    return build_device_common(h, d_bases, d_read_off, n_reads, n_bases, read_idx_base, threshold, flags, nullptr, 0);
}

// host entry points: the offsets go over first (everything that needs only them starts at once); the
// bases follow in slices inside build_impl, overlapped with k_pack / k_masks
static int stage_host_reads(dbg_t *h, const uint8_t *bases, const uint64_t *read_off, uint64_t n_reads, uint64_t n_bases) {
    TRY(reserve(h, h->b_bases, n_bases + 32));
    TRY(reserve(h, h->b_off, (n_reads + 1) * 8));
    CK(cudaMemcpyAsync(h->b_off.p, read_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    h->host_bases = n_bases ? bases : nullptr;
    return DBG_OK;
}

// The caller's buffers are its own again when dbg_build* returns -- also when it returns an error from the middle
// of the slice loop, with copies of later slices still queued on the copy stream.
static void release_host_reads(dbg_t *h, int rc) {
    if (rc != DBG_OK) {
        cudaStreamSynchronize(h->copy_stream);
        cudaStreamSynchronize(h->stream);
    }
    h->host_bases = nullptr;
}

DBG_API int dbg_build_batched(dbg_t *h, const uint8_t *bases, const uint64_t *read_off, uint64_t n_reads,
                              const uint32_t *read_batch, uint32_t n_batches, double threshold, uint32_t flags) {
    if (!h) return DBG_ERR_ARG;
    if (!read_off || (n_reads && !read_batch)) return fail(h, DBG_ERR_ARG, "dbg_build_batched: NULL input");
    if (n_batches == 0) return fail(h, DBG_ERR_ARG, "dbg_build_batched: n_batches is 0");
    const uint64_t n_bases = read_off[n_reads];
    if (n_bases && !bases) return fail(h, DBG_ERR_ARG, "dbg_build_batched: bases is NULL");
    for (uint64_t r = 0; r < n_reads; r++)
        if (read_batch[r] >= n_batches || (r && read_batch[r] < read_batch[r - 1]))
            return fail(h, DBG_ERR_ARG, "dbg_build_batched: read_batch must be non-decreasing and < n_batches (read %llu)",
                        (unsigned long long)r);
    CK(cudaSetDevice(h->device));
    TRY(stage_host_reads(h, bases, read_off, n_reads, n_bases));
    TRY(reserve(h, h->b_readbatch, (n_reads + 1) * 4));
    if (n_reads) CK(cudaMemcpyAsync(h->b_readbatch.p, read_batch, n_reads * 4, cudaMemcpyHostToDevice, h->stream));
    const int rc = build_device_common(h, ptr<const uint8_t>(h->b_bases), ptr<const uint64_t>(h->b_off), n_reads, n_bases,
                                       0, threshold, flags, ptr<const uint32_t>(h->b_readbatch), n_batches);
    release_host_reads(h, rc);
    return rc;
}

DBG_API int dbg_build_batched_device(dbg_t *h, const uint8_t *d_bases, const uint64_t *d_read_off, uint64_t n_reads,
                                     uint64_t n_bases, const uint32_t *d_read_batch, uint32_t n_batches,
                                     double threshold, uint32_t flags) {
    if (!h) return DBG_ERR_ARG;
    if (n_batches == 0 || (n_reads && !d_read_batch))
        return fail(h, DBG_ERR_ARG, "dbg_build_batched_device: bad argument");
    CK(cudaSetDevice(h->device));
    return build_device_common(h, d_bases, d_read_off, n_reads, n_bases, 0, threshold, flags, d_read_batch, n_batches);
}

// DBG_BOTH_STRANDS: build from the reads on the device followed by their reverse complements (read n + i = rc of read i)
static int build_both_strands(dbg_t *h, const uint8_t *d_bases, const uint64_t *d_read_off, uint64_t n_reads, uint64_t n_bases,
                              uint64_t read_idx_base, double threshold, uint32_t flags) {
    if (h->generic) return fail(h, DBG_ERR_ARG, "DBG_BOTH_STRANDS exists for the DNA engine only");
    if (read_idx_base || (flags & (DBG_TABLE_ONLY | DBG_CHUNKS_ONLY)))
        return fail(h, DBG_ERR_ARG, "DBG_BOTH_STRANDS is a single-handle build (read_idx_base 0, no DBG_TABLE_ONLY / DBG_CHUNKS_ONLY)");
    if (2 * n_reads >= 0xFFFFFFFFull) return fail(h, DBG_ERR_OVERFLOW, "read index exceeds 32 bits");
    TRY(reserve(h, h->b_bases2, 2 * n_bases + 32));
    TRY(reserve(h, h->b_off2, (2 * n_reads + 1) * 8));
    uint64_t grid = cdiv(n_reads, DBG_WARPS);
    if (grid > 148 * 16) grid = 148 * 16;
    if (grid < 1) grid = 1;
    LAUNCH(k_strand_reads, grid, DBG_BLOCK, h->stream, (const u8 *)d_bases, (const u64 *)d_read_off, (u64)n_reads, (u64)n_bases, ptr<u8>(h->b_bases2), ptr<u64>(h->b_off2));
    CKK();
    return build_device_common(h, ptr<const uint8_t>(h->b_bases2), ptr<const uint64_t>(h->b_off2), 2 * n_reads, 2 * n_bases, 0,
                               threshold, flags & ~DBG_BOTH_STRANDS, nullptr, 0);
}

DBG_API int dbg_build(dbg_t *h, const uint8_t *bases, const uint64_t *read_off, uint64_t n_reads,
              uint64_t read_idx_base, double threshold, uint32_t flags) {
    if (!h) return DBG_ERR_ARG;
    if (!read_off) return fail(h, DBG_ERR_ARG, "dbg_build: read_off is NULL");
    const uint64_t n_bases = read_off[n_reads];
    if (n_bases && !bases) return fail(h, DBG_ERR_ARG, "dbg_build: bases is NULL");
    CK(cudaSetDevice(h->device));
    if (flags & DBG_BOTH_STRANDS) {                                      // (no slice streaming in this mode: everything over first)
        TRY(reserve(h, h->b_bases, n_bases + 32));
        TRY(reserve(h, h->b_off, (n_reads + 1) * 8));
        CK(cudaMemcpyAsync(h->b_off.p, read_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, h->stream));
        if (n_bases) CK(cudaMemcpyAsync(h->b_bases.p, bases, n_bases, cudaMemcpyHostToDevice, h->stream));
        h->host_bases = nullptr;
        const int rc2 = build_both_strands(h, ptr<const uint8_t>(h->b_bases), ptr<const uint64_t>(h->b_off), n_reads, n_bases,
                                           read_idx_base, threshold, flags);
        if (rc2 != DBG_OK) cudaStreamSynchronize(h->stream);             // the caller's buffers are its own again
        return rc2;
    }
    TRY(stage_host_reads(h, bases, read_off, n_reads, n_bases));
    const int rc = build_device_common(h, ptr<const uint8_t>(h->b_bases), ptr<const uint64_t>(h->b_off), n_reads, n_bases,
                                       read_idx_base, threshold, flags, nullptr, 0);
    release_host_reads(h, rc);
    return rc;
}

#define NEED_TABLE() do { if (!h) return DBG_ERR_ARG; if (h->generic) return fail(h, DBG_ERR_STATE, "%s exists for the DNA engine only", __func__); if (!h->table_ready) return fail(h, DBG_ERR_STATE, "%s needs a node table (dbg_build* or dbg_table_reset first)", __func__); } while (0)

DBG_API int dbg_graph_stage(dbg_t *h, int stage, double threshold, uint32_t flags) {
    NEED_TABLE();
    CK(cudaSetDevice(h->device));
    if (!h->tm.empty()) { CK(cudaStreamSynchronize(h->stream)); tm_flush(h); }   // the previous stage's times
    h->n_ev = 0;
    if (stage == 0) {
        if (threshold != threshold) return fail(h, DBG_ERR_ARG, "threshold is NaN");
        h->threshold = threshold;
        h->build_flags = flags;
        h->built = h->walked = false;
        h->up_ready = false;
        return DISPATCH(graph_finalize, h);
    }
    if (stage == 1) return DISPATCH(graph_order_scan, h);
    if (stage == 2) return DISPATCH(graph_complete, h);
    return fail(h, DBG_ERR_ARG, "dbg_graph_stage: stage %d out of range (0..2)", stage);
}

DBG_API int dbg_record_words(const dbg_t *h) { return h ? h->W + 3 : 0; }
DBG_API int dbg_order_words(const dbg_t *h) { return h ? h->W + 4 : 0; }

DBG_API int dbg_export_counts(dbg_t *h, uint32_t n_parts, uint64_t *counts) {
    NEED_TABLE();
    if (n_parts == 0 || n_parts > DBG_MAX_PARTS || !counts) return fail(h, DBG_ERR_ARG, "dbg_export_counts: bad argument");
    CK(cudaSetDevice(h->device));
    return DISPATCH(export_counts_impl, h, n_parts, counts);
}

DBG_API int dbg_export_records(dbg_t *h, uint32_t n_parts, const uint64_t *counts, uint64_t *d_out) {
    NEED_TABLE();
    if (n_parts == 0 || n_parts > DBG_MAX_PARTS || !counts) return fail(h, DBG_ERR_ARG, "dbg_export_records: bad argument");
    CK(cudaSetDevice(h->device));
    return DISPATCH(export_records_impl, h, n_parts, counts, d_out);
}

DBG_API int dbg_table_reset(dbg_t *h, uint64_t expected_nodes) {
    if (!h) return DBG_ERR_ARG;
    if (h->generic) return fail(h, DBG_ERR_STATE, "dbg_table_reset exists for the DNA engine only");
    CK(cudaSetDevice(h->device));
    return DISPATCH(table_reset_impl, h, expected_nodes);
}

DBG_API int dbg_merge_records(dbg_t *h, const uint64_t *d_records, uint64_t n) {
    NEED_TABLE();
    if (n && !d_records) return fail(h, DBG_ERR_ARG, "NULL records");
    CK(cudaSetDevice(h->device));
    return DISPATCH(merge_records_impl, h, d_records, n);
}

DBG_API int dbg_order_export(dbg_t *h, uint64_t *d_out, uint64_t capacity, uint64_t *n_out) {
    NEED_TABLE();
    if (!n_out) return fail(h, DBG_ERR_ARG, "NULL n_out");
    CK(cudaSetDevice(h->device));
    return DISPATCH(order_export_impl, h, d_out, capacity, n_out);
}

DBG_API int dbg_order_merge(dbg_t *h, const uint64_t *d_records, uint64_t n) {
    NEED_TABLE();
    CK(cudaSetDevice(h->device));
    return DISPATCH(order_merge_impl, h, d_records, n);
}

#define NEED_CHUNKS() do { if (!h) return DBG_ERR_ARG; if (h->generic) return fail(h, DBG_ERR_STATE, "%s exists for the DNA engine only", __func__); if (!h->chunks_valid) return fail(h, DBG_ERR_STATE, "%s needs a chunk table (dbg_build* with DBG_CHUNKS_ONLY, or dbg_chunk_table_reset, first)", __func__); } while (0)

DBG_API int dbg_chunk_record_words(const dbg_t *h) { return h ? (h->W == 1 ? 4 : 8) : 0; }

DBG_API int dbg_chunk_export_counts(dbg_t *h, uint32_t n_parts, uint64_t *counts) {
    NEED_CHUNKS();
    if (n_parts == 0 || n_parts > DBG_MAX_PARTS || !counts) return fail(h, DBG_ERR_ARG, "dbg_chunk_export_counts: bad argument");
    CK(cudaSetDevice(h->device));
    return DISPATCH(chunk_export_counts_impl, h, n_parts, counts);
}

DBG_API int dbg_chunk_export_records(dbg_t *h, uint32_t n_parts, const uint64_t *counts, uint64_t *d_out) {
    NEED_CHUNKS();
    if (n_parts == 0 || n_parts > DBG_MAX_PARTS || !counts) return fail(h, DBG_ERR_ARG, "dbg_chunk_export_records: bad argument");
    CK(cudaSetDevice(h->device));
    return DISPATCH(chunk_export_records_impl, h, n_parts, counts, d_out);
}

DBG_API int dbg_chunk_table_reset(dbg_t *h, uint64_t expected_chunks) {
    if (!h) return DBG_ERR_ARG;
    if (h->generic) return fail(h, DBG_ERR_STATE, "dbg_chunk_table_reset exists for the DNA engine only");
    CK(cudaSetDevice(h->device));
    return DISPATCH(chunk_table_reset_impl, h, expected_chunks);
}

DBG_API int dbg_chunk_merge_records(dbg_t *h, const uint64_t *d_records, uint64_t n) {
    NEED_CHUNKS();
    if (n && !d_records) return fail(h, DBG_ERR_ARG, "NULL records");
    CK(cudaSetDevice(h->device));
    return DISPATCH(chunk_merge_records_impl, h, d_records, n);
}

DBG_API int dbg_chunk_expand(dbg_t *h) {
    NEED_CHUNKS();
    CK(cudaSetDevice(h->device));
    return DISPATCH(chunk_expand_impl, h);
}

DBG_API int dbg_stats(dbg_t *h, dbg_stats_t *out) {
    if (!h || !out) return DBG_ERR_ARG;
    if (!h->tm.empty()) {                               // stage times still sitting in their events
        CK(cudaSetDevice(h->device));
        CK(cudaStreamSynchronize(h->stream));
        tm_flush(h);
    }
    *out = h->st;
    return DBG_OK;
}

#define NEED_BUILT() do { if (!h) return DBG_ERR_ARG; if (!h->built) return fail(h, DBG_ERR_STATE, "%s before a successful dbg_build", __func__); } while (0)

template <int W>
static int export_impl(dbg_t *h, uint32_t *node_id, uint64_t *key_lo, uint64_t *key_hi, uint64_t *first_tag,
                       uint32_t *succ_cnt, uint8_t *flags, uint16_t *kept) {
    const uint64_t n = h->st.n_nodes;
    if (n == 0) return DBG_OK;
    DevBuf *s = h->b_scratch;
    TRY(reserve(h, s[0], (n + 1) * 4));
    TRY(reserve(h, s[1], n * 8));
    TRY(reserve(h, s[2], n * 8));
    TRY(reserve(h, s[3], n * 8));
    TRY(reserve(h, s[4], n * 16));
    TRY(reserve(h, s[5], n * 4));        // flags (n bytes) then kept (2n bytes) at offset n rounded to 2
    u8 *d_flags = ptr<u8>(s[5]);
    u16 *d_kept = (u16 *)(ptr<u8>(s[5]) + ((n + 1) & ~1ull));
    DevCounters *d_ctr = ptr<DevCounters>(h->b_ctr);
    // the occupied slots in first-occurrence order (= the reference's dict order, debruijn.py:121-143): sorted on the
    // device, so a host caller does not sort and permute seven arrays (30 ms of numpy for 3 x 10^5 nodes)
    TRY(collect_sorted<W>(h, F_OCC, n, s[0], &d_ctr->n_list_a));
    LAUNCH(k_export_gather<W>, blocks_for(n), DBG_BLOCK, h->stream, ptr<const Slot<W>>(h->b_table), ptr<const u8>(h->b_flags),
           ptr<const u16>(h->b_kept), ptr<const u32>(s[0]), n, ptr<u64>(s[1]), ptr<u64>(s[2]), ptr<u64>(s[3]), ptr<u32>(s[4]),
           d_flags, d_kept);
    CKK();
    if (node_id) CK(cudaMemcpyAsync(node_id, s[0].p, n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (key_lo) CK(cudaMemcpyAsync(key_lo, s[1].p, n * 8, cudaMemcpyDeviceToHost, h->stream));
    if (key_hi) CK(cudaMemcpyAsync(key_hi, s[2].p, n * 8, cudaMemcpyDeviceToHost, h->stream));
    if (first_tag) CK(cudaMemcpyAsync(first_tag, s[3].p, n * 8, cudaMemcpyDeviceToHost, h->stream));
    if (succ_cnt) CK(cudaMemcpyAsync(succ_cnt, s[4].p, n * 16, cudaMemcpyDeviceToHost, h->stream));
    if (flags) CK(cudaMemcpyAsync(flags, d_flags, n, cudaMemcpyDeviceToHost, h->stream));
    if (kept) CK(cudaMemcpyAsync(kept, d_kept, n * 2, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DBG_OK;
}

DBG_API int dbg_export_nodes(dbg_t *h, uint32_t *node_id, uint64_t *key_lo, uint64_t *key_hi, uint64_t *first_tag,
                     uint32_t *succ_cnt, uint8_t *flags, uint16_t *kept) {
    NEED_BUILT();
    CK(cudaSetDevice(h->device));
    return DISPATCH(export_impl, h, node_id, key_lo, key_hi, first_tag, succ_cnt, flags, kept);
}

DBG_API int dbg_export_edges(dbg_t *h, uint32_t *src_id, uint32_t *dst_id, uint8_t *symbol, uint32_t *count, uint8_t *kept) {
    NEED_BUILT();
    if (!h->generic) return fail(h, DBG_ERR_STATE, "dbg_export_edges is the generic-alphabet engine's export; the DNA engine "
                                             "reports successor counts per node (dbg_export_nodes)");
    const uint64_t n = h->n_earr;
    if (n == 0) return DBG_OK;
    CK(cudaSetDevice(h->device));
    DevBuf *s = h->b_scratch;
    TRY(reserve(h, s[0], n * 4));
    TRY(reserve(h, s[1], n * 4));
    TRY(reserve(h, s[2], n * 4));
    TRY(reserve(h, s[3], n));
    TRY(reserve(h, s[4], n));
    LAUNCH(k_g_export_edges, blocks_for(n), DBG_BLOCK, h->stream, ptr<const GEdge>(h->b_earr), n, ptr<const u32>(h->b_estart),
           ptr<const u32>(h->b_enk), ptr<const u8>(h->b_lut) + 256, ptr<u32>(s[0]), ptr<u32>(s[1]), ptr<u8>(s[3]), ptr<u32>(s[2]),
           ptr<u8>(s[4]));
    CKK();
    if (src_id) CK(cudaMemcpyAsync(src_id, s[0].p, n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (dst_id) CK(cudaMemcpyAsync(dst_id, s[1].p, n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (count) CK(cudaMemcpyAsync(count, s[2].p, n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (symbol) CK(cudaMemcpyAsync(symbol, s[3].p, n, cudaMemcpyDeviceToHost, h->stream));
    if (kept) CK(cudaMemcpyAsync(kept, s[4].p, n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DBG_OK;
}

DBG_API int dbg_branch_nodes(dbg_t *h, uint32_t *node_id) {
    NEED_BUILT();
    if (h->branch.size() && !node_id) return fail(h, DBG_ERR_ARG, "NULL output");
    if (h->branch.size()) memcpy(node_id, h->branch.data(), h->branch.size() * 4);
    return DBG_OK;
}

DBG_API int dbg_pulled_nodes(dbg_t *h, uint32_t *node_id) {
    NEED_BUILT();
    if (h->pulled.size() && !node_id) return fail(h, DBG_ERR_ARG, "NULL output");
    if (h->pulled.size()) memcpy(node_id, h->pulled.data(), h->pulled.size() * 4);
    return DBG_OK;
}

DBG_API int dbg_pull_reads(dbg_t *h, uint8_t *read_flags) {
    NEED_BUILT();
    if (h->n_reads && !read_flags) return fail(h, DBG_ERR_ARG, "NULL output");
    if (!h->read_flags_valid) {                         // DBG_FINAL / DBG_NO_PULL_READS: the list is empty
        memset(read_flags, 0, h->n_reads);
        return DBG_OK;
    }
    if (h->n_reads) memcpy(read_flags, h->read_flags.data(), h->n_reads);
    return DBG_OK;
}

// ---- sharded node table (dbg_shard.cuh) ----
#define NEED_SHARD(stage_min) do { NEED_BUILT(); if (h->shard_n < 2) return fail(h, DBG_ERR_STATE, "%s: the handle is not sharded (dbg_shard_config)", __func__); \
    if (h->shard_stage < (stage_min)) return fail(h, DBG_ERR_STATE, "%s called out of order", __func__); } while (0)

DBG_API int dbg_shard_config(dbg_t *h, uint32_t rank, uint32_t n_ranks) {
    if (!h) return DBG_ERR_ARG;
    if (n_ranks > DBG_MAX_PARTS || (n_ranks && rank >= n_ranks)) return fail(h, DBG_ERR_ARG, "dbg_shard_config: bad rank %u of %u", rank, n_ranks);
    if (n_ranks > 1 && n_ranks > 0xFFFFFFu) return fail(h, DBG_ERR_ARG, "too many ranks");
    h->shard_rank = n_ranks > 1 ? rank : 0;
    h->shard_n = n_ranks > 1 ? n_ranks : 0;
    h->shard_stage = 0;
    return DBG_OK;
}

DBG_API int dbg_shard_exit_counts(dbg_t *h, uint64_t *counts) {
    NEED_SHARD(0);
    if (!counts) return fail(h, DBG_ERR_ARG, "NULL counts");
    CK(cudaSetDevice(h->device));
    return DISPATCH(shard_exit_counts_impl, h, counts);
}

DBG_API int dbg_shard_exit_keys(dbg_t *h, uint64_t *d_keys_out) {
    NEED_SHARD(1);
    if (h->n_exit && !d_keys_out) return fail(h, DBG_ERR_ARG, "NULL output");
    CK(cudaSetDevice(h->device));
    return DISPATCH(shard_exit_keys_impl, h, d_keys_out);
}

DBG_API int dbg_shard_mark_entries(dbg_t *h, const uint64_t *d_keys, uint64_t n) {
    NEED_SHARD(1);
    if (n && !d_keys) return fail(h, DBG_ERR_ARG, "NULL keys");
    CK(cudaSetDevice(h->device));
    return DISPATCH(shard_mark_entries_impl, h, d_keys, n);
}

DBG_API int dbg_shard_rank_local(dbg_t *h, uint64_t *n_reduced) {
    NEED_SHARD(2);
    if (!n_reduced) return fail(h, DBG_ERR_ARG, "NULL n_reduced");
    CK(cudaSetDevice(h->device));
    return DISPATCH(shard_rank_local_impl, h, n_reduced);
}

DBG_API int dbg_shard_entry_ordinals(dbg_t *h, uint32_t *d_out) {
    NEED_SHARD(3);
    if (h->n_req && !d_out) return fail(h, DBG_ERR_ARG, "NULL output");
    CK(cudaSetDevice(h->device));
    return shard_entry_ordinals_impl(h, d_out);
}

DBG_API int dbg_shard_reduced_records(dbg_t *h, const uint32_t *d_replies, const uint64_t *rid_base, uint64_t *d_out) {
    NEED_SHARD(3);
    if (!rid_base || (h->n_exit && !d_replies) || (h->n_rv && !d_out)) return fail(h, DBG_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(h->device));
    return shard_reduced_records_impl(h, d_replies, rid_base, d_out);
}

DBG_API int dbg_shard_walk(dbg_t *h, const uint64_t *d_all_records, uint64_t n_total, uint64_t my_base, uint64_t *n_contigs) {
    NEED_SHARD(4);
    if (!n_contigs || (n_total && !d_all_records) || my_base + h->n_rv > n_total) return fail(h, DBG_ERR_ARG, "dbg_shard_walk: bad argument");
    CK(cudaSetDevice(h->device));
    return DISPATCH(shard_walk_impl, h, d_all_records, n_total, my_base, n_contigs);
}

DBG_API int dbg_shard_contigs(dbg_t *h, dbg_shard_contig_t *out) {
    NEED_SHARD(5);
    if (h->n_sview && !out) return fail(h, DBG_ERR_ARG, "NULL output");
    if (h->n_sview) memcpy(out, h->sview, h->n_sview * sizeof(dbg_shard_contig_t));
    return DBG_OK;
}

template <int W>
static int node_keys_impl(dbg_t *h, const uint32_t *node_id, uint64_t n, uint64_t *key_lo, uint64_t *key_hi) {
    DevBuf *s = h->b_scratch;
    TRY(reserve(h, s[0], n * 4));
    TRY(reserve(h, s[1], n * 8));
    TRY(reserve(h, s[2], n * 8));
    CK(cudaMemcpyAsync(s[0].p, node_id, n * 4, cudaMemcpyHostToDevice, h->stream));
    LAUNCH(k_gather_keys<W>, blocks_for(n), DBG_BLOCK, h->stream, ptr<const Slot<W>>(h->b_table), ptr<const u32>(s[0]),
           n, ptr<u64>(s[1]), ptr<u64>(s[2]));
    CKK();
    if (key_lo) CK(cudaMemcpyAsync(key_lo, s[1].p, n * 8, cudaMemcpyDeviceToHost, h->stream));
    if (key_hi) CK(cudaMemcpyAsync(key_hi, s[2].p, n * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DBG_OK;
}

DBG_API int dbg_node_keys(dbg_t *h, const uint32_t *node_id, uint64_t n, uint64_t *key_lo, uint64_t *key_hi) {
    NEED_BUILT();
    if (n == 0) return DBG_OK;
    if (!node_id) return fail(h, DBG_ERR_ARG, "NULL input");
    for (uint64_t i = 0; i < n; i++)
        if (node_id[i] >= h->n_slots) return fail(h, DBG_ERR_ARG, "node id %u out of range", node_id[i]);
    CK(cudaSetDevice(h->device));
    return DISPATCH(node_keys_impl, h, node_id, n, key_lo, key_hi);
}

template <int W>
static int twin_impl(dbg_t *h, const uint32_t *node_id, uint64_t n, uint32_t *twin, uint8_t *canonical) {
    DevBuf *s = h->b_scratch;
    TRY(reserve(h, s[0], n * 4));
    TRY(reserve(h, s[1], n * 4));
    TRY(reserve(h, s[2], n));
    CK(cudaMemcpyAsync(s[0].p, node_id, n * 4, cudaMemcpyHostToDevice, h->stream));
    LAUNCH(k_twin_nodes<W>, blocks_for(n), DBG_BLOCK, h->stream, ptr<const Slot<W>>(h->b_table), geom_of(h), h->k,
           ptr<const u32>(s[0]), n, ptr<u32>(s[1]), ptr<u8>(s[2]));
    CKK();
    if (twin) CK(cudaMemcpyAsync(twin, s[1].p, n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (canonical) CK(cudaMemcpyAsync(canonical, s[2].p, n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DBG_OK;
}

DBG_API int dbg_twin_nodes(dbg_t *h, const uint32_t *node_id, uint64_t n, uint32_t *twin_id, uint8_t *canonical) {
    NEED_BUILT();
    if (n == 0) return DBG_OK;
    if (!node_id) return fail(h, DBG_ERR_ARG, "NULL input");
    if (h->generic || h->batch_bits) return fail(h, DBG_ERR_STATE, "dbg_twin_nodes: DNA handles, unbatched builds");
    for (uint64_t i = 0; i < n; i++)
        if (node_id[i] >= h->n_slots) return fail(h, DBG_ERR_ARG, "node id %u out of range", node_id[i]);
    CK(cudaSetDevice(h->device));
    return DISPATCH(twin_impl, h, node_id, n, twin_id, canonical);
}

template <int W>
static int degree_impl(dbg_t *h, const uint32_t *node_id, uint64_t n, uint8_t *indeg, uint8_t *outdeg, int8_t *balance) {
    DevBuf *s = h->b_scratch;
    TRY(reserve(h, s[0], h->n_slots * 4));
    TRY(reserve(h, s[1], n * 4));
    TRY(reserve(h, s[2], n));
    TRY(reserve(h, s[3], n));
    TRY(reserve(h, s[4], n));
    CK(cudaMemsetAsync(s[0].p, 0, h->n_slots * 4, h->stream));
    CK(cudaMemcpyAsync(s[1].p, node_id, n * 4, cudaMemcpyHostToDevice, h->stream));
    uint64_t grid = blocks_for(h->n_slots);
    if (grid > 148 * 16) grid = 148 * 16;
    LAUNCH(k_true_indegree<W>, grid, DBG_BLOCK, h->stream, ptr<const Slot<W>>(h->b_table), geom_of(h), h->k, ptr<u32>(s[0]));
    CKK();
    LAUNCH(k_degree_gather<W>, blocks_for(n), DBG_BLOCK, h->stream, ptr<const Slot<W>>(h->b_table), ptr<const u32>(s[0]),
           ptr<const u32>(s[1]), n, ptr<u8>(s[2]), ptr<u8>(s[3]), ptr<signed char>(s[4]));
    CKK();
    if (indeg) CK(cudaMemcpyAsync(indeg, s[2].p, n, cudaMemcpyDeviceToHost, h->stream));
    if (outdeg) CK(cudaMemcpyAsync(outdeg, s[3].p, n, cudaMemcpyDeviceToHost, h->stream));
    if (balance) CK(cudaMemcpyAsync(balance, s[4].p, n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DBG_OK;
}

DBG_API int dbg_degree_balance(dbg_t *h, const uint32_t *node_id, uint64_t n, uint8_t *indeg, uint8_t *outdeg, int8_t *balance) {
    NEED_BUILT();
    if (n == 0) return DBG_OK;
    if (!node_id) return fail(h, DBG_ERR_ARG, "NULL input");
    if (h->generic) return fail(h, DBG_ERR_STATE, "dbg_degree_balance: DNA handles only (a generic-alphabet handle lists its edges with dbg_export_edges)");
    if (h->shard_n > 1) return fail(h, DBG_ERR_STATE, "dbg_degree_balance needs the complete graph, this handle holds a shard");
    for (uint64_t i = 0; i < n; i++)
        if (node_id[i] >= h->n_slots) return fail(h, DBG_ERR_ARG, "node id %u out of range", node_id[i]);
    CK(cudaSetDevice(h->device));
    return DISPATCH(degree_impl, h, node_id, n, indeg, outdeg, balance);
}

template <int W>
static int probe_stats_impl(dbg_t *h, int which, uint64_t *hist, uint64_t *n_entries, uint64_t *sum_len, uint64_t *max_len,
                            uint64_t *n_slots_out) {
    DevBuf *s = h->b_scratch;
    TRY(reserve(h, s[0], (PROBE_HIST + 3) * 8));
    CK(cudaMemsetAsync(s[0].p, 0, (PROBE_HIST + 3) * 8, h->stream));
    u64 *d_hist = ptr<u64>(s[0]), *d_acc = ptr<u64>(s[0]) + PROBE_HIST;
    const uint64_t n_slots = which == 0 ? h->n_slots : h->n_cslots;
    uint64_t grid = blocks_for(n_slots);
    if (grid > 148 * 8) grid = 148 * 8;
    if (which == 0)
        LAUNCH(k_probe_stats_nodes<W>, grid, DBG_BLOCK, h->stream, ptr<const Slot<W>>(h->b_table), geom_of(h), d_hist, d_acc);
    else
        LAUNCH(k_probe_stats_chunks<W>, grid, DBG_BLOCK, h->stream, ptr<const ChunkSlot<W>>(h->b_ctable), n_slots, d_hist, d_acc);
    CKK();
    uint64_t host[PROBE_HIST + 3];
    CK(cudaMemcpyAsync(host, s[0].p, sizeof host, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (hist) memcpy(hist, host, PROBE_HIST * 8);
    if (n_entries) *n_entries = host[PROBE_HIST];
    if (sum_len) *sum_len = host[PROBE_HIST + 1];
    if (max_len) *max_len = host[PROBE_HIST + 2];
    if (n_slots_out) *n_slots_out = n_slots;
    return DBG_OK;
}

DBG_API int dbg_probe_stats(dbg_t *h, int which, uint64_t *hist64, uint64_t *n_entries, uint64_t *sum_len, uint64_t *max_len,
                            uint64_t *n_slots) {
    NEED_BUILT();
    if (which != 0 && which != 1) return fail(h, DBG_ERR_ARG, "dbg_probe_stats: which = 0 (node table) or 1 (chunk table)");
    if (h->generic) return fail(h, DBG_ERR_STATE, "dbg_probe_stats exists for the DNA engine only");
    if (which == 0 && !h->table_ready) return fail(h, DBG_ERR_STATE, "dbg_probe_stats: no node table on this handle");
    if (which == 1 && (!h->chunks_valid || !h->n_cslots))
        return fail(h, DBG_ERR_STATE, "dbg_probe_stats: no chunk table on this handle (the build did not go through chunks)");
    CK(cudaSetDevice(h->device));
    return DISPATCH(probe_stats_impl, h, which, hist64, n_entries, sum_len, max_len, n_slots);
}

template <int W>
static int lookup_impl(dbg_t *h, const uint8_t *q, const uint32_t *batch, uint64_t n, uint32_t *counts) {
    DevBuf *s = h->b_scratch;
    const uint64_t bytes = n * (uint64_t)(h->k + 1);
    TRY(reserve(h, s[0], bytes));
    TRY(reserve(h, s[1], n * 4));
    TRY(reserve(h, s[2], n * 4));
    CK(cudaMemcpyAsync(s[0].p, q, bytes, cudaMemcpyHostToDevice, h->stream));
    if (batch) CK(cudaMemcpyAsync(s[2].p, batch, n * 4, cudaMemcpyHostToDevice, h->stream));
    if (h->generic)
        LAUNCH(k_g_lookup_edges, blocks_for(n), DBG_BLOCK, h->stream, ptr<const Slot<2>>(h->b_etable), h->n_eslots, h->k,
               h->sym_bits, ptr<const u8>(h->b_lut), ptr<const u8>(s[0]), n, ptr<u32>(s[1]));
    else
    LAUNCH(k_lookup_edges<W>, blocks_for(n), DBG_BLOCK, h->stream, ptr<const Slot<W>>(h->b_table), geom_of(h), h->k,
           ptr<const u8>(s[0]), batch ? ptr<const u32>(s[2]) : (const u32 *)nullptr, n, ptr<u32>(s[1]));
    CKK();
    CK(cudaMemcpyAsync(counts, s[1].p, n * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DBG_OK;
}

DBG_API int dbg_lookup_edges(dbg_t *h, const uint8_t *kplus1mers, uint64_t n, uint32_t *counts) {
    NEED_BUILT();
    if (n == 0) return DBG_OK;
    if (!kplus1mers || !counts) return fail(h, DBG_ERR_ARG, "NULL argument");
    CK(cudaSetDevice(h->device));
    return DISPATCH(lookup_impl, h, kplus1mers, (const uint32_t *)nullptr, n, counts);
}

DBG_API int dbg_lookup_edges_batched(dbg_t *h, const uint8_t *kplus1mers, const uint32_t *batch, uint64_t n,
                                     uint32_t *counts) {
    NEED_BUILT();
    if (n == 0) return DBG_OK;
    if (!kplus1mers || !counts || !batch) return fail(h, DBG_ERR_ARG, "NULL argument");
    if (h->generic || (!h->batch_bits && h->n_batches == 0)) return fail(h, DBG_ERR_STATE, "the graph was not built by dbg_build_batched");
    for (uint64_t i = 0; i < n; i++)
        if (batch[i] >= h->n_batches) return fail(h, DBG_ERR_ARG, "batch id %u out of range", batch[i]);
    CK(cudaSetDevice(h->device));
    return DISPATCH(lookup_impl, h, kplus1mers, batch, n, counts);
}

DBG_API int dbg_walk(dbg_t *h, int stop_at_branch, uint64_t *n_contigs) {
    NEED_BUILT();
    CK(cudaSetDevice(h->device));
    h->walked = false;
    TRY(DISPATCH(walk_impl, h, stop_at_branch));
    h->walked = true;
    if (n_contigs) *n_contigs = h->n_cview;
    return DBG_OK;
}

#define NEED_WALKED() do { if (!h) return DBG_ERR_ARG; if (!h->walked) return fail(h, DBG_ERR_STATE, "%s before a successful dbg_walk", __func__); } while (0)

DBG_API int dbg_contigs(dbg_t *h, dbg_contig_t *out) {
    NEED_WALKED();
    if (h->n_cview && !out) return fail(h, DBG_ERR_ARG, "NULL output");
    if (h->n_cview) memcpy(out, h->cview, h->n_cview * sizeof(dbg_contig_t));
    return DBG_OK;
}

DBG_API int dbg_segments(dbg_t *h, dbg_segment_t *out, uint64_t *n_segments) {
    NEED_WALKED();
    if (n_segments) *n_segments = h->segments.size();
    if (out && h->segments.size()) memcpy(out, h->segments.data(), h->segments.size() * sizeof(dbg_segment_t));
    return DBG_OK;
}

// segment list of contig c: its own segments, or the single chain (start, n_nodes)
static void contig_segments(const dbg_t *h, const dbg_contig_t &c, std::vector<dbg_segment_t> &out) {
    if (c.seg_end > c.seg_begin)
        for (uint64_t s = c.seg_begin; s < c.seg_end; s++) out.push_back(h->segments[s]);
    else
        out.push_back(dbg_segment_t{c.start, c.n_nodes});
}

// up[j][v] = succ^(2^(j+1))(v) for the strings / scores kernels; once per build, on first use
static int ensure_lift(dbg_t *h) {
    if (h->up_ready) return DBG_OK;
    const uint64_t n = h->n_slots;
    TRY(reserve(h, h->b_up, (uint64_t)LIFT_LEVELS * n * 4));
    const u32 *prev = ptr<const u32>(h->b_succ);
    for (int j = 0; j < LIFT_LEVELS; j++) {
        u32 *next = ptr<u32>(h->b_up) + (uint64_t)j * n;
        LAUNCH(k_lift, blocks_for(n), DBG_BLOCK, h->stream, prev, next, n);
        CKK();
        prev = next;
    }
    h->up_ready = true;
    return DBG_OK;
}

template <int W>
static int materialize_impl(dbg_t *h, const uint64_t *idx, uint64_t n, char *buf, const uint64_t *off) {
    std::vector<SegJob> jobs;
    std::vector<dbg_segment_t> segs;
    uint64_t total = 0;
    for (uint64_t i = 0; i < n; i++) {
        const dbg_contig_t &c = h->cview[idx[i]];
        segs.clear();
        contig_segments(h, c, segs);
        uint64_t o = off[i];
        bool first = true;
        for (const dbg_segment_t &s : segs) {
            if (s.n_nodes == 0) continue;
            SegJob j;
            j.entry = s.entry; j.n_nodes = s.n_nodes; j.out = o; j.first = first ? 1u : 0u; j.pad = 0;
            jobs.push_back(j);
            o += first ? (uint64_t)h->k + s.n_nodes - 1 : s.n_nodes;
            first = false;
        }
        total = std::max(total, o);
    }
    if (jobs.empty()) return DBG_OK;
    DevBuf *s = h->b_scratch;
    TRY(reserve(h, s[0], jobs.size() * sizeof(SegJob)));
    TRY(reserve(h, s[1], total));
    CK(cudaMemcpyAsync(s[0].p, jobs.data(), jobs.size() * sizeof(SegJob), cudaMemcpyHostToDevice, h->stream));
    TRY(ensure_lift(h));
    LAUNCH((k_materialize<W>), jobs.size(), LIFT_BLOCK, h->stream, ptr<const Slot<W>>(h->b_table), h->n_slots, h->k,
           ptr<const u32>(h->b_succ), ptr<const u32>(h->b_up),
           h->generic ? ptr<const u8>(h->b_lut) + 256 : (const u8 *)nullptr, ptr<const SegJob>(s[0]), ptr<char>(s[1]));
    CKK();
    // strings may be requested with gaps between them: copy each back to where the caller wants it
    uint64_t lo = off[0], hi = 0;
    for (uint64_t i = 0; i < n; i++) { lo = std::min(lo, off[i]); }
    hi = total;
    CK(cudaMemcpyAsync(buf + lo, ptr<char>(s[1]) + lo, hi - lo, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DBG_OK;
}

DBG_API int dbg_materialize(dbg_t *h, const uint64_t *idx, uint64_t n, char *buf, const uint64_t *off) {
    NEED_WALKED();
    if (n == 0) return DBG_OK;
    if (!idx || !buf || !off) return fail(h, DBG_ERR_ARG, "NULL argument");
    for (uint64_t i = 0; i < n; i++)
        if (idx[i] >= h->n_cview) return fail(h, DBG_ERR_ARG, "contig index out of range");
    CK(cudaSetDevice(h->device));
    return DISPATCH(materialize_impl, h, idx, n, buf, off);
}

template <int W>
static int scores_impl(dbg_t *h, const uint64_t *idx, uint64_t n, uint64_t *scores) {
    std::vector<u32> entry, nodes;
    std::vector<ScoreJob> jobs(n);
    std::vector<dbg_segment_t> segs;
    for (uint64_t i = 0; i < n; i++) {
        segs.clear();
        contig_segments(h, h->cview[idx[i]], segs);
        jobs[i].seg_begin = entry.size();
        for (const dbg_segment_t &s : segs) { entry.push_back(s.entry); nodes.push_back(s.n_nodes); }
        jobs[i].seg_end = entry.size();
    }
    DevBuf *s = h->b_scratch;
    TRY(reserve(h, s[0], entry.size() * 4));
    TRY(reserve(h, s[1], nodes.size() * 4));
    TRY(reserve(h, s[2], n * sizeof(ScoreJob)));
    TRY(reserve(h, s[3], n * 8));
    CK(cudaMemcpyAsync(s[0].p, entry.data(), entry.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(s[1].p, nodes.data(), nodes.size() * 4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(s[2].p, jobs.data(), n * sizeof(ScoreJob), cudaMemcpyHostToDevice, h->stream));
    TRY(ensure_lift(h));
    const bool gen = h->generic;
    LAUNCH((k_scores<W>), n, LIFT_BLOCK, h->stream, ptr<const Slot<W>>(h->b_table), h->n_slots, ptr<const u16>(h->b_kept),
           ptr<const u32>(h->b_succ), ptr<const u32>(h->b_up), gen ? ptr<const GEdge>(h->b_earr) : (const GEdge *)nullptr,
           gen ? ptr<const u32>(h->b_estart) : (const u32 *)nullptr, gen ? ptr<const u32>(h->b_edeg) : (const u32 *)nullptr,
           ptr<const u32>(s[0]), ptr<const u32>(s[1]), ptr<const ScoreJob>(s[2]), ptr<u64>(s[3]));
    CKK();
    CK(cudaMemcpyAsync(scores, s[3].p, n * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return DBG_OK;
}

DBG_API int dbg_contig_scores(dbg_t *h, const uint64_t *idx, uint64_t n, uint64_t *scores) {
    NEED_WALKED();
    if (n == 0) return DBG_OK;
    if (!idx || !scores) return fail(h, DBG_ERR_ARG, "NULL argument");
    for (uint64_t i = 0; i < n; i++)
        if (idx[i] >= h->n_cview) return fail(h, DBG_ERR_ARG, "contig index out of range");
    CK(cudaSetDevice(h->device));
    return DISPATCH(scores_impl, h, idx, n, scores);
}

#include "dbg_concat_host.cuh"
