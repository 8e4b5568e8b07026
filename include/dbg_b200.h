/*
 * dbg_b200.h  --  C ABI of the B200-native de Bruijn assembly hot path.
 *
 * This is the drop-in boundary for the path `II_assembleFromReads.py:58-63` drives in the
 * reference (`debruijn.construct_graph` + `debruijn.output_contigs`).  The reference is pure
 * Python, so "the FFI a maintainer would bind" is ctypes: INTEGRATION.md shows the stub, and
 * `py-debruijn_b200/_native.py` is that stub in full.  Plain pointers and sizes only; no
 * torch / CUDA types in any signature (a stream is passed as void*).
 *
 * Conventions
 *   - every call returns DBG_OK (0) or a negative dbg_status; dbg_last_error() gives text
 *   - the caller owns every buffer it passes; the handle owns all device memory
 *   - one handle per host thread / stream; the library starts no host threads
 *   - node ids are table-slot indices, valid until the next dbg_build* on the handle
 *   - alphabet: dbg_create / dbg_create_ex handles take DNA, exactly the bytes 'A','C','G','T' (2-bit codes
 *     0..3 in that order): k-mers ("nodes", debruijn.py:127) are 2k-bit integers, first base most
 *     significant; k <= 31 uses 64-bit keys, k <= 63 uses 128-bit keys (key_hi is 0 for k <= 31).
 *     dbg_create_alpha handles take ANY alphabet of up to 255 byte values (the reference's real inputs are
 *     peptides, II_assembleFromReads.py:43-47; debruijn.py itself works on any characters): codes = position in the
 *     symbol list, b = ceil(log2(n_symbols)) bits each (at least 1), 128-bit keys, b (k + 1) <= 125 and k <= 63 --
 *     peptides (5 bits) k <= 24, ACGT + N (3 bits) k <= 40, 8-bit codes k <= 14
 *   - a tag is (global read index << 32 | position in read) of a node's FIRST occurrence
 *     in the reference's iteration order (debruijn.py:123-126); sorting by tag reproduces
 *     the reference's dict order
 */
#ifndef DBG_B200_H
#define DBG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dbg_handle dbg_t;

typedef enum {
    DBG_OK = 0,
    DBG_ERR_CUDA = -1,       /* CUDA runtime error, or no usable device                    */
    DBG_ERR_ARG = -2,        /* bad argument (k out of range, null pointer, ...)            */
    DBG_ERR_ALPHABET = -3,   /* a read holds a byte other than A/C/G/T                      */
    DBG_ERR_NOMEM = -4,      /* device or host allocation failed                            */
    DBG_ERR_STATE = -5,      /* call out of order (e.g. dbg_walk before dbg_build)          */
    DBG_ERR_ZERODIV = -6,    /* threshold == 0 with a multi-successor node (debruijn.py:163
                                raises ZeroDivisionError there)                             */
    DBG_ERR_OVERFLOW = -7    /* read index / position / node count exceeds 32 bits          */
} dbg_status;

/* dbg_build flags */
#define DBG_FINAL        1u  /* construct_graph(..., final=True), debruijn.py:281-283        */
#define DBG_EXACT_ORDER  2u  /* also compute first-appearance order of each node's successors
                                (Counter.most_common tie order, debruijn.py:159-164); costs a
                                second pass over the reads only when multi-successor nodes exist */
#define DBG_NO_PULL_READS 4u /* skip pull_out_read (debruijn.py:274-278)                      */
#define DBG_TABLE_ONLY   8u  /* stop after the node table (window loop, debruijn.py:123-144): the
                                multi-GPU build exchanges node records before the graph stages   */

#define DBG_CHUNKS_ONLY 16u  /* stop after phase 1 of the chunk build (the distinct super-k-mers of the
                                reads): the multi-GPU build exchanges CHUNK records, before any node table */

#define DBG_BOTH_STRANDS 32u /* EXTENSION without a reference line (north-star "canonical k-mers"; the reference has no reverse
                                complement anywhere, SURVEY.md section 9 delta 1): build from the reads FOLLOWED BY THEIR REVERSE
                                COMPLEMENTS (read n_reads + i = reverse complement of read i), i.e. exactly what the reference
                                computes on that longer read list -- which is its oracle.  The graph then holds every k-mer
                                next to its reverse complement (twin nodes, dbg_twin_nodes): the canonical graph, unfolded.
                                dbg_build / dbg_build_device, DNA handles, read_idx_base 0; dbg_stats / dbg_pull_reads count
                                2 x n_reads reads.                                                                         */

/* node flag bits (dbg_export_nodes) */
#define DBG_NODE_START   0x02u  /* reference indegree == 0 (debruijn.py:137-143 quirk)       */
#define DBG_NODE_BRANCH  0x04u  /* > 1 successor survives pruningEdges (debruijn.py:230-236) */
#define DBG_NODE_DEADEND 0x08u  /* reference outdegree == 0                                   */
#define DBG_NODE_PULLED  0x10u  /* in already_pull_out (debruijn.py:249-254)                  */
#define DBG_NODE_MULTI   0x20u  /* > 1 distinct successor before pruning                      */

/* contig terminal kinds (debruijn.py:288-316) */
#define DBG_TERM_DEADEND 0u  /* stopped at a node without successors; node is on the path     */
#define DBG_TERM_BRANCH  1u  /* stopped at a branch node; node is on the path                 */
#define DBG_TERM_PULLED  2u  /* stopped in front of a pulled node; node is NOT on the path    */

typedef struct {
    uint64_t n_reads;
    uint64_t n_bases;
    uint64_t n_windows;      /* (k+1)-mer instances = iterations of debruijn.py:126-144       */
    uint64_t n_nodes;        /* distinct k-mers        -> "number of {k}mer:"                 */
    uint64_t n_edges;        /* distinct (k+1)-mers    = len(edge_count_table)                */
    uint64_t n_multi;        /* nodes with > 1 distinct successor                             */
    uint64_t n_branch;       /* -> "branch number:" (before `final` empties the list)         */
    uint64_t n_starts;       /* -> "Number of kmers have no income edges:"                    */
    uint64_t n_pulled;       /* len(already_pull_out)                                         */
    uint64_t n_pull_reads;   /* len(pull_out_read) (0 when DBG_FINAL)                         */
    uint64_t n_contigs;      /* after dbg_walk                                                */
    uint64_t n_cycle_starts; /* starts whose walk hits a cycle and emits nothing (:289)       */
    uint64_t table_slots;
    uint64_t table_bytes;
    uint64_t hll_estimate;   /* distinct-node estimate used to size the table                 */
    uint64_t n_launches;     /* kernels this library launched in the last dbg_build* + dbg_walk */
    uint64_t n_chunk_instances; /* super-k-mer (chunk) instances the window loop was grouped into  */
    uint64_t n_chunk_list;   /* first / last chunks of reads (kept in a list, not deduplicated)    */
    uint64_t chunk_slots;    /* slots of the chunk table                                           */
    uint32_t key_words;      /* 1 (k <= 31) or 2                                              */
    uint32_t jump_rounds;    /* pointer-jumping rounds run by dbg_walk                        */
    uint32_t n_insert_passes;/* L2-resident partition passes of the table build               */
    uint32_t n_regions;      /* node table laid out by minimizer: regions (1 = plain hash layout)   */
    /* device time per stage, CUDA events on the handle's stream, milliseconds */
    float ms_h2d, ms_pack, ms_estimate, ms_table_init, ms_insert, ms_finalize, ms_order,
          ms_lists, ms_tips, ms_pull_reads, ms_build_total, ms_walk_rank, ms_walk_emit,
          ms_walk_total,
          /* parts of ms_insert when the chunk build runs (dbg_chunks.cuh): */
          ms_chunk_insert, ms_expand_sketch, ms_expand, ms_region_order;
} dbg_stats_t;

typedef struct {
    uint32_t start;          /* node id of the first node on the path                          */
    uint32_t n_nodes;        /* nodes on the path; contig length = k + n_nodes - 1             */
    uint32_t terminal;       /* node id at which the walk stopped                              */
    uint32_t kind;           /* DBG_TERM_*                                                     */
    uint64_t seg_begin;      /* [seg_begin, seg_end) into the segment list: > 1 segment only   */
    uint64_t seg_end;        /*   for final=True paths that run through multi-successor nodes  */
} dbg_contig_t;

typedef struct {
    uint32_t entry;          /* node id at which the segment starts                            */
    uint32_t n_nodes;        /* nodes taken from this segment                                  */
} dbg_segment_t;

/* ---- lifetime ---- */
int  dbg_create(dbg_t **out, int device, int k);
/* key_words: 0 = by k (1 for k <= 31, else 2); 2 forces 128-bit keys for a small k, which leaves
 * room for many batch ids above the k-mer (dbg_build_batched)                               */
int  dbg_create_ex(dbg_t **out, int device, int k, int key_words);
/* generic-alphabet engine: `symbols` = the n_symbols <= 255 distinct byte values the reads may hold; a node key is
 * k x b bits (symbol codes, first symbol most significant; b as above).  Same calls as a DNA handle except the batched and
 * multi-GPU ones; successors are reported by dbg_export_edges instead of 4 counters per node.                */
int  dbg_create_alpha(dbg_t **out, int device, int k, const uint8_t *symbols, int n_symbols);
void dbg_destroy(dbg_t *h);
const char *dbg_last_error(const dbg_t *h);          /* h may be NULL: error of a failed create */
int  dbg_set_stream(dbg_t *h, void *cuda_stream);    /* cudaStream_t; NULL = handle's own stream */
int  dbg_version(void);
/* findSupportReadScore of IV_sortOutputs.py:10-15 for MANY contigs: out_scores[c] = sum, in read order and in
 * float64 (the reference's left-to-right `score +=`), of read_score[r] over the reads r that are substrings of
 * contig c (a read that occurs more than once counts once, as `read in contig` does; an empty read is in every
 * contig).  reads = the keys of the reference's score_table in dict order, concatenated bytes + n_reads + 1
 * offsets; contigs likewise; all HOST pointers.  Any alphabet; the handle only lends its device, stream and
 * scratch memory (its k does not matter).                                                                  */
int  dbg_support_scores(dbg_t *h, const uint8_t *reads, const uint64_t *read_off, uint64_t n_reads,
                        const double *read_score, const uint8_t *contigs, const uint64_t *contig_off,
                        uint64_t n_contigs, double *out_scores);
/* EXTENSION without a reference line (SURVEY.md section 9 delta 5; BASELINE.json configs[2] "abundance filter
 * min-count=2"): the following builds on this handle skip every window whose (k+1)-mer occurs fewer than
 * min_count times over all reads, as if `if total[read[i:i+k+1]] < min_count: continue` stood at the top of the
 * loop body of debruijn.py:126; everything downstream is the reference's semantics on what is left.  1 = off
 * (the reference).  Pinned by oracle/dbg_oracle.py build(min_count) and oracle/dbg_oracle.c build_filtered only. */
int  dbg_set_min_count(dbg_t *h, uint32_t min_count);

/* ---- construct_graph (debruijn.py:206-285) ----
 * reads = concatenated ASCII bases + n_reads+1 offsets; read i is bases[off[i] .. off[i+1]).
 * read_idx_base is added to every read index in the tags (multi-GPU / batched callers).
 * threshold is pruningEdges' divisor (debruijn.py:163), as a double.
 * dbg_build takes HOST pointers (copies are part of the call); dbg_build_device takes
 * DEVICE pointers already resident in HBM.  Both run the whole graph stage: pack, table
 * build, degrees/prune/branch flags, tip pruning, pull-out read flags.                      */
int dbg_build(dbg_t *h, const uint8_t *bases, const uint64_t *read_off, uint64_t n_reads,
              uint64_t read_idx_base, double threshold, uint32_t flags);
int dbg_build_device(dbg_t *h, const uint8_t *d_bases, const uint64_t *d_read_off,
                     uint64_t n_reads, uint64_t n_bases, uint64_t read_idx_base,
                     double threshold, uint32_t flags);
/* Batched construct_graph: MANY independent read sets ("batches": e.g. one per template,
 * V_FindTemplateGroups.py:152-190) assembled in ONE pass.  read_batch[r] = batch of read r,
 * non-decreasing, < n_batches.  The batch id rides in the key bits above the k-mer, so equal
 * k-mers of different batches are different nodes and no edge joins two batches: the result is
 * exactly the union of the graphs construct_graph would build per batch, with the tags (global
 * read index) keeping every per-batch list in the reference's order.  Needs
 * 2k + ceil(log2(n_batches + 1)) <= 64 (k <= 31) or <= 128 key bits.  In dbg_export_nodes /
 * dbg_node_keys the batch id is key >> 2k.                                                    */
int dbg_build_batched(dbg_t *h, const uint8_t *bases, const uint64_t *read_off, uint64_t n_reads,
                      const uint32_t *read_batch, uint32_t n_batches, double threshold, uint32_t flags);
/* DEVICE pointers (reads and batch ids already in HBM); the caller vouches for read_batch's order */
int dbg_build_batched_device(dbg_t *h, const uint8_t *d_bases, const uint64_t *d_read_off, uint64_t n_reads,
                             uint64_t n_bases, const uint32_t *d_read_batch, uint32_t n_batches,
                             double threshold, uint32_t flags);
int dbg_stats(dbg_t *h, dbg_stats_t *out);

/* ---- multi-GPU building blocks: the window loop is data parallel over reads,
 * its result per rank is a set of NODE RECORDS { key | first-occurrence tag | 4 successor counts }
 * (dbg_record_words() u64 words each) whose merge is associative (min tag, summed counts).  A
 * rank builds its table with DBG_TABLE_ONLY, exports its records grouped by owner
 * ((slot hash * n_parts) >> 32), the caller exchanges them (NCCL all-to-all), merges what it
 * owns, all-gathers the merged shards and merges those into a fresh table, then runs the graph
 * stages: 0 = degrees/prune/branch flags, 1 = tie-order scan of the LOCAL reads (order records
 * are then exchanged with dbg_order_export / dbg_order_merge), 2 = order apply, lists, tips,
 * pull-out flags of the LOCAL reads.  All record pointers are DEVICE pointers.               */
int dbg_graph_stage(dbg_t *h, int stage, double threshold, uint32_t flags);
int dbg_record_words(const dbg_t *h);
int dbg_order_words(const dbg_t *h);
int dbg_export_counts(dbg_t *h, uint32_t n_parts, uint64_t *counts /* host, n_parts */);
int dbg_export_records(dbg_t *h, uint32_t n_parts, const uint64_t *counts, uint64_t *d_out);
int dbg_table_reset(dbg_t *h, uint64_t expected_nodes);
int dbg_merge_records(dbg_t *h, const uint64_t *d_records, uint64_t n);
int dbg_order_export(dbg_t *h, uint64_t *d_out, uint64_t capacity, uint64_t *n_out);
int dbg_order_merge(dbg_t *h, const uint64_t *d_records, uint64_t n);

/* The same exchange one level earlier (less volume, and only the owner expands): what the window loop
 * of a rank's reads leaves BEFORE any node table is the set of its distinct chunks { string | min tag |
 * multiplicity } (dbg_chunk_record_words() u64 words each), whose merge is associative too.  A chunk's
 * owner is a function of its minimizer, which all its nodes share and every occurrence of those nodes
 * has, so everything about one node meets on one rank.  dbg_build* with DBG_CHUNKS_ONLY, export grouped
 * by owner, all-to-all, dbg_chunk_table_reset + dbg_chunk_merge_records, then dbg_chunk_expand leaves
 * the rank's shard of the node table, ready for dbg_export_records(1 part) and the steps above.    */
int dbg_chunk_record_words(const dbg_t *h);
int dbg_chunk_export_counts(dbg_t *h, uint32_t n_parts, uint64_t *counts /* host, n_parts */);
int dbg_chunk_export_records(dbg_t *h, uint32_t n_parts, const uint64_t *counts, uint64_t *d_out);
int dbg_chunk_table_reset(dbg_t *h, uint64_t expected_chunks);
int dbg_chunk_merge_records(dbg_t *h, const uint64_t *d_records, uint64_t n);
int dbg_chunk_expand(dbg_t *h);

/* ---- owner-computes graph stages and contig walk on a SHARDED node table ----
 * After dbg_chunk_expand a rank holds the shard of the node table whose minimizers it owns.  With
 * dbg_shard_config(rank, n_ranks) set, dbg_graph_stage 0 / 1 / 2 run on the shard alone (degrees,
 * pruningEdges, tie order, branch flags and start list are per-node work, debruijn.py:129-166, 230-236; a kept
 * successor owned by another rank makes its node an EXIT instead of an internal-consistency fault).  Stage 2
 * refuses a shard with branch nodes: ordered tip pruning (debruijn.py:241-254) needs the neighbourhoods on one
 * rank -- gather the shards (dbg_export_records / dbg_merge_records) for such inputs.
 * The contig walk (debruijn.py:288-347) then runs as a two-level list ranking, the caller moving three small
 * buffers between ranks (device pointers; in this order):
 *   dbg_shard_exit_counts   exits grouped by the rank that owns their successor (host counts, n_ranks)
 *   dbg_shard_exit_keys     the successors' keys in that order, key_words u64 each      -> all-to-all
 *   dbg_shard_mark_entries  received keys: these nodes are ENTRIES of this shard
 *   dbg_shard_rank_local    list ranking of the shard with exits as roots; *n_reduced = entries + starts
 *   dbg_shard_entry_ordinals one u32 per received key, in order                          -> all-to-all back
 *   dbg_shard_reduced_records  replies (exit order) + the first reduced id of every rank (host, n_ranks: prefix
 *                           sums of all ranks' n_reduced) -> 2 u64 per reduced vertex     -> all-gather
 *   dbg_shard_walk          every rank's records back to back, my_base = this rank's first reduced id: ranks the
 *                           reduced list and leaves one dbg_shard_contig_t per start of THIS shard that emits a
 *                           contig, sorted by tag; all ranks' records merged by tag are the reference's order. */
typedef struct {
    uint64_t tag;            /* first-occurrence tag of the start node                                */
    uint32_t start_slot, start_rank;       /* node id = (rank, slot in that rank's shard)             */
    uint32_t terminal_slot, terminal_rank;
    uint32_t n_nodes;        /* contig length = k + n_nodes - 1                                       */
    uint32_t kind;           /* DBG_TERM_*                                                            */
} dbg_shard_contig_t;
int dbg_shard_config(dbg_t *h, uint32_t rank, uint32_t n_ranks);     /* n_ranks <= 1: not sharded    */
int dbg_shard_exit_counts(dbg_t *h, uint64_t *counts /* host, n_ranks */);
int dbg_shard_exit_keys(dbg_t *h, uint64_t *d_keys_out);
int dbg_shard_mark_entries(dbg_t *h, const uint64_t *d_keys, uint64_t n);
int dbg_shard_rank_local(dbg_t *h, uint64_t *n_reduced);
int dbg_shard_entry_ordinals(dbg_t *h, uint32_t *d_out);
int dbg_shard_reduced_records(dbg_t *h, const uint32_t *d_replies, const uint64_t *rid_base /* host, n_ranks */,
                              uint64_t *d_out);
int dbg_shard_walk(dbg_t *h, const uint64_t *d_all_records, uint64_t n_total, uint64_t my_base, uint64_t *n_contigs);
int dbg_shard_contigs(dbg_t *h, dbg_shard_contig_t *out);            /* n_contigs entries            */

/* nodes in first-occurrence order (ascending first_tag = the reference's dict order, debruijn.py:121-143; sorted
 * on the device).  Any pointer may be
 * NULL.  succ_cnt: 4 per node = edge_count_table[node + "ACGT"[c]]; kept: low 8 bits = kept
 * successor bases in E order (2 bits each, first in bits 0-1), bits 8-10 = how many are kept
 * (pruningEdges result, debruijn.py:150-166).                                               */
int dbg_export_nodes(dbg_t *h, uint32_t *node_id, uint64_t *key_lo, uint64_t *key_hi,
                     uint64_t *first_tag, uint32_t *succ_cnt, uint8_t *flags, uint16_t *kept);
/* generic-alphabet handles: the n_edges distinct (k+1)-mers, grouped by source node and, inside a node, in
 * the reference's E order (count descending, ties by first appearance, debruijn.py:159-164): src / dst node
 * ids, the (k+1)-mer's last symbol (byte), its multiplicity = edge_count_table value, kept = 1 if it survives
 * pruningEdges (a prefix of every node's run).  Any pointer may be NULL.                                    */
int dbg_export_edges(dbg_t *h, uint32_t *src_id, uint32_t *dst_id, uint8_t *symbol, uint32_t *count, uint8_t *kept);
int dbg_branch_nodes(dbg_t *h, uint32_t *node_id);   /* n_branch ids, reference order        */
int dbg_pulled_nodes(dbg_t *h, uint32_t *node_id);   /* n_pulled ids, reference order        */
int dbg_pull_reads(dbg_t *h, uint8_t *read_flags);   /* n_reads bytes, 1 = in pull_out_read  */
int dbg_node_keys(dbg_t *h, const uint32_t *node_id, uint64_t n, uint64_t *key_lo, uint64_t *key_hi);
/* EXTRA output, no reference line (north-star kernel 4 "in/out-degree and Eulerian balance"; SURVEY.md section 9
 * delta 4): the reference keeps outdegree = distinct successors and a 0/1 first-seen indegree FLAG
 * (debruijn.py:129-143, reported by dbg_export_nodes) and never a balance.  For the given nodes of the unpruned
 * graph: indegree = distinct predecessors (distinct (k+1)-mers that end in the node), outdegree = distinct
 * successors, balance = outdegree - indegree.  DNA handles holding the complete graph; any pointer may be NULL. */
int dbg_degree_balance(dbg_t *h, const uint32_t *node_id, uint64_t n, uint8_t *indegree, uint8_t *outdegree,
                       int8_t *balance);
/* Measurement aid (north-star: "probe length" as evidence for the hash-table design; bench.py reports it): probe-length
 * statistics of the node table (which = 0) or the chunk table (which = 1) the last build left.  Both tables probe
 * linearly from a home slot; an entry's probe length is the number of slots a lookup of its key touches.
 * hist64[j] = entries of length j + 1 (hist64[63]: everything longer); any pointer may be NULL.  DNA handles.        */
int dbg_probe_stats(dbg_t *h, int which, uint64_t *hist64, uint64_t *n_entries, uint64_t *sum_len, uint64_t *max_len,
                    uint64_t *n_slots);
/* twin_id[j] = node id of the reverse complement of node node_id[j], 0xFFFFFFFF if the graph does not hold it (it always
 * does after a DBG_BOTH_STRANDS build); canonical[j] = 1 iff the node's k-mer <= its reverse complement (the canonical
 * representative of the pair; both 1 for a palindrome, which is its own twin).  DNA handles, unbatched builds.        */
int dbg_twin_nodes(dbg_t *h, const uint32_t *node_id, uint64_t n, uint32_t *twin_id, uint8_t *canonical);

/* edge_count_table[(k+1)-mer] for n ASCII (k+1)-mers laid back to back; 0 = absent
 * (II_assembleFromReads.py:17 indexes it per contig window)                                 */
int dbg_lookup_edges(dbg_t *h, const uint8_t *kplus1mers, uint64_t n, uint32_t *counts);
/* the same in batch batch[i] of a graph built by dbg_build_batched */
int dbg_lookup_edges_batched(dbg_t *h, const uint8_t *kplus1mers, const uint32_t *batch, uint64_t n,
                             uint32_t *counts);

/* ---- output_contigs (debruijn.py:326-347) ----
 * stop_at_branch = 1 when branch_kmer is the list construct_graph returned with final=False,
 * 0 when it is empty (final=True).  Contigs come out in the reference's order.             */
int dbg_walk(dbg_t *h, int stop_at_branch, uint64_t *n_contigs);
int dbg_contigs(dbg_t *h, dbg_contig_t *out);                      /* n_contigs entries      */
int dbg_segments(dbg_t *h, dbg_segment_t *out, uint64_t *n_segments); /* out may be NULL      */
/* contig strings idx[0..n) written back to back into buf at off[i] (host pointers);
 * string i has length k + n_nodes - 1, no terminator                                        */
int dbg_materialize(dbg_t *h, const uint64_t *idx, uint64_t n, char *buf, const uint64_t *off);
/* getScore (II_assembleFromReads.py:14-18) of contigs idx[0..n): sum of edge counts along
 * the path                                                                                  */
int dbg_contig_scores(dbg_t *h, const uint64_t *idx, uint64_t n, uint64_t *scores);

/* ---- contig concatenation, III_concatenateOutpus.py:57-142 ----
 * The reference's randomised greedy merge (`for k in 20..6: up to 100000 times: current = random.choice(...)`,
 * :57-137) as ONE persistent kernel per k: the contigs live in a byte arena in HBM, Python's Mersenne Twister runs
 * on the device, no host round trip between iterations.  The handle only lends its device, stream and error text.
 *   dbg_concat_begin    contigs (any bytes) = concatenated bytes + n + 1 offsets, HOST pointers; slot i = contig i.
 *   dbg_concat_run      one k (:58-138).  mt_state = the 625 words of random.Random.getstate()[1] (624 state words +
 *                       position), read and written back: the draws are exactly those of `random.choice` (:65) on
 *                       that generator.  Returns with status
 *                         DBG_CONCAT_DONE         the k is finished; n_alive = len(contigs), max_len = the longest
 *                         DBG_CONCAT_NEED_CHOICE  the iteration in flight met swap-tolerant matches (:85-131), which
 *                                                 the reference decides by findSupportReadScore of two candidate
 *                                                 strings: read the matches with dbg_concat_pending (slot order =
 *                                                 the reference's print order; `current` = the slot drawn), the
 *                                                 strings with dbg_concat_lengths / dbg_concat_strings, answer with
 *                                                 dbg_concat_choose and call dbg_concat_run(resume = 1).
 *   dbg_concat_pending  the n_pending matched slots of that iteration, ascending, with their codes
 *                       (1 head-on-tail :76, 2 tail-on-head :80, 3 swapped head-on-tail :85, 4 swapped tail-on-head :108).
 *   dbg_concat_choose   one byte per pending match; for codes 3 / 4: 1 = result_1, 2 = result_2 (:91-98, :114-121).
 *   dbg_concat_lengths  length of every slot's current string and whether the slot is still in the list.
 *   dbg_concat_strings  strings of the given slots back to back at bytes + off[j] (off[j+1] - off[j] = its length).
 *   dbg_concat_dedup    :139-142: keep[i] = 1 iff slot i is in the list and no earlier slot's string contains it.
 */
#define DBG_CONCAT_DONE 1u
#define DBG_CONCAT_NEED_CHOICE 2u
typedef struct {
    uint32_t status;         /* DBG_CONCAT_*                                                            */
    uint32_t current;        /* NEED_CHOICE: the slot random.choice drew                                */
    uint64_t iterations;     /* iterations of this k finished so far                                    */
    uint32_t n_alive;        /* len(contigs)                                                            */
    uint32_t max_len;        /* len(max(contigs, key=len))                                              */
    uint32_t n_pending;      /* NEED_CHOICE: matches of the iteration in flight                         */
    uint32_t pad;
    uint64_t arena_bytes;    /* bytes of string storage in use                                          */
} dbg_concat_status_t;
int dbg_concat_begin(dbg_t *h, const uint8_t *contigs, const uint64_t *contig_off, uint64_t n_contigs);
int dbg_concat_run(dbg_t *h, int k, uint64_t max_iter, uint32_t *mt_state /* 625 */, int resume, dbg_concat_status_t *out);
int dbg_concat_pending(dbg_t *h, uint32_t *slots, uint8_t *codes);
int dbg_concat_choose(dbg_t *h, const uint8_t *choice);
int dbg_concat_lengths(dbg_t *h, uint32_t *len, uint8_t *alive);
int dbg_concat_strings(dbg_t *h, const uint32_t *slots, uint64_t n, uint8_t *bytes, const uint64_t *off);
int dbg_concat_dedup(dbg_t *h, uint8_t *keep);
int dbg_concat_end(dbg_t *h);

#ifdef __cplusplus
}
#endif
#endif /* DBG_B200_H */
