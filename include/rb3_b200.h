/*
 * rb3_b200.h -- C ABI of the B200-native BWT-merge engine (librb3b200.so).
 *
 * Drop-in boundary for the merge path of `ropebwt3 build` (lh3/ropebwt3
 * v3.10-r281).  Plain pointers and sizes only; every entry point names the
 * reference interface it replaces (file:line relative to the reference tree).
 * The dynamic B+-tree rope (mrope_t) is replaced by an opaque, device-resident
 * static index that is rebuilt by a streaming merge at every batch.
 *
 * Conventions: symbols are nt6 codes 0..5 = $ACGTN (io.c:12-21).  All functions
 * returning int return 0 on success and a negative code on failure; the message
 * is available from rb3b_last_error().  There is NO CPU fallback: without a
 * usable CUDA device every call fails with RB3B_ENODEV.
 *
 * Threading: like the reference, an index must not be used from two
 * threads at once; different indices may be.  Every call runs in the execution
 * context (device, stream, scratch arena, counters) that is current on the
 * calling host thread; a thread that never made one current gets its own
 * default context on first use, so calls from different threads run on
 * different streams and overlap on the device -- the reference's pipeline mode
 * (kt_pipeline of rb3_build_sais and rb3_fmi_merge_plain, build.c:55-83) is two
 * host threads calling rb3b_build_bwt_dev and rb3b_merge_plain_dev.
 */
#ifndef RB3_B200_H
#define RB3_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB3B_ASIZE 6

#define RB3B_OK        0
#define RB3B_ENODEV   -1   /* no CUDA device / CUDA runtime error */
#define RB3B_ENOMEM   -2
#define RB3B_EINVAL   -3   /* bad argument (symbol >= 6, missing sentinel, ...) */
#define RB3B_EIO      -4
#define RB3B_EFORMAT  -5

typedef struct rb3b_index_s rb3b_index_t;
typedef struct rb3b_ctx_s rb3b_ctx_t;
typedef struct rb3b_batch_s rb3b_batch_t;

/* ---- runtime -------------------------------------------------------------- */
int         rb3b_init(int device);                 /* select device, create stream + memory pool */
const char *rb3b_last_error(void);
const char *rb3b_version(void);
int         rb3b_set_stream(void *cuda_stream);    /* run on a caller-owned cudaStream_t (NULL = own stream) */
int         rb3b_sync(void);
int         rb3b_trim(void);                       /* give the calling context's scratch memory back to the device */
/* explicit contexts: several devices or streams driven from one process (one host thread per context at a time) */
rb3b_ctx_t *rb3b_ctx_create(int device);
int         rb3b_ctx_make_current(rb3b_ctx_t *ctx);   /* bind to the calling thread; NULL = the thread's default context */
void        rb3b_ctx_destroy(rb3b_ctx_t *ctx);
int         rb3b_set_param(const char *key, int64_t value);   /* tuning knobs, all optional: "seg_len" (walk slice length, 0 = by batch size), "fine_len", "halo_segments", "index_kind", "bitmap_max_symbols", "fmd_threads", ... */
int64_t     rb3b_get_stat(const char *key);        /* counters of the last call: "kernel_launches", "n_segments", "fix_rounds", "unresolved_rows", "n_blocks" ... */

/* ---- index life cycle (replaces mr_init / mr_destroy, mrope.c:15-34) ------- */
rb3b_index_t *rb3b_index_create(void);
void          rb3b_index_destroy(rb3b_index_t *idx);
/* optional size hint (like mr_init's pool pre-sizing has no equivalent in the reference): the index is expected to grow
 * to n_symbols; device buffers are sized once instead of being regrown by the merges */
int           rb3b_index_reserve(rb3b_index_t *idx, int64_t n_symbols);

/* sorting order of the collection = third argument of mr_init (mrope.c:15, mrope.h:6-8): 0 input order (default),
 * 1 RLO, 2 RCLO.  Only rb3b_insert_multi honours it; it is stored in / restored from the .fmr header. */
int           rb3b_index_set_order(rb3b_index_t *idx, int sorting_order);
int           rb3b_index_get_order(const rb3b_index_t *idx);

/* ---- building blocks of `build` -------------------------------------------- */
/* rb3_enc_plain2fmr (fm-index.c:114-137): first batch, BWT in host memory. */
int rb3b_index_from_plain(rb3b_index_t *idx, int64_t len, const uint8_t *bwt);
/* same with the BWT already resident in device memory */
int rb3b_index_from_plain_dev(rb3b_index_t *idx, int64_t len, const uint8_t *d_bwt);
/* rb3_enc_fmd2fmr (fm-index.c:56-85) / mr_restore (mrope.c:161): load a run list (need not be coalesced). */
int rb3b_index_from_runs(rb3b_index_t *idx, int64_t n_runs, const uint8_t *sym, const int64_t *len);
int rb3b_index_from_runs_device(rb3b_index_t *idx, int64_t n_runs, const uint8_t *d_sym, const int64_t *d_len);   /* run list in device memory */

/* rb3_fmi_merge_plain (fm-index.c:279-303): merge the partial BWT `bwt` of a new
 * batch (host memory, caller-owned) into the index in place. */
int rb3b_merge_plain(rb3b_index_t *idx, int64_t len, const uint8_t *bwt);
/* Optional: start copying a batch to the device now, for a later rb3b_merge_plain(idx, len, host_bwt) or
 * rb3b_batch_prepare(len, host_text) with the SAME pointer and length, so that the copy of batch i+1 overlaps the merge of
 * batch i (the reference overlaps reading batch i+1 with merging batch i the same way: kt_pipeline, build.c:55-83,186-201).
 * Up to two batches may be in flight; the host buffer must stay unchanged until the consuming call returns and should be
 * pinned (rb3b_host_alloc_pinned) for the copy to be asynchronous.  Not consuming a prefetched batch is harmless. */
int rb3b_prefetch_batch(int64_t len, const uint8_t *host);
/* same, partial BWT already in device memory (no host<->device copies) */
int rb3b_merge_plain_dev(rb3b_index_t *idx, int64_t len, const uint8_t *d_bwt);

/* mr_insert_multi (mrope.c:300-385), the `build -2/-s/-r` path (build.c:214-218): insert the strings of `text`
 * (concatenated 0-terminated nt6 strings exactly as rb3_seq_read delivers them; the reference reverses them first with
 * rb3_reverse_all, this call does not need that) into the index in the index's sorting order. */
int rb3b_insert_multi(rb3b_index_t *idx, int64_t len, const uint8_t *text);
int rb3b_insert_multi_dev(rb3b_index_t *idx, int64_t len, const uint8_t *d_text);

/* rb3_mg_rank_plain (fm-index.c:202-225): the interleave array only.  rb[i] =
 * (ka+i)<<6 | B[i]<<3 | bucket(i) exactly as fm-index.c:168; acc[7] = C[] of the batch. */
int rb3b_mg_rank_plain(const rb3b_index_t *idx, int64_t len, const uint8_t *bwt, int64_t *rb, int64_t acc[RB3B_ASIZE + 1]);
int rb3b_mg_rank_plain_dev(const rb3b_index_t *idx, int64_t len, const uint8_t *d_bwt, int64_t *d_rb, int64_t acc[RB3B_ASIZE + 1]);

/* Multi-device building blocks.  The reference parallelises rb3_mg_rank_plain over the new sequences with kt_for
 * (fm-index.c:220); here the index is replicated and device `part` of `n_parts` computes the interleave positions of
 * its share of every sequence.  d_ka[len] (device): position of every row this part resolved, -1 elsewhere; the
 * element-wise MAX over all parts (ncclAllReduce) is the array of rb3b_mg_rank_plain without the packing.
 * Returns 0; 1 when the part could not resolve all of its rows locally (fall back to part 0 of 1); < 0 on error. */
int rb3b_mg_rank_part(const rb3b_index_t *idx, int64_t len, const uint8_t *d_bwt, int part, int n_parts, int64_t *d_ka);
/* second half of rb3_fmi_merge_plain (worker_mgins, fm-index.c:237-249 / :295) given the complete d_ka[len] */
int rb3b_merge_with_ka(rb3b_index_t *idx, int64_t len, const uint8_t *d_bwt, const int64_t *d_ka);

/* Multi-device communicator (NCCL over NVLink, bound at run time): the calling thread's context becomes rank `rank` of
 * `world`.  id128: 128 bytes from rb3b_dist_unique_id on one rank, passed to all by the host program (MPI, a file,
 * torch.distributed, or plain memory between the threads of one process). */
int rb3b_dist_unique_id(void *id128);
int rb3b_dist_init(int rank, int world, const void *id128);
int rb3b_dist_finalize(void);
int rb3b_dist_rank(void);
int rb3b_dist_world(void);
int rb3b_dist_nccl_version(void);

/* rb3_fmi_merge_plain (fm-index.c:279-303) across the ranks of the communicator: every rank holds a replica of the index
 * and passes the same batch; the rank phase -- what the reference spreads over host threads with kt_for (fm-index.c:220) --
 * is split over the devices, the interleave positions are exchanged over NVLink and every replica is merged.  Returns 0,
 * 1 when the ranks had to fall back to the unsharded rank phase (still exact), < 0 on error. */
int rb3b_merge_plain_dist_dev(rb3b_index_t *idx, int64_t len, const uint8_t *d_bwt);
/* same with the batch in host memory visible to every rank: each rank copies 1/world of it over PCIe, NVLink does the rest */
int rb3b_merge_plain_dist(rb3b_index_t *idx, int64_t len, const uint8_t *bwt);

/* rb3_fmi_merge (fm-index.c:251-277) for `ropebwt3 merge`: B is another index. */
int rb3b_merge_index(rb3b_index_t *idx, const rb3b_index_t *other);

/* ---- queries --------------------------------------------------------------- */
/* mr_rank1a (mrope.h:70, mrope.c:71-121) / rld_rank1a (rld0.c:416-437), batched:
 * ok[q*6+c] = |{i<k[q] : B[i]=c}|, sym[q] = B[k[q]] or -1 when k[q] >= n. */
int rb3b_rank1a(const rb3b_index_t *idx, int64_t nq, const int64_t *k, int64_t *ok, int8_t *sym);
int rb3b_rank1a_dev(const rb3b_index_t *idx, int64_t nq, const int64_t *d_k, int64_t *d_ok, int8_t *d_sym);
/* LF-walk flavour used by the merge: out[q] = C[c[q]] + rank(c[q], k[q]) */
int rb3b_lf_dev(const rb3b_index_t *idx, int64_t nq, const int64_t *d_k, const uint8_t *d_c, int64_t *d_out, int variant);

/* rb3_fmi_get_acc (fm-index.c:544-550): acc[c] = #symbols < c; returns total length */
int64_t rb3b_get_acc(const rb3b_index_t *idx, int64_t acc[RB3B_ASIZE + 1]);
int64_t rb3b_index_bytes(const rb3b_index_t *idx);   /* device bytes held by the index */
/* With rb3b_set_param("async_merge", 1), rb3b_merge_plain[_dev] into a bitmap index only QUEUES the streaming merge (on the
 * context's second stream) and returns once the interleave positions are known: the next call overlaps with it until it
 * needs the merged cells.  The index's host-side state (rb3b_get_acc ...) is already that of the merged index.
 * rb3b_index_wait blocks until the merge has finished and reports what its kernels found (a batch that is not a BWT leaves
 * the index unusable).  Every other entry point waits by itself.  Off by default (measured gain: 3 %). */
int     rb3b_index_wait(rb3b_index_t *idx);

/* ---- export ---------------------------------------------------------------- */
/* canonical (coalesced) run list; call with sym==NULL to get the count. mr_itr_next_block loop (fm-index.c:41-51). */
int64_t rb3b_export_runs(const rb3b_index_t *idx, uint8_t *sym, int64_t *len, int64_t cap);
/* rb3_enc_fmr2fmd + rld_dump (fm-index.c:31-54, rld0.c:222-243): byte-identical .fmd; fn "-" = stdout */
int rb3b_dump_fmd(const rb3b_index_t *idx, const char *fn);
/* mr_dump (mrope.c:152-159): a legal .fmr (loadable by mr_restore and extendable by the CPU code) */
int rb3b_dump_fmr(const rb3b_index_t *idx, const char *fn, int max_nodes, int block_len);
/* mr_print_bwt (mrope.c:195-210): plain text + '\n' */
int rb3b_dump_plain(const rb3b_index_t *idx, const char *fn);
/* rb3_fmi_restore (fm-index.h:123-133): sniff FMD magic first, then FMR */
int rb3b_restore(rb3b_index_t *idx, const char *fn);
/* host-only encoders behind the two dumps (no device needed): run list -> malloc'd file image, returns its size.
 * rld_enc/rld_enc_finish/rld_rank_index/rld_dump (rld0.c:137-243) and mr_dump/rope_dump_node (mrope.c:152, rope.c:265-287). */
int64_t rb3b_fmd_image(int64_t n_runs, const uint8_t *sym, const int64_t *len, uint8_t **out);
int64_t rb3b_fmr_image(int64_t n_runs, const uint8_t *sym, const int64_t *len, int max_nodes, int block_len, uint8_t **out);
/* host-only reader behind rb3b_restore (rld_restore + the decode loop of rb3_enc_fmd2fmr, fm-index.c:56-85; mr_restore,
 * mrope.c:161-177): a .fmd or .fmr file image -> canonical run list in malloc'd arrays (rb3b_host_free), returns the
 * number of runs; *sorting_order = the .fmr's order byte (0 for .fmd) */
int64_t rb3b_runs_from_image(const uint8_t *image, int64_t n_bytes, uint8_t **sym, int64_t **len, int *sorting_order);
void    rb3b_host_free(void *p);

/* ---- sampled suffix array (`ropebwt3 ssa`, ssa.c) ---------------------------------- */
/* rb3_ssa_gen (ssa.c:55-81): sizes of the two arrays for sample shift ss: m = #strings, n_ssa, ms = bits of a string id */
int rb3b_ssa_sizes(const rb3b_index_t *idx, int ssa_shift, int64_t *m, int64_t *n_ssa, int *ms);
/* rb3_ssa_gen into device arrays d_r2i[m], d_ssa[n_ssa] with the layout of rb3_ssa_t (fm-index.h, ssa.c:28-38) */
int rb3b_ssa_gen_dev(const rb3b_index_t *idx, int ssa_shift, uint64_t *d_r2i, uint64_t *d_ssa);
/* rb3_ssa_gen + rb3_ssa_dump (ssa.c:198-213): byte-identical .ssa file; fn "-" = stdout */
int rb3b_ssa_dump(const rb3b_index_t *idx, int ssa_shift, const char *fn);

/* ---- partial BWT of a batch (rb3_build_sais, sais-ss.c:50-56) on the GPU ----- */
/* text: concatenated 0-terminated nt6 strings (host); bwt_out: host, len bytes (may alias text). */
int rb3b_build_bwt(int64_t len, const uint8_t *text, uint8_t *bwt_out);
int rb3b_build_bwt_dev(int64_t len, const uint8_t *d_text, uint8_t *d_bwt_out);
/* same in RLO (1) / RCLO (2) order: the BWT that mr_insert_multi (mrope.c:300) produces for this batch on an empty rope */
int rb3b_build_bwt_so(int64_t len, const uint8_t *text, int sorting_order, uint8_t *bwt_out);
int rb3b_build_bwt_so_dev(int64_t len, const uint8_t *d_text, int sorting_order, uint8_t *d_bwt_out);

/* Prepared batches: the two steps of the reference's pipeline mode (kt_pipeline in build.c:55-83 -- step 0 rb3_build_sais,
 * step 1 rb3_fmi_merge_plain / rb3_enc_plain2fmr).  Step 0 needs no index, so a host thread with its own context can
 * prepare batch i+1 while another merges batch i.  The prepared batch holds the partial BWT AND the batch in walk order
 * (its text read backwards with the inverse suffix array alongside, which the suffix sort yields for free), so step 1 skips
 * the LF chase the BWT-only seam call needs.  text: concatenated 0-terminated nt6 strings, as for rb3b_build_bwt. */
rb3b_batch_t  *rb3b_batch_prepare(int64_t len, const uint8_t *text);
rb3b_batch_t  *rb3b_batch_prepare_dev(int64_t len, const uint8_t *d_text);
void           rb3b_batch_destroy(rb3b_batch_t *b);
int64_t        rb3b_batch_len(const rb3b_batch_t *b);
const uint8_t *rb3b_batch_bwt_dev(const rb3b_batch_t *b);   /* the partial BWT in device memory (owned by the batch) */
int            rb3b_merge_prepared(rb3b_index_t *idx, const rb3b_batch_t *b);

/* Largest batch (in symbols) that fits the device next to an index of index_symbols: lets a caller clamp -m (build.c:39,
 * default 7G) to the device.  The .fmd is the same for any batching. */
int64_t rb3b_max_batch_symbols(int64_t index_symbols);

/* ---- device memory helpers for callers without a CUDA runtime of their own -- */
void *rb3b_dev_alloc(int64_t bytes);
void  rb3b_dev_free(void *p);
void *rb3b_host_alloc_pinned(int64_t bytes);   /* page-locked host memory: batch buffers that rb3b_h2d copies at full PCIe rate */
void  rb3b_host_free_pinned(void *p);
int   rb3b_h2d(void *dst, const void *src, int64_t bytes);
int   rb3b_d2h(void *dst, const void *src, int64_t bytes);

#ifdef __cplusplus
}
#endif
#endif
