/*
 * minialign_b200.h -- C ABI of the B200-native read->reference mapping hot path (drop-in for minialign's
 * per-batch mapper).  Plain pointers and sizes only; no CUDA or torch types cross this boundary.
 *
 * What each entry point replaces in the reference (ocxtal/minialign @ /root/reference):
 *
 *   mab_init / mab_destroy      mm_align_init / mm_align_destroy            minialign.c:4671-4718, 4651-4664
 *                               (+ gaba_init / gaba_dp_init,                gaba_wrap.h:245-295, gaba.c:3848-3928)
 *   mab_map_batch               pt_worker_t mm_align_worker: for each read  minialign.c:4589-4601
 *                               of a bseq_t batch call mm_align_seq          minialign.c:4427-4474
 *   mab_result / release,       mm_reg_t / mm_aln_t / gaba_alignment_t views minialign.c:3260-3267, gaba.h:193-219
 *   mab_detach_batch /          that mm_align_drain_intl hands the printer   minialign.c:4607-4626
 *   mab_results_get / _free     (detached = owned by the caller like the batch's lmm arena, 4615-4623)
 *   mab_sketch                  mm_sketch                                   minialign.c:2410-2435
 *   mab_seed_chain              mm_seed + mm_chain                          minialign.c:3500-3541, 3702-3721
 *   mab_extend_pair             the body of the mm_extend loop:             minialign.c:4134-4154
 *                               gaba_dp_fill_root/fill/search_max/trace     gaba.c:2110-2203, 2776-2817, 3372-3393
 *
 * Ownership mirrors the reference: the caller owns the index blob and the read block; results of one batch live in
 * the context until mab_release_batch (the reference frees its lmm arena in the drain, minialign.c:4615-4623).
 * One context per GPU, one host thread per context.  Errors are negative return codes; nothing aborts.
 * There is NO CPU fallback: every call fails with MAB_ENODEV when no sm_100 device is usable.
 */
#ifndef MINIALIGN_B200_H
#define MINIALIGN_B200_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAB_OK        0
#define MAB_ENODEV   -1		/* no usable CUDA device / CUDA runtime error (message via mab_last_error) */
#define MAB_EINVAL   -2		/* bad argument / unsupported parameter set (e.g. non-combined gap model) */
#define MAB_ENOMEM   -3		/* device or host allocation failed */
#define MAB_EOVERFLOW -4	/* a batch-level device buffer (result pool, workspace) overflowed even after it was grown and the batch re-run */
#define MAB_EFORMAT  -5		/* text path: a record layout the device reader does not take (see mab_text_begin) */

/* mapping parameters: the subset of mm_align_params_t (minialign.c:2517-2524) + gaba_params_t (gaba.h:90-110) the hot
 * path reads.  k, w, b, occ[] come from the index itself. */
typedef struct {
	int32_t wlen, glen;			/* chainable window edge length, linkable gap length (-W, -G; default 7000) */
	uint32_t min_score;			/* -s */
	float min_ratio;			/* -m */
	int8_t score_matrix[16];	/* [a | b << 2] */
	int8_t gi, ge, gfa, gfb;	/* positive penalties; combined (piecewise-affine) model: all non-zero */
	int8_t xdrop;				/* -Y */
	uint8_t _pad[3];
	uint32_t flags;				/* MAB_FLAG_* */
} mab_params_t;

/* the caller keeps the index image alive and unchanged for the lifetime of the context (and its clones): no host copy is made */
#define MAB_FLAG_BORROW_INDEX 1u

typedef struct mab_ctx mab_ctx;

/* blob = raw (inflated) .mai payload after the 12-byte {magic,size} header: the relocatable mm_idx_t image
 * (minialign.c:3070-3167).  Copied to HBM; the host copy is only read during the call. device = CUDA ordinal. */
mab_ctx *mab_init(const void *mai_blob, uint64_t size, const mab_params_t *params, int device);
/* another context on the same device sharing the parent's index image (several batches in flight per GPU cost one index copy);
 * the parent must outlive its clones.  One host thread per context at a time. */
mab_ctx *mab_clone(mab_ctx *parent);
/* Staged set-up, for a host that produces the image piece by piece (the .mai container is a sequence of independently deflated
 * 1 MiB frames, minialign.c:1137-1290, 3128-3167): every piece goes to all the listed devices while the rest is still being
 * inflated, so a human-sized index (15-17 GB) is on the GPUs when the last frame is done.
 *   ld = mab_load_begin(max_size, params, devices, n)     device buffers of max_size bytes, a ring of page-locked staging slots
 *   mab_load_put(ld, offset, src, n)                       any thread, any order; returns when src may be reused
 *   mab_load_end(ld, blob, size, ctx_out)                  waits for the copies; one context per device (as from mab_init) in ctx_out[n]
 *   mab_load_abort(ld)                                     instead of mab_load_end: drops everything
 * mab_load_end consumes the loader whatever it returns. */
typedef struct mab_loader mab_loader;
mab_loader *mab_load_begin(uint64_t max_size, const mab_params_t *params, const int *devices, int n_devices);
int mab_load_put(mab_loader *ld, uint64_t offset, const void *src, uint64_t n);
int mab_load_end(mab_loader *ld, const void *blob, uint64_t size, mab_ctx **ctx_out);
void mab_load_abort(mab_loader *ld);
void mab_destroy(mab_ctx *ctx);
const char *mab_last_error(void);

/* index facts the host printer needs (mm_idx_seq_t, minialign.c:2464-2470) */
uint32_t mab_n_ref(const mab_ctx *ctx);
int mab_ref_info(const mab_ctx *ctx, uint32_t rid, const char **name, uint32_t *l_name, uint32_t *l_seq, const uint8_t **seq);
int mab_index_params(const mab_ctx *ctx, uint32_t *k, uint32_t *w, uint32_t *b, uint32_t *n_occ, uint32_t *occ /* [7] */);

/* Map one batch.  seq_block holds n_seq reads, 1 byte/base codes A,C,G,T = 0..3, N = 4 (minialign.c:214-232), read i at
 * seq_block[seq_ofs[i] .. seq_ofs[i] + seq_len[i]).  Host memory (pinned or pageable).
 * On success the batch's results are readable through mab_result() until mab_release_batch(). */
int mab_map_batch(mab_ctx *ctx, const uint8_t *seq_block, uint64_t block_size,
	const uint64_t *seq_ofs, const uint32_t *seq_len, uint32_t n_seq);

/* Result of read i as a flat little-endian u32 stream, the same layout the reference harness and the oracle dump:
 *   [0] n_all  [1] n_uniq   then n_all alignments, each
 *   16 x u32 header : score(lo,hi) identity(double bits lo,hi) agcnt bgcnt dcnt slen plen npathwords rank mapq 0 0 0 0
 *   slen x 8 x u32  : aid bid apos bpos alen blen ppos(lo,hi)          (gaba_segment_s, gaba.h:193-198)
 *   npathwords x u32: path bit string, LSB first, sentinel bit at plen (gaba.h:205-219)
 * Returns the number of u32 words (0 = unmapped, the reference's NULL mm_reg_t). */
uint64_t mab_result(const mab_ctx *ctx, uint32_t i, const uint32_t **words);
void mab_release_batch(mab_ctx *ctx);

/* Detach the results of the last batch from the context: the returned object owns them (like the batch's lmm arena in the
 * reference, freed by the drain, minialign.c:4615-4623), stays valid across further mab_map_batch calls and may be read from
 * another thread, so that printing batch i overlaps mapping batch i + 1.  mab_results_get has the semantics of mab_result. */
typedef struct mab_results mab_results;
mab_results *mab_detach_batch(mab_ctx *ctx);
uint64_t mab_results_get(const mab_results *r, uint32_t i, const uint32_t **words);
void mab_results_free(mab_results *r);

/* The one word of state the reference's worker thread carries from read to read (`rlen`, see the text path below): mab_map_batch
 * chains it through the context; a caller that spreads consecutive batches over several contexts moves it by hand. */
uint32_t mab_get_rlen(const mab_ctx *ctx);
void mab_set_rlen(mab_ctx *ctx, uint32_t rlen);

/* device-side statistics of the last batch (for bench.py / roofline): kernel milliseconds measured with CUDA events
 * on the context's stream, DP vectors filled, bytes moved each way */
typedef struct {
	float ms_total, ms_h2d, ms_seed, ms_sortchain, ms_extend, ms_d2h, ms_post;
	float ms_extend_r0;			/* the round-0 k_extend launch alone (the dominant kernel), CUDA events on its stream */
	uint64_t n_vectors;			/* anti-diagonal vectors filled (down + up + replays) */
	uint64_t n_fill_calls, n_trace;
	uint64_t h2d_bytes, d2h_bytes;
	uint32_t n_launches;		/* kernels launched */
	uint32_t n_retry;
	float ms_wall, ms_wall_sizing, ms_wall_wait;	/* host wall clock: whole call, workspace sizing between the two device phases, blocked in stream syncs */
	float ms_wall_submit;						/* host wall clock spent issuing copies and launches (driver calls that should not block) */
	uint32_t n_failed;							/* reads given up on because a per-read device structure overflowed (reported unmapped) */
	uint32_t _pad;
} mab_stats_t;
int mab_last_stats(const mab_ctx *ctx, mab_stats_t *out);

/* Memory: free / total bytes of the context's device; the HBM the DP arenas of this context may take (default 40 GB; with very
 * long reads or many contexts per GPU fewer warps stay resident instead of over-allocating: 8 MB per resident warp at 25 kb reads). */
int mab_device_memory(const mab_ctx *ctx, uint64_t *free_bytes, uint64_t *total_bytes);
void mab_set_arena_budget(mab_ctx *ctx, uint64_t bytes);

/* When set, mab_map_batch takes seq_block as a DEVICE pointer already resident in HBM (bench.py's kernel-only arm). */
int mab_set_device_input(mab_ctx *ctx, int on);

/* ---- text path: FASTA / FASTQ bytes in, SAM bytes out ----------------------------------------------------------------------
 * One call sequence replaces, for one chunk of the input file, the reference's source -> worker -> drain pipeline stage
 * (mm_align_source: bseq_read, minialign.c:4565-4583, 1996-2164; mm_align_worker, 4589-4601; mm_align_drain_intl ->
 * mm_print_mapped, 4607-4626, 5095-5426): the records are parsed, mapped, post-processed and printed on the device; the host
 * hands over text and gets text back.
 *
 * text: n_bytes of FASTA or FASTQ holding whole records (the caller cuts the file at record boundaries), starting with the
 * delimiter ('>' or '@').  Taken as is: '\n' line ends, FASTA sequences on any number of lines, FASTQ with four lines per
 * record; anything else gives MAB_EFORMAT and the caller falls back to its own reader + mab_map_batch.
 * flags: the tag bits of mab_sam.h (MAB_TAG_*, MAB_OMIT_REP) | MAB_TEXT_* below.
 *
 * The reference worker thread carries one word of state from read to read (`rlen`, minialign.c:3865 vs 3873; with -t1 that is
 * file order).  Chunks mapped concurrently (several contexts, several GPUs) stay byte-identical to the -t1 run by passing it on:
 *   mab_text_begin(ctx, ..., rlen_prev, rlen_known = 0)   map without knowing what the previous chunk left behind
 *   mab_text_commit(ctx, rlen_prev, &rlen_next)           ... once the previous chunk's value is known (re-maps the rare read
 *                                                         whose first seed test depended on it); rlen_next goes to the next chunk
 *   mab_text_finish(ctx, ...)                             post-processing + SAM formatting + copy out
 * A single context used sequentially calls mab_map_text, which chains the value through the context. */
enum { MAB_TAG_RG = 1 << 0, MAB_TAG_NH = 1 << 2, MAB_TAG_IH = 1 << 3, MAB_TAG_AS = 1 << 4, MAB_TAG_XS = 1 << 5, MAB_TAG_NM = 1 << 6, MAB_TAG_SA = 1 << 7, MAB_TAG_MD = 1 << 8,
       MAB_OMIT_REP = 1 << 30 };		/* optional SAM tags (-T, minialign.c:2527-2537) and -R */
#define MAB_TEXT_KEEP_QUAL   0x01000000u	/* -Q: print FASTQ qualities (default: '*', minialign.c:6145, 5134, 5187) */
#define MAB_TEXT_DEVICE_OUT  0x02000000u	/* leave the SAM text on the device (kernel-side measurements): no copy out */
typedef struct {
	uint64_t n_reads, n_bases;			/* records kept (non-empty sequence) and their bases */
	uint64_t sam_bytes;					/* valid after mab_text_finish */
	uint32_t rlen_valid, rlen_next;		/* the value this chunk leaves behind (rlen_valid = 0: it loaded no chain, pass the previous one on) */
} mab_text_info_t;
/* optional, before the first chunk: the largest chunk the caller will pass (the reference's -N / batch size, minialign.c:6145).  The
 * context and its clones size their device buffers for it at once; without it they follow the chunks they see, and growing a buffer
 * later waits for the whole device. */
int mab_text_reserve(mab_ctx *ctx, uint64_t max_chunk_bytes);
int mab_text_begin(mab_ctx *ctx, const char *text, uint64_t n_bytes, uint32_t flags, uint32_t rlen_prev, int rlen_known, mab_text_info_t *info);
int mab_text_commit(mab_ctx *ctx, uint32_t rlen_prev, mab_text_info_t *info);
/* sam_out = NULL: the text is left in a pinned buffer owned by the context, *sam_ptr points at it (valid until the next begin) */
int mab_text_finish(mab_ctx *ctx, char *sam_out, uint64_t sam_cap, const char **sam_ptr, mab_text_info_t *info);
int mab_map_text(mab_ctx *ctx, const char *text, uint64_t n_bytes, uint32_t flags, char *sam_out, uint64_t sam_cap, const char **sam_ptr, mab_text_info_t *info);
/* the SAM header (@HD, @SQ, @PG; minialign.c:5095-5120) for this context's index; returns the length, writes at most cap bytes */
uint64_t mab_sam_header_text(const mab_ctx *ctx, const char *version, const char *cmdline, char *out, uint64_t cap);
/* page-locked host memory for text / SAM buffers (copies from / to pageable memory are staged by the driver and block) */
void *mab_host_alloc(uint64_t bytes);
void *mab_host_alloc_on(int device, uint64_t bytes);	/* the same from a thread that has not used a device yet (selects `device` first) */
void mab_host_free(void *p);
/* page-lock memory of the caller's own (page-aligned; touch it first, on as many threads as you like: that is the slow part and
 * unlike mab_host_alloc it does not hold up the other CUDA calls of the process) */
int mab_host_register(int device, void *p, uint64_t bytes);
void mab_host_unregister(void *p);

/* ---- stage-level entry points (parity tests; same semantics as the reference functions named above) ---- */
/* sketch of one read: writes the minimizer words followed by the 4-word cap; returns #words (may exceed cap) */
uint64_t mab_sketch(mab_ctx *ctx, const uint8_t *seq, uint32_t len, uint64_t *out, uint64_t cap);
/* seeds (+ leaves) and roots after rounds 0..round of mm_seed + mm_chain; returns n_seed */
uint64_t mab_seed_chain(mab_ctx *ctx, const uint8_t *seq, uint32_t len, uint32_t round,
	uint32_t *seeds, uint64_t seed_cap, uint64_t *n_total, uint32_t *roots, uint64_t root_cap, uint64_t *n_root);
/* n <= 32767 elements of 16 bytes (key = the first 8) through the reference's unstable sort (ksort.h:82-131) as k_sortchain walks
 * it and through the parallel form k_sort uses: both must give the reference's order, ties included */
int mab_sort_check(mab_ctx *ctx, const uint32_t *elems, uint32_t n, uint32_t *out_exact, uint32_t *out_walk);
/* n independent extension problems: pair i = (a_i, b_i, apos, bpos, brev, narrow); res = 16 u32 per pair, alignment in
 * the flat layout above at aln_out + aln_ofs[i] (aln_ofs[n] = total).  See oracle/ref_harness.c refh_extend. */
typedef struct {
	uint64_t a_ofs, b_ofs;		/* offsets into seq_block */
	uint32_t alen, blen, apos, bpos, brev, narrow;
	int64_t min_score;
} mab_pair_t;
int mab_extend_pairs(mab_ctx *ctx, const uint8_t *seq_block, uint64_t block_size, const mab_pair_t *pairs, uint32_t n,
	uint32_t *res /* 16 x n */, uint32_t *aln_out, uint64_t aln_cap, uint64_t *aln_ofs /* n + 1 */);

/* integer roofline of the DP step: runs the bulk block loop of the fill (the product's own step code) on register-resident
 * synthetic state with k_extend's launch shape, n_blocks x 32 anti-diagonals per warp; *vectors_per_s = achieved ceiling
 * (bench.py reports k_extend against it) */
int mab_fill_peak(mab_ctx *ctx, int masks, uint32_t n_blocks, double *vectors_per_s);

/* runs every packed-SIMD / permute / warp primitive the DP uses on fixed inputs; out = 64 x 32 words (test hook) */
int mab_selftest(mab_ctx *ctx, uint32_t *out);

#ifdef __cplusplus
}
#endif
#endif
