/*
 * ref_gpuworker.c -- the reference-side binding of INTEGRATION.md, built for real (test infrastructure: lives under oracle/,
 * is built into oracle/_ref/minialign-gpuworker, never part of the product path).
 *
 * A translation unit that #includes the UNMODIFIED reference source (ocxtal/minialign, minialign.c, from where it lies under
 * /root/reference, -DREF_SRC) and adds exactly what a maintainer would add to make the reference binary map on the GPU:
 *
 *   gw_align_worker     a pt_worker_t with the signature and contract of mm_align_worker (minialign.c:4589-4601): for every
 *                       read of the bseq_t batch set seq[i].u64 = mm_reg_t* (or NULL), allocated from the batch's lmm arena
 *                       -- but the mapping is ONE call to mab_map_batch() of libminialign_b200.so over the whole batch
 *   gw_reg_from_words   the flat result words of mab_result() -> mm_reg_t / mm_aln_t / gaba_alignment_t in the layout the
 *                       reference's printer and drain expect (minialign.c:3260-3267, 4364-4396; gaba.c:3244-3291)
 *   gw_align_file       mm_align_file (4725-4731) with the worker swapped
 *   main                mm_opt_init + the body of main_align (6365-6440) for a prebuilt .mai index
 *
 * Everything else -- the option parser, the FASTA/FASTQ reader (bseq_read), the thread pool (pt_stream), the strictly ordered
 * drain and the SAM printer -- is the reference's own code.  The SAM it prints must equal the stock binary's (tests/test_gpu_text.py).
 */
#define NAMESPACE ref
#ifndef UNITTEST
#define UNITTEST 0
#endif
#define main ref_main_unused
#include REF_SRC
#undef main

#include <pthread.h>
#include "../include/minialign_b200.h"

static struct {
	pthread_mutex_t mu;
	void const *blob; uint64_t blob_size;
	mab_params_t prm;
	mab_ctx *ctx[MAX_THREADS], *parent;
	int device;
} gw = { .mu = PTHREAD_MUTEX_INITIALIZER };

/* one context per worker thread, like the reference's mm_tbuf_t per thread (4709-4711); clones share the index image */
static mab_ctx *gw_ctx(uint32_t tid)
{
	pthread_mutex_lock(&gw.mu);
	if(gw.ctx[tid] == NULL) {
		gw.ctx[tid] = gw.parent ? mab_clone(gw.parent) : mab_init(gw.blob, gw.blob_size, &gw.prm, gw.device);
		if(gw.ctx[tid] == NULL) { fprintf(stderr, "[E::gw_ctx] %s\n", mab_last_error()); exit(1); }
		if(gw.parent == NULL) { gw.parent = gw.ctx[tid]; }
	}
	pthread_mutex_unlock(&gw.mu);
	return(gw.ctx[tid]);
}

/* flat words of one read (include/minialign_b200.h) -> mm_reg_t in the batch's arena */
static mm_reg_t const *gw_reg_from_words(lmm_t *lmm, uint32_t const *w, uint64_t nw)
{
	if(nw == 0) { return(NULL); }
	uint32_t n_all = w[0], n_uniq = w[1];
	mm_reg_t *reg = lmm_malloc(lmm, sizeof(mm_reg_t) + n_all * sizeof(mm_aln_t *));
	reg->n_all = n_all; reg->n_uniq = n_uniq;
	uint32_t const *p = w + 2;
	for(uint32_t i = 0; i < n_all; i++) {
		uint32_t slen = p[7], plen = p[8], npw = p[9];
		uint64_t pn = (uint64_t)npw + 8;
		/* gaba_alignment_s, path words (zero above the sentinel), segments: the layout of trace_init (gaba.c:3255-3278); lmm_malloc
		 * leaves sizeof(mm_aln_t) of head margin in front of the block (lmm_init_margin, 4579) */
		struct gaba_alignment_s *a = lmm_malloc(lmm, sizeof(struct gaba_alignment_s) + sizeof(uint32_t) * _roundup(pn, 8) + sizeof(struct gaba_segment_s) * (slen + 1));
		memset(a, 0, sizeof(struct gaba_alignment_s) + sizeof(uint32_t) * _roundup(pn, 8));
		memcpy(&a->score, &p[0], 8); memcpy(&a->identity, &p[2], 8);
		a->agcnt = p[4]; a->bgcnt = p[5]; a->dcnt = p[6]; a->slen = slen; a->plen = plen; a->padding = 0x40000000;
		struct gaba_segment_s *seg = (struct gaba_segment_s *)(a->path + _roundup(pn, 8));
		memcpy(seg, p + 16, sizeof(struct gaba_segment_s) * slen);
		a->seg = seg;
		memcpy(a->path, p + 16 + 8 * (uint64_t)slen, sizeof(uint32_t) * npw);
		mm_aln_t *ma = (mm_aln_t *)a - 1;							/* .head_margin = sizeof(mm_aln_t), as in mm_pack_reg (4384) */
		ma->aid = p[10]; ma->mapq = p[11];
		reg->aln[i] = ma;
		p += 16 + 8 * (uint64_t)slen + npw;
	}
	return(reg);
}

/* drop-in for mm_align_worker (4589-4601) */
static void *gw_align_worker(uint32_t tid, void *arg, void *item)
{
	(void)arg;
	mm_align_step_t *s = (mm_align_step_t *)item;
	bseq_t *r = (bseq_t *)s;
	mab_ctx *ctx = gw_ctx(tid);
	uint64_t *ofs = malloc(sizeof(uint64_t) * (r->n_seq + 1));
	uint32_t *len = malloc(sizeof(uint32_t) * (r->n_seq + 1));
	for(uint64_t i = 0; i < r->n_seq; i++) { ofs[i] = (uint64_t)(r->seq[i].seq - (uint8_t *)r->base); len[i] = r->seq[i].l_seq; }
	if(mab_map_batch(ctx, (uint8_t const *)r->base, r->size, ofs, len, r->n_seq) != MAB_OK) { fprintf(stderr, "[E::gw_align_worker] %s\n", mab_last_error()); exit(1); }
	for(uint64_t i = 0; i < r->n_seq; i++) {
		uint32_t const *w = NULL;
		uint64_t nw = mab_result(ctx, (uint32_t)i, &w);
		r->seq[i].u64 = (uintptr_t)gw_reg_from_words(s->lmm, w, nw);
	}
	mab_release_batch(ctx);
	free(ofs); free(len);
	return(s);
}

/* mm_align_file (4725-4731) with the worker swapped */
static int gw_align_file(mm_align_t *b, bseq_file_t *fp, mm_print_t *pr)
{
	if(fp == NULL || pr == NULL) { return(-1); }
	b->fp = fp; b->pr = pr;
	pt_stream(b->pt, b, mm_align_source, gw_align_worker, mm_align_drain);
	return(fp->is_eof > 2 ? 1 : 0);
}

/* the raw (inflated) .mai payload: what mab_init takes (the reference's own container reader, minialign.c:1295-1502) */
static void *gw_load_blob(char const *path, pt_t *pt, uint64_t *size)
{
	pg_t *pg = pg_init(fopen(path, "rb"), pt);
	if(pg == NULL) { return(NULL); }
	struct { uint32_t magic; uint32_t pad; uint64_t size; } h;
	uint8_t hdr[12];
	if(pgread(pg, hdr, 12) != 12) { pg_destroy(pg); return(NULL); }
	memcpy(&h.magic, hdr, 4); memcpy(&h.size, hdr + 4, 8);
	uint8_t *blob = malloc(h.size + 64);
	if(pgread(pg, blob, h.size) != h.size) { free(blob); pg_destroy(pg); return(NULL); }
	pg_destroy(pg);
	*size = h.size;
	return(blob);
}

int main(int argc, char *argv[])
{
	(void)argc;
	mm_opt_t *o = mm_opt_init((char const *const *)argv);
	if(o == NULL || o->parg.n < 2 || !mm_endswith(*o->parg.a, ".mai")) { fprintf(stderr, "usage: minialign-gpuworker [options] index.mai reads.fa [...] > out.sam\n"); return(1); }
	o->b.batch_size = 64ull << 20;								/* larger batches than the reference's 512 KB: one mab_map_batch per batch */
	if(getenv("GW_BATCH_KB")) { o->b.batch_size = (uint64_t)atol(getenv("GW_BATCH_KB")) << 10; }
	gw.device = getenv("GW_DEVICE") ? atoi(getenv("GW_DEVICE")) : 0;
	gw.blob = gw_load_blob(*o->parg.a, o->pt, &gw.blob_size);
	if(gw.blob == NULL) { fprintf(stderr, "[E::main] failed to read `%s'\n", *o->parg.a); return(1); }
	/* the subset of the parameters the hot path reads (mm_align_params_t 2517-2524, gaba_params_t gaba.h:90-110) */
	gw.prm.wlen = o->a.wlen; gw.prm.glen = o->a.glen; gw.prm.min_score = o->a.min_score; gw.prm.min_ratio = o->a.min_ratio;
	for(int i = 0; i < 16; i++) { gw.prm.score_matrix[i] = o->a.p.score_matrix[i]; }
	gw.prm.gi = o->a.p.gi; gw.prm.ge = o->a.p.ge; gw.prm.gfa = o->a.p.gfa; gw.prm.gfb = o->a.p.gfb; gw.prm.xdrop = o->a.p.xdrop;

	/* the body of main_align (6365-6440) for a prebuilt index */
	pg_t *pg = pg_init(fopen(*o->parg.a, "rb"), o->pt);
	if(pg == NULL) { return(1); }
	mm_idx_t *mi = mm_idx_load(pg, (read_t const)pgread);
	pg_freeze(pg);
	if(mi == NULL) { return(1); }
	mm_align_t *aln = mm_align_init(&o->a, mi, o->pt);
	mm_print_t *pr = mm_print_init(&o->r);
	if(aln == NULL || pr == NULL) { return(1); }
	mm_print_header(pr, mi->n_seq, mi->s);
	bseq_params_t bq = o->b;
	int ret = 0;
	for(char const *const *q = (char const *const *)&o->parg.a[1]; *q; q++) {
		bseq_file_t *fp = bseq_open(&bq, *q);
		if(fp == NULL) { fprintf(stderr, "[E::main] failed to open `%s'\n", *q); ret = 1; break; }
		int err = gw_align_file(aln, fp, pr);
		bseq_close(fp);
		if(err) { ret = 1; break; }
	}
	mm_align_destroy(aln); mm_idx_destroy(mi); mm_print_destroy(pr); pg_destroy(pg);
	for(uint32_t i = 0; i < MAX_THREADS; i++) { if(gw.ctx[i] != NULL && gw.ctx[i] != gw.parent) { mab_destroy(gw.ctx[i]); } }	/* clones before their parent */
	mab_destroy(gw.parent);
	mm_opt_destroy(o);
	return(ret);
}
