/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into, or called by, the product path).
 *
 * A thin translation unit that #includes the UNMODIFIED reference source (ocxtal/minialign, minialign.c) from
 * where it lies under /root/reference (path passed as -DREF_SRC) so that its `static` hot-path functions
 * (mm_sketch, mm_seed, mm_chain, mm_extend, mm_align_seq, ...; minialign.c:2410-4474) can be driven from the
 * parity tests through ctypes and their intermediate arrays dumped.  No reference source is copied here: only
 * calls into it.  Built by oracle/Makefile.ref into oracle/_ref/libref_harness.so (git-ignored).
 *
 * Everything exported is prefixed refh_.  All outputs are flat little-endian arrays the tests compare 1:1
 * with the output of oracle/mm_oracle.c (the restatement) and of the CUDA path.
 */
#define NAMESPACE ref
#ifndef UNITTEST
#define UNITTEST 0
#endif
#include REF_SRC

#include <stdint.h>
#include <string.h>

typedef struct {
	mm_opt_t *o;
	pg_t *pg;
	mm_idx_t *mi;
	mm_align_t *aln;
	lmm_t *lmm;
} refh_t;

/* refh_open: parse argv exactly like the reference's main (presets included), load the .mai given as the first
 * positional argument (minialign.c:6365-6417), create the alignment context with one worker (minialign.c:4671). */
void *refh_open(int argc, char **argv)
{
	refh_t *h = calloc(1, sizeof(refh_t));
	h->o = mm_opt_init((char const *const *)argv);
	if(h->o == NULL || h->o->parg.n == 0) { free(h); return(NULL); }
	h->pg = pg_init(fopen(*h->o->parg.a, "rb"), h->o->pt);
	if(h->pg == NULL) { free(h); return(NULL); }
	h->mi = mm_idx_load(h->pg, (read_t const)pgread);
	pg_freeze(h->pg);
	if(h->mi == NULL) { free(h); return(NULL); }
	h->aln = mm_align_init(&h->o->a, h->mi, h->o->pt);
	if(h->aln == NULL) { free(h); return(NULL); }
	h->lmm = lmm_init_margin(NULL, 512 * 1024, sizeof(mm_aln_t), 0);
	return(h);
}

void refh_close(void *_h)
{
	refh_t *h = (refh_t *)_h;
	if(h == NULL) { return; }
	lmm_clean(h->lmm);
	mm_align_destroy(h->aln);
	mm_idx_destroy(h->mi);
	pg_destroy(h->pg);
	mm_opt_destroy(h->o);
	free(h);
}

/* parameters as the reference resolved them (so tests can hand the same numbers to the oracle / C-ABI) */
void refh_params(void *_h, int32_t *out /* [32] */)
{
	refh_t *h = (refh_t *)_h;
	mm_align_params_t const *a = &h->o->a;
	int i = 0;
	out[i++] = h->mi->k; out[i++] = h->mi->w; out[i++] = h->mi->b; out[i++] = h->mi->n_occ;
	for(int j = 0; j < 4; j++) { out[i++] = (int32_t)h->mi->occ[j]; }
	out[i++] = a->wlen; out[i++] = a->glen; out[i++] = (int32_t)a->min_score;
	memcpy(&out[i++], &a->min_ratio, 4);
	out[i++] = a->p.gi; out[i++] = a->p.ge; out[i++] = a->p.gfa; out[i++] = a->p.gfb; out[i++] = a->p.xdrop;
	for(int j = 0; j < 16; j++) { out[i + j] = a->p.score_matrix[j]; }
}

/* refh_sketch: mm_sketch (minialign.c:2410) on one encoded sequence; returns #words incl. the 4-word cap */
uint64_t refh_sketch(void *_h, uint8_t const *seq, uint32_t len, uint64_t *out, uint64_t cap)
{
	refh_t *h = (refh_t *)_h;
	uint64_v b = { 0 };
	mm_sketch_t sk;
	mm_sketch_init(&sk, h->mi->w, h->mi->k, &b);
	mm_sketch(&sk, seq, len);
	uint64_t n = b.n;
	if(n <= cap) { memcpy(out, b.a, n * sizeof(uint64_t)); }
	free(b.a);
	return(n);
}

/* refh_get: mm_idx_get (minialign.c:2728); copies up to cap occurrences, returns n */
uint32_t refh_get(void *_h, uint64_t minier, uint64_t *out, uint32_t cap)
{
	refh_t *h = (refh_t *)_h;
	uint32_t n = 0;
	v2u32_t const *r = mm_idx_get(h->mi, minier, &n);
	for(uint32_t i = 0; i < n && i < cap; i++) { out[i] = r[i].u64[0]; }
	return(n);
}

/* refh_seed_chain: run mm_seed(i)+mm_chain(i) for rounds 0..round (without mm_extend in between, which is what
 * the reference does whenever the previous rounds produced no result) and dump the sorted seed array (+ leaves)
 * and the sorted root array.  returns n_seed (without sentinel); *n_total = seed.n after chaining, *n_root */
uint64_t refh_seed_chain(void *_h, uint8_t const *seq, uint32_t len, uint32_t round,
	uint32_t *seeds, uint64_t seed_cap, uint64_t *n_total, uint32_t *roots, uint64_t root_cap, uint64_t *n_root)
{
	refh_t *h = (refh_t *)_h;
	mm_tbuf_t *t = h->aln->t[0];
	mm_tbuf_clear(t, h->lmm);
	mm_init_query(t, len, seq, 0, 0);
	uint64_t ns = 0, nr = 0;
	for(uint32_t i = 0; i <= round && i < t->mi.n_occ; i++) {
		ns = mm_seed(t, i);
		nr = ns ? mm_chain(t, i) : 0;
	}
	*n_total = ns ? t->seed.n : 0; *n_root = nr;
	if(ns && t->seed.n <= seed_cap) { memcpy(seeds, t->seed.a, t->seed.n * sizeof(mm_seed_t)); }
	if(nr && nr <= root_cap) { memcpy(roots, t->root.a, nr * sizeof(mm_root_t)); }
	return(ns ? t->n_seed : 0);
}

/* flat serialisation of one alignment: the layout shared by the reference harness, the oracle and the C-ABI.
 * header (16 x u32): score_lo, score_hi, identity_lo, identity_hi, agcnt, bgcnt, dcnt, slen, plen, npathwords, aid(rank), mapq, 0,0,0,0
 * then slen segments (8 x u32): aid, bid, apos, bpos, alen, blen, ppos_lo, ppos_hi
 * then npathwords path words, npathwords = (plen + 31) / 32 + 1 */
static uint64_t refh_dump_aln(mm_aln_t const *ma, uint32_t *out, uint64_t cap)
{
	gaba_alignment_t const *a = ma->a;
	uint32_t npw = (a->plen + 31) / 32 + 1;
	uint64_t need = 16 + 8 * (uint64_t)a->slen + npw;
	if(need > cap) { return(need); }
	uint32_t *p = out;
	memcpy(p, &a->score, 8); p += 2;
	memcpy(p, &a->identity, 8); p += 2;
	*p++ = a->agcnt; *p++ = a->bgcnt; *p++ = a->dcnt; *p++ = a->slen; *p++ = a->plen; *p++ = npw;
	*p++ = ma->aid; *p++ = ma->mapq; *p++ = 0; *p++ = 0; *p++ = 0; *p++ = 0;
	for(uint32_t i = 0; i < a->slen; i++) {
		gaba_path_section_t const *s = &a->seg[i];
		*p++ = s->aid; *p++ = s->bid; *p++ = s->apos; *p++ = s->bpos; *p++ = s->alen; *p++ = s->blen;
		memcpy(p, &s->ppos, 8); p += 2;
	}
	memcpy(p, a->path, sizeof(uint32_t) * npw); p += npw;
	/* clear bits beyond plen + 1 (sentinel) so comparisons are well-defined */
	return(need);
}

/* refh_align: mm_align_seq (minialign.c:4427).  out[0]=n_all, out[1]=n_uniq, then n_all alignments.
 * returns #u32 words needed (0 => unmapped) */
uint64_t refh_align(void *_h, uint8_t const *seq, uint32_t len, uint32_t qid, uint32_t *out, uint64_t cap)
{
	refh_t *h = (refh_t *)_h;
	mm_reg_t const *reg = mm_align_seq(h->aln->t[0], len, seq, qid, h->lmm);
	if(reg == NULL) { return(0); }
	uint64_t n = 2;
	if(cap >= 2) { out[0] = reg->n_all; out[1] = reg->n_uniq; }
	for(uint32_t i = 0; i < reg->n_all; i++) {
		n += refh_dump_aln(reg->aln[i], n < cap ? out + n : out, n < cap ? cap - n : 0);
	}
	for(uint32_t i = 0; i < reg->n_all; i++) { lmm_free(h->lmm, (void *)reg->aln[i]->a); }
	lmm_free(h->lmm, (void *)reg);
	return(n);
}

/* refh_extend: one iteration body of the mm_extend loop (minialign.c:4134-4154) on explicit sequences:
 * downward mm_extend_core from (apos,bpos) with band index `narrow`, gaba_dp_search_max, clip (minialign.c:3963-3966),
 * upward mm_extend_core over the reversed sections, gaba_dp_trace.  brev selects the reverse-complement query.
 * res[0..15]: down.max(lo,hi) down.status down.apos down.bpos | pos.aid pos.bid pos.apos pos.bpos pos.plen |
 *             up.max(lo,hi) up.status | traced(0/1) | tp.apos tp.bpos
 * aln_out receives the alignment in refh_dump_aln layout when traced. returns words needed for aln. */
uint64_t refh_extend(void *_h, uint8_t const *a, uint32_t alen, uint8_t const *b, uint32_t blen,
	uint32_t apos, uint32_t bpos, uint32_t brev, uint32_t narrow, int64_t min_score, uint32_t *res, uint32_t *aln_out, uint64_t cap)
{
	refh_t *h = (refh_t *)_h;
	mm_tbuf_t *t = h->aln->t[0];
	t->alloc.opaque = (void *)h->lmm;
	gaba_section_t r[2] = { _sec_fw(0, a, alen), _sec_rv(0, a, alen) };
	gaba_section_t q[3] = { _sec_fw(0, b, blen), _sec_rv(0, b, blen), _sec_fw(0, b, blen) };
	gaba_dp_t *dp = &t->dp[narrow];
	memset(res, 0, 16 * sizeof(uint32_t));
	gaba_dp_flush(t->dp);
	gaba_fill_t const *f = mm_extend_core(dp, &r[0], t->t, &q[brev], t->qtp + brev, ((mm_pos_pair_t){ apos, bpos }));
	memcpy(&res[0], &f->max, 8); res[2] = f->status; res[3] = (uint32_t)f->apos; res[4] = (uint32_t)f->bpos;
	if(f->max == 0) { return(0); }
	gaba_pos_pair_t const *cp = gaba_dp_search_max(dp, f);
	res[5] = cp->aid; res[6] = cp->bid; res[7] = cp->apos; res[8] = cp->bpos; res[9] = (uint32_t)cp->plen;
	int32_t ta = MAX2(1, MIN2((int32_t)cp->apos, (int32_t)alen)), tb = MAX2(1, MIN2((int32_t)cp->bpos, (int32_t)blen));
	res[14] = ta; res[15] = tb;
	f = mm_extend_core(dp, &r[1], t->t + 1, &q[1 - brev], t->qtp + 1 - brev, ((mm_pos_pair_t){ alen - ta, blen - tb }));
	memcpy(&res[10], &f->max, 8); res[12] = f->status;
	if(f->max < min_score) { return(0); }
	gaba_alignment_t const *al = gaba_dp_trace(dp, f, &t->alloc);
	if(al == NULL) { return(0); }
	res[13] = 1;
	mm_aln_t *ma = (mm_aln_t *)al - 1;
	ma->aid = 0; ma->mapq = 0;
	uint64_t n = refh_dump_aln(ma, aln_out, cap);
	lmm_free(h->lmm, (void *)al);
	return(n);
}

/* refh_extend_dump: the seed-chain-extend loop of mm_align_seq (minialign.c:4444-4449) WITHOUT the post-processing;
 * dumps n_res then per result {score, iid, n_aln, plen, lb, ub} (6 x u32) so the extend state machine can be compared
 * before pruning / MAPQ.  Alignments are left to the arena. */
uint64_t refh_extend_dump(void *_h, uint8_t const *seq, uint32_t len, uint32_t *out, uint64_t cap)
{
	refh_t *h = (refh_t *)_h;
	mm_tbuf_t *t = h->aln->t[0];
	mm_tbuf_clear(t, h->lmm);
	mm_init_query(t, len, seq, 0, 0);
	for(uint64_t i = 0; i < t->mi.n_occ; i++) {
		if(mm_seed(t, i) == 0) { continue; }
		if(mm_chain(t, i) == 0) { continue; }
		if(mm_extend(t, i) > 0) { break; }
	}
	uint64_t n = 1;
	out[0] = t->n_res;
	mm_res_t *r = (mm_res_t *)t->root.a;
	for(uint32_t i = 0; i < t->n_res && n + 6 <= cap; i++) {
		mm_bin_t *bin = (mm_bin_t *)&t->bin.a[r[i].iid];
		out[n++] = r[i].score; out[n++] = r[i].iid; out[n++] = bin->n_aln; out[n++] = bin->plen; out[n++] = bin->lb; out[n++] = bin->ub;
	}
	return(n);
}

/* refh_align_many: mm_align_seq (minialign.c:4427) over n reads on n_threads of the reference's own worker buffers (the
 * context must have been opened with -t<n_threads>): thread j maps reads j, j + n_threads, ...; results are freed, nothing is
 * read or printed.  Returns the wall-clock seconds of the parallel section: the reference's hot path alone, no I/O
 * (bench.py's cpu_baseline.hot_path). */
#include <pthread.h>
#include <time.h>
typedef struct { refh_t *h; uint32_t tid, nth, n; uint8_t const *block; uint64_t const *ofs; uint32_t const *len; uint64_t n_mapped; } refh_job_t;
static void *refh_align_many_worker(void *arg)
{
	refh_job_t *j = (refh_job_t *)arg;
	lmm_t *lmm = lmm_init_margin(NULL, 512 * 1024, sizeof(mm_aln_t), 0);
	for(uint32_t i = j->tid; i < j->n; i += j->nth) {
		mm_reg_t const *reg = mm_align_seq(j->h->aln->t[j->tid], j->len[i], j->block + j->ofs[i], i, lmm);
		if(reg == NULL) { continue; }
		j->n_mapped++;
		for(uint32_t k = 0; k < reg->n_all; k++) { lmm_free(lmm, (void *)reg->aln[k]->a); }
		lmm_free(lmm, (void *)reg);
	}
	lmm_clean(lmm);
	return(NULL);
}
double refh_align_many(void *_h, uint8_t const *block, uint64_t const *ofs, uint32_t const *len, uint32_t n, uint32_t n_threads, uint64_t *n_mapped)
{
	refh_t *h = (refh_t *)_h;
	if(n_threads == 0 || n_threads > pt_nth(h->o->pt)) { return(-1.0); }
	pthread_t *th = calloc(n_threads, sizeof(pthread_t));
	refh_job_t *job = calloc(n_threads, sizeof(refh_job_t));
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for(uint32_t t = 0; t < n_threads; t++) {
		job[t] = (refh_job_t){ .h = h, .tid = t, .nth = n_threads, .n = n, .block = block, .ofs = ofs, .len = len, .n_mapped = 0 };
		pthread_create(&th[t], NULL, refh_align_many_worker, &job[t]);
	}
	uint64_t m = 0;
	for(uint32_t t = 0; t < n_threads; t++) { pthread_join(th[t], NULL); m += job[t].n_mapped; }
	clock_gettime(CLOCK_MONOTONIC, &t1);
	if(n_mapped) { *n_mapped = m; }
	free(th); free(job);
	return((double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec));
}
