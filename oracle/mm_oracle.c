/*
 * mm_oracle.c -- TEST INFRASTRUCTURE ONLY (see mm_oracle.h).  Citations are into /root/reference/minialign.c.
 */
#include "mm_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define MAX2(a, b) ((a) > (b) ? (a) : (b))
#define MIN2(a, b) ((a) < (b) ? (a) : (b))

/* ---------------------------------------------------------------- mm_extend_core (4075-4112) */
static uint8_t const ntail[128] = {
	4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,
	4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4
};
static ora_section_t const tail_sec = { 0xfffffffe, 96, ntail, 0 };			/* 4512-4518 */

static int64_t extend_core(ora_dp_t *dp, ora_section_t const *a, ora_section_t const *at,
	ora_section_t const *b, ora_section_t const *bt, uint32_t apos, uint32_t bpos)
{
	int64_t f = ora_dp_fill_root(dp, a, apos, b, bpos, 0);
	int64_t m = f;
	uint32_t flag = ORA_TERM;
	while((flag & ora_fill(dp, f)->status) == 0) {
		uint32_t st = ora_fill(dp, f)->status;
		if(st & ORA_UPDATE_A) { a = at; }
		if(st & ORA_UPDATE_B) { b = bt; }
		flag |= st & (ORA_UPDATE_A | ORA_UPDATE_B);
		f = ora_dp_fill(dp, f, a, b, 0);
		m = ora_fill(dp, f)->max > ora_fill(dp, m)->max ? f : m;
	}
	return(m);
}

static uint64_t dump_aln(ora_aln_t const *a, uint32_t rank, uint32_t mapq, uint32_t *out, uint64_t cap)
{
	uint64_t need = 16 + 8 * (uint64_t)a->slen + a->npath;
	if(need > cap) { return(need); }
	uint32_t *p = out;
	memcpy(p, &a->score, 8); p += 2;
	memcpy(p, &a->identity, 8); p += 2;
	*p++ = a->agcnt; *p++ = a->bgcnt; *p++ = a->dcnt; *p++ = a->slen; *p++ = a->plen; *p++ = a->npath;
	*p++ = rank; *p++ = mapq; *p++ = 0; *p++ = 0; *p++ = 0; *p++ = 0;
	for(uint32_t i = 0; i < a->slen; i++) {
		ora_seg_t const *s = &a->seg[i];
		*p++ = s->aid; *p++ = s->bid; *p++ = s->apos; *p++ = s->bpos; *p++ = s->alen; *p++ = s->blen;
		memcpy(p, &s->ppos, 8); p += 2;
	}
	memcpy(p, a->path, sizeof(uint32_t) * a->npath);
	return(need);
}

/* one body of the mm_extend loop (4134-4154) on explicit sequences; mirrors refh_extend in ref_harness.c */
uint64_t mmo_extend(mmo_params_t const *p, uint8_t const *a, uint32_t alen, uint8_t const *b, uint32_t blen,
	uint32_t apos, uint32_t bpos, uint32_t brev, uint32_t narrow, int64_t min_score, uint32_t *res, uint32_t *aln_out, uint64_t cap)
{
	static int const bw[3] = { 64, 32, 16 };
	ora_dp_t dp;
	uint64_t n = 0;
	memset(res, 0, 16 * sizeof(uint32_t));
	if(ora_dp_init(&dp, &p->gp, bw[narrow]) != 0) { return(0); }
	ora_section_t r[2] = { { 0, alen, a, 0 }, { 1, alen, a, 1 } };
	ora_section_t q[3] = { { 0, blen, b, 0 }, { 1, blen, b, 1 }, { 0, blen, b, 0 } };
	ora_dp_flush(&dp);
	int64_t f = extend_core(&dp, &r[0], &tail_sec, &q[brev], &tail_sec, apos, bpos);
	ora_fill_t const *ff = ora_fill(&dp, f);
	memcpy(&res[0], &ff->max, 8); res[2] = ff->status; res[3] = (uint32_t)ff->apos; res[4] = (uint32_t)ff->bpos;
	if(ff->max == 0) { goto done; }
	ora_pos_t cp = ora_dp_search_max(&dp, f);
	res[5] = cp.aid; res[6] = cp.bid; res[7] = cp.apos; res[8] = cp.bpos; res[9] = (uint32_t)cp.plen;
	int32_t ta = MAX2(1, MIN2((int32_t)cp.apos, (int32_t)alen)), tb = MAX2(1, MIN2((int32_t)cp.bpos, (int32_t)blen));
	res[14] = (uint32_t)ta; res[15] = (uint32_t)tb;
	f = extend_core(&dp, &r[1], &tail_sec, &q[1 - brev], &tail_sec, alen - (uint32_t)ta, blen - (uint32_t)tb);
	ff = ora_fill(&dp, f);
	memcpy(&res[10], &ff->max, 8); res[12] = ff->status;
	if(ff->max < min_score) { goto done; }
	ora_aln_t *al = ora_dp_trace(&dp, f);
	if(al == NULL) { goto done; }
	res[13] = 1;
	n = dump_aln(al, 0, 0, aln_out, cap);
	ora_aln_free(al);
done:
	ora_dp_clean(&dp);
	return(n);
}
