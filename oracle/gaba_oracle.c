/*
 * gaba_oracle.c -- TEST INFRASTRUCTURE ONLY (see gaba_oracle.h).
 *
 * Scalar restatement of the reference's GABA engine for the COMBINED (piecewise-affine) gap model, any band
 * width W in {16,32,64}.  One lane at a time, plain int8/int16 casts where the reference wraps or saturates.
 * Citations are into /root/reference/gaba.c unless noted.
 *
 * Differences in *representation* (never in observable results):
 *   - the 16 MB stack of gaba_block_s / gaba_joint_tail_s objects is a pair of growable arrays (blk[], tl[]);
 *     phantom blocks are entries with ORA_X_HEAD set and a `link` index;
 *   - the sequence prefetch buffers (bufa/bufb, gaba.c:957-1144) are replaced by two append-only streams of
 *     consumed bases; the W-lane window of a vector is the last W entries of each stream;
 *   - mirrored ("phantom") section pointers are an explicit `rev` flag.
 */
#include "gaba_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define MAX2(a, b) ((a) > (b) ? (a) : (b))
#define MIN2(a, b) ((a) < (b) ? (a) : (b))

static inline int8_t w8(int x) { return((int8_t)(uint8_t)x); }						/* wrapping int8 */
static inline int8_t s8(int x) { return((int8_t)(x > 127 ? 127 : (x < -128 ? -128 : x))); }	/* saturating int8 */
static inline int16_t w16(int x) { return((int16_t)(uint16_t)x); }

/* ---------------------------------------------------------------- init (gaba.c:3613-3842) */
static int gap_h(ora_params_t const *p, int l) { return(MAX2(-1 * (l > 0) * p->gi - p->ge * l, -1 * p->gfb * l)); }	/* gaba.c:834 */
static int gap_v(ora_params_t const *p, int l) { return(MAX2(-1 * (l > 0) * p->gi - p->ge * l, -1 * p->gfa * l)); }	/* gaba.c:835 */
static int gap_e(ora_params_t const *p, int l) { return(-1 * (l > 0) * p->gi - p->ge * l); }							/* gaba.c:837 */

static int max_match(ora_params_t const *p) { int m = -128; for(int i = 0; i < 16; i++) { m = MAX2(m, p->score_matrix[i]); } return(m); }
static int min_match(ora_params_t const *p) { int m = 127; for(int i = 0; i < 16; i++) { m = MIN2(m, p->score_matrix[i]); } return(m); }

/* gaba_init_check_score (gaba.c:3613-3638); the wrapper runs it with the W=16 object only (gaba_wrap.h:286-292) */
static int check_score(ora_params_t const *p)
{
	int M = max_match(p), X = min_match(p);
	if(M <= 0 || M > 6 || X >= 0 || X < -7) { return(-1); }
	if(X < -2 * (p->gi + p->ge)) { return(-1); }
	if(p->gfa != 0 && p->gfb != 0 && X <= -1 * (p->gfa + p->gfb)) { return(-1); }
	if(p->ge <= 0 || p->gi < 0) { return(-1); }
	if(p->gfa < 0 || (p->gfa != 0 && p->gfa <= p->ge)) { return(-1); }
	if(p->gfb < 0 || (p->gfb != 0 && p->gfb <= p->ge)) { return(-1); }
	if((p->gfa == 0) ^ (p->gfb == 0)) { return(-1); }
	int ofs = p->gi + p->ge;
	for(int i = 0; i < 16 / 2; i++) {
		int t1 = ofs + gap_h(p, i*2 + 1) - gap_h(p, i*2);
		int t2 = ofs + (M + gap_v(p, i*2 + 1)) - gap_v(p, (i + 1) * 2);
		int t3 = ofs + (M + gap_h(p, i*2 + 1)) - gap_h(p, (i + 1) * 2);
		int t4 = ofs + gap_h(p, i*2 + 1) - gap_h(p, i*2);
		if(MAX2(MAX2(t1, t2), MAX2(t3, t4)) > 127) { return(-1); }
		if(MIN2(MIN2(t2, t2), MIN2(t3, t4)) < 0) { return(-1); }
	}
	return(0);
}

int ora_dp_init(ora_dp_t *dp, ora_params_t const *p, int W)
{
	memset(dp, 0, sizeof(*dp));
	if(!(W == 16 || W == 32 || W == 64)) { return(-1); }
	if(p->gi == 0 || p->gfa == 0 || p->gfb == 0) { return(-2); }		/* COMBINED only (gaba_wrap.h:207-222) */
	if(check_score(p) != 0) { return(-3); }
	dp->W = W;
	int M = max_match(p), ofs = p->gi + p->ge;

	/* gaba_init_score_vector (gaba.c:3644-3678), arch_util.h:166-210 */
	for(int i = 0; i < 16; i++) { dp->sb[i] = w8(p->score_matrix[i] + 2 * ofs); }
	dp->adjh = dp->adjv = p->gi;
	dp->ofsh = dp->ofsv = w8(-ofs);
	dp->gfh = w8(ofs - p->gfb);
	dp->gfv = w8(ofs - p->gfa);
	dp->tx = w8(p->xdrop - 128);								/* gaba.c:3823 */
	dp->gi = p->gi; dp->ge = p->ge; dp->gfa = p->gfa; dp->gfb = p->gfb;

	/* imx, xmx (gaba.c:3797-3827) */
	int64_t acc[2] = { 0, 0 };
	for(int i = 0; i < 16; i++) { acc[(i & 3) == (i >> 2)] += p->score_matrix[i]; }
	double m = (double)acc[1] / 4.0, x = (double)acc[0] / 12.0;
	dp->imx = 1 / (m - x); dp->xmx = x / (m - x);

	/* gaba_init_diff_vectors (gaba.c:3705-3732) */
	ora_block_t *rb = &dp->rootblk;
	uint8_t dh[ORA_WMAX], dv[ORA_WMAX], de[ORA_WMAX], df[ORA_WMAX];
	for(int i = 0; i < W / 2; i++) {
		dh[W/2 - 1 - i] = (uint8_t)(ofs + gap_h(p, i*2 + 1) - gap_h(p, i*2));
		dh[W/2     + i] = (uint8_t)(ofs + M + gap_v(p, i*2 + 1) - gap_v(p, (i + 1) * 2));
		dv[W/2 - 1 - i] = (uint8_t)(ofs + M + gap_h(p, i*2 + 1) - gap_h(p, (i + 1) * 2));
		dv[W/2     + i] = (uint8_t)(ofs + gap_v(p, i*2 + 1) - gap_v(p, i*2));
		de[W/2 - 1 - i] = (uint8_t)(p->gi + dv[W/2 - 1 - i] + gap_e(p, i*2 + 1) - gap_h(p, i*2 + 1));
		de[W/2     + i] = (uint8_t)(p->gi + dv[W/2     + i] - p->gi);
		df[W/2 - 1 - i] = (uint8_t)(p->gi + dh[W/2 - 1 - i] - p->gi);
		df[W/2     + i] = (uint8_t)(p->gi + dh[W/2     + i] + gap_e(p, i*2 + 1) - gap_v(p, i*2 + 1));
	}
	for(int q = 0; q < W; q++) {
		rb->dh[q] = w8(0 - (int8_t)dh[q]);						/* negated, gaba.c:3727-3730 */
		rb->dv[q] = (int8_t)dv[q]; rb->de[q] = (int8_t)de[q]; rb->df[q] = (int8_t)df[q];
	}
	rb->acc = 0; rb->xstat = ORA_X_ROOT; rb->acnt = rb->bcnt = 0; rb->link = -1;

	/* gaba_init_phantom (gaba.c:3739-3791), gaba_init_middle_delta (3684-3694) */
	ora_tail_t *rt = &dp->root;
	int64_t init_max = -(M + gap_h(p, 1));
	rt->f.max = init_max; rt->f.status = ORA_UPDATE_A | ORA_UPDATE_B;
	rt->f.apos = -W / 2; rt->f.bpos = -W / 2;
	rt->tail = -1; rt->last_blk = 0;
	rt->mdrop = w16((int)init_max - 128);
	rt->cha[0] = 0x0c; rt->chb[W - 1] = 0x03;
	for(int q = 0; q < W; q++) { rt->xd[q] = -128; }
	for(int i = 0; i < W / 2; i++) {
		rt->md[W/2 - 1 - i] = w16(-(i + 1) * M + gap_h(p, i*2 + 1));
		rt->md[W/2     + i] = w16(-(i + 1) * M + gap_v(p, i*2 + 1));
	}
	ora_dp_flush(dp);
	return(0);
}

void ora_dp_clean(ora_dp_t *dp)
{
	free(dp->blk); free(dp->tl); free(dp->sa); free(dp->sb_);
	memset(dp, 0, sizeof(*dp));
}

/* gaba_dp_flush (gaba.c:3969-4002): the stack restarts; entry 0 of each array is the root template */
void ora_dp_flush(ora_dp_t *dp)
{
	if(dp->mblk == 0) { dp->mblk = 64; dp->blk = malloc(sizeof(ora_block_t) * dp->mblk); }
	if(dp->mtl == 0) { dp->mtl = 16; dp->tl = malloc(sizeof(ora_tail_t) * dp->mtl); }
	dp->blk[0] = dp->rootblk; dp->nblk = 1;
	dp->tl[0] = dp->root; dp->ntl = 1;
	dp->na = dp->nb = 0;
}

static int64_t push_blk(ora_dp_t *dp)
{
	if(dp->nblk == dp->mblk) { dp->mblk *= 2; dp->blk = realloc(dp->blk, sizeof(ora_block_t) * dp->mblk); }
	memset(&dp->blk[dp->nblk], 0, sizeof(ora_block_t));
	return((int64_t)dp->nblk++);
}
static int64_t push_tail(ora_dp_t *dp)
{
	if(dp->ntl == dp->mtl) { dp->mtl *= 2; dp->tl = realloc(dp->tl, sizeof(ora_tail_t) * dp->mtl); }
	memset(&dp->tl[dp->ntl], 0, sizeof(ora_tail_t));
	return((int64_t)dp->ntl++);
}
static void push_a(ora_dp_t *dp, uint8_t c)
{
	if(dp->na == dp->msa) { dp->msa = dp->msa ? 2 * dp->msa : 4096; dp->sa = realloc(dp->sa, dp->msa); }
	dp->sa[dp->na++] = c;
}
static void push_b(ora_dp_t *dp, uint8_t c)
{
	if(dp->nb == dp->msb) { dp->msb = dp->msb ? 2 * dp->msb : 4096; dp->sb_ = realloc(dp->sb_, dp->msb); }
	dp->sb_[dp->nb++] = c;
}

/* sequence fetch with the reference's code tables (gaba.c:864-879, 957-1118) */
static uint8_t fetch_a(ora_section_t const *s, uint64_t i)
{
	static uint8_t const comp[16] = { 3, 2, 1, 0, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4 };
	if(!s->rev) { return(s->base[i]); }
	uint8_t c = s->base[s->len - 1 - i];
	return((c & 0x80) ? 0 : comp[c & 15]);
}
static uint8_t fetch_b(ora_section_t const *s, uint64_t i)
{
	static uint8_t const shift[16] = { 0, 4, 8, 12, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2 };
	static uint8_t const compshift[16] = { 12, 8, 4, 0, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2 };
	uint8_t c = s->rev ? s->base[s->len - 1 - i] : s->base[i];
	return((c & 0x80) ? 0 : (s->rev ? compshift : shift)[c & 15]);
}

/* ---------------------------------------------------------------- vector update (gaba.c:1604-1699) */
typedef struct {
	int8_t dh[ORA_WMAX], dv[ORA_WMAX], de[ORA_WMAX], df[ORA_WMAX], delta[ORA_WMAX], drop[ORA_WMAX];
	int32_t acc; uint32_t dir;
} vec_t;

/* one anti-diagonal; the stream heads (dp->na, dp->nb) must already include the base this step consumes */
static void vec_step(ora_dp_t *dp, vec_t *v, int down, ora_mask_t *m)
{
	int const W = dp->W;
	if(!down) {				/* _fill_right: bsl dh, df (lane q <- q-1, lane 0 <- 0) */
		for(int q = W - 1; q > 0; q--) { v->dh[q] = v->dh[q - 1]; v->df[q] = v->df[q - 1]; }
		v->dh[0] = 0; v->df[0] = 0;
	} else {				/* _fill_down: bsr dv, de (lane q <- q+1, lane W-1 <- 0) */
		for(int q = 0; q < W - 1; q++) { v->dv[q] = v->dv[q + 1]; v->de[q] = v->de[q + 1]; }
		v->dv[W - 1] = 0; v->de[W - 1] = 0;
	}
	uint64_t mh = 0, mv = 0, me = 0, mf = 0;
	int8_t d[ORA_WMAX];
	for(int q = 0; q < W; q++) {
		uint8_t idx = dp->sa[dp->na - 1 - q] | dp->sb_[dp->nb - W + q];
		int8_t t = (idx & 0x80) ? 0 : dp->sb[idx & 15];						/* pshufb */
		int8_t dh = v->dh[q], dv = v->dv[q], de = v->de[q], df = v->df[q];
		int8_t dfh = w8(dv + dp->gfh), dfv = w8(dp->gfv - dh);
		if(getenv("ORA_DEBUG2")) { int x1 = dv + dp->gfh, x2 = dp->gfv - dh, x3 = de + dp->adjh, x4 = df + dp->adjv; if(x1 != (int8_t)x1 || x2 != (int8_t)x2 || x3 != (int8_t)x3 || x4 != (int8_t)x4) fprintf(stderr, "WRAP q %d: %d %d %d %d (dh %d dv %d de %d df %d)\n", q, x1, x2, x3, x4, dh, dv, de, df); }
		int8_t s = MAX2(de, df); s = MAX2(s, dfh); t = MAX2(t, dfv); t = MAX2(t, s);
		uint64_t gfh = t == dfh, gh = t == de, gfv = t == dfv, gv = t == df;
		mh |= (gfh | gh) << q; gh &= ~gfh;
		mv |= (gfv | gv) << q; gv &= ~gfv;
		de = w8(de + dp->adjh); int8_t te = MAX2(de, t);
		me |= (gh | (uint64_t)(te == t)) << q;
		de = w8(te + dh); dh = w8(dh + t);
		df = w8(df + dp->adjv); int8_t tf = MAX2(df, t);
		mf |= (gv | (uint64_t)(tf == t)) << q;
		df = w8(tf - dv); t = w8(dv - t);
		dv = dh; dh = t;
		v->dh[q] = dh; v->dv[q] = dv; v->de[q] = de; v->df[q] = df;
		d[q] = down ? w8(dp->ofsv + dv) : w8(dp->ofsh - dh);					/* _fill_update_delta */
		v->delta[q] = w8(v->delta[q] + d[q]);
		v->drop[q] = s8(v->drop[q] - d[q]);
	}
	v->acc += d[0] - d[W - 1];												/* _dir_update */
	if(m) { m->h = mh; m->v = mv; m->e = me; m->f = mf; }
	dp->vec_count++;
}

static void vec_load(ora_dp_t *dp, vec_t *v, ora_block_t const *prev)
{
	memcpy(v->dh, prev->dh, ORA_WMAX); memcpy(v->dv, prev->dv, ORA_WMAX);
	memcpy(v->de, prev->de, ORA_WMAX); memcpy(v->df, prev->df, ORA_WMAX);
	memset(v->delta, 0, ORA_WMAX);
	memcpy(v->drop, dp->xd, ORA_WMAX);
	v->acc = prev->acc; v->dir = 0;											/* _dir_init */
}

/* _fill_store_context (gaba.c:1734-1778) */
static void blk_store(ora_dp_t *dp, ora_block_t *b, vec_t const *v, int acnt, int bcnt)
{
	int const W = dp->W;
	memcpy(b->dh, v->dh, ORA_WMAX); memcpy(b->dv, v->dv, ORA_WMAX);
	memcpy(b->de, v->de, ORA_WMAX); memcpy(b->df, v->df, ORA_WMAX);
	b->dir_mask = v->dir; b->acc = (int8_t)v->acc;								/* _dir_save: int8 store */
	if(getenv("ORA_DEBUG")) { fprintf(stderr, "ORA blk acnt %d bcnt %d dir %08x acc %d drop_c %d delta_c %d d0 %d dW %d\n", acnt, bcnt, v->dir, v->acc, v->drop[dp->W/2], v->delta[dp->W/2], v->delta[0], v->delta[dp->W-1]); }
	b->xstat = (int8_t)(((int)dp->tx - (int)v->drop[W / 2]) & ORA_X_TERM);
	int32_t cofs = v->delta[W / 2];
	b->acnt = (int8_t)acnt; b->bcnt = (int8_t)bcnt;
	dp->ofsd += cofs; dp->rem[0] -= acnt; dp->rem[1] -= bcnt;
	uint64_t mm = 0;
	for(int q = 0; q < W; q++) {
		int8_t sum = w8(v->drop[q] + v->delta[q]);
		if(sum > dp->xd[q]) { mm |= 1ULL << q; }
	}
	b->max_mask = mm;
	cofs += 0x0100;
	for(int q = 0; q < W; q++) {
		int8_t drop = v->drop[q], delta = v->delta[q];
		int16_t md = w16(dp->md[q] + delta);
		int8_t ov = (int8_t)(~w8(drop + delta) & (drop & delta));				/* andn(add(drop,delta), and(drop,delta)) */
		md = w16(md + (0x0100 & (int16_t)ov));
		int8_t uv = (int8_t)(s8(delta - 0x40) | drop);
		md = w16(md + (0x0100 & (int16_t)uv));
		md = w16(md - cofs);
		dp->md[q] = md;
		dp->xd[q] = drop;
	}
}

/* ---------------------------------------------------------------- section / tail bookkeeping */
/* fill_load_section (gaba.c:1269-1308); breakpoint masks are always zero on this path (minialign never merges) */
static void load_section(ora_dp_t *dp, int64_t tail, ora_section_t const *a, ora_section_t const *b, uint32_t pridx)
{
	ora_tail_t const *t = &dp->tl[tail];
	ora_section_t const *s[2] = { a, b };
	for(int i = 0; i < 2; i++) {
		uint32_t ridx = t->ridx[i] == 0 ? s[i]->len : t->ridx[i];
		dp->rem[i] = ridx; dp->sridx[i] = ridx; dp->ids[i] = s[i]->id; dp->sec[i] = *s[i];
	}
	dp->pridx = pridx; dp->ofsd = 0;
	dp->wtail = tail;
}

/* fill_load_vectors + fill_create_phantom (gaba.c:1315-1332, 1376-1399) */
static int64_t load_vectors(ora_dp_t *dp, int64_t tail)
{
	int const W = dp->W;
	ora_tail_t const *t = &dp->tl[tail];
	for(int q = W - 1; q >= 0; q--) { push_a(dp, t->cha[q]); }
	for(int q = 0; q < W; q++) { push_b(dp, t->chb[q]); }
	memcpy(dp->xd, t->xd, sizeof(dp->xd)); memcpy(dp->md, t->md, sizeof(dp->md));
	int64_t prev = t->last_blk;
	int64_t ph = push_blk(dp);
	ora_block_t *p = &dp->blk[ph], *pb = &dp->blk[prev];
	memcpy(p->dh, pb->dh, ORA_WMAX); memcpy(p->dv, pb->dv, ORA_WMAX);
	memcpy(p->de, pb->de, ORA_WMAX); memcpy(p->df, pb->df, ORA_WMAX);
	p->acc = pb->acc; p->xstat = (int8_t)((pb->xstat & ORA_X_ROOT) | ORA_X_HEAD);
	p->acnt = p->bcnt = 0; p->link = prev; p->na = dp->na; p->nb = dp->nb;
	return(ph);
}

/* consume `n` bases of side i into the stream (the reference prefetches them; consumption order is the same) */
static void consume(ora_dp_t *dp, int i, uint32_t already, uint32_t n)
{
	ora_section_t const *s = &dp->sec[i];
	for(uint32_t k = 0; k < n; k++) {
		uint64_t pos = (uint64_t)s->len - dp->rem[i] + already + k;
		if(i == 0) { push_a(dp, fetch_a(s, pos)); } else { push_b(dp, fetch_b(s, pos)); }
	}
}

/* fill_init_fetch (gaba.c:1168-1210): the first W/2-1 bases of each side enter the window without a DP step */
static int64_t init_fetch(ora_dp_t *dp, int64_t ph, int64_t apos, int64_t bpos)
{
	int32_t irem[2] = { (int32_t)(-1 - (int32_t)apos), (int32_t)(-1 - (int32_t)bpos) };
	int32_t srem[2] = { (int32_t)dp->rem[0], (int32_t)dp->rem[1] };
	int32_t adj[2] = { 1, 0 };
	int32_t len[2];
	for(int i = 0; i < 2; i++) {
		int32_t x = MIN2(irem[i], srem[i]);
		int32_t y = (srem[1 - i] - irem[1 - i]) + (adj[i] + irem[i]);
		len[i] = MIN2(x, y);
	}
	consume(dp, 0, 0, (uint32_t)len[0]); consume(dp, 1, 0, (uint32_t)len[1]);
	dp->blk[ph].acnt = (int8_t)len[0]; dp->blk[ph].bcnt = (int8_t)len[1];
	dp->rem[0] = (uint32_t)(srem[0] - len[0]); dp->rem[1] = (uint32_t)(srem[1] - len[1]);
	return(bpos + len[1]);
}

/* fill_create_tail (gaba.c:1405-1499); `last` is the last processed entry, xstat its status */
static int64_t create_tail(ora_dp_t *dp, int64_t last)
{
	int const W = dp->W;
	ora_block_t *b = &dp->blk[last];
	int xstat = b->xstat;
	int cnt = ((uint8_t)b->acnt) | (((uint8_t)b->bcnt) << 8);
	if(cnt == 0 && !(b->xstat & ORA_X_HEAD)) { dp->nblk = (size_t)last; last--; }	/* squash the empty block */
	else if(cnt == 0) { last--; }													/* (head with nothing fetched: reference reads the slot before it) */
	int64_t ti = push_tail(dp);
	ora_tail_t *t = &dp->tl[ti];
	ora_tail_t const *prev = &dp->tl[dp->wtail];
	t->last_blk = last;
	/* fill_save_vectors */
	for(int q = 0; q < W; q++) { t->cha[q] = dp->sa[dp->na - 1 - q]; t->chb[q] = dp->sb_[dp->nb - W + q]; }
	memcpy(t->xd, dp->xd, sizeof(t->xd)); memcpy(t->md, dp->md, sizeof(t->md));
	int16_t mdrop = -32768;
	for(int q = 0; q < W; q++) { int16_t s = w16(dp->md[q] + dp->xd[q]); mdrop = MAX2(mdrop, s); }
	/* fill_save_section */
	t->mdrop = mdrop; t->istat = 0; t->pridx = dp->pridx;
	uint32_t upd = 0;
	for(int i = 0; i < 2; i++) {
		t->ridx[i] = dp->rem[i]; t->adv[i] = dp->sridx[i] - dp->rem[i];
		if(t->ridx[i] == 0) { upd |= i == 0 ? ORA_UPDATE_A : ORA_UPDATE_B; }
	}
	t->tail = dp->wtail;
	t->f.aid = dp->ids[0]; t->f.bid = dp->ids[1];
	t->f.ascnt = prev->f.ascnt + (t->ridx[0] == 0); t->f.bscnt = prev->f.bscnt + (t->ridx[1] == 0);
	t->f.apos = prev->f.apos + t->adv[0]; t->f.bpos = prev->f.bpos + t->adv[1];
	t->f.max = (prev->f.max - prev->mdrop) + dp->ofsd + mdrop;
	t->f.status = ((uint32_t)(xstat & ORA_X_TERM) << 8) | upd;
	return(ti);
}

/* fill_section_seq_bounded -> fill_seq_bounded -> bulk / cap loops (gaba.c:1821-2103) */
static int64_t fill_blocks(ora_dp_t *dp, int64_t cur)
{
	int cap = 0;
	while(1) {
		if(dp->blk[cur].xstat < 0) { break; }								/* TERM (sign bit) */
		if(!cap && (dp->rem[0] < ORA_BLK || dp->rem[1] < ORA_BLK || dp->pridx < ORA_BLK)) { cap = 1; }
		int64_t bi = push_blk(dp);
		ora_block_t *b = &dp->blk[bi];
		b->na = dp->na; b->nb = dp->nb;
		vec_t v; vec_load(dp, &v, &dp->blk[bi - 1]);
		int acnt = 0, bcnt = 0, i = 0;
		for(; i < ORA_BLK; i++) {
			v.dir = (v.dir << 1) | (uint32_t)(v.acc < 0);						/* _dir_fetch */
			int down = (int)(v.dir & 1);
			if(cap) {															/* _fill_cap_test_idx */
				int64_t ar = (int64_t)dp->rem[0] - (acnt + !down), br = (int64_t)dp->rem[1] - (bcnt + down);
				int64_t pr = ar + br + (int64_t)dp->pridx;
				if((ar | br | pr) < 0) { v.dir >>= 1; break; }				/* windback */
			}
			if(down) { consume(dp, 1, (uint32_t)bcnt, 1); bcnt++; } else { consume(dp, 0, (uint32_t)acnt, 1); acnt++; }
			b = &dp->blk[bi];
			vec_step(dp, &v, down, &b->mask[i]);
		}
		dp->pridx -= (uint32_t)i;
		if(i < ORA_BLK) { v.dir = (i == 0) ? v.dir : (v.dir << (ORA_BLK - i)); }	/* _dir_adjust_remainder (x86 shl by 32 is a no-op) */
		blk_store(dp, &dp->blk[bi], &v, acnt, bcnt);
		cur = bi;
		if(i != ORA_BLK) { break; }
	}
	return(cur);
}

int64_t ora_dp_fill_root(ora_dp_t *dp, ora_section_t const *a, uint32_t apos, ora_section_t const *b, uint32_t bpos, uint32_t pridx)
{
	/* fill_create_bridge (gaba.c:1339-1369) */
	int64_t bi = push_tail(dp);
	ora_tail_t *brg = &dp->tl[bi];
	*brg = dp->tl[0];
	brg->istat = 1;
	brg->ridx[0] = a->len - apos; brg->ridx[1] = b->len - bpos;
	brg->adv[0] = apos; brg->adv[1] = bpos;
	brg->tail = 0;
	brg->f.aid = a->id; brg->f.bid = b->id;
	load_section(dp, bi, a, b, pridx == 0 ? UINT32_MAX : pridx);
	int64_t ph = load_vectors(dp, 0);
	if(init_fetch(dp, ph, dp->tl[0].f.apos, dp->tl[0].f.bpos) < -1) { return(create_tail(dp, ph)); }
	return(create_tail(dp, fill_blocks(dp, ph)));
}

int64_t ora_dp_fill(ora_dp_t *dp, int64_t prev, ora_section_t const *a, ora_section_t const *b, uint32_t pridx)
{
	load_section(dp, prev, a, b, pridx == 0 ? dp->tl[prev].pridx : pridx);
	int64_t ph = load_vectors(dp, prev);
	if(dp->tl[prev].f.bpos < -1) {
		if(init_fetch(dp, ph, dp->tl[prev].f.apos, dp->tl[prev].f.bpos) < -1) { return(create_tail(dp, ph)); }
	}
	return(create_tail(dp, fill_blocks(dp, ph)));
}

/* ---------------------------------------------------------------- max search (gaba.c:2604-2817) */
typedef struct {
	int64_t blk; uint32_t p, q;
	int32_t gidx[2], sgidx[2];
} leaf_t;

static int popcnt32(uint32_t x) { return(__builtin_popcount(x)); }
static uint64_t tz64(uint64_t x) { return(x ? (uint64_t)__builtin_ctzll(x) : 64); }
static uint64_t lz64(uint64_t x) { return(x ? (uint64_t)__builtin_clzll(x) : 64); }

/* leaf_search (gaba.c:2708-2771): returns plen, fills lf with (block, p, q) and the grid indices */
static uint64_t leaf_search(ora_dp_t *dp, int64_t ti, leaf_t *lf)
{
	int const W = dp->W;
	ora_tail_t const *t = &dp->tl[ti];
	uint64_t max_mask = 0;													/* leaf_load_max_mask (2609-2629) */
	for(int q = 0; q < W; q++) { if(w16(t->md[q] + t->xd[q]) == t->mdrop) { max_mask |= 1ULL << q; } }
	int64_t b = t->last_blk + 1;
	int32_t ridx[2] = { (int32_t)t->ridx[0], (int32_t)t->ridx[1] };
	while(1) {
		--b;
		if((dp->blk[b].xstat & ORA_X_ROOT) == ORA_X_ROOT) { return(0); }
		while(dp->blk[b].xstat & ORA_X_HEAD) { b = dp->blk[b].link; }
		ridx[0] += dp->blk[b].acnt; ridx[1] += dp->blk[b].bcnt;
		if((max_mask & ~dp->blk[b].max_mask) == 0) { break; }
		max_mask &= ~dp->blk[b].max_mask;
	}
	/* leaf_detect_pos (2663-2695): replay the block from its start state (diffs of the entry physically before it) */
	ora_block_t const *blk = &dp->blk[b];
	uint64_t marr[ORA_BLK + 1];
	int n = blk->acnt + blk->bcnt;
	memset(marr, 0, sizeof(marr));
	{
		uint64_t sna = dp->na, snb = dp->nb;
		int8_t sxd[ORA_WMAX]; memcpy(sxd, dp->xd, ORA_WMAX);
		vec_t v; vec_load(dp, &v, &dp->blk[b - 1]);
		int8_t mx[ORA_WMAX]; memset(mx, 0, sizeof(mx));
		dp->na = blk->na; dp->nb = blk->nb;									/* fill_restore_fetch equivalent (1217-1263) */
		for(int i = 0; i < n; i++) {
			v.dir = (v.dir << 1) | (uint32_t)(v.acc < 0);
			int down = (int)(v.dir & 1);
			if(down) { dp->nb++; } else { dp->na++; }
			vec_step(dp, &v, down, NULL);
			uint64_t m = 0;
			for(int q = 0; q < W; q++) { if(v.delta[q] > mx[q]) { m |= 1ULL << q; mx[q] = v.delta[q]; } }
			marr[i] = m;
		}
		dp->na = sna; dp->nb = snb;
		dp->vec_count -= (uint64_t)n;
	}
	/* leaf_search_pos (2635-2650): while(m > mask_arr && (max_mask & ~(--m)->all) != 0) { max_mask &= ~m->all; } */
	int m = n;
	for(;;) {
		if(!(m > 0)) { break; }
		m--;
		if((max_mask & ~marr[m]) == 0) { break; }
		max_mask &= ~marr[m];
	}
	uint32_t p = (uint32_t)m, q = (uint32_t)tz64(marr[m] & max_mask);
	lf->blk = b; lf->p = p & 0xff; lf->q = q & 0xff;

	/* restore reverse indices (2745-2762) */
	int32_t fcnt = (int32_t)p + 1;
	uint32_t dir_mask = blk->dir_mask >> (ORA_BLK - fcnt);
	ridx[0] -= (int32_t)((uint32_t)(fcnt - popcnt32(dir_mask)) - (1 + q));
	ridx[1] -= (int32_t)((uint32_t)(0 + popcnt32(dir_mask)) - ((uint32_t)W - q));
	for(int i = 0; i < 2; i++) { lf->gidx[i] = lf->sgidx[i] = 1 - ridx[i] + (int32_t)t->ridx[i]; }
	int32_t rem[2] = { ridx[0] - (int32_t)t->ridx[0], ridx[1] - (int32_t)t->ridx[1] };
	uint64_t plen = (uint64_t)t->f.apos + (uint64_t)t->f.bpos + 2 + (uint64_t)W - (uint64_t)(int64_t)rem[1] - (uint64_t)(int64_t)rem[0];
	return(plen);
}

/* gaba_dp_search_max (gaba.c:2776-2817) */
ora_pos_t ora_dp_search_max(ora_dp_t *dp, int64_t ti)
{
	leaf_t lf; memset(&lf, 0, sizeof(lf));
	ora_pos_t pos;
	pos.plen = leaf_search(dp, ti, &lf);
	int32_t gidx[2] = { lf.gidx[0], lf.gidx[1] }, acc[2] = { 0, 0 };
	ora_tail_t const *t = &dp->tl[ti];
	uint32_t id[2] = { t->f.aid, t->f.bid };
	while(t->tail >= 0) {
		int upd[2] = { 1 > gidx[0], 1 > gidx[1] };
		if(!upd[0] && !upd[1]) { break; }
		uint32_t nid[2] = { t->f.aid, t->f.bid };
		acc[0] += (int32_t)t->adv[0]; acc[1] += (int32_t)t->adv[1];
		t = &dp->tl[t->tail];
		for(int i = 0; i < 2; i++) {
			int mask = upd[i] && t->ridx[i] == 0;
			if(mask) { gidx[i] += acc[i]; id[i] = nid[i]; acc[i] = 0; }
		}
	}
	pos.aid = id[0]; pos.bid = id[1]; pos.apos = (uint32_t)gidx[0]; pos.bpos = (uint32_t)gidx[1];
	return(pos);
}

/* ---------------------------------------------------------------- traceback (gaba.c:2820-3393) */
#define TS_H 1
#define TS_V 2
#define TS_S 4
enum { ts_d = TS_H | TS_V, ts_v0 = TS_V, ts_v1 = TS_V | TS_S, ts_h0 = TS_H, ts_h1 = TS_H | TS_S };
enum { P_D_HEAD, P_D_MID, P_D_TAIL, P_H_HEAD, P_H_BODY, P_H_TAIL, P_V_HEAD, P_V_BODY, P_V_TAIL };

typedef struct {
	int64_t blk; int32_t mi; uint32_t q, state;
	int32_t gidx[2], sgidx[2]; uint32_t ofs[2], id[2];
	int64_t tail[2];
	uint32_t gi[2], ge[2], gf[2];
	uint8_t *pops; uint64_t npop, mpop;			/* popped moves, backward order: 1 = v (down), 0 = h (right) */
	ora_seg_t *seg; uint32_t nseg, mseg;		/* pushed backward */
} trace_t;

static inline uint64_t mbit(ora_dp_t const *dp, uint64_t m, uint32_t q)
{
	/* x86 shift semantics of `mask->x.all >> q` for the mask word width of each W (gaba.c:2952-2973) */
	if(dp->W == 64) { return((m >> (q & 63)) & 1); }
	if(dp->W == 32) { return(((uint32_t)m >> (q & 31)) & 1); }
	return((q & 31) >= 16 ? 0 : ((m >> (q & 31)) & 1));
}

static void trace_pop(trace_t *w, int v)
{
	if(w->npop == w->mpop) { w->mpop = w->mpop ? 2 * w->mpop : 4096; w->pops = realloc(w->pops, w->mpop); }
	w->pops[w->npop++] = (uint8_t)v;
}

/* trace_reload_section (gaba.c:2826-2859) */
static void trace_reload_section(ora_dp_t *dp, trace_t *w, int i)
{
	int64_t tail = w->tail[i], prev = tail;
	int32_t gidx = w->gidx[i];
	while(gidx <= 0) {
		do {
			gidx += dp->tl[tail].istat ? 0 : (int32_t)dp->tl[tail].adv[i];
			prev = tail; tail = dp->tl[tail].tail;
		} while(dp->tl[tail].ridx[i] != 0);
	}
	w->tail[i] = tail;
	w->id[i] = i == 0 ? dp->tl[prev].f.aid : dp->tl[prev].f.bid;
	w->ofs[i] = dp->tl[prev].istat ? dp->tl[prev].adv[i] : 0;
	w->gidx[i] = gidx; w->sgidx[i] = gidx;
}

/* trace_core (gaba.c:3111-3232) as an explicit state machine; returns after _trace_term */
static void trace_core(ora_dp_t *dp, trace_t *w)
{
	int const W = dp->W;
	uint32_t const HEAD_CNT = (uint32_t)(W / ORA_BLK + (W == 16));
	int64_t b = w->blk; int32_t mi = w->mi; uint32_t q = w->q, save = HEAD_CNT;
	uint32_t dir = dp->blk[b].dir_mask >> (ORA_BLK - (mi + 1));
	int32_t *gidx = w->gidx;
	int bulk = 0, pos;
	switch(w->state) {
		case ts_d:  pos = P_D_HEAD; break;
		case ts_v0: pos = P_V_HEAD; break;
		case ts_v1: pos = P_V_TAIL; break;
		case ts_h0: pos = P_H_HEAD; break;
		case ts_h1: pos = P_H_TAIL; break;
		default: return;
	}
	#define MASK()		(&dp->blk[b].mask[mi])
	#define HBIT()		mbit(dp, MASK()->h, q)
	#define VBIT()		mbit(dp, MASK()->v, q)
	#define EBIT()		mbit(dp, MASK()->e, q)
	#define FBIT()		mbit(dp, MASK()->f, q)
	#define NHE()		mbit(dp, ~MASK()->h & MASK()->e, q)
	#define NVF()		mbit(dp, ~MASK()->v & MASK()->f, q)
	/* _trace_test_bulk (3052-3060) */
	#define TEST_BULK() ({ \
		int32_t _ga = gidx[0] - dp->blk[b].acnt, _gb = gidx[1] - dp->blk[b].bcnt; \
		int _ok = !(W > _ga) && !(W > _gb); \
		if(_ok) { gidx[0] = _ga; gidx[1] = _gb; } \
		_ok; \
	})
	/* _trace_reload_block (3032-3043) */
	#define RELOAD_BLOCK() { \
		b--; mi = ORA_BLK - 1; dir = dp->blk[b].dir_mask; \
		if(dp->blk[b].xstat & ORA_X_HEAD) { \
			do { b = dp->blk[b].link; } while(dp->blk[b].xstat & ORA_X_HEAD); \
			int _cnt = dp->blk[b].acnt + dp->blk[b].bcnt; \
			mi = _cnt - 1; dir = dp->blk[b].dir_mask >> (ORA_BLK - _cnt); \
		} \
	}
	/* _pop_vector: update index (tail mode only), push a path bit, move q, then the mode-specific block reload */
	#define POP(_v) { \
		if(!bulk) { gidx[_v]--; } \
		trace_pop(w, _v); mi--; \
		q += (dir & 1) - (uint32_t)(_v); dir >>= 1; \
		if(mi < 0) { \
			if(bulk) {										/* _trace_bulk_load_n (3070-3082) */ \
				RELOAD_BLOCK(); \
				if(!TEST_BULK()) { \
					if(q >= (uint32_t)W) { goto term; } \
					gidx[1] += (int32_t)(q - save); gidx[0] += (int32_t)(save - q); \
					save = HEAD_CNT; bulk = 0; \
				} \
			} else {										/* _trace_tail_load_n (3083-3099) */ \
				if(dp->blk[b - 1].xstat & ORA_X_HEAD) {		/* _trace_reload_tail (3009-3026) */ \
					b--; do { b = dp->blk[b].link; } while(dp->blk[b].xstat & ORA_X_HEAD); \
					int _cnt = dp->blk[b].acnt + dp->blk[b].bcnt; \
					mi = _cnt - 1; dir = dp->blk[b].dir_mask >> (ORA_BLK - _cnt); \
				} else { \
					RELOAD_BLOCK(); \
					if(--save >= HEAD_CNT && TEST_BULK()) { save = q; bulk = 1; } \
				} \
			} \
		} \
	}
	while(1) {
		switch(pos) {
		case P_D_HEAD:
			if(HBIT() != 0) { pos = P_H_HEAD; break; }
			if(!bulk && (gidx[0] == 0 || gidx[1] == 0)) { w->state = ts_d; goto term; }
			POP(0); pos = P_D_MID; break;
		case P_D_MID:
			POP(1); pos = P_D_TAIL; break;
		case P_D_TAIL:
			if(VBIT() != 0) { pos = P_V_HEAD; break; }
			pos = P_D_HEAD; break;
		case P_H_HEAD:
			if(EBIT() == 0) {								/* short gap (gf) */
				if(!bulk && gidx[0] == 0) { w->state = ts_h0; goto term; }
				w->gf[0]++; POP(0); pos = P_D_HEAD; break;
			}
			w->gi[0]++; pos = P_H_BODY; break;
		case P_H_BODY:
			if(!bulk && gidx[0] == 0) { w->state = ts_h1; goto term; }
			w->ge[0]++; POP(0); pos = P_H_TAIL; break;
		case P_H_TAIL:
			pos = (NHE() == 0) ? P_H_BODY : P_D_HEAD; break;
		case P_V_HEAD:
			if(FBIT() == 0) {
				if(!bulk && gidx[1] == 0) { w->state = ts_v0; goto term; }
				w->gf[1]++; POP(1); pos = P_D_TAIL; break;
			}
			w->gi[1]++; pos = P_V_BODY; break;
		case P_V_BODY:
			if(!bulk && gidx[1] == 0) { w->state = ts_v1; goto term; }
			w->ge[1]++; POP(1); pos = P_V_TAIL; break;
		case P_V_TAIL:
			pos = (NVF() == 0) ? P_V_BODY : P_D_TAIL; break;
		}
	}
term:
	w->blk = b; w->mi = mi; w->q = q & 0xff;				/* uint8 store (gaba.c:493, 3227) */
	#undef MASK
	#undef HBIT
	#undef VBIT
	#undef EBIT
	#undef FBIT
	#undef NHE
	#undef NVF
	#undef TEST_BULK
	#undef RELOAD_BLOCK
	#undef POP
}

void ora_aln_free(ora_aln_t *a) { if(a) { free(a->seg); free(a->path); free(a); } }

/* gaba_dp_trace / trace_body / trace_init / trace_push_segment (gaba.c:2865-2895, 3244-3393) */
ora_aln_t *ora_dp_trace(ora_dp_t *dp, int64_t ti)
{
	ora_tail_t const *t = &dp->tl[ti];
	leaf_t lf; memset(&lf, 0, sizeof(lf));
	uint64_t plen = t->f.bpos < -1 ? 0 : leaf_search(dp, ti, &lf);
	trace_t w; memset(&w, 0, sizeof(w));
	w.blk = lf.blk; w.mi = (int32_t)lf.p; w.q = lf.q; w.state = ts_d;
	for(int i = 0; i < 2; i++) { w.gidx[i] = lf.gidx[i]; w.sgidx[i] = lf.sgidx[i]; w.tail[i] = ti; }
	while(w.npop < plen) {									/* path + ofs > aln->path */
		if(w.gidx[0] < (int32_t)((w.state & TS_H) != 0)) { trace_reload_section(dp, &w, 0); }
		if(w.gidx[1] < (int32_t)((w.state & TS_V) != 0)) { trace_reload_section(dp, &w, 1); }
		trace_core(dp, &w);
		if(w.q >= (uint32_t)dp->W) { free(w.pops); free(w.seg); return(NULL); }
		/* trace_push_segment */
		if(w.nseg == w.mseg) { w.mseg = w.mseg ? 2 * w.mseg : 8; w.seg = realloc(w.seg, sizeof(ora_seg_t) * w.mseg); }
		ora_seg_t *s = &w.seg[w.nseg++];
		s->aid = w.id[0]; s->bid = w.id[1];
		s->apos = w.ofs[0] + (uint32_t)w.gidx[0]; s->bpos = w.ofs[1] + (uint32_t)w.gidx[1];
		s->alen = (uint32_t)(w.sgidx[0] - w.gidx[0]); s->blen = (uint32_t)(w.sgidx[1] - w.gidx[1]);
		s->ppos = plen - w.npop;
		w.sgidx[0] = w.gidx[0]; w.sgidx[1] = w.gidx[1];
	}
	ora_aln_t *a = calloc(1, sizeof(ora_aln_t));
	uint32_t gcnt[2] = { w.ge[0] + w.gf[0], w.ge[1] + w.gf[1] };
	/* gaba.c:3340-3349: _mul_v2i32 is _mm_mul_epi32 (v2i32.h:106), a 32x32->64 multiply of the LOW lane only, so the
	 * b-side (v-gap) counters never enter the sum: the high lane receives the (zero) upper halves of the products */
	int32_t g[2] = { (int32_t)(dp->gi * w.gi[0] + dp->ge * w.ge[0] + dp->gfa * w.gf[0]), 0 };
	uint64_t dlen = ((uint32_t)plen - gcnt[1] - gcnt[0]) >> 1;
	int64_t dsc = t->f.max + g[1] + g[0];
	a->score = t->f.max;
	a->identity = dlen == 0 ? 0.0 : (((double)dsc / (double)dlen) * dp->imx - dp->xmx);
	a->agcnt = gcnt[0]; a->bgcnt = gcnt[1]; a->dcnt = (uint32_t)dlen; a->plen = (uint32_t)plen;
	a->slen = w.nseg;
	a->seg = malloc(sizeof(ora_seg_t) * (w.nseg ? w.nseg : 1));
	for(uint32_t i = 0; i < w.nseg; i++) { a->seg[i] = w.seg[w.nseg - 1 - i]; }
	a->npath = (uint32_t)((plen + 31) / 32 + 1);
	a->path = calloc(a->npath + 2, sizeof(uint32_t));
	for(uint64_t i = 0; i < plen; i++) { if(w.pops[plen - 1 - i]) { a->path[i >> 5] |= 1u << (i & 31); } }
	a->path[plen >> 5] |= 1u << (plen & 31);				/* sentinel (gaba.c:3287) */
	free(w.pops); free(w.seg);
	return(a);
}

/* ---------------------------------------------------------------- CIGAR (gaba_parse.h:107-263) */
/* 64 path bits starting at (possibly negative) bit position pos; bits below zero come from the two header words
 * that precede path[] in gaba_alignment_s: plen, then padding = 0x40000000 (gaba.c:3277-3278, gaba.h:217) */
static uint64_t path_bits(uint32_t const *path, uint32_t plen, int64_t pos)
{
	uint64_t r = 0;
	int64_t w0 = pos >> 5;									/* arithmetic shift */
	uint32_t sh = (uint32_t)(pos & 31);
	uint32_t wd[3];
	for(int k = 0; k < 3; k++) {
		int64_t wi = w0 + k;
		wd[k] = wi >= 0 ? path[wi] : (wi == -1 ? 0x40000000u : (wi == -2 ? plen : 0));
	}
	r = ((uint64_t)wd[0] | ((uint64_t)wd[1] << 32)) >> sh;
	if(sh) { r |= (uint64_t)wd[2] << (64 - sh); }
	return(r);
}

static uint64_t dump_num(char *buf, uint64_t len, char ch)
{
	char tmp[24]; int n = 0;
	do { tmp[n++] = (char)('0' + len % 10); len /= 10; } while(len);
	for(int i = 0; i < n; i++) { buf[i] = tmp[n - 1 - i]; }
	buf[n] = ch;
	return((uint64_t)n + 1);
}

/* _parser_loop_rv (gaba_parse.h:162-184) with the cigar callbacks (240-263); path must carry >= 2 readable words past
 * the sentinel (ora_aln_t does) */
uint64_t ora_dump_cigar_reverse(char *buf, uint32_t const *path, uint64_t offset, uint64_t len)
{
	char *b = buf;
	uint32_t plen_hdr = 0;									/* only reached for len < 64 at offset 0; value irrelevant beyond bit 30 of padding */
	int64_t ofs = (int64_t)offset - 64; uint64_t idx = len;
	while((int64_t)idx > 0) {
		uint64_t m, c;
		m = lz64(path_bits(path, plen_hdr, ofs + (int64_t)idx));
		c = MIN2(idx, m - (m > 0)); idx -= c; if(c) { b += dump_num(b, c, 'D'); }
		m = lz64(~path_bits(path, plen_hdr, ofs + (int64_t)idx));
		c = MIN2(idx, m); idx -= c; if(c) { b += dump_num(b, c, 'I'); }
		uint64_t sidx = idx;
		do {
			m = lz64(path_bits(path, plen_hdr, ofs + (int64_t)idx) ^ 0x5555555555555555ULL);
			c = MIN2(idx, m) & ~0x01ULL; idx -= c;
		} while(c == 64);
		if((sidx - idx) >> 1) { b += dump_num(b, (sidx - idx) >> 1, 'M'); }
	}
	*b = '\0';
	return((uint64_t)(b - buf));
}

/* _parser_loop_fw (gaba_parse.h:143-161) */
uint64_t ora_dump_cigar_forward(char *buf, uint32_t const *path, uint64_t offset, uint64_t len)
{
	char *b = buf;
	uint64_t lim = offset + len, ridx = len;
	while((int64_t)ridx > 0) {
		uint64_t m, c;
		m = tz64(~path_bits(path, 0, (int64_t)(lim - ridx)));
		c = MIN2(ridx, m - (m > 0)); ridx -= c; if(c) { b += dump_num(b, c, 'I'); }
		m = tz64(path_bits(path, 0, (int64_t)(lim - ridx)));
		c = MIN2(ridx, m); ridx -= c; if(c) { b += dump_num(b, c, 'D'); }
		uint64_t sridx = ridx;
		do {
			m = tz64(path_bits(path, 0, (int64_t)(lim - ridx)) ^ 0x5555555555555555ULL);
			c = MIN2(ridx, m) & ~0x01ULL; ridx -= c;
		} while(c == 64);
		if((sridx - ridx) >> 1) { b += dump_num(b, (sridx - ridx) >> 1, 'M'); }
	}
	*b = '\0';
	return((uint64_t)(b - buf));
}
