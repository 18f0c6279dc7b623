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

/* ================================================================ index blob access (2476-2483, 2693-2697, 3070-3167) */
struct mmo_s {
	uint8_t const *blob; uint64_t size;
	mmo_params_t p;
	uint64_t bkt_ofs, mask, seq_ofs;
	uint32_t b, w, k, n_occ, occ[8], n_seq;
	double mcoef, xcoef;
	uint32_t twlen, tglen;
	ora_dp_t dp[3];

	/* per-read work (mm_tbuf_t, 3301-3334) */
	uint32_t rid, qid, rlen, qlen;
	uint8_t const *qseq;
	uint64_t *sk; uint64_t nsk, msk;						/* sketch words */
	uint32_t *resc; uint64_t nresc, mresc, presc;			/* {qs, n, p_lo, p_hi} */
	uint32_t *seed; uint64_t nseed_arr, mseed, n_seed;		/* {upos, rid, vpos, lid} x seed.n */
	uint32_t *root; uint64_t nroot, mroot;					/* {plen, lid}; reused as res {score, iid} */
	uint32_t *next; uint64_t nnext, mnext;
	uint32_t n_res;
	uint64_t *bin; uint64_t nbin, mbin;						/* slots: bin header (2 slots) or alignment index */
	ora_aln_t **alns; uint64_t naln, maln;					/* alignment store ("lmm") */
	uint64_t *kh; uint32_t kh_mask, kh_max, kh_cnt, kh_ub;	/* dedup hash `pos`: {key, val} pairs */
};

static inline uint64_t ld64(uint8_t const *p) { uint64_t v; memcpy(&v, p, 8); return(v); }
static inline uint32_t ld32(uint8_t const *p) { uint32_t v; memcpy(&v, p, 4); return(v); }
static inline uint16_t ld16(uint8_t const *p) { uint16_t v; memcpy(&v, p, 2); return(v); }

typedef struct { uint8_t const *seq; char const *name; uint32_t l_seq, l_name, circular; } mmo_ref_t;
static mmo_ref_t mmo_ref(mmo_t const *m, uint32_t rid)
{
	uint8_t const *s = m->blob + m->seq_ofs + 24 * (uint64_t)rid;
	mmo_ref_t r = { m->blob + ld64(s), (char const *)(m->blob + ld64(s + 8)), ld32(s + 16), ld16(s + 20), ld16(s + 22) };
	return(r);
}

uint64_t mmo_vec_count(mmo_t *m) { return(m->dp[0].vec_count + m->dp[1].vec_count + m->dp[2].vec_count); }

static void kh_reset(mmo_t *m);

mmo_t *mmo_init(void const *blob, uint64_t size, mmo_params_t const *p)
{
	mmo_t *m = calloc(1, sizeof(mmo_t));
	m->blob = blob; m->size = size; m->p = *p;
	m->bkt_ofs = ld64(m->blob); m->mask = ld64(m->blob + 8);
	m->b = m->blob[16]; m->w = m->blob[17]; m->k = m->blob[18]; m->n_occ = m->blob[19];
	for(int i = 0; i < 7; i++) { m->occ[i] = ld32(m->blob + 20 + 4 * i); }
	m->n_seq = ld32(m->blob + 48); m->seq_ofs = ld64(m->blob + 56);
	/* mm_align_init (4676-4681): both branches add score_matrix[0] */
	double mc = 0.0, xc = 0.0;
	for(int i = 0; i < 16; i++) { if((i & 3) == (i >> 3)) { mc += p->gp.score_matrix[0]; } else { xc += p->gp.score_matrix[0]; } }
	m->mcoef = mc / 4.0; m->xcoef = xc / 12.0;
	m->twlen = (uint32_t)((p->wlen << 1) - p->wlen); m->tglen = (uint32_t)((p->glen << 1) - p->glen);	/* 4506 */
	static int const bw[3] = { 64, 32, 16 };
	for(int i = 0; i < 3; i++) { if(ora_dp_init(&m->dp[i], &p->gp, bw[i]) != 0) { free(m); return(NULL); } }
	m->kh_max = 256; m->kh = malloc(16 * (size_t)m->kh_max);	/* kh_init_static(&t->pos, 128) -> 256 (4520, 369-373) */
	kh_reset(m);
	return(m);
}

void mmo_destroy(mmo_t *m)
{
	if(m == NULL) { return; }
	for(int i = 0; i < 3; i++) { ora_dp_clean(&m->dp[i]); }
	for(uint64_t i = 0; i < m->naln; i++) { ora_aln_free(m->alns[i]); }
	free(m->sk); free(m->resc); free(m->seed); free(m->root); free(m->next); free(m->bin); free(m->alns); free(m->kh);
	free(m);
}

/* ================================================================ sketch (2349-2435) */
static uint32_t crc32c_u64(uint32_t crc, uint64_t data)		/* _mm_crc32_u64: reflected Castagnoli, no pre/post inversion */
{
	static uint32_t tab[256]; static int init = 0;
	if(!init) {
		for(uint32_t i = 0; i < 256; i++) { uint32_t c = i; for(int j = 0; j < 8; j++) { c = (c >> 1) ^ (0x82f63b78u & (0u - (c & 1))); } tab[i] = c; }
		init = 1;
	}
	for(int i = 0; i < 8; i++) { crc = tab[(crc ^ (uint32_t)(data >> (8 * i))) & 0xff] ^ (crc >> 8); }
	return(crc);
}

#define RESERVE(type, arr, cap, need) { if((need) > (cap)) { (cap) = MAX2(2 * (cap), (uint64_t)(need) + 256); (arr) = realloc((arr), sizeof(type) * (cap)); } }

static void sketch(mmo_t *m, uint8_t const *seq, uint32_t len)
{
	uint64_t const k = m->k, w = m->w, kk = k - 1, shift1 = 2 * kk, mask = (1ULL << 2 * k) - 1;
	RESERVE(uint64_t, m->sk, m->msk, m->nsk + 4 * (uint64_t)len / w + 256 + 8);
	uint64_t *q = m->sk + m->nsk;
	uint64_t r[64]; for(int i = 0; i < 64; i++) { r[i] = UINT64_MAX; }
	uint64_t u = 0, k0 = 0, k1 = 0, p = 0, f;
	#define PUSH_KMER() { uint64_t c = seq[p++]; k0 = (k0 << 2 | c) & mask; k1 = (k1 >> 2) | ((3ULL ^ c) << shift1); }
	#define CORE(i) ({ \
		PUSH_KMER(); \
		uint64_t km = k0 < k1 ? k0 : k1, kx = k0 < k1 ? k1 : k0, mm = k0 < k1 ? 0 : 0x80; \
		uint64_t h = (((uint64_t)crc32c_u64((uint32_t)kx, kx) ^ km) & mask) << 8 | (i) | mm; \
		f = MIN2(f, h); uint64_t v = MIN2(f, r[(i) + 1]); \
		if((v == h) | (v - u)) { *q++ = v; } \
		u = v; h; \
	})
	for(uint64_t i = 0; i < kk && p < len; i++) { PUSH_KMER(); }
	while((int64_t)(len - p) >= (int64_t)w) {
		f = UINT64_MAX; for(uint64_t i = 0; i < w; i++) { r[i] = CORE(i); }
		uint64_t rr = UINT64_MAX; for(uint64_t i = 0; i < w; i++) { rr = MIN2(rr, r[w - i - 1]); r[w - i - 1] = rr; }
	}
	uint64_t l = len - p;
	if(l > 0) {
		f = UINT64_MAX; for(uint64_t i = 0; i < l; i++) { r[w + i] = CORE(i) + w; }
		uint64_t rr = UINT64_MAX; for(uint64_t i = 0; i < w; i++) { rr = MIN2(rr, r[w + l - i - 1]); r[w + l - i - 1] = rr; }
		for(uint64_t i = 0; i < w; i++) { r[i] = r[l + i] - l; }
		u += w - l;
	}
	*q++ = 0xffffffffffff0000ULL; *q++ = u; *q++ = k0; *q++ = k1;		/* cap */
	m->nsk = (uint64_t)(q - m->sk);
	#undef PUSH_KMER
	#undef CORE
}

uint64_t mmo_sketch(mmo_t *m, uint8_t const *seq, uint32_t len, uint64_t *out, uint64_t cap)
{
	m->nsk = 0; sketch(m, seq, len);
	if(m->nsk <= cap) { memcpy(out, m->sk, 8 * m->nsk); }
	return(m->nsk);
}

/* ================================================================ index probe (2727-2748, 634-643) */
static uint8_t const *idx_get(mmo_t const *m, uint64_t minier, uint32_t *n)
{
	uint8_t const *bk = m->blob + m->bkt_ofs + 32 * (minier & m->mask);
	uint64_t hmask = ld32(bk), a = ld64(bk + 16), pofs = ld64(bk + 24);
	*n = 0;
	if(a == 0) { return(NULL); }
	uint64_t key = minier >> m->b, pos = key & hmask, kk;
	uint8_t const *val = NULL;
	do {
		uint8_t const *slot = m->blob + a + 16 * pos;
		if((kk = ld64(slot)) == key) { val = slot + 8; break; }
		pos = hmask & (pos + 1);
	} while(kk + 1 != 0);
	if(val == NULL) { return(NULL); }
	uint64_t v = ld64(val);
	if((int64_t)v >= 0) { *n = 1; return(val); }
	*n = (uint32_t)v;
	return(m->blob + pofs + 8 * ((v >> 32) & 0x7fffffff));
}

uint32_t mmo_get(mmo_t *m, uint64_t minier, uint64_t *out, uint32_t cap)
{
	uint32_t n; uint8_t const *r = idx_get(m, minier, &n);
	for(uint32_t i = 0; i < n && i < cap; i++) { out[i] = ld64(r + 8 * i); }
	return(n);
}

/* ================================================================ exact radix sort (ksort.h:82-131) */
/* American-flag MSD radix sort on 8-bit digits with an insertion-sort cutoff at 64 elements; restated with index
 * arithmetic.  The visiting order of the permutation cycles is what fixes the order of equal keys, so it follows the
 * reference step by step.  esz = element size in u32 words (4: key = words 0,1 as u64; 2: key = word 0). */
static inline uint64_t rs_key(uint32_t const *e, int esz) { return(esz == 4 ? ((uint64_t)e[1] << 32 | e[0]) : e[0]); }

static void rs_insertion(uint32_t *a, uint64_t n, int esz)
{
	uint32_t tmp[4];
	for(uint64_t i = 1; i < n; i++) {
		if(rs_key(a + esz * i, esz) < rs_key(a + esz * (i - 1), esz)) {
			memcpy(tmp, a + esz * i, 4 * esz);
			uint64_t j;
			for(j = i; j > 0 && rs_key(tmp, esz) < rs_key(a + esz * (j - 1), esz); j--) { memcpy(a + esz * j, a + esz * (j - 1), 4 * esz); }
			memcpy(a + esz * j, tmp, 4 * esz);
		}
	}
}

static void rs_flag(uint32_t *a, uint64_t n, int esz, int s)
{
	uint64_t head[256], end[256];
	memset(end, 0, sizeof(end));
	for(uint64_t i = 0; i < n; i++) { end[(rs_key(a + esz * i, esz) >> s) & 0xff]++; }
	head[0] = 0;
	for(int k = 1; k < 256; k++) { end[k] += end[k - 1]; head[k] = end[k - 1]; }
	for(int k = 0; k < 256;) {
		if(head[k] != end[k]) {
			int l = (int)((rs_key(a + esz * head[k], esz) >> s) & 0xff);
			if(l != k) {
				uint32_t tmp[4], swp[4];
				memcpy(tmp, a + esz * head[k], 4 * esz);
				do {
					memcpy(swp, tmp, 4 * esz); memcpy(tmp, a + esz * head[l], 4 * esz); memcpy(a + esz * head[l], swp, 4 * esz); head[l]++;
					l = (int)((rs_key(tmp, esz) >> s) & 0xff);
				} while(l != k);
				memcpy(a + esz * head[k], tmp, 4 * esz); head[k]++;
			} else { head[k]++; }
		} else { k++; }
	}
	if(s) {
		s = s > 8 ? s - 8 : 0;
		uint64_t beg = 0;
		for(int k = 0; k < 256; k++) {
			uint64_t sz = end[k] - beg;
			if(sz > 64) { rs_flag(a + esz * beg, sz, esz, s); }
			else if(sz > 1) { rs_insertion(a + esz * beg, sz, esz); }
			beg = end[k];
		}
	}
}

static void radix_sort(uint32_t *a, uint64_t n, int esz)
{
	if(n <= 64) { rs_insertion(a, n, esz); } else { rs_flag(a, n, esz, esz == 4 ? 56 : 24); }
}

/* ================================================================ seeds (3340-3541) */
#define OFS0 0x40000000u
static inline uint32_t u_of(uint32_t x, uint32_t y) { return(((x << 1) - y) + OFS0); }
static inline uint32_t v_of(uint32_t x, uint32_t y) { return(((y << 1) - x) + OFS0); }
static inline int32_t as_of(uint32_t const *s) { return((int32_t)(((s[0] - OFS0) << 1) + (s[2] - OFS0)) / 3); }
static inline int32_t bs_of(uint32_t const *s) { return((int32_t)(((s[2] - OFS0) << 1) + (s[0] - OFS0)) / 3); }

static void expand(mmo_t *m, uint32_t n, uint8_t const *r, uint32_t qs)
{
	if(n == 0) { return; }
	RESERVE(uint32_t, m->seed, m->mseed, 4 * (m->nseed_arr + n + 2));
	for(uint32_t i = 0; i < n; i++) {
		uint32_t rid = ld32(r + 8 * i + 4);
		if(rid < m->qid) { continue; }
		uint32_t rs = ld32(r + 8 * i), rmask = 0u - (rid & 1);
		uint32_t _rs = rs + (m->k & rmask), _qs = qs ^ rmask;
		uint32_t *s = m->seed + 4 * m->nseed_arr++;
		s[0] = u_of(_rs, _qs); s[1] = rid >> 1; s[2] = v_of(_rs, _qs); s[3] = INT32_MAX;
	}
}

static void collect_seed(mmo_t *m)
{
	m->nsk = 0; sketch(m, m->qseq, m->qlen);
	RESERVE(uint32_t, m->resc, m->mresc, 4 * (m->nsk + 1));
	m->nresc = 0;
	uint32_t const max_occ = m->occ[m->n_occ - 1], resc_occ = m->occ[0];
	uint64_t w = m->w, base = 0 - w, v = w;
	for(uint64_t const *p = m->sk; !(((int64_t)*p) >> 16 == (int64_t)-1); p++) {
		uint64_t u = *p & 0x7f, fr = (*p >> 7) & 1, h = *p >> 8;
		base += u <= v ? w : 0; v = u;
		uint32_t n; uint8_t const *r = idx_get(m, h, &n);
		if(n > max_occ) { continue; }
		uint32_t pos = (uint32_t)((base + u + (m->k & (0 - fr))) ^ (0 - fr));
		if(n > resc_occ) {
			uint32_t *s = m->resc + 4 * m->nresc++;
			uint64_t ofs = (uint64_t)(r - m->blob);
			s[0] = pos; s[1] = n; s[2] = (uint32_t)ofs; s[3] = (uint32_t)(ofs >> 32);
			continue;
		}
		expand(m, n, r, pos);
	}
	m->presc = 0;
}

static uint64_t mm_seed(mmo_t *m, uint64_t cnt)
{
	if(cnt == 0) {
		m->nseed_arr = 0; m->n_seed = 0;
		collect_seed(m);
	} else {
		if(cnt == 1) { radix_sort(m->resc, m->nresc, 4); }
		m->nseed_arr = m->n_seed;
		for(uint64_t i = 0; i < m->nseed_arr; i++) { m->seed[4 * i + 3] = INT32_MAX; }
		while(m->presc < m->nresc && m->resc[4 * m->presc + 1] <= m->occ[cnt]) {
			uint32_t const *p = m->resc + 4 * m->presc;
			expand(m, p[1], m->blob + ((uint64_t)p[2] | (uint64_t)p[3] << 32), p[0]);
			m->presc++;
		}
	}
	m->n_seed = m->nseed_arr;
	if(m->nseed_arr == 0) { return(0); }
	RESERVE(uint32_t, m->seed, m->mseed, 4 * (m->nseed_arr + 2));
	uint32_t *s = m->seed + 4 * m->nseed_arr++;
	s[0] = (uint32_t)INT32_MIN; s[1] = INT32_MAX; s[2] = (uint32_t)INT32_MIN; s[3] = INT32_MAX;	/* sentinel */
	radix_sort(m->seed, m->nseed_arr, 4);
	return(m->nseed_arr);
}

/* ================================================================ chain (3372-3402, 3547-3721) */
/* window vector lanes: (uub, rid, vub, vlb); position vector lanes: (upos, rid, vpos, vpos); compares are signed */
typedef struct { int32_t l[4]; } v4_t;
static inline v4_t load_pv(uint32_t const *s) { v4_t r = { { (int32_t)s[0], (int32_t)s[1], (int32_t)s[2], (int32_t)s[2] } }; return(r); }
static inline v4_t load_wv(uint32_t const *s, int32_t len)
{
	v4_t r = load_pv(s);
	r.l[0] = (int32_t)((uint32_t)r.l[0] + (uint32_t)len); r.l[2] = (int32_t)((uint32_t)r.l[2] + (uint32_t)len);
	return(r);
}
static inline uint32_t inside_mask(v4_t w, v4_t d)
{
	uint32_t mk = 0;
	for(int i = 0; i < 4; i++) { if(d.l[i] > w.l[i]) { mk |= 0xfu << (4 * i); } }
	return(mk);
}
#define inside_wv(w, d)		( inside_mask(w, d) == 0xf000 )
#define inside_uub(w, d)	( (inside_mask(w, d) & 0xff) == 0 )
static inline v4_t update_wv(v4_t w, v4_t f)
{
	int32_t d0 = (int32_t)((uint32_t)w.l[0] - (uint32_t)f.l[0]), d2 = (int32_t)((uint32_t)w.l[2] - (uint32_t)f.l[2]);
	w.l[0] = (int32_t)((uint32_t)w.l[0] - (uint32_t)d2); w.l[2] = (int32_t)((uint32_t)w.l[2] - (uint32_t)d0);
	return(w);
}
static inline int32_t pdiff(v4_t w, v4_t f)
{
	return((int32_t)(((uint32_t)w.l[0] - (uint32_t)f.l[0]) + ((uint32_t)w.l[2] - (uint32_t)f.l[2])));
}

static uint64_t chain_seeds(mmo_t *m)
{
	uint32_t *s = m->seed, *c = m->root;
	uint32_t ncid = 0, nlid = (uint32_t)m->n_seed + 1, nlsid = 0, tsid = (uint32_t)m->n_seed;
	while(nlsid < tsid) {
		uint32_t lid = nlid++;
		uint32_t *lf = s + 4 * lid;								/* leaf: {rsid, rid, lsid, cid} */
		lf[0] = nlsid; lf[2] = nlsid; lf[1] = s[4 * nlsid + 1]; lf[3] = UINT32_MAX;
		uint32_t plen = s[4 * nlsid] + s[4 * nlsid + 2], scnt = 1;
		uint64_t nrsid = nlsid; nlsid = UINT32_MAX;
		while(1) {
			uint32_t rsid = (uint32_t)nrsid; nrsid = 0;
			v4_t wv = load_wv(s + 4 * rsid, (int32_t)m->twlen);
			for(uint32_t sid = rsid + 1; ; sid++) {
				v4_t fv = load_pv(s + 4 * sid);
				if(!inside_wv(wv, fv)) {
					nlsid = MIN2(nlsid, sid);
					if(inside_uub(wv, fv)) { continue; }
					break;
				}
				wv = update_wv(wv, fv);
				int64_t di = (int64_t)(((uint64_t)(int64_t)pdiff(wv, fv) << 32) | sid);
				nrsid = (uint64_t)MAX2((int64_t)nrsid, di);
			}
			if(nrsid == 0) { nrsid = rsid; break; }
			if(s[4 * (uint32_t)nrsid + 3] != INT32_MAX) { nrsid = (uint32_t)nrsid; break; }
			s[4 * (uint32_t)nrsid + 3] = lid; scnt++;
			if(nlsid <= nrsid) { nlsid = UINT32_MAX; }
		}
		if(nrsid == lf[2]) { continue; }
		uint32_t cid = UINT32_MAX;
		if(s[4 * nrsid + 3] < lid) {
			nrsid = s[4 * s[4 * nrsid + 3] + 0];					/* leaf[seed.lid].rsid */
			cid = s[4 * s[4 * nrsid + 3] + 3];						/* leaf[seed(new nrsid).lid].cid */
		}
		if(cid == UINT32_MAX) { cid = ncid++; c[2 * cid] = OFS0; c[2 * cid + 1] = lid; }
		lf[3] = cid; lf[0] = (uint32_t)nrsid;
		uint32_t ps = s[4 * nrsid] + s[4 * nrsid + 2];
		plen = (uint32_t)((int32_t)OFS0 - (int32_t)((uint32_t)((1.0 - 1.0 / (double)scnt) * (double)(uint32_t)(ps - plen))));
		if(plen < c[2 * cid]) { c[2 * cid] = plen; c[2 * cid + 1] = lid; }
	}
	m->nroot = ncid; m->nseed_arr = nlid;
	return(ncid);
}

static uint64_t mm_chain(mmo_t *m)
{
	RESERVE(uint32_t, m->seed, m->mseed, 4 * (2 * m->nseed_arr + 2));
	RESERVE(uint32_t, m->root, m->mroot, 2 * (m->nseed_arr + 2));
	RESERVE(uint32_t, m->next, m->mnext, 2 * (m->nseed_arr + 2));
	m->nroot = 0; m->nnext = 0;
	if(chain_seeds(m) == 0) { return(0); }
	/* mm_circularize (3632-3695) is a no-op unless the index marks a reference circular (-c), which this path does not support */
	radix_sort(m->root, m->nroot, 2);
	return(m->nroot);
}

uint64_t mmo_seed_chain(mmo_t *m, uint8_t const *seq, uint32_t len, uint32_t round,
	uint32_t *seeds, uint64_t seed_cap, uint64_t *n_total, uint32_t *roots, uint64_t root_cap, uint64_t *n_root)
{
	m->qid = 0; m->qlen = len; m->qseq = seq; m->nresc = 0; m->presc = 0; m->nseed_arr = 0; m->n_seed = 0; m->nroot = 0;
	uint64_t ns = 0, nr = 0;
	for(uint32_t i = 0; i <= round && i < m->n_occ; i++) {
		ns = mm_seed(m, i);
		nr = ns ? mm_chain(m) : 0;
	}
	*n_total = ns ? m->nseed_arr : 0; *n_root = nr;
	if(ns && m->nseed_arr <= seed_cap) { memcpy(seeds, m->seed, 16 * m->nseed_arr); }
	if(nr && nr <= root_cap) { memcpy(roots, m->root, 8 * nr); }
	return(ns ? m->n_seed : 0);
}

/* ================================================================ dedup hash `pos` (kh_t, 346-613) */
/* Ordered linear probing with Robin-Hood displacement.  Restated literally because two of its properties are
 * observable: a key that shares its home slot with a later-inserted key is not recognised as a duplicate
 * (kh_allocate stops at the first resident whose home >= ours), and the value pointer returned for the head key
 * can go stale when the tail key's insertion displaces it (4027-4029). */
#define KH_EMPTY	UINT64_MAX
#define KH_MOVED	(UINT64_MAX - 1)
#define KH_INIT		UINT64_MAX

static void kh_reset(mmo_t *m)													/* kh_clear (481-495) */
{
	m->kh_mask = 255; m->kh_cnt = 0; m->kh_ub = (uint32_t)(256 * 0.4);
	for(uint64_t i = 0; i < 256; i++) { m->kh[2 * i] = KH_EMPTY; m->kh[2 * i + 1] = KH_INIT; }
}

typedef struct { uint64_t idx, n; } kh_bidx_t;
static kh_bidx_t kh_allocate(uint64_t *a, uint64_t k, uint64_t v, uint64_t mask)	/* 503-536 */
{
	#define POLL(_i, _b0) ({ \
		int64_t _b = (int64_t)(_b0); uint64_t _k1; \
		while(1) { \
			_k1 = a[2 * (_i)]; \
			if(_b <= (int64_t)(_k1 & mask) + (int64_t)(_k1 + 2 < 2)) { break; } \
			_b -= (int64_t)(((_i) + 1) & (mask + 1)); \
			(_i) = ((_i) + 1) & mask; \
		} \
		_k1; \
	})
	uint64_t i = k & mask, k0 = k, v0 = v;
	uint64_t k1 = POLL(i, i);
	if(k0 == k1) { return((kh_bidx_t){ i, 0 }); }
	uint64_t j = i;
	a[2 * i] = k0;
	while(k1 + 2 >= 2) {
		uint64_t v1 = a[2 * i + 1];
		a[2 * i + 1] = v0;
		k0 = k1; v0 = v1;
		i = (i + 1) & mask;
		k1 = POLL(i, k0 & mask);
		a[2 * i] = k0;
	}
	a[2 * i + 1] = v0;
	return((kh_bidx_t){ j, 1 });
	#undef POLL
}

static void kh_extend(mmo_t *m)													/* 543-579 */
{
	uint64_t prev_size = (uint64_t)m->kh_mask + 1, size = 2 * prev_size, mask = size - 1;
	m->kh_mask = (uint32_t)mask; m->kh_ub = (uint32_t)(size * 0.4);
	if(size > m->kh_max) { m->kh = realloc(m->kh, 16 * size); m->kh_max = (uint32_t)size; }
	for(uint64_t i = 0; i < prev_size; i++) { m->kh[2 * (i + prev_size)] = KH_EMPTY; m->kh[2 * (i + prev_size) + 1] = KH_INIT; }
	for(uint64_t i = 0; i < size; i++) {
		uint64_t k = m->kh[2 * i];
		if(k + 2 < 2 || (k & mask) == i) { continue; }
		uint64_t v = m->kh[2 * i + 1];
		m->kh[2 * i] = KH_MOVED; m->kh[2 * i + 1] = KH_INIT;
		kh_allocate(m->kh, k, v, mask);
	}
}

static uint64_t kh_put_ptr(mmo_t *m, uint64_t key, int extend)					/* 604-613; returns the slot index */
{
	if(extend && m->kh_cnt >= m->kh_ub) { kh_extend(m); }
	kh_bidx_t b = kh_allocate(m->kh, key, KH_INIT, m->kh_mask);
	m->kh_cnt += (uint32_t)b.n;
	return(b.idx);
}
#define KH_VAL(_m, _slot)	( (_m)->kh[2 * (_slot) + 1] )

static inline uint64_t pos_key(uint64_t x, uint64_t y) { return(x ^ (x >> 29) ^ y ^ __builtin_bswap64(y)); }	/* 3362 */

/* ================================================================ extend state machine (3778-4173) */
#define MM_CREM 50000
#define MM_SREM 8
typedef struct {
	uint32_t cp[2], tp[2];
	uint32_t aid, bid, iid, eid, sid, rev;
	int64_t prem; uint32_t pacc, crem, srem, narrow, min_score;
} search_t;

/* result bins live in a slot array like the reference's ptr_v: a 2-slot header {n_aln, plen | lb, ub} then one
 * slot per alignment (3252-3258) */
#define BIN_NALN(_m, _iid)	( ((uint32_t *)&(_m)->bin[_iid])[0] )
#define BIN_PLEN(_m, _iid)	( ((uint32_t *)&(_m)->bin[_iid])[1] )
#define BIN_LB(_m, _iid)	( ((uint32_t *)&(_m)->bin[(_iid) + 1])[0] )
#define BIN_UB(_m, _iid)	( ((uint32_t *)&(_m)->bin[(_iid) + 1])[1] )
#define BIN_ALN(_m, _iid, _j)	( (_m)->alns[(_m)->bin[(_iid) + 2 + (_j)]] )

static uint64_t bin_push(mmo_t *m, uint64_t v)
{
	RESERVE(uint64_t, m->bin, m->mbin, m->nbin + 4);
	m->bin[m->nbin] = v;
	return(m->nbin++);
}

static void load_pos(mmo_t const *m, uint32_t const *p, uint32_t *rev, uint32_t *cp)		/* 3817-3833 */
{
	int32_t bs = bs_of(p);
	*rev = bs < 0;
	cp[0] = (uint32_t)as_of(p);
	cp[1] = (uint32_t)bs + ((uint32_t)(bs >> 31) & m->qlen);
	if(cp[0] >= m->rlen || cp[1] >= m->qlen) {
		cp[0] -= MIN2(cp[0], m->k);
		cp[1] -= MIN2(cp[1], m->k);
	}
}

static int load_root(mmo_t *m, search_t *st, uint32_t cid)									/* 3838-3881 */
{
	uint32_t const *s = m->seed;
	uint32_t lid = m->root[2 * cid + 1];
	uint32_t plen = (uint32_t)((int32_t)OFS0 - (int32_t)m->root[2 * cid]);
	if(plen * m->mcoef < 2.0 * m->p.min_score) { return(1); }
	m->nnext = 0;
	uint32_t iid = (uint32_t)bin_push(m, 0); bin_push(m, 0);
	/* The source initialises the header as (mm_bin_t){ .lb = UINT32_MAX } but copies it through a void** type-pun
	 * (3855, kvec.h:105-115); the reference's AVX2 build (gcc 13.3 -O3, strict aliasing) drops the dead .lb store and
	 * emits two `movq $0` stores, i.e. lb starts at 0 (checked in the disassembly of oracle/_ref/minialign).  Parity is
	 * defined against that build, so the header starts all-zero here as well. */
	BIN_NALN(m, iid) = 0; BIN_PLEN(m, iid) = 0; BIN_LB(m, iid) = 0; BIN_UB(m, iid) = 0;
	uint32_t eid = m->n_res++;
	m->root[2 * eid] = OFS0; m->root[2 * eid + 1] = iid;
	uint32_t rsid = s[4 * lid + 0];
	uint32_t const *p = s + 4 * rsid;
	st->aid = p[1]; st->bid = m->qid;
	/* mm_init_ref (3742-3755) precedes nothing that reads rlen in load_pos?  No: load_pos reads self->rlen of the
	 * PREVIOUS chain (3865 runs before 3873).  Keep that order. */
	load_pos(m, p, &st->rev, st->cp);
	st->tp[0] = st->cp[0]; st->tp[1] = st->cp[1];
	st->iid = iid; st->eid = eid; st->sid = rsid;
	st->prem = plen; st->pacc = 0; st->srem = MM_SREM; st->narrow = 0;
	mmo_ref_t r = mmo_ref(m, st->aid);
	m->rid = st->aid; m->rlen = r.l_seq;
	return(0);
}

static uint64_t load_next(mmo_t *m, search_t *st)											/* 3887-3944 */
{
	if(st->srem == 0) { return(0); }
	st->srem--;
	uint32_t const *s = m->seed;
	uint32_t *n = m->next;
	uint64_t ncnt = m->nnext, ofs = 2 * (uint64_t)m->tglen;
	uint32_t fa = st->cp[0], fb = st->cp[1] - (st->rev ? m->qlen : 0);
	v4_t fv = { { (int32_t)u_of(fa, fb), (int32_t)st->aid, (int32_t)v_of(fa, fb), (int32_t)v_of(fa, fb) } };
	uint64_t plim = ofs - st->pacc;
	if(st->pacc > ofs) { ncnt = 0; }
	for(uint64_t i = 0; i < ncnt; i++) {
		if(n[2 * i] >= plim) { ncnt = i; break; }
		n[2 * i] += st->pacc;
	}
	uint64_t sid = st->sid;
	for(uint64_t rcnt = 2 * (uint64_t)st->srem; sid > 0 && rcnt > 0; sid--) {
		v4_t wv = load_wv(s + 4 * (sid - 1), (int32_t)m->tglen), zv = load_wv(s + 4 * (sid - 1), 128);
		if(!inside_uub(wv, fv)) { break; }
		if(!inside_wv(wv, fv) || inside_wv(zv, fv)) { continue; }
		n[2 * ncnt] = (uint32_t)pdiff(wv, fv); n[2 * ncnt + 1] = (uint32_t)(sid - 1); ncnt++; rcnt--;
	}
	st->sid = (uint32_t)sid;
	m->nnext = ncnt;
	if(ncnt == 0) { st->pacc = 0; st->srem = 0; return(0); }
	radix_sort(n, ncnt, 2);
	m->nnext--;
	uint32_t nsid = n[2 * m->nnext + 1];
	st->pacc = (uint32_t)(ofs - n[2 * m->nnext]);
	load_pos(m, s + 4 * nsid, &st->rev, st->cp);
	return(st->srem);
}

static int test_dup(mmo_t *m, search_t *st, ora_pos_t const *cp)							/* 3952-3981 */
{
	uint64_t k = pos_key((uint64_t)cp->apos | (uint64_t)cp->bpos << 32, (uint64_t)st->aid | (uint64_t)st->bid << 32);
	uint64_t t = kh_put_ptr(m, k, 1);
	uint64_t prev = KH_VAL(m, t);
	int32_t pa = MAX2(1, MIN2((int32_t)cp->apos, (int32_t)m->rlen)), pb = MAX2(1, MIN2((int32_t)cp->bpos, (int32_t)m->qlen));
	st->tp[0] = (uint32_t)pa; st->tp[1] = (uint32_t)pb;
	KH_VAL(m, t) = (uint64_t)st->eid | (uint64_t)UINT32_MAX << 32;
	if(prev == KH_INIT) { return(0); }
	uint32_t eid = (uint32_t)KH_VAL(m, t);								/* reads back what was just stored (3969-3973) */
	if(eid != st->eid && cp->plen < BIN_PLEN(m, m->root[2 * eid + 1])) { st->srem = 0; }
	else { st->narrow = MIN2(st->narrow + 1, 2); }
	return(1);
}

static int record(mmo_t *m, search_t *st, uint64_t ai)										/* 3986-4067 */
{
	ora_aln_t const *a = m->alns[ai];
	ora_seg_t const *sl = &a->seg[a->slen - 1], *s0 = &a->seg[0];
	uint32_t p[4] = {
		m->rlen - (sl->apos + sl->alen), m->qlen - (sl->bpos + sl->blen),
		m->rlen - s0->apos, m->qlen - s0->bpos
	};
	st->cp[0] = p[0]; st->cp[1] = p[1];
	st->prem -= a->plen; st->pacc = a->plen;
	uint64_t id = (uint64_t)st->aid | (uint64_t)st->bid << 32;
	uint64_t hk = pos_key((uint64_t)p[0] | (uint64_t)p[1] << 32, id), tk = pos_key((uint64_t)p[2] | (uint64_t)p[3] << 32, id);
	uint64_t h = kh_put_ptr(m, hk, 1);
	uint64_t t = kh_put_ptr(m, tk, 0);
	int new = (uint32_t)(KH_VAL(m, h) >> 32) == UINT32_MAX;
	uint32_t nid = new ? (uint32_t)bin_push(m, ai) : (uint32_t)(KH_VAL(m, h) >> 32);
	uint32_t iid = st->iid;
	uint32_t ovl = MAX2(BIN_LB(m, iid), p[1]) - MIN2(BIN_UB(m, iid), p[3]) - p[1] + p[3];
	m->root[2 * st->eid] = (uint32_t)((int64_t)m->root[2 * st->eid] - (a->score + (int64_t)(uint32_t)((ovl * 2) * a->identity)));
	BIN_NALN(m, iid) += (uint32_t)new;
	BIN_PLEN(m, iid) += a->plen;
	BIN_LB(m, iid) = MIN2(BIN_LB(m, iid), p[1]);
	BIN_UB(m, iid) = MAX2(BIN_UB(m, iid), p[3]);
	uint64_t eid_hi = (uint64_t)UINT32_MAX << 32;
	if(m->alns[m->bin[nid]]->score > a->score) {
		KH_VAL(m, t) = (uint64_t)st->eid | eid_hi;
	} else {
		if(m->bin[nid] != ai) { m->bin[nid] = ai; }
		KH_VAL(m, h) = KH_VAL(m, t) = (uint64_t)st->eid | (uint64_t)nid << 32;
	}
	st->srem = MM_SREM; st->narrow = 0;
	float ms = (float)a->score * m->p.min_ratio;
	st->min_score = (uint32_t)(((float)st->min_score > ms) ? (float)st->min_score : ms);
	return(new && st->prem > 0 ? 0 : 1);
}

static int finish_root(mmo_t *m, search_t *st)												/* 3794-3811 */
{
	if(BIN_NALN(m, st->iid) == 0 || m->root[2 * st->eid] > (uint32_t)((int32_t)OFS0 - (int32_t)m->p.min_score)) {
		m->nbin = st->iid; m->n_res--; st->crem--;
	} else {
		st->crem = st->crem != 0 ? MM_CREM : 0;
	}
	return(st->crem == 0);
}

static uint64_t mm_extend(mmo_t *m)															/* 4118-4173 */
{
	search_t st; memset(&st, 0, sizeof(st));
	st.crem = MM_CREM; st.min_score = m->p.min_score;
	for(uint64_t k = 0; k < m->nroot; k++) {
		if(load_root(m, &st, (uint32_t)k)) { break; }
		mmo_ref_t ref = mmo_ref(m, st.aid);
		ora_section_t r[2] = { { m->rid << 1, ref.l_seq, ref.seq, 0 }, { (m->rid << 1) + 1, ref.l_seq, ref.seq, 1 } };
		ora_section_t q[3] = { { 0, m->qlen, m->qseq, 0 }, { 1, m->qlen, m->qseq, 1 }, { 0, m->qlen, m->qseq, 0 } };
		for(; st.srem > 0 && st.prem > 0; load_next(m, &st)) {
			for(int i = 0; i < 3; i++) { ora_dp_flush(&m->dp[i]); }
			ora_dp_t *dp = &m->dp[st.narrow];
			int64_t f = extend_core(dp, &r[0], &tail_sec, &q[st.rev], &tail_sec, st.cp[0], st.cp[1]);
			if(ora_fill(dp, f)->max == 0) { continue; }
			ora_pos_t cp = ora_dp_search_max(dp, f);
			if(test_dup(m, &st, &cp) != 0) { continue; }
			dp = &m->dp[st.narrow];
			f = extend_core(dp, &r[1], &tail_sec, &q[1 - st.rev], &tail_sec, ref.l_seq - st.tp[0], m->qlen - st.tp[1]);
			if(ora_fill(dp, f)->max < (int64_t)m->p.min_score) { continue; }
			ora_aln_t *a = ora_dp_trace(dp, f);
			if(a == NULL) { continue; }
			RESERVE(ora_aln_t *, m->alns, m->maln, m->naln + 1);
			m->alns[m->naln] = a;
			if(record(m, &st, m->naln++)) { break; }
		}
		if(finish_root(m, &st)) { break; }
	}
	return(m->n_res);
}

/* ================================================================ post-processing (4175-4396) */
#define SC(_x)	( (int32_t)OFS0 - (int32_t)(_x) )		/* _ofs() */

static uint64_t collect_supp(mmo_t *m, uint32_t n_res)										/* 4214-4263 */
{
	uint32_t *res = m->root;
	#define SWAP_RES(x, y) { uint32_t t0 = res[2*(x)], t1 = res[2*(x)+1]; res[2*(x)] = res[2*(y)]; res[2*(x)+1] = res[2*(y)+1]; res[2*(y)] = t0; res[2*(y)+1] = t1; }
	uint64_t p, q;
	for(p = 1, q = n_res; p < q; p++) {
		uint64_t max = 0;
		for(uint64_t i = p; i < q; i++) {
			uint32_t si = res[2 * i + 1];
			int64_t lb = BIN_LB(m, si), ub = BIN_UB(m, si), span = ub - lb;
			int covered = 0;
			for(uint64_t j = 0; j < p; j++) {
				uint32_t tj = res[2 * j + 1];
				if(BIN_UB(m, tj) < ub) { lb = MAX2(lb, (int64_t)BIN_UB(m, tj)); } else { ub = MIN2(ub, (int64_t)BIN_LB(m, tj)); }
				if(1.2 * (ub - lb) < span) { q--; SWAP_RES(i, q); i--; covered = 1; break; }
			}
			if(covered) { continue; }
			max = MAX2(max, ((uint64_t)(2 * (ub - lb) - span) << 32) | i);
		}
		if(max & 0xffffffff) { SWAP_RES(p, max & 0xffffffff); }
	}
	p = MIN2(p, q);
	#undef SWAP_RES
	return(p);
}

static uint32_t clip_mapq(double x) { uint32_t v = (uint32_t)x; return(MIN2(v, 60u * 16)); }	/* _clip (4177) */

static uint64_t post_map(mmo_t *m)															/* 4270-4325 */
{
	uint32_t *res = m->root;
	uint64_t p = collect_supp(m, m->n_res);
	int64_t usc = 0, lsc = INT64_MAX, tsc = 0;
	for(uint64_t i = p; i < m->n_res; i++) {
		usc = MAX2(usc, (int64_t)SC(res[2 * i])); lsc = MIN2(lsc, (int64_t)SC(res[2 * i])); tsc += SC(res[2 * i]);
	}
	lsc = (lsc == INT32_MAX) ? 0 : lsc;
	double tpc = 1.0, x = m->xcoef, mx = m->mcoef + m->xcoef;
	for(uint64_t i = 0; i < p; i++) {
		uint32_t score = (uint32_t)SC(res[2 * i]), iid = res[2 * i + 1];
		double pid = 0.0; uint64_t len = 0;
		for(uint64_t j = 0; j < BIN_NALN(m, iid); j++) {
			len += BIN_ALN(m, iid, j)->plen;
			pid += (double)BIN_ALN(m, iid, j)->plen * BIN_ALN(m, iid, j)->identity;
		}
		pid /= (double)len;
		double ec = 2.0 / (pid * mx - x);
		double ulen = ec * MAX2((int64_t)score - usc, 0), pe = 1.0 / (ulen * ulen + 1);
		BIN_PLEN(m, iid) = clip_mapq(-10.0 * 16 * log10(pe));
		tpc *= 1.0 - pe;
	}
	double tpe = MIN2(1.0 - tpc, 1.0);
	for(uint64_t i = p; i < m->n_res; i++) {
		uint32_t iid = res[2 * i + 1];
		BIN_PLEN(m, iid) = clip_mapq(-10.0 * 16 * log10(1.0 - tpe * (double)(res[2 * i] - lsc + 1) / (double)tsc));
	}
	return(p);
}

/* mm_align_seq (4427-4474) + mm_pack_reg (4364-4396) into the flat layout */
uint64_t mmo_align(mmo_t *m, uint8_t const *seq, uint32_t len, uint32_t qid, uint32_t *out, uint64_t cap)
{
	(void)qid;
	if(len < m->k || len * m->mcoef < (double)m->p.min_score) { return(0); }
	/* mm_tbuf_clear + mm_init_query */
	m->nresc = 0; m->presc = 0; m->nseed_arr = 0; m->n_seed = 0; m->nroot = 0; m->nnext = 0; m->n_res = 0; m->nbin = 0;
	kh_reset(m);
	for(uint64_t i = 0; i < m->naln; i++) { ora_aln_free(m->alns[i]); }
	m->naln = 0;
	m->qid = 0; m->qlen = len; m->qseq = seq;
	for(uint64_t i = 0; i < m->n_occ; i++) {
		if(mm_seed(m, i) == 0) { continue; }
		if(mm_chain(m) == 0) { continue; }
		if(mm_extend(m) > 0) { break; }
	}
	if(m->n_res == 0) { return(0); }
	radix_sort(m->root, m->n_res, 2);
	/* mm_prune_regs (4185-4207) */
	uint32_t *res = m->root;
	uint64_t q = m->n_res;
	uint32_t min = (uint32_t)SC((uint32_t)(SC(res[0]) * m->p.min_ratio));
	while(res[2 * --q] > min) {}
	m->n_res = (uint32_t)(q + 1);
	uint32_t n_all = m->n_res, n_uniq = (uint32_t)post_map(m);
	/* pack */
	uint64_t n = 2; uint32_t cnt = 0, uniq = 0;
	for(uint64_t i = 0; i < n_all; i++) {
		uint32_t iid = res[2 * i + 1];
		for(uint64_t j = 0; j < BIN_NALN(m, iid); j++) {
			n += dump_aln(BIN_ALN(m, iid, j), (uint32_t)i, BIN_PLEN(m, iid), n < cap ? out + n : out, n < cap ? cap - n : 0);
			cnt++;
		}
		if(i == (uint64_t)n_uniq - 1) { uniq = cnt; }
	}
	if(cap >= 2) { out[0] = cnt; out[1] = uniq; }
	return(n);
}

/* debugging twin of refh_extend_dump (ref_harness.c) */
uint64_t mmo_extend_dump(mmo_t *m, uint8_t const *seq, uint32_t len, uint32_t *out, uint64_t cap)
{
	m->nresc = 0; m->presc = 0; m->nseed_arr = 0; m->n_seed = 0; m->nroot = 0; m->nnext = 0; m->n_res = 0; m->nbin = 0;
	kh_reset(m);
	for(uint64_t i = 0; i < m->naln; i++) { ora_aln_free(m->alns[i]); }
	m->naln = 0; m->qid = 0; m->qlen = len; m->qseq = seq;
	for(uint64_t i = 0; i < m->n_occ; i++) {
		if(mm_seed(m, i) == 0) { continue; }
		if(mm_chain(m) == 0) { continue; }
		if(mm_extend(m) > 0) { break; }
	}
	uint64_t n = 1;
	out[0] = m->n_res;
	for(uint32_t i = 0; i < m->n_res && n + 6 <= cap; i++) {
		uint32_t iid = m->root[2 * i + 1];
		out[n++] = m->root[2 * i]; out[n++] = iid; out[n++] = BIN_NALN(m, iid); out[n++] = BIN_PLEN(m, iid); out[n++] = BIN_LB(m, iid); out[n++] = BIN_UB(m, iid);
	}
	return(n);
}
