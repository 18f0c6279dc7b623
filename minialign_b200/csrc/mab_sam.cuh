/*
 * mab_sam.cuh -- the SAM printer on the device (mm_print_sam_mapped / _unmapped and helpers, minialign.c:5095-5426; path parser
 * gaba_parse.h:107-263).  Same text, byte for byte, as minialign_b200/csrc/host/mab_sam.cpp produces from the flat records (the
 * two are compared against each other and against the reference's SAM in the tests); here the input is what the device already
 * holds: the result pool with the output plan of mab_post.cuh, the read block, the text chunk (names, qualities) and the index
 * image (reference names and bases for RNAME / MD).
 *
 * One warp per read, two passes over the same code: k_sam<false> only counts bytes (per-read lengths -> exclusive scan ->
 * offsets), k_sam<true> writes.  The formatting itself (numbers, CIGAR run lengths out of the path bits, MD walk) is a serial
 * byte stream per read, executed uniformly by the warp with lane 0 storing; the 20 kb of bases (and qualities) of a line are
 * decoded by all lanes.
 */
#pragma once
#include "mab_post.cuh"

namespace mab {

/* tag bits: include/minialign_b200.h (MAB_TAG_*, MAB_OMIT_REP) */
#define MAB_TAG_NH_ 0x04u
#define MAB_TAG_IH_ 0x08u
#define MAB_TAG_AS_ 0x10u
#define MAB_TAG_XS_ 0x20u
#define MAB_TAG_NM_ 0x40u
#define MAB_TAG_SA_ 0x80u
#define MAB_TAG_MD_ 0x100u
#define MAB_OMIT_REP_ 0x40000000u

template <bool W>
struct SamW {
	uint8_t *p; uint64_t n; int lane;
	__device__ __forceinline__ void c(uint32_t ch) { if(W && lane == 0) { p[n] = (uint8_t)ch; } n++; }
	/* up to 8 characters packed little-endian */
	__device__ __forceinline__ void lit(unsigned long long s, int cnt) { for(int i = 0; i < cnt; i++) { c((uint32_t)(s >> (8 * i)) & 0xff); } }
	template <int N> __device__ __forceinline__ void str(const char (&s)[N]) { _Pragma("unroll") for(int i = 0; i < N - 1; i++) { c((uint32_t)(uint8_t)s[i]); } }
	__device__ __forceinline__ void bytes(const uint8_t *s, uint32_t len) { if(W) { for(uint32_t t = lane; t < len; t += 32) { p[n + t] = s[t]; } } n += len; }
	/* read name: a tab inside it was turned into a space by the reader (_escape, minialign.c:2012, 2032) */
	__device__ __forceinline__ void qname(const uint8_t *s, uint32_t len) { if(W) { for(uint32_t t = lane; t < len; t += 32) { uint8_t ch = s[t]; p[n + t] = ch == '\t' ? ' ' : ch; } } n += len; }
	__device__ __forceinline__ void small(uint32_t v)						/* v < 10^8 */
	{
		unsigned long long buf = 0; int nd = 0;
		do { uint32_t q = v / 10; buf = (buf << 8) | ('0' + (v - 10 * q)); v = q; nd++; } while(v);
		lit(buf, nd);
	}
	__device__ __forceinline__ void pad8(uint32_t v)							/* exactly eight digits */
	{
		unsigned long long buf = 0;
		for(int i = 0; i < 8; i++) { uint32_t q = v / 10; buf = (buf << 8) | ('0' + (v - 10 * q)); v = q; }
		lit(buf, 8);
	}
	__device__ __forceinline__ void num(unsigned long long v)
	{
		if(v < 100000000ull) { small((uint32_t)v); return; }
		unsigned long long hi = v / 100000000ull;
		if(hi < 100000000ull) { small((uint32_t)hi); } else { small((uint32_t)(hi / 100000000ull)); pad8((uint32_t)(hi % 100000000ull)); }
		pad8((uint32_t)(v - hi * 100000000ull));
	}
	/* decoded bases: forward (decaf) or reversed + complemented (decar), minialign.c:229-230 */
	__device__ __forceinline__ void seq(const uint8_t *b, uint32_t len, bool rc)
	{
		if(W) {
			const unsigned long long lf = 0x4e54474341ull, lr = 0x4e41434754ull;
			for(uint32_t t = lane; t < len; t += 32) {
				uint32_t code = rc ? b[len - 1 - t] : b[t];
				p[n + t] = code < 5 ? (uint8_t)((rc ? lr : lf) >> (8 * code)) : 0;
			}
		}
		n += len;
	}
	__device__ __forceinline__ void rev(const uint8_t *s, uint32_t len) { if(W) { for(uint32_t t = lane; t < len; t += 32) { p[n + t] = s[len - 1 - t]; } } n += len; }
};

/* 64 path bits from (possibly negative) bit position pos.  Below bit 0 the reference reads the two words in front of path[] in
 * gaba_alignment_s: `plen` and `padding` = 0x40000000 (gaba.h:217, gaba.c:3277-3278); above the sentinel the words are zero */
struct PathBits {
	const uint32_t *p; uint32_t npw, plen;
	__device__ __forceinline__ uint32_t word(long long i) const { return i >= 0 ? ((unsigned long long)i < npw ? p[i] : 0u) : (i == -1 ? 0x40000000u : (i == -2 ? plen : 0u)); }
	__device__ __forceinline__ unsigned long long at(long long pos) const
	{
		long long w0 = pos >> 5; uint32_t sh = (uint32_t)(pos & 31);
		unsigned long long r = ((unsigned long long)word(w0) | (unsigned long long)word(w0 + 1) << 32) >> sh;
		if(sh) { r |= (unsigned long long)word(w0 + 2) << (64 - sh); }
		return r;
	}
};
__device__ __forceinline__ unsigned long long sam_lz(unsigned long long x) { return (unsigned long long)__clzll((long long)x); }

/* the reverse path parser (_parser_loop_rv, gaba_parse.h:162-184): del(c), ins(c), diag(c) in the reference's order */
template <class D, class I, class M>
__device__ __forceinline__ void parse_rv(const PathBits &pb, unsigned long long offset, unsigned long long len, D del, I ins, M diag)
{
	long long ofs = (long long)offset - 64; unsigned long long idx = len;
	while((long long)idx > 0) {
		unsigned long long m = sam_lz(pb.at(ofs + (long long)idx));
		unsigned long long c = m - (m > 0); c = idx < c ? idx : c; idx -= c; del(c);
		m = sam_lz(~pb.at(ofs + (long long)idx));
		c = idx < m ? idx : m; idx -= c; ins(c);
		do {
			m = sam_lz(pb.at(ofs + (long long)idx) ^ 0x5555555555555555ull);
			c = (idx < m ? idx : m) & ~1ull; idx -= c; diag(c >> 1);
		} while(c == 64);
	}
}

/* decimal digits of v at p (per-lane helper of the warp-parallel CIGAR); returns their number */
__device__ __forceinline__ uint32_t sam_ndigits(uint32_t v)
{
	return v < 10u ? 1u : v < 100u ? 2u : v < 1000u ? 3u : v < 10000u ? 4u : v < 100000u ? 5u : v < 1000000u ? 6u : v < 10000000u ? 7u : v < 100000000u ? 8u : v < 1000000000u ? 9u : 10u;
}
__device__ __forceinline__ void sam_put_dec(uint8_t *p, uint32_t v, uint32_t nd)
{
	for(uint32_t i = nd; i > 0; i--) { uint32_t q = v / 10; p[i - 1] = (uint8_t)('0' + (v - 10 * q)); v = q; }
}

/* CIGAR of path bits [ppos, ppos + plen), written from the top bit down (the reference parses alignments in reverse:
 * _parser_loop_rv, gaba_parse.h:162-184, driven by gaba_dump_cigar_reverse).
 *
 * What the serial parser (parse_rv above, used for the MD walk) produces is a run-length encoding of a purely local
 * classification of the path bits, scanning downwards: a 0 directly above a 1 and that 1 form a diagonal step (M); every other 1 is
 * an insertion (I), every other 0 a deletion (D).  (The parser counts a run of zeros and keeps the last one for the pair it forms
 * with the following 1 -- `m - (m > 0)` --, counts a run of ones as insertions, then eats "01" pairs; below bit 0 it reads the
 * words in front of the path, a 0 then a 1 (PathBits), so zeros at the bottom are deletions.)  Consecutive steps of one kind make
 * one CIGAR operation.  That makes the CIGAR data-parallel: every lane classifies 32 positions with a few bit operations, marks the
 * positions that start a new run, and formats the runs that end at its own starts; run lengths come from the previous start
 * (warp scan, carried from chunk to chunk), output offsets from a warp scan of the text lengths.  1024 positions per iteration. */
template <bool W>
__device__ __forceinline__ void sam_cigar_rv(SamW<W> &o, const PathBits &pb, unsigned long long ppos, unsigned long long plen)
{
	const int lane = o.lane;
	const long long len = (long long)plen;
	if(len <= 0) { return; }
	long long prev_q = -1; uint32_t prev_cls = 0;				/* the latest run start so far (downward coordinate q = len - 1 - position) and its class */
	uint32_t cls_above = 3;										/* class of the position just above the chunk (3 = none: the top position always starts a run) */
	for(long long hi = len - 1; hi >= 0; hi -= 1024) {
		const long long base = hi - 32ll * lane - 31;			/* position of bit 0 of this lane's 32 positions; bit 31 is the topmost */
		unsigned long long x = base > -32 ? pb.at((long long)ppos + base - 1) : 0ull;		/* bit 0: lower neighbour of position base; bit 33: upper neighbour of base + 31 */
		uint32_t Wd = (uint32_t)(x >> 1), Lw = (uint32_t)x, Uw = (uint32_t)(x >> 2);
		if(base + 31 == len - 1) { Uw |= 0x80000000u; }			/* nothing above the top of the segment: a 1 there is an insertion */
		const uint32_t valid = base >= 0 ? 0xffffffffu : (base <= -32 ? 0u : (0xffffffffu << (uint32_t)(-base)));
		const uint32_t Mm = ((Wd & ~Uw) | (~Wd & Lw)) & valid, Im = (Wd & Uw) & valid, Dm = (~Wd & ~Lw) & valid;
		const uint32_t cls0 = (Mm & 1u) ? 0u : (Im & 1u) ? 1u : (Dm & 1u) ? 2u : 3u;
		uint32_t up = __shfl_up_sync(0xffffffffu, cls0, 1);
		if(lane == 0) { up = cls_above; }
		cls_above = __shfl_sync(0xffffffffu, cls0, 31);
		const uint32_t start = valid & ~((Mm & ((Mm >> 1) | ((uint32_t)(up == 0) << 31))) | (Im & ((Im >> 1) | ((uint32_t)(up == 1) << 31))) | (Dm & ((Dm >> 1) | ((uint32_t)(up == 2) << 31))));
		/* the latest start above this lane's positions: exclusive maximum over the lanes (q grows downwards), else the carry */
		unsigned long long mine = 0;
		if(start) {
			int k = __ffs((int)start) - 1;
			uint32_t c = (Mm >> k) & 1u ? 0u : (Im >> k) & 1u ? 1u : 2u;
			mine = ((unsigned long long)(len - 1 - base - k) + 1) << 2 | c;
		}
		unsigned long long inc = mine;
		for(int d = 1; d < 32; d <<= 1) { unsigned long long y = __shfl_up_sync(0xffffffffu, inc, d); if(lane >= d && y > inc) { inc = y; } }
		unsigned long long exc = __shfl_up_sync(0xffffffffu, inc, 1);
		if(lane == 0) { exc = 0; }
		const unsigned long long last = __shfl_sync(0xffffffffu, inc, 31);
		long long pq0 = exc ? (long long)(exc >> 2) - 1 : prev_q; uint32_t pc0 = exc ? (uint32_t)(exc & 3) : prev_cls;
		/* text bytes of the runs that end at this lane's starts */
		uint32_t bytes = 0;
		{
			long long pq = pq0; uint32_t pc = pc0;
			for(uint32_t m = start; m; ) {
				int k = 31 - __clz((int)m); m &= ~(1u << k);
				long long q = len - 1 - base - k;
				if(pq >= 0) { uint32_t run = (uint32_t)(q - pq); bytes += sam_ndigits(pc == 0 ? run >> 1 : run) + 1; }
				pq = q; pc = (Mm >> k) & 1u ? 0u : (Im >> k) & 1u ? 1u : 2u;
			}
		}
		uint32_t tot, ofs = warp_excl_scan(bytes, lane, &tot);
		if(W) {
			uint8_t *p = o.p + o.n + ofs;
			long long pq = pq0; uint32_t pc = pc0;
			for(uint32_t m = start; m; ) {
				int k = 31 - __clz((int)m); m &= ~(1u << k);
				long long q = len - 1 - base - k;
				if(pq >= 0) {
					uint32_t run = (uint32_t)(q - pq), v = pc == 0 ? run >> 1 : run, nd = sam_ndigits(v);
					sam_put_dec(p, v, nd); p[nd] = pc == 0 ? 'M' : pc == 1 ? 'I' : 'D'; p += nd + 1;
				}
				pq = q; pc = (Mm >> k) & 1u ? 0u : (Im >> k) & 1u ? 1u : 2u;
			}
		}
		o.n += tot;
		if(last) { prev_q = (long long)(last >> 2) - 1; prev_cls = (uint32_t)(last & 3); }
	}
	if(prev_q >= 0) {											/* the last run reaches the bottom of the segment */
		uint32_t run = (uint32_t)(len - prev_q);
		o.num(prev_cls == 0 ? run >> 1 : run); o.c(prev_cls == 0 ? 'M' : prev_cls == 1 ? 'I' : 'D');
	}
}

struct SamRead { const uint8_t *name; uint32_t l_name; const uint8_t *seq; uint32_t l_seq; const uint8_t *qual; };
struct SamRef { const uint8_t *name; uint32_t l_name; const uint8_t *seq; uint32_t l_seq; };
__device__ __forceinline__ SamRef sam_ref(const DevParams &P, uint32_t rid)
{
	const uint8_t *s = P.idx + P.seq_ofs + 24ull * rid;
	SamRef r; r.seq = P.idx + ldg64(s); r.name = P.idx + ldg64(s + 8); r.l_seq = ldg32(s + 16); r.l_name = ldg16(s + 20);
	return r;
}

/* mm_print_sam_mapped_core (5146-5197) */
template <bool W>
__device__ __forceinline__ void sam_core(SamW<W> &o, const DevParams &P, const SamRead &q, const uint32_t *s, const PathBits &pb, uint32_t flag, uint32_t mapq)
{
	uint32_t aid = s[0], bid = s[1], apos = s[2], bpos = s[3], alen = s[4], blen = s[5];
	unsigned long long ppos = (unsigned long long)s[6] | (unsigned long long)s[7] << 32;
	SamRef r = sam_ref(P, aid >> 1);
	uint32_t rs = r.l_seq - apos - alen;
	uint32_t hl = q.l_seq - bpos - blen, tl = bpos;
	uint32_t qs = (flag & 0x900) ? hl : 0, qe = q.l_seq - ((flag & 0x900) ? tl : 0);
	o.qname(q.name, q.l_name); o.c('\t');
	o.num(flag | ((~bid & 1) << 4)); o.c('\t');
	o.bytes(r.name, r.l_name); o.c('\t');
	o.num((unsigned long long)rs + 1); o.c('\t');
	o.num(mapq >> 4); o.c('\t');
	if(hl) { o.num(hl); o.c((flag & 0x900) ? 'H' : 'S'); }
	sam_cigar_rv(o, pb, ppos, (unsigned long long)alen + blen);
	if(tl) { o.num(tl); o.c((flag & 0x900) ? 'H' : 'S'); }
	o.str("\t*\t0\t0\t");
	if(bid & 1) { o.seq(q.seq + qs, qe - qs, false); } else { o.seq(q.seq + (q.l_seq - qe), qe - qs, true); }
	o.c('\t');
	if(q.qual != nullptr && q.l_seq != 0) {
		if(bid & 1) { o.bytes(q.qual + qs, qe - qs); } else { o.rev(q.qual + (q.l_seq - qe), qe - qs); }
	} else { o.c('*'); }
}

/* mm_print_sam_md (5239-5298) */
template <bool W>
__device__ __forceinline__ void sam_md(SamW<W> &o, const DevParams &P, const SamRead &q, const uint32_t *s, const PathBits &pb)
{
	const unsigned long long lf = 0x4e54474341ull;
	uint32_t aid = s[0], bid = s[1], apos = s[2], bpos = s[3], alen = s[4], blen = s[5];
	unsigned long long ppos = (unsigned long long)s[6] | (unsigned long long)s[7] << 32;
	uint32_t rev = ~bid & 1;
	SamRef r = sam_ref(P, aid >> 1);
	const uint8_t *rp = r.seq + (r.l_seq - apos - alen), *rb = rp;
	const uint8_t *qp = rev ? q.seq + (q.l_seq - bpos) : q.seq + (q.l_seq - bpos - blen);
	o.str("\tMD:Z:");
	auto dec = [&](uint32_t code) -> uint32_t { code &= 15; return code < 5 ? (uint32_t)(lf >> (8 * code)) & 0xff : 0; };
	auto del = [&](unsigned long long c) {
		if(c > 0) { o.num((unsigned long long)(rp - rb)); o.c('^'); rb = rp + c; for(unsigned long long i = 0; i < c; i++) { o.c(dec(rp[i])); } rp += c; }
	};
	/* _match_ff / _match_fr: 16 columns at a time, report the first mismatch of each chunk and restart behind it */
	if(rev == 0) {
		parse_rv(pb, ppos, (unsigned long long)alen + blen, del,
			[&](unsigned long long c) { qp += c; },
			[&](unsigned long long c) {
				rp += c; qp += c;
				for(unsigned long long i = c, l = 0; i > 0; i -= l) {
					l = i < 16 ? i : 16;
					unsigned long long mc = 0;
					while(mc < 16 && rp[(long long)mc - (long long)i] == qp[(long long)mc - (long long)i]) { mc++; }
					if(mc < l) { o.num((unsigned long long)(rp - i + mc - rb)); o.c(dec(rp[(long long)mc - (long long)i])); rb = rp - i + mc + 1; l = mc + 1; }
				}
			});
	} else {
		parse_rv(pb, ppos, (unsigned long long)alen + blen, del,
			[&](unsigned long long c) { qp -= c; },
			[&](unsigned long long c) {
				rp += c; qp -= c;
				for(unsigned long long i = c, l = 0; i > 0; i -= l) {
					l = i < 16 ? i : 16;
					unsigned long long mc = 0;
					while(mc < l && rp[(long long)mc - (long long)i] == (uint8_t)(qp[(long long)i - 1 - (long long)mc] ^ 0x03)) { mc++; }
					if(mc < l) { o.num((unsigned long long)(rp - i + mc - rb)); o.c(dec(rp[(long long)mc - (long long)i])); rb = rp - i + mc + 1; l = mc + 1; }
				}
			});
	}
	o.num((unsigned long long)(rp - rb));
}

struct SamAln { const uint32_t *a, *seg, *path; uint32_t slen, plen, npw, mapq; };
__device__ __forceinline__ SamAln sam_aln(const uint32_t *pool, const uint32_t *item)
{
	SamAln x; x.a = pool + ((unsigned long long)item[0] | (unsigned long long)item[1] << 32);
	x.slen = x.a[7]; x.plen = x.a[8]; x.npw = x.a[9];
	uint32_t sn = x.a[10];
	x.seg = x.a + MAB_ALN_HDR + 8ull * (sn - x.slen); x.path = x.a + MAB_ALN_HDR + 8ull * sn; x.mapq = item[3];
	return x;
}
__device__ __forceinline__ uint32_t sam_edit_dist(const uint32_t *a)				/* 5333-5337 */
{
	unsigned long long ib = (unsigned long long)a[2] | (unsigned long long)a[3] << 32;
	double identity; memcpy(&identity, &ib, 8);
	return (uint32_t)__double2ll_rz(__dmul_rn(__uint2double_rn(a[6]), __dsub_rn(1.0, identity))) + a[4] + a[5];
}

/* all SAM lines of one read (mm_print_sam_unmapped 5126-5141, mm_print_sam_mapped 5389-5426); plan = nullptr: unmapped */
template <bool W>
__device__ __forceinline__ void sam_read(SamW<W> &o, const DevParams &P, const uint32_t *pool, const uint32_t *plan, const SamRead &q, uint32_t tags)
{
	if(plan == nullptr) {
		o.qname(q.name, q.l_name); o.str("\t4\t*\t0\t0\t*\t*\t0\t0\t");
		o.seq(q.seq, q.l_seq, false); o.c('\t');
		if(q.qual != nullptr && q.l_seq != 0) { o.bytes(q.qual, q.l_seq); } else { o.c('*'); }
		o.c('\n');
		return;
	}
	const uint32_t n_all = plan[0], n_uniq = plan[1];
	const uint32_t *items = plan + 2;
	/* the reference keeps its MM_OMIT_REP flag (0x08, 2489) in the same word as the tag bits, where bit 3 is the IH tag (2533):
	 * asking for IH also drops the secondary records, and -R (omit) also prints IH (5323, 5401, 6110) */
	unsigned long long n = (tags & (MAB_TAG_IH_ | MAB_OMIT_REP_)) ? n_uniq : n_all;
	uint32_t flag = 0;
	for(unsigned long long i = 0; i < n; i++) {
		if(i >= n_uniq) { flag = 0x100; }
		SamAln a = sam_aln(pool, items + 4 * i);
		PathBits pb = { a.path, a.npw, a.plen };
		for(unsigned long long j = a.slen; j > 0; j--) {
			const uint32_t *s = a.seg + 8 * (j - 1);
			sam_core(o, P, q, s, pb, flag, a.mapq);
			/* general tags (5303-5340) */
			if(tags & MAB_TAG_NH_) { o.str("\tNH:i:"); o.num(n_all); }
			if(tags & (MAB_TAG_IH_ | MAB_OMIT_REP_)) { o.str("\tIH:i:"); o.num(i); }
			if(tags & MAB_TAG_AS_) { o.str("\tAS:i:"); o.num(a.a[0]); }
			if(tags & MAB_TAG_NM_) { o.str("\tNM:i:"); o.num(sam_edit_dist(a.a)); }
			if(tags & MAB_TAG_MD_) { sam_md(o, P, q, s, pb); }
			bool skip = false;
			if(i == 0 && j == a.slen) {											/* primary-only tags (5346-5384) */
				flag = 0x800;
				if(tags & MAB_TAG_XS_) { o.str("\tXS:i:"); o.num(n_all > 1 ? sam_aln(pool, items + 4).a[0] : 0u); }
				if((tags & MAB_TAG_SA_) && (n_uniq > 1 || a.slen > 1)) {
					o.str("\tSA:Z:");
					SamRef r0 = sam_ref(P, 0);
					for(unsigned long long x = 0; x < n_uniq; x++) {
						SamAln b = sam_aln(pool, items + 4 * x);
						PathBits pbb = { b.path, b.npw, b.plen };
						for(unsigned long long y = b.slen; y > 0; y--) {
							if(x == 0 && y == b.slen) { continue; }
							const uint32_t *t = b.seg + 8 * (y - 1);					/* mm_print_sam_supp (5203-5233) */
							SamRef rr = sam_ref(P, t[0] >> 1);
							uint32_t rs = rr.l_seq - t[2] - t[4], hl = q.l_seq - t[3] - t[5], tl = t[3];
							o.bytes(r0.name, r0.l_name); o.c(',');						/* the reference prints r->name, i.e. the FIRST sequence's (5217) */
							o.num((unsigned long long)rs + 1); o.c(',');
							o.c((t[1] & 1) ? '+' : '-'); o.c(',');
							if(hl) { o.num(hl); o.c('H'); }
							sam_cigar_rv(o, pbb, (unsigned long long)t[6] | (unsigned long long)t[7] << 32, (unsigned long long)t[4] + t[5]);
							if(tl) { o.num(tl); o.c('H'); }
							o.c(','); o.num(b.mapq); o.c(','); o.num(sam_edit_dist(b.a)); o.c(';');
						}
					}
					skip = true;
				}
			}
			o.c('\n');
			if(skip) { i = n; break; }
		}
		flag = 0x800;
	}
}

/* one warp per read; WRITE = false: recs[i].sam_len = bytes of the read's lines; WRITE = true: the lines at out + recs[i].sam_ofs */
template <bool WRITE>
__global__ void k_sam(DevParams P, const uint32_t *pool, const ReadRec *reads, TextRec *recs, uint32_t n_reads, const uint8_t *text, const uint8_t *base,
	uint32_t tags, uint32_t keep_qual, uint8_t *out)
{
	int lane = threadIdx.x & 31;
	uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
	for(uint32_t i = gw; i < n_reads; i += nw) {
		TextRec *tr = &recs[i];
		const ReadRec *r = &reads[i];
		SamW<WRITE> o; o.p = WRITE ? out + tr->sam_ofs : nullptr; o.n = 0; o.lane = lane;
		if(!(tr->flags & MAB_TR_DROPPED)) {
			SamRead q; q.name = text + tr->name_ofs; q.l_name = tr->name_len; q.seq = base + r->seq_ofs; q.l_seq = r->len;
			q.qual = (keep_qual && (tr->flags & MAB_TR_HASQUAL)) ? text + tr->qual_ofs : nullptr;
			const uint32_t *plan = nullptr;
			if(r->result_words != 0 && r->err == 0) {
				const uint32_t *rec = pool + r->result_ofs;
				plan = post_scratch(const_cast<uint32_t *>(rec), r->result_words) + 4ull * rec[0];
			}
			sam_read(o, P, pool, plan, q, tags);
		}
		if(!WRITE && lane == 0) { tr->sam_len = o.n; }
		__syncwarp();
	}
}

/* exclusive scan of the per-read lengths into the output offsets (single CTA); tc->sam_total = bytes */
__global__ void k_sam_offsets(TextRec *recs, uint32_t n, TextCounters *tc)
{
	__shared__ uint64_t sm[34];
	uint64_t carry = 0;
	for(uint32_t i0 = 0; i0 < n; i0 += blockDim.x) {
		uint32_t i = i0 + threadIdx.x;
		uint64_t x = i < n ? recs[i].sam_len : 0, tot;
		uint64_t e = block_excl_sum(x, sm, &tot);
		if(i < n) { recs[i].sam_ofs = carry + e; }
		carry += tot;
	}
	if(threadIdx.x == 0) { tc->sam_total = carry; }
}

}  // namespace mab
