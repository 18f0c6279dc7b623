/* mab_sam.cpp -- see mab_sam.h.  Citations are into /root/reference/minialign.c unless noted. */
#include "mab_sam.h"
#include <cstdlib>
#include <cstring>

namespace {
const char DEC_F[16] = { 'A', 'C', 'G', 'T', 'N', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };		/* decaf (229) */
const char DEC_R[16] = { 'T', 'G', 'C', 'A', 'N', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };		/* decar (230) */

inline void put_num(std::string &o, uint64_t v) { char b[24]; int n = 0; do { b[n++] = (char)('0' + v % 10); v /= 10; } while(v); while(n) { o.push_back(b[--n]); } }

/* view of one alignment inside the flat result stream */
struct Aln {
	int64_t score; double identity; uint32_t agcnt, bgcnt, dcnt, slen, plen, npw, rank, mapq;
	const uint32_t *seg, *path;
};
const uint32_t *parse_aln(const uint32_t *w, Aln &a)
{
	a.score = (int64_t)((uint64_t)w[0] | (uint64_t)w[1] << 32);
	uint64_t ib = (uint64_t)w[2] | (uint64_t)w[3] << 32; memcpy(&a.identity, &ib, 8);
	a.agcnt = w[4]; a.bgcnt = w[5]; a.dcnt = w[6]; a.slen = w[7]; a.plen = w[8]; a.npw = w[9]; a.rank = w[10]; a.mapq = w[11];
	a.seg = w + 16; a.path = a.seg + 8ull * a.slen;
	return a.path + a.npw;
}

/* 64 path bits from (possibly negative) bit position pos.  Below bit 0 the reference reads the two words in front of
 * path[] in gaba_alignment_s: `plen` and `padding` = 0x40000000 (gaba.h:217, gaba.c:3277-3278); above the sentinel the
 * words are zero (gaba.c:3287-3289). */
struct PathBits {
	const uint32_t *p; uint32_t npw, plen;
	uint32_t word(int64_t i) const { return i >= 0 ? ((uint64_t)i < npw ? p[i] : 0u) : (i == -1 ? 0x40000000u : (i == -2 ? plen : 0u)); }
	uint64_t at(int64_t pos) const
	{
		int64_t w0 = pos >> 5; uint32_t sh = (uint32_t)(pos & 31);
		uint64_t r = ((uint64_t)word(w0) | (uint64_t)word(w0 + 1) << 32) >> sh;
		if(sh) { r |= (uint64_t)word(w0 + 2) << (64 - sh); }
		return r;
	}
};
inline uint64_t lz(uint64_t x) { return x ? (uint64_t)__builtin_clzll(x) : 64; }

/* the reverse path parser (_parser_loop_rv, gaba_parse.h:162-184): calls del(c), ins(c), diag(c) in the reference's order */
template <class D, class I, class M>
void parse_rv(const PathBits &pb, uint64_t offset, uint64_t len, D del, I ins, M diag)
{
	int64_t ofs = (int64_t)offset - 64; uint64_t idx = len;
	while((int64_t)idx > 0) {
		uint64_t m = lz(pb.at(ofs + (int64_t)idx));
		uint64_t c = m - (m > 0); c = idx < c ? idx : c; idx -= c; del(c);
		m = lz(~pb.at(ofs + (int64_t)idx));
		c = idx < m ? idx : m; idx -= c; ins(c);
		do {
			m = lz(pb.at(ofs + (int64_t)idx) ^ 0x5555555555555555ull);
			c = (idx < m ? idx : m) & ~1ull; idx -= c; diag(c >> 1);
		} while(c == 64);
	}
}

void put_cigar_rv(std::string &o, const PathBits &pb, uint64_t ppos, uint64_t plen)
{
	uint64_t mrun = 0;
	auto flush_m = [&]() { if(mrun) { put_num(o, mrun); o.push_back('M'); mrun = 0; } };
	parse_rv(pb, ppos, plen,
		[&](uint64_t c) { if(c) { flush_m(); put_num(o, c); o.push_back('D'); } },
		[&](uint64_t c) { if(c) { flush_m(); put_num(o, c); o.push_back('I'); } },
		[&](uint64_t c) { mrun += c; });
	/* the reference emits one M per diagonal run (_diag_end, gaba_parse.h:181): runs are separated by a del/ins call with c > 0 or by the loop restart */
	flush_m();
}

/* mm_print_sam_mapped_core (5146-5197) */
void put_core(std::string &o, const MabSamRef *r, const MabSamRead *q, const uint32_t *s, const PathBits &pb, uint32_t flag, uint32_t mapq)
{
	uint32_t aid = s[0], bid = s[1], apos = s[2], bpos = s[3], alen = s[4], blen = s[5];
	uint64_t ppos = (uint64_t)s[6] | (uint64_t)s[7] << 32;
	uint32_t rid = aid >> 1;
	uint32_t rs = r[rid].l_seq - apos - alen;
	uint32_t hl = q->l_seq - bpos - blen, tl = bpos;
	uint32_t qs = (flag & 0x900) ? hl : 0, qe = q->l_seq - ((flag & 0x900) ? tl : 0);
	o.append(q->name, q->l_name); o.push_back('\t');
	put_num(o, flag | ((~bid & 1) << 4)); o.push_back('\t');
	o.append(r[rid].name, r[rid].l_name); o.push_back('\t');
	put_num(o, (uint64_t)rs + 1); o.push_back('\t');
	put_num(o, mapq >> 4); o.push_back('\t');
	if(hl) { put_num(o, hl); o.push_back((flag & 0x900) ? 'H' : 'S'); }
	put_cigar_rv(o, pb, ppos, (uint64_t)alen + blen);
	if(tl) { put_num(o, tl); o.push_back((flag & 0x900) ? 'H' : 'S'); }
	o.append("\t*\t0\t0\t");
	{	/* sequence, decoded in bulk (forward: decaf, reverse strand: reversed + complemented, decar) */
		size_t o0 = o.size(), n = qe - qs;
		o.resize(o0 + n);
		char *d = &o[o0];
		if(bid & 1) { const uint8_t *b = q->seq + qs; for(size_t i = 0; i < n; i++) { d[i] = DEC_F[b[i] & 15]; } }
		else { const uint8_t *b = q->seq + (q->l_seq - qe); for(size_t i = 0; i < n; i++) { d[i] = DEC_R[b[n - 1 - i] & 15]; } }
	}
	o.push_back('\t');
	if(q->qual && q->qual[0] != '\0') {
		if(bid & 1) { o.append(q->qual + qs, qe - qs); }
		else { const char *b = q->qual + (q->l_seq - qe); for(uint32_t i = qe - qs; i > 0; i--) { o.push_back(b[i - 1]); } }
	} else { o.push_back('*'); }
}

/* mm_print_sam_md (5239-5298) */
void put_md(std::string &o, const MabSamRef *r, const MabSamRead *q, const uint32_t *s, const PathBits &pb)
{
	uint32_t aid = s[0], bid = s[1], apos = s[2], bpos = s[3], alen = s[4], blen = s[5];
	uint64_t ppos = (uint64_t)s[6] | (uint64_t)s[7] << 32;
	uint32_t rev = ~bid & 1, rid = aid >> 1;
	const uint8_t *rp = r[rid].seq + (r[rid].l_seq - apos - alen), *rb = rp;
	const uint8_t *qp = rev ? q->seq + (q->l_seq - bpos) : q->seq + (q->l_seq - bpos - blen);
	o.append("\tMD:Z:");
	auto del = [&](uint64_t c) {
		if(c > 0) { put_num(o, (uint64_t)(rp - rb)); o.push_back('^'); rb = rp + c; for(uint64_t i = 0; i < c; i++) { o.push_back(DEC_F[rp[i] & 15]); } rp += c; }
	};
	auto ins_f = [&](uint64_t c) { qp += c; };
	auto ins_r = [&](uint64_t c) { qp -= c; };
	/* _match_ff / _match_fr: scan 16 columns at a time, report the first mismatch of each chunk and restart behind it */
	auto match_ff = [&](uint64_t c) {
		rp += c; qp += c;
		for(uint64_t i = c, l = 0; i > 0; i -= l) {
			l = i < 16 ? i : 16;
			uint64_t mc = 0;
			while(mc < 16 && rp[(int64_t)mc - (int64_t)i] == qp[(int64_t)mc - (int64_t)i]) { mc++; }
			if(mc < l) { put_num(o, (uint64_t)(rp - i + mc - rb)); o.push_back(DEC_F[rp[(int64_t)mc - (int64_t)i] & 15]); rb = rp - i + mc + 1; l = mc + 1; }
		}
	};
	auto match_fr = [&](uint64_t c) {
		rp += c; qp -= c;
		for(uint64_t i = c, l = 0; i > 0; i -= l) {
			l = i < 16 ? i : 16;
			/* query chunk = the l bytes at qp + i - l, reversed and complemented by xor 3 (_rvbp_v16i8, 5258); lanes >= l compare junk and are ignored */
			uint64_t mc = 0;
			while(mc < l && rp[(int64_t)mc - (int64_t)i] == (uint8_t)(qp[(int64_t)i - 1 - (int64_t)mc] ^ 0x03)) { mc++; }
			if(mc < l) { put_num(o, (uint64_t)(rp - i + mc - rb)); o.push_back(DEC_F[rp[(int64_t)mc - (int64_t)i] & 15]); rb = rp - i + mc + 1; l = mc + 1; }
		}
	};
	if(rev == 0) { parse_rv(pb, ppos, (uint64_t)alen + blen, del, ins_f, match_ff); }
	else { parse_rv(pb, ppos, (uint64_t)alen + blen, del, ins_r, match_fr); }
	put_num(o, (uint64_t)(rp - rb));
}

uint32_t edit_dist(const Aln &a) { return (uint32_t)((double)a.dcnt * (1.0 - a.identity)) + a.agcnt + a.bgcnt; }		/* 5333-5337 */
}  // namespace

extern "C" void mab_sam_header(std::string &out, const MabSamRef *refs, uint32_t n_ref, const char *version, const char *cmdline)
{
	out.append("@HD\tVN:1.0\tSO:unsorted\n");
	for(uint32_t i = 0; i < n_ref; i++) { out.append("@SQ\tSN:"); out.append(refs[i].name, refs[i].l_name); out.append("\tLN:"); put_num(out, refs[i].l_seq); out.push_back('\n'); }
	out.append("@PG\tID:minialign\tPN:minialign\tVN:"); out.append(version); out.append("\tCL:"); out.append(cmdline); out.push_back('\n');
}

extern "C" void mab_sam_record(std::string &o, const MabSamRef *refs, const MabSamRead *q, const uint32_t *words, uint64_t n_words, uint32_t tags)
{
	if(n_words == 0) {										/* mm_print_sam_unmapped (5126-5141) */
		o.append(q->name, q->l_name); o.append("\t4\t*\t0\t0\t*\t*\t0\t0\t");
		{ size_t o0 = o.size(); o.resize(o0 + q->l_seq); char *d = &o[o0]; for(uint32_t i = 0; i < q->l_seq; i++) { d[i] = DEC_F[q->seq[i] & 15]; } }
		o.push_back('\t');
		if(q->qual && q->qual[0] != '\0') { o.append(q->qual, q->l_seq); } else { o.push_back('*'); }
		o.push_back('\n');
		return;
	}
	uint32_t n_all = words[0], n_uniq = words[1];
	std::vector<Aln> al(n_all);
	const uint32_t *p = words + 2;
	for(uint32_t i = 0; i < n_all; i++) { p = parse_aln(p, al[i]); }
	/* the reference keeps its MM_OMIT_REP flag (0x08, 2489) in the same word as the tag bits, where bit 3 is the IH tag (2533):
	 * asking for IH also drops the secondary records, and -R (omit) also prints IH (5323, 5401, 6110) */
	uint64_t n = (tags & (MAB_TAG_IH | MAB_OMIT_REP)) ? n_uniq : n_all;
	uint32_t flag = 0;
	for(uint64_t i = 0; i < n; i++) {							/* mm_print_sam_mapped (5389-5426) */
		if(i >= n_uniq) { flag = 0x100; }
		const Aln &a = al[i];
		PathBits pb = { a.path, a.npw, a.plen };
		for(uint64_t j = a.slen; j > 0; j--) {
			const uint32_t *s = a.seg + 8 * (j - 1);
			put_core(o, refs, q, s, pb, flag, a.mapq);
			/* general tags (5303-5340) */
			if(tags & MAB_TAG_NH) { o.append("\tNH:i:"); put_num(o, n_all); }
			if(tags & (MAB_TAG_IH | MAB_OMIT_REP)) { o.append("\tIH:i:"); put_num(o, i); }
			if(tags & MAB_TAG_AS) { o.append("\tAS:i:"); put_num(o, (uint32_t)a.score); }
			if(tags & MAB_TAG_NM) { o.append("\tNM:i:"); put_num(o, edit_dist(a)); }
			if(tags & MAB_TAG_MD) { put_md(o, refs, q, s, pb); }
			bool skip = false;
			if(i == 0 && j == a.slen) {							/* primary-only tags (5346-5384) */
				flag = 0x800;
				if(tags & MAB_TAG_XS) { o.append("\tXS:i:"); put_num(o, n_all > 1 ? (uint32_t)al[1].score : 0); }
				if((tags & MAB_TAG_SA) && (n_uniq > 1 || al[0].slen > 1)) {
					o.append("\tSA:Z:");
					for(uint64_t x = 0; x < n_uniq; x++) {
						const Aln &b = al[x];
						PathBits pbb = { b.path, b.npw, b.plen };
						for(uint64_t y = b.slen; y > 0; y--) {
							if(x == 0 && y == b.slen) { continue; }
							const uint32_t *t = b.seg + 8 * (y - 1);		/* mm_print_sam_supp (5203-5233) */
							uint32_t rid = t[0] >> 1, rs = refs[rid].l_seq - t[2] - t[4], hl = q->l_seq - t[3] - t[5], tl = t[3];
							o.append(refs[0].name, refs[0].l_name); o.push_back(',');	/* the reference prints r->name, i.e. the FIRST sequence's (5217) */
							put_num(o, (uint64_t)rs + 1); o.push_back(',');
							o.push_back((t[1] & 1) ? '+' : '-'); o.push_back(',');
							if(hl) { put_num(o, hl); o.push_back('H'); }
							put_cigar_rv(o, pbb, (uint64_t)t[6] | (uint64_t)t[7] << 32, (uint64_t)t[4] + t[5]);
							if(tl) { put_num(o, tl); o.push_back('H'); }
							o.push_back(','); put_num(o, b.mapq); o.push_back(','); put_num(o, edit_dist(b)); o.push_back(';');
						}
					}
					skip = true;
				}
			}
			o.push_back('\n');
			if(skip) { i = n; break; }
		}
		flag = 0x800;
	}
}

extern "C" char *mab_sam_format_c(const MabSamRef *refs, uint32_t n_ref, const MabSamRead *read, const uint32_t *words, uint64_t n_words, uint32_t tags, uint64_t *len)
{
	(void)n_ref;
	std::string s;
	mab_sam_record(s, refs, read, words, n_words, tags);
	char *p = (char *)malloc(s.size() + 1);
	memcpy(p, s.data(), s.size()); p[s.size()] = 0; *len = s.size();
	return p;
}
extern "C" void mab_sam_free(char *p) { free(p); }

extern "C" uint32_t mab_sam_parse_tags(const char *list)
{
	uint32_t t = 0;
	for(const char *p = list; p && *p;) {
		const char *e = p; while(*e && !strchr(",;:/", *e)) { e++; }
		if(e - p == 2) {
			if(!strncmp(p, "RG", 2)) t |= MAB_TAG_RG; else if(!strncmp(p, "NH", 2)) t |= MAB_TAG_NH; else if(!strncmp(p, "IH", 2)) t |= MAB_TAG_IH;
			else if(!strncmp(p, "AS", 2)) t |= MAB_TAG_AS; else if(!strncmp(p, "XS", 2)) t |= MAB_TAG_XS; else if(!strncmp(p, "NM", 2)) t |= MAB_TAG_NM;
			else if(!strncmp(p, "SA", 2)) t |= MAB_TAG_SA; else if(!strncmp(p, "MD", 2)) t |= MAB_TAG_MD;
		}
		p = *e ? e + 1 : e;
	}
	return t;
}
