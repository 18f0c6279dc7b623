/*
 * mab_text.cuh -- the sequence reader on the device: a chunk of FASTA / FASTQ text (whole records, as the host cut it from the
 * file) is indexed and packed into the mapper's batch layout (1 byte/base codes with 64 B zero margins, bseq_t,
 * minialign.c:2109-2146) without the host touching a base.  What the reference's reader does per record
 * (bseq_read_fasta, minialign.c:1996-2088) is reproduced for well-formed input:
 *
 *   header   delimiter ('>' / '@') at a line start; up to 32 spaces behind it are skipped (state 2); the name runs to the first
 *            space or newline, one '\r' in front of the newline is dropped (state 3); the rest of the line is the comment
 *   sequence every byte except '\n' up to the next record, encoded through the low-nibble table encaf (214-232)
 *   FASTQ    four lines per record (header, sequence, '+', quality); the quality line is kept only on request (-Q)
 *   records with an empty sequence are dropped (min_len = 1, 2077, 6145)
 *
 * Not taken here (TextCounters::err = MAB_TXT_EFORMAT, the caller falls back to the host reader): text that does not start with
 * the delimiter, FASTQ with wrapped sequence / quality lines or blank lines between records.
 *
 *   k_text_count / k_text_mark   record starts (FASTA: '>' after '\n') or line ends (FASTQ), counted per 4 KB tile, then written
 *                                in order behind an exclusive scan of the tile counts (k_scan_u32)
 *   k_text_index                 one warp per record: name, sequence range, base count
 *   k_text_layout                offsets of the reads in the base block (scan), ReadRec array, totals for the host
 *   k_text_pack                  one warp per record: encode + squeeze the newlines out (ballot compaction), coalesced both ways
 */
#pragma once
#include "mab_pipe.cuh"

namespace mab {

#define MAB_TXT_TILE 4096			/* bytes per CTA of the count / mark kernels: 256 threads x 16 B */

/* 16-bit mask of the marked bytes among the 16 at p (aligned); bytes at or beyond n read as 0 */
__device__ __forceinline__ uint32_t text_mask16(const uint8_t *text, uint64_t n, uint64_t p, uint32_t fastq)
{
	if(p >= n) { return 0; }
	uint4 v = *(const uint4 *)(text + p);
	uint32_t w[4] = { v.x, v.y, v.z, v.w };
	uint32_t prev = p == 0 ? '\n' : text[p - 1], m = 0;
	#pragma unroll
	for(int j = 0; j < 16; j++) {
		uint32_t c = (w[j >> 2] >> (8 * (j & 3))) & 0xff;
		uint32_t hit = fastq ? (c == '\n') : (c == '>' && prev == '\n');
		if(p + j < n) { m |= hit << j; }
		prev = c;
	}
	return m;
}

__global__ void k_text_count(const uint8_t *text, uint64_t n, uint32_t fastq, uint32_t *tile_cnt)
{
	__shared__ uint64_t sm[34];
	uint64_t p = (uint64_t)blockIdx.x * MAB_TXT_TILE + 16ull * threadIdx.x;
	uint64_t tot;
	block_excl_sum((uint64_t)__popc(text_mask16(text, n, p, fastq)), sm, &tot);
	if(threadIdx.x == 0) { tile_cnt[blockIdx.x] = (uint32_t)tot; }
}

/* exclusive scan of m counters in place (single CTA); *total = sum */
__global__ void k_scan_u32(uint32_t *a, uint32_t m, unsigned long long *total)
{
	__shared__ uint64_t sm[34];
	uint64_t carry = 0;
	for(uint32_t i0 = 0; i0 < m; i0 += blockDim.x) {
		uint32_t i = i0 + threadIdx.x;
		uint64_t x = i < m ? a[i] : 0, tot;
		uint64_t e = block_excl_sum(x, sm, &tot);
		if(i < m) { a[i] = (uint32_t)(carry + e); }
		carry += tot;
	}
	if(threadIdx.x == 0) { *total = carry; }
}

__global__ void k_text_mark(const uint8_t *text, uint64_t n, uint32_t fastq, const uint32_t *tile_ofs, uint32_t *marks, uint64_t mark_cap, TextCounters *tc)
{
	__shared__ uint64_t sm[34];
	uint64_t p = (uint64_t)blockIdx.x * MAB_TXT_TILE + 16ull * threadIdx.x;
	uint32_t m = text_mask16(text, n, p, fastq);
	uint64_t tot;
	uint64_t o = tile_ofs[blockIdx.x] + block_excl_sum((uint64_t)__popc(m), sm, &tot);
	while(m) {
		int j = __ffs((int)m) - 1; m &= m - 1;
		if(o < mark_cap) { marks[o] = (uint32_t)(p + j); } else { atomicOr(&tc->err, MAB_TXT_EMARKS); }
		o++;
	}
}

/* first position in [p, e) whose byte satisfies pred (0: ' ' or '\n', 1: '\n', 2: not ' '), or e; warp-cooperative */
__device__ __forceinline__ uint64_t text_find(const uint8_t *text, uint64_t p, uint64_t e, int pred, int lane)
{
	for(; p < e; p += 32) {
		uint64_t q = p + lane;
		uint32_t c = q < e ? text[q] : '\n';
		int hit = pred == 0 ? (c == ' ' || c == '\n') : pred == 1 ? (c == '\n') : (c != ' ');
		uint32_t b = __ballot_sync(0xffffffffu, hit);
		if(b) { uint64_t r = p + (uint32_t)(__ffs((int)b) - 1); return r < e ? r : e; }
	}
	return e;
}

__global__ void k_text_index(const uint8_t *text, uint64_t n, const uint32_t *marks, TextCounters *tc, TextRec *recs, uint64_t rec_cap)
{
	int lane = threadIdx.x & 31;
	uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
	const uint32_t fastq = tc->fastq;
	uint64_t n_mark = tc->n_mark;
	if(tc->err & MAB_TXT_EMARKS) { return; }
	uint64_t n_rec = fastq ? n_mark / 4 : n_mark;
	if(gw == 0 && lane == 0) {
		uint32_t err = 0;
		if(fastq && (n_mark & 3) != 0) { err |= MAB_TXT_EFORMAT; }
		if(!fastq && (n_mark == 0 || marks[0] != 0)) { err |= MAB_TXT_EFORMAT; }
		if(n_rec > rec_cap) { err |= MAB_TXT_ERECS; }
		if(err) { atomicOr(&tc->err, err); }
		tc->n_rec = (uint32_t)n_rec;
	}
	if(n_rec > rec_cap) { return; }
	for(uint64_t i = gw; i < n_rec; i += nw) {
		uint64_t s, e, seq_beg, seq_end, qual = 0;
		uint32_t bad = 0;
		if(fastq) {
			s = i ? (uint64_t)marks[4 * i - 1] + 1 : 0;
			uint64_t nl0 = marks[4 * i], nl1 = marks[4 * i + 1], nl2 = marks[4 * i + 2], nl3 = marks[4 * i + 3];
			e = nl0 + 1; seq_beg = nl0 + 1; seq_end = nl1; qual = nl2 + 1;
			if(text[s] != '@' || text[nl1 + 1] != '+' || nl3 - nl2 != nl1 - nl0) { bad = 1; }
		} else {
			s = marks[i]; e = i + 1 < n_rec ? (uint64_t)marks[i + 1] : n;
			seq_beg = 0; seq_end = e;
		}
		/* header line: [s + 1, hdr_end) with hdr_end at its '\n' (FASTA: searched; FASTQ: known) */
		uint64_t p = text_find(text, s + 1, (s + 33 < e ? s + 33 : e), 2, lane);			/* _strip: one 32-byte window of spaces */
		uint64_t ne = text_find(text, p, e, 0, lane);
		uint64_t name_len = ne - p;
		uint64_t hdr_nl = (ne < e && text[ne] == '\n') ? ne : text_find(text, ne, e, 1, lane);
		if(ne == hdr_nl && name_len > 0 && text[ne - 1] == '\r') { name_len--; }
		uint32_t len;
		if(fastq) { len = (uint32_t)(seq_end - seq_beg); }
		else {
			seq_beg = hdr_nl < e ? hdr_nl + 1 : e;
			uint32_t cnt = 0;
			for(uint64_t q0 = seq_beg; q0 < seq_end; q0 += 128) {
				#pragma unroll
				for(int u = 0; u < 4; u++) { uint64_t q = q0 + 32 * u + lane; cnt += (q < seq_end && text[q] == '\n'); }
			}
			cnt = __reduce_add_sync(0xffffffffu, cnt);
			len = (uint32_t)(seq_end - seq_beg) - cnt;
		}
		if(lane == 0) {
			TextRec r;
			r.name_ofs = p; r.seq_beg = seq_beg; r.seq_end = seq_end; r.qual_ofs = qual; r.name_len = (uint32_t)name_len; r.len = len;
			r.flags = (len < 1 ? MAB_TR_DROPPED : 0) | (fastq ? MAB_TR_HASQUAL : 0); r._pad = 0; r.sam_ofs = 0; r.sam_len = 0;
			recs[i] = r;
			if(bad) { atomicOr(&tc->err, MAB_TXT_EFORMAT); }
		}
	}
}

/* offsets of the reads in the base block: read i at 64 + sum over j < i of (len_j + 64) (dropped records take no room) */
__global__ void k_text_layout(const TextRec *recs, ReadRec *reads, TextCounters *tc)
{
	__shared__ uint64_t sm[34];
	__shared__ uint32_t smax[32];
	if(tc->err) { return; }
	uint32_t n = tc->n_rec, mx = 0;
	uint64_t carry = 64, tot_len = 0;
	for(uint32_t i0 = 0; i0 < n; i0 += blockDim.x) {
		uint32_t i = i0 + threadIdx.x;
		uint32_t len = i < n ? recs[i].len : 0;
		uint64_t sz = len ? (uint64_t)len + 64 : 0, tot;
		uint64_t e = block_excl_sum(sz, sm, &tot);
		if(i < n) {
			ReadRec r; memset(&r, 0, sizeof(r));
			r.seq_ofs = carry + e; r.len = len; r.rlen_in = MAB_RLEN_OWN;
			reads[i] = r;
		}
		carry += tot; mx = len > mx ? len : mx;
		uint64_t tl; block_excl_sum(len, sm, &tl); tot_len += tl;
	}
	mx = (uint32_t)__reduce_max_sync(0xffffffffu, (int)mx);
	if((threadIdx.x & 31) == 0) { smax[threadIdx.x >> 5] = mx; }
	__syncthreads();
	if(threadIdx.x == 0) {
		for(uint32_t w = 1; w < (blockDim.x >> 5); w++) { mx = smax[w] > mx ? smax[w] : mx; }
		tc->tot_len = tot_len; tc->span = carry + 64; tc->maxlen = mx;
	}
}

__device__ __forceinline__ uint32_t text_encode(uint32_t c)								/* encaf[c & 15] (minialign.c:225-229) */
{
	const unsigned long long lut = (1ull << 12) | (2ull << 28) | (3ull << 16) | (3ull << 20) | (4ull << 56);
	return (uint32_t)(lut >> (4 * (c & 15))) & 15u;
}

__global__ void k_text_pack(const uint8_t *text, const TextRec *recs, const ReadRec *reads, uint32_t n_rec, uint8_t *base)
{
	int lane = threadIdx.x & 31;
	uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
	const uint32_t below = (1u << lane) - 1;
	for(uint32_t i = gw; i < n_rec; i += nw) {
		if(recs[i].len == 0) { continue; }
		uint64_t q0 = recs[i].seq_beg, e = recs[i].seq_end;
		uint8_t *o = base + reads[i].seq_ofs;
		for(; q0 < e; q0 += 32) {
			uint64_t q = q0 + lane;
			uint32_t c = q < e ? text[q] : '\n';
			uint32_t keep = __ballot_sync(0xffffffffu, c != '\n');
			if(c != '\n') { o[__popc(keep & below)] = (uint8_t)text_encode(c); }
			o += __popc(keep);
		}
	}
}

}  // namespace mab
