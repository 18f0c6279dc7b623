/*
 * mab_index.cpp -- host-side index construction (minialign.c:2750-2997, `minialign -d` / FASTA references on the command
 * line) and the .mai container writer (minialign.c:1135-1502, 3040-3127).  Offline, CPU: outside the GPU hot path (SURVEY.md
 * section 8 f2), but part of the drop-in command line.
 *
 * The mapping results depend on the index through (a) the minimizer set, (b) the ORDER of the occurrences of a minimizer
 * (seeds are generated in that order and the unstable seed sort keeps ties in generation order) and (c) the occurrence
 * thresholds occ[].  (b) is fixed by the reference's unstable radix sort of every first-stage bucket (ksort.h:82-131), whose
 * permutation cycles are followed step by step below; the second-stage hash tables only have to be valid linear-probing tables
 * for the probe in mab_scalar.cuh (idx_get), their slot order is not observable.
 *
 * Blob layout = the payload of a .mai index block: mm_idx_t (64 B) | buckets 32 B x 2^b | sequence
 * table 24 B x n_seq | per non-empty bucket {slots 16 B x (mask + 1), u64 p[p[0] + 1]} | names and 1 B/base sequences with
 * 64 B margins.  All "pointers" are byte offsets from the start of the blob.
 */
#include "mab_index.h"
#include <zlib.h>
#include <algorithm>
#include <cstdio>
#include <cstring>

namespace {

struct Mini { uint64_t hrem; uint32_t pos, rid; };		/* mm_mini_t (minialign.c:2658-2664) */

/* CRC32C step of hash64 (minialign.c:2353): seed = low word of the data, so it vanishes for k <= 16 */
uint32_t crc32c_u64(uint32_t crc, uint64_t data)
{
	if((uint32_t)data == crc && (data >> 32) == 0) { return 0; }
	for(int i = 0; i < 64; i++) {
		uint32_t bit = (crc ^ (uint32_t)(data >> i)) & 1u;
		crc = (crc >> 1) ^ (0x82f63b78u & (0u - bit));
	}
	return crc;
}

/* (w,k)-minimizers of one sequence in emission order: {hash, strand, position} (mm_sketch, minialign.c:2410-2435, decoded like
 * mm_idx_drain_intl, 2826-2841).  Same formulation as the device sketch (k_seed_scan): candidate j = hash << 8 | strand << 7 |
 * j mod w, window minimum over the last w candidates, a word is emitted when the minimum changes or a new equal minimum
 * appears; the position is rebuilt from the in-window index exactly like the reference's decoder. */
void sketch(const uint8_t *seq, uint32_t len, uint32_t k, uint32_t w, std::vector<uint64_t> &words)
{
	words.clear();
	if(len < k) { return; }
	const uint64_t kk = k - 1, shift1 = 2 * kk, mask = (1ull << (2 * k)) - 1;
	const uint32_t npos = len - (uint32_t)kk;
	std::vector<uint64_t> ring(64, ~0ull);
	uint64_t k0 = 0, k1 = 0, u = 0;
	for(uint64_t i = 0; i < kk; i++) { uint64_t c = seq[i]; k0 = (k0 << 2 | c) & mask; k1 = (k1 >> 2) | ((3ull ^ c) << shift1); }
	for(uint32_t j = 0; j < npos; j++) {
		uint64_t c = seq[j + kk];
		k0 = (k0 << 2 | c) & mask; k1 = (k1 >> 2) | ((3ull ^ c) << shift1);						/* _push_kmer: N (4) leaks into the neighbour */
		uint64_t km = k0 < k1 ? k0 : k1, kx = k0 < k1 ? k1 : k0, mm = k0 < k1 ? 0 : 0x80;
		uint64_t h = ((uint64_t)crc32c_u64((uint32_t)kx, kx) ^ km) & mask;
		uint64_t enc = h << 8 | (uint64_t)(j % w) | mm;
		ring[j & 63] = enc;
		uint64_t v = ~0ull;
		uint32_t lo = j + 1 >= w ? j + 1 - w : 0;
		for(uint32_t t = lo; t <= j; t++) { v = std::min(v, ring[t & 63]); }
		if(v == enc || v != u) { words.push_back(v); }
		u = v;
	}
}

/* ---- bucket sort of the minimizers by their key remainder ----
 * The order in which equal keys (the occurrences of one minimizer) end up is observable: it is the order of the occurrence array
 * in the index and from there of the seeds.  The reference sorts with an in-place MSD radix sort whose distribution pass is not
 * stable, so the permutation it performs is followed here step by step, in the same index-based form as the device and host
 * sorts of the mapping path (mab_scalar.cuh radix_sort_exact_warp, mab_host.inl rs_sort64): digit histogram -> [head, tail)
 * range per digit -> for every digit in turn, the element at the head is carried along its displacement cycle until one that
 * belongs here comes back -> ranges above 64 elements recurse on the next digit, smaller ones are finished by insertion. */
inline unsigned digit_of(const Mini &m, int shift) { return (unsigned)(m.hrem >> shift) & 255u; }

void insertion_by_key(Mini *a, size_t n)
{
	for(size_t i = 1; i < n; i++) {
		if(!(a[i].hrem < a[i - 1].hrem)) { continue; }
		Mini t = a[i];
		size_t j = i;
		while(j > 0 && t.hrem < a[j - 1].hrem) { a[j] = a[j - 1]; j--; }
		a[j] = t;
	}
}

void flag_sort(Mini *a, size_t n, int shift)
{
	size_t head[256], tail[256];
	memset(tail, 0, sizeof(tail));
	for(size_t i = 0; i < n; i++) { tail[digit_of(a[i], shift)]++; }
	head[0] = 0;
	for(int d = 1; d < 256; d++) { tail[d] += tail[d - 1]; head[d] = tail[d - 1]; }
	for(int d = 0; d < 256; ) {
		if(head[d] == tail[d]) { d++; continue; }
		unsigned t = digit_of(a[head[d]], shift);
		if(t == (unsigned)d) { head[d]++; continue; }
		Mini carried = a[head[d]];								/* displacement cycle: put it where it belongs, pick up what was there */
		while(t != (unsigned)d) { std::swap(carried, a[head[t]]); head[t]++; t = digit_of(carried, shift); }
		a[head[d]++] = carried;
	}
	if(shift == 0) { return; }
	const int next = shift > 8 ? shift - 8 : 0;
	size_t beg = 0;
	for(int d = 0; d < 256; d++) {
		size_t sz = tail[d] - beg;
		if(sz > 64) { flag_sort(a + beg, sz, next); } else if(sz > 1) { insertion_by_key(a + beg, sz); }
		beg = tail[d];
	}
}
void radix_sort_128x(Mini *p, size_t l) { if(l <= 64) { insertion_by_key(p, l); } else { flag_sort(p, l, 56); } }

void put64(std::vector<uint8_t> &v, size_t ofs, uint64_t x) { memcpy(v.data() + ofs, &x, 8); }
void put32(std::vector<uint8_t> &v, size_t ofs, uint32_t x) { memcpy(v.data() + ofs, &x, 4); }
void put16(std::vector<uint8_t> &v, size_t ofs, uint16_t x) { memcpy(v.data() + ofs, &x, 2); }

}  // namespace

bool mab_build_index(const std::vector<MabIdxSeq> &refs, const MabIdxParams &prm, std::vector<uint8_t> &blob, std::string &err)
{
	if(prm.k == 0 || prm.k > 31 || prm.w == 0 || prm.w >= 32 || prm.n_frq == 0 || prm.n_frq > 7) { err = "bad index parameters"; return false; }
	if(refs.empty()) { err = "no reference sequence"; return false; }
	const uint32_t b = std::min<uint32_t>(prm.k * 2, prm.b);									/* clip bucket size (2950) */
	const uint64_t nb = 1ull << b, bmask = nb - 1;
	std::vector<std::vector<Mini>> bkt(nb);
	/* sketch every sequence, push the minimizers to the first-stage buckets in sequence order (2826-2841) */
	std::vector<uint64_t> words;
	for(size_t si = 0; si < refs.size(); si++) {
		if(refs[si].seq.size() >= 0x7fffffffu) { err = "reference sequence too long"; return false; }
		sketch(refs[si].seq.data(), (uint32_t)refs[si].seq.size(), prm.k, prm.w, words);
		uint64_t base = 0 - (uint64_t)prm.w, v = prm.w;
		for(uint64_t p : words) {
			uint64_t u = p & 0x7f, fr = (p >> 7) & 1, h = p >> 8;
			base += u <= v ? prm.w : 0; v = u;
			Mini m; m.hrem = h >> b; m.pos = (uint32_t)(base + u); m.rid = (uint32_t)((si << 1) + fr);
			bkt[h & bmask].push_back(m);
		}
	}
	/* sort every bucket, count keys and per-key occurrences (mm_idx_count_occ, 2868-2901) */
	std::vector<uint32_t> cnt;
	std::vector<uint32_t> n_keys(nb, 0), n_single(nb, 0);
	for(uint64_t i = 0; i < nb; i++) {
		std::vector<Mini> &a = bkt[i];
		if(a.empty()) { continue; }
		radix_sort_128x(a.data(), a.size());
		uint32_t n = 1, keys = 1, single = 0;
		for(size_t j = 1; j < a.size(); j++) {
			if(a[j].hrem != a[j - 1].hrem) { single += n == 1; cnt.push_back(n); n = 0; keys++; }
			n++;
		}
		single += n == 1; cnt.push_back(n);
		n_keys[i] = keys; n_single[i] = single;
	}
	if(cnt.empty()) { err = "no minimizer found (sequences shorter than k?)"; return false; }
	/* occurrence thresholds (2980-2986): (k-th smallest count) + 1 */
	uint32_t occ[7] = { 0 };
	for(uint32_t i = 0; i < prm.n_frq; i++) {
		if(prm.frq[i] <= 0.0f) { occ[i] = 0xffffffffu; continue; }
		size_t kth = (size_t)(uint32_t)((1.0 - prm.frq[i]) * (double)cnt.size());
		if(kth >= cnt.size()) { kth = cnt.size() - 1; }
		std::vector<uint32_t> c(cnt);
		std::nth_element(c.begin(), c.begin() + kth, c.end());
		occ[i] = c[kth] + 1;
	}
	const uint64_t max_cnt = occ[prm.n_frq - 1];
	/* layout */
	const uint64_t bkt_ofs = 64, s_ofs = bkt_ofs + 32 * nb;
	uint64_t ofs = s_ofs + 24ull * refs.size();
	struct BL { uint64_t a_ofs, p_ofs, size, np; };
	std::vector<BL> bl(nb, BL{ 0, 0, 0, 0 });
	for(uint64_t i = 0; i < nb; i++) {
		if(bkt[i].empty()) { continue; }
		uint64_t need = (uint64_t)(1.1 * n_keys[i] / 0.4), size = 256;							/* kh_init_static (369-389) */
		while(size < need) { size <<= 1; }
		bl[i].size = size; bl[i].np = bkt[i].size() - n_single[i];								/* r[0] (2944) */
		bl[i].a_ofs = ofs; ofs += 16 * size;
		bl[i].p_ofs = ofs; ofs += 8 * (bl[i].np + 1);
	}
	std::vector<uint64_t> seq_ofs(refs.size()), name_ofs(refs.size());
	for(size_t si = 0; si < refs.size(); si++) {
		name_ofs[si] = ofs; ofs += (refs[si].name.size() + 1 + 7) & ~7ull;
		ofs += 64; seq_ofs[si] = ofs; ofs += refs[si].seq.size(); ofs += 64; ofs = (ofs + 7) & ~7ull;
	}
	blob.assign(ofs, 0);
	/* mm_idx_t header (2476-2483) */
	put64(blob, 0, bkt_ofs); put64(blob, 8, bmask);
	blob[16] = (uint8_t)b; blob[17] = (uint8_t)prm.w; blob[18] = (uint8_t)prm.k; blob[19] = (uint8_t)prm.n_frq;
	for(int i = 0; i < 7; i++) { put32(blob, 20 + 4 * i, occ[i]); }
	put32(blob, 48, (uint32_t)refs.size()); put32(blob, 52, 1); put64(blob, 56, s_ofs);
	/* second-stage tables (mm_idx_build_hash, 2907-2948) */
	for(uint64_t i = 0; i < nb; i++) {
		const std::vector<Mini> &a = bkt[i];
		if(a.empty()) { continue; }
		const uint64_t size = bl[i].size, hmask = size - 1;
		size_t so = bl[i].a_ofs, po = bl[i].p_ofs;
		for(uint64_t j = 0; j < size; j++) { put64(blob, so + 16 * j, ~0ull); put64(blob, so + 16 * j + 8, ~0ull); }
		uint64_t sp = 0, n_put = 0;
		/* Literal port of the fill loop (2931-2937): a run is stored when it ends and holds at most max_cnt occurrences.  `q`
		 * (start of the pending run) only advances when a run is stored, so after one over-frequent minimizer every later key
		 * of the bucket fails the length test as well and is dropped: reference behaviour, reproduced. */
		auto fill = [&](size_t &q, size_t p) {
			uint64_t key = a[q].hrem, val = (uint64_t)a[q].pos | (uint64_t)a[q].rid << 32;
			if(++q < p) {
				put64(blob, po + 8 * (++sp), val); val = sp << 32 | 1ull << 63 | 1;
				do { put64(blob, po + 8 * (++sp), (uint64_t)a[q].pos | (uint64_t)a[q].rid << 32); val++; } while(++q < p);
			}
			uint64_t pos = key & hmask;
			while(true) {																		/* plain linear probing: valid for idx_get's scan */
				uint64_t kk; memcpy(&kk, blob.data() + so + 16 * pos, 8);
				if(kk == ~0ull || kk == key) { break; }
				pos = (pos + 1) & hmask;
			}
			put64(blob, so + 16 * pos, key); put64(blob, so + 16 * pos + 8, val);
			n_put++;
		};
		size_t q = 0, p = 1;
		for(; p < a.size(); p++) {
			if(a[p - 1].hrem != a[p].hrem && (uint64_t)(p - q) <= max_cnt) { fill(q, p); }
		}
		if((uint64_t)(p - q) <= max_cnt) { fill(q, p); }
		put64(blob, po, bl[i].np);
		size_t bo = bkt_ofs + 32 * i;															/* kh_t {mask, max, cnt, ub, a} + p */
		put32(blob, bo, (uint32_t)hmask); put32(blob, bo + 4, (uint32_t)size); put32(blob, bo + 8, (uint32_t)n_put); put32(blob, bo + 12, (uint32_t)(size * 0.4));
		put64(blob, bo + 16, bl[i].a_ofs); put64(blob, bo + 24, bl[i].p_ofs);
	}
	/* sequence table and bodies (mm_idx_seq_t, 2464-2470) */
	for(size_t si = 0; si < refs.size(); si++) {
		size_t so = s_ofs + 24 * si;
		put64(blob, so, seq_ofs[si]); put64(blob, so + 8, name_ofs[si]);
		put32(blob, so + 16, (uint32_t)refs[si].seq.size()); put16(blob, so + 20, (uint16_t)refs[si].name.size()); put16(blob, so + 22, 0);
		memcpy(blob.data() + name_ofs[si], refs[si].name.data(), refs[si].name.size());
		if(!refs[si].seq.empty()) { memcpy(blob.data() + seq_ofs[si], refs[si].seq.data(), refs[si].seq.size()); }
	}
	return true;
}

/* "PG00" framed zlib stream: frames of <= 1 MiB raw, deflate level 1, terminator frame with length 0xffffffff */
bool mab_write_mai(const char *path, const std::vector<uint8_t> &blob)
{
	FILE *fp = fopen(path, "wb");
	if(!fp) { return false; }
	std::vector<uint8_t> raw(12 + blob.size());
	uint32_t magic = 0x0849414du; uint64_t size = blob.size();
	memcpy(raw.data(), &magic, 4); memcpy(raw.data() + 4, &size, 8); memcpy(raw.data() + 12, blob.data(), blob.size());
	std::vector<uint8_t> cbuf(compressBound(1 << 20));
	bool ok = true;
	for(size_t p = 0; p < raw.size() && ok; p += 1 << 20) {
		size_t n = std::min<size_t>(1 << 20, raw.size() - p);
		uLongf cl = (uLongf)cbuf.size();
		if(compress2(cbuf.data(), &cl, raw.data() + p, (uLong)n, 1) != Z_OK) { ok = false; break; }
		uint32_t len = (uint32_t)cl;
		ok = fwrite("PG00", 1, 4, fp) == 4 && fwrite(&len, 4, 1, fp) == 1 && fwrite(cbuf.data(), 1, cl, fp) == cl;
	}
	uint32_t term = 0xffffffffu;
	ok = ok && fwrite("PG00", 1, 4, fp) == 4 && fwrite(&term, 4, 1, fp) == 1;
	return fclose(fp) == 0 && ok;
}
