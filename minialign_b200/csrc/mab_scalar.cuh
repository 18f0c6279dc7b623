/*
 * mab_scalar.cuh -- single-thread device routines of the mapper: index probe, exact radix sort, array chaining, the
 * position hash.  These are the data-dependent, branchy parts of minialign's map.c (minialign.c:3171-4067); they run on
 * one lane per read (thousands of reads in flight) because their control flow is strictly sequential per read and
 * bit-exactness depends on reproducing it step by step.
 */
#pragma once
#include "mab_types.h"

namespace mab {

#define MAB_OFS0 0x40000000u

/* index image loads: explicit ld.global.nc.  Besides the read-only path this matters for code generation: a load through the
 * generic pointer stored in DevParams counts as thread-varying for the compiler's uniformity analysis, a global load from a
 * warp-uniform address does not, and everything derived from it (section lengths -> loop bounds) stays warp-uniform. */
__device__ __forceinline__ uint64_t ldg64(const uint8_t *p) { return __ldg((const unsigned long long *)p); }
__device__ __forceinline__ uint32_t ldg32(const uint8_t *p) { return __ldg((const uint32_t *)p); }
__device__ __forceinline__ uint16_t ldg16(const uint8_t *p) { return __ldg((const uint16_t *)p); }

/* reference sequence descriptor (mm_idx_seq_t, minialign.c:2464-2470) */
struct RefSeq { const uint8_t *seq; uint32_t l_seq; uint32_t circular; };
__device__ __forceinline__ RefSeq ref_seq(const DevParams &P, uint32_t rid)
{
	const uint8_t *s = P.idx + P.seq_ofs + 24ull * rid;
	RefSeq r; r.seq = P.idx + ldg64(s); r.l_seq = ldg32(s + 16); r.circular = ldg16(s + 22);
	return r;
}

/* CRC32C step of the minimizer hash (minialign.c:2353).  The same 64-bit word is both the seed (low half) and the data,
 * so for k <= 16 the result is 0; the loop is only reached for k >= 17. */
__device__ __forceinline__ uint32_t crc32c_u64(uint32_t crc, uint64_t data)
{
	if((uint32_t)data == crc && (data >> 32) == 0) { return 0; }
	for(int i = 0; i < 64; i++) {
		uint32_t bit = (crc ^ (uint32_t)(data >> i)) & 1u;
		crc = (crc >> 1) ^ (0x82f63b78u & (0u - bit));
	}
	return crc;
}

/* mm_idx_get (minialign.c:2727-2748) + kh_get_ptr (634-643): 2-level probe.  Returns the occurrence array (8 B each)
 * and n; the dependent loads are bucket header -> hash slot(s) -> value array. */
__device__ __forceinline__ const uint8_t *idx_get(const DevParams &P, uint64_t minier, uint32_t *n)
{
	const uint8_t *bk = P.idx + P.bkt_ofs + 32ull * (minier & P.bkt_mask);
	uint64_t hmask = ldg32(bk), a = ldg64(bk + 16), pofs = ldg64(bk + 24);
	*n = 0;
	if(a == 0) { return nullptr; }
	uint64_t key = minier >> P.b, pos = key & hmask, kk;
	const uint8_t *val = nullptr;
	do {
		const uint8_t *slot = P.idx + a + 16ull * pos;
		kk = ldg64(slot);
		if(kk == key) { val = slot + 8; break; }
		pos = hmask & (pos + 1);
	} while(kk + 1 != 0);
	if(val == nullptr) { return nullptr; }
	uint64_t v = ldg64(val);
	if((int64_t)v >= 0) { *n = 1; return val; }
	*n = (uint32_t)v;
	return P.idx + pofs + 8ull * ((v >> 32) & 0x7fffffff);
}

/* ---------------------------------------------------------------- exact radix sort (ksort.h:82-131) */
/* In-place American-flag sort on 8-bit digits, insertion sort at <= 64 elements.  The reference's sort is unstable and
 * the order it leaves equal keys in is observable downstream (chaining walks the array in order), so the permutation
 * cycles are followed in exactly the reference's order.  Iterative, with an explicit frame stack in `frames`
 * (8 levels x 257 u32).  esz = element size in u32 words: 4 (key = words 1:0) or 2 (key = word 0). */
/* ESZ = element size in u32 words: 4 (seeds / rescue entries, key = words 1:0) or 2 (roots / next, key = word 0) */
template <int ESZ> __device__ __forceinline__ uint64_t rs_key(const uint32_t *e) { return ESZ == 4 ? ((uint64_t)e[1] << 32 | e[0]) : (uint64_t)e[0]; }
template <int ESZ> __device__ __forceinline__ void rs_copy(uint32_t *d, const uint32_t *s)
{
	if(ESZ == 4) { *(uint4 *)d = *(const uint4 *)s; } else { *(uint2 *)d = *(const uint2 *)s; }
}

/* insertion sort of the reference (ksort.h:82-95): stable, used below 65 elements */
template <int ESZ>
__device__ __forceinline__ void rs_insertion(uint32_t *a, uint32_t n)
{
	for(uint32_t i = 1; i < n; i++) {
		if(rs_key<ESZ>(a + ESZ * i) < rs_key<ESZ>(a + ESZ * (i - 1))) {
			uint32_t tmp[4] __attribute__((aligned(16)));
			rs_copy<ESZ>(tmp, a + ESZ * i);
			uint64_t tk = rs_key<ESZ>(tmp);
			uint32_t j;
			for(j = i; j > 0 && tk < rs_key<ESZ>(a + ESZ * (j - 1)); j--) { rs_copy<ESZ>(a + ESZ * j, a + ESZ * (j - 1)); }
			rs_copy<ESZ>(a + ESZ * j, tmp);
		}
	}
}

/* Warp-cooperative exact sort.  The reference's sort is unstable and the order it leaves equal keys in is observable
 * downstream (chaining walks the array in order), so the permutation cycles of every distribution pass are followed by
 * lane 0 in exactly the reference's order; the digit histogram, the prefix sum, the scan that skips levels on which all
 * keys of a range share the digit (such a pass moves nothing in the reference and recurses into the same range) and the
 * insertion sorts of the (independent) small buckets are spread over the 32 lanes.  Buckets larger than 64 go on an explicit
 * stack of {beg, cnt | shift/8 << 29} pairs; sub-ranges are disjoint, so the order they are processed in does not matter.
 * `sm` = 512 u32 of shared memory owned by this warp, `stack` = MAB_RS_STACK pairs of (global) scratch owned by this warp.
 * Structured control flow only: this is inlined into k_extend, whose shuffles rely on provable warp convergence. */
#define MAB_RS_FRAME 260						/* scratch u32 per "frame"; callers allocate 8 frames per warp */
#define MAB_RS_STACK (8 * MAB_RS_FRAME / 2)
template <int ESZ>
__device__ __forceinline__ void radix_sort_exact_warp(uint32_t *a, uint32_t n, uint32_t *stack, uint32_t *sm, int lane, uint32_t *err)
{
	constexpr int esz = ESZ;
	if(n <= 64) { if(lane == 0) { rs_insertion<ESZ>(a, n); } __syncwarp(); return; }
	uint32_t *cnt = sm, *head = sm + 256;
	uint32_t sp = 1;
	if(lane == 0) { stack[0] = 0; stack[1] = n | ((esz == 4 ? 7u : 3u) << 29); }
	__syncwarp();
	while(sp > 0) {
		sp--;
		uint32_t beg = stack[2 * sp], w1 = stack[2 * sp + 1];
		uint32_t cntn = w1 & 0x1fffffffu, s = (w1 >> 29) * 8;
		uint32_t *base = a + (uint64_t)esz * beg;
		__syncwarp();
		uint64_t k0 = rs_key<ESZ>(base), diff = 0;
		for(uint32_t i = 1 + lane; i < cntn; i += 32) { diff |= rs_key<ESZ>(base + esz * i) ^ k0; }
		uint32_t dlo = __reduce_or_sync(0xffffffffu, (uint32_t)diff), dhi = __reduce_or_sync(0xffffffffu, (uint32_t)(diff >> 32));
		diff = (uint64_t)dhi << 32 | dlo;
		while(s > 0 && ((diff >> s) & 0xff) == 0) { s -= 8; }
		if(((diff >> s) & 0xff) != 0) {				/* else: s == 0 and uniform, nothing left to order */
			for(int k = lane; k < 256; k += 32) { cnt[k] = 0; }
			__syncwarp();
			/* digit histogram: lanes holding the same digit elect a leader that adds their count (no shared-memory atomics: the
			 * compiler's atomic aggregation code defeats the convergence analysis of the whole kernel) */
			for(uint32_t i0 = 0; i0 < cntn; i0 += 32) {
				uint32_t i = i0 + lane;
				uint32_t d = i < cntn ? (uint32_t)((rs_key<ESZ>(base + esz * i) >> s) & 0xff) : 0x100u + (uint32_t)lane;
				uint32_t m = __match_any_sync(0xffffffffu, d);
				if(d < 0x100u && lane == __ffs((int)m) - 1) { cnt[d] += (uint32_t)__popc(m); }
				__syncwarp();
			}
			/* inclusive prefix sum over 256 counters: 8 per lane + warp scan */
			uint32_t loc[8], sum = 0;
			for(int j = 0; j < 8; j++) { sum += cnt[8 * lane + j]; loc[j] = sum; }
			uint32_t inc = sum;
			for(int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, inc, d); if(lane >= d) { inc += y; } }
			uint32_t excl = inc - sum;
			__syncwarp();
			for(int j = 0; j < 8; j++) { uint32_t e = excl + loc[j]; cnt[8 * lane + j] = e; head[8 * lane + j] = e - (loc[j] - (j ? loc[j - 1] : 0)); }
			__syncwarp();
			if(lane == 0) {
				/* the permutation: cycle leaders in bucket order, exactly as the reference walks them (ksort.h:97-118) */
				for(int k = 0; k < 256;) {
					if(head[k] != cnt[k]) {
						int l = (int)((rs_key<ESZ>(base + esz * head[k]) >> s) & 0xff);
						if(l != k) {
							uint32_t tmp[4] __attribute__((aligned(16))), swp[4] __attribute__((aligned(16)));
							rs_copy<ESZ>(tmp, base + esz * head[k]);
							do {
								rs_copy<ESZ>(swp, tmp); rs_copy<ESZ>(tmp, base + esz * head[l]); rs_copy<ESZ>(base + esz * head[l], swp); head[l]++;
								l = (int)((rs_key<ESZ>(tmp) >> s) & 0xff);
							} while(l != k);
							rs_copy<ESZ>(base + esz * head[k], tmp); head[k]++;
						} else { head[k]++; }
					} else { k++; }
				}
			}
			__syncwarp();
			if(s > 0) {
				uint32_t ns = s - 8;
				for(int j = 0; j < 8; j++) {
					int k = 32 * j + lane;
					uint32_t b0 = k == 0 ? 0 : cnt[k - 1], sz = cnt[k] - b0;
					if(sz > 1 && sz <= 64) { rs_insertion<ESZ>(base + esz * b0, sz); }		/* independent small buckets, one per lane */
					uint32_t m = __ballot_sync(0xffffffffu, sz > 64);
					if(sz > 64) {
						uint32_t pos = sp + (uint32_t)__popc(m & ((1u << lane) - 1));
						if(pos < MAB_RS_STACK) { stack[2 * pos] = beg + b0; stack[2 * pos + 1] = sz | ((ns >> 3) << 29); }
					}
					sp += (uint32_t)__popc(m);
					if(sp > MAB_RS_STACK) { sp = MAB_RS_STACK; *err |= 1u; }				/* warp-uniform */
				}
			}
		}
		__syncwarp();
	}
}

/* ---------------------------------------------------------------- the same sort, parallel form */
/* One distribution pass of the reference's sort moves every element once, but in an order that depends on the data: bucket k's
 * pointer walks its region, an element that belongs elsewhere starts a cycle of swaps that ends with the first element belonging
 * to k (ksort.h:97-118).  What the order of equal keys after the pass depends on is only WHICH element ends up in which slot, and
 * that follows from much less than the elements themselves.  Call an element foreign when it sits in another bucket's region.
 * Then (tests/test_sort_model.py checks this against the literal pass):
 *   - the walk only ever touches foreign elements, region by region in position order: an arrival in bucket l (from a cycle that
 *     started in an earlier bucket) takes the slot at l's pointer, the natives between it and l's next foreign element move up by
 *     one, and that foreign element is the next to travel; in bucket k's own phase a foreign element at the pointer starts a
 *     cycle and the element that closes the cycle lands in its slot, natives stay;
 *   - so the a-th early arrival in l lands at l's begin (a = 1) or one behind l's (a-1)-th foreign slot, a closing element lands
 *     in the slot its cycle started from, a native moves up by one iff it sits in front of the t-th foreign slot of its region,
 *     t = the number of early arrivals (the region's pointer position when its own phase begins).
 * The sequential part is therefore a walk over one byte per foreign element (its destination digit) and a 16-bit counter per
 * bucket: two dependent shared-memory loads per element instead of a 16-byte swap through three, and a footprint of n bytes
 * instead of 16 n, so that 3-5 times as many reads are resident per SM while the elements themselves stay in global memory and
 * are only streamed (histogram, classification, placement: coalesced loads, one scattered 16-byte store per element and pass).
 * Elements ping-pong between two global buffers A (the seed array) and B (the free second half of it); small buckets are
 * insertion-sorted in a shared-memory tile like before and always end in A.
 * sm: 1280 u32 {begin[256], end[256], q[256] = foreign-list start | popped << 16, qe[256] u16, tl[256] u16} + tile (256 elements)
 * + fdig (cap bytes).  fpos / where: 2 x cap u16 of global scratch.  Frames of more than 32767 elements are not taken (caller). */
#define MAB_WK_TILE 256u
#define MAB_WK_SM_WORDS (256u * 3u + 256u + 4u * MAB_WK_TILE)
__device__ __forceinline__ void wk_copy_range(uint32_t *dst, const uint32_t *src, uint32_t cnt, int lane)
{
	for(uint32_t t = lane; t < cnt; t += 32) { ((uint4 *)dst)[t] = ((const uint4 *)src)[t]; }
}
__device__ __forceinline__ void radix_sort_walk_warp(uint32_t *A, uint32_t *Bf, uint32_t n, uint32_t *stack, uint32_t *sm, uint8_t *fdig, uint16_t *fpos, uint16_t *where, int lane, uint32_t *err)
{
	uint32_t *begin = sm, *end = sm + 256, *q = sm + 512;
	uint16_t *qe = (uint16_t *)(sm + 768), *tl = qe + 256;
	uint32_t *tile = sm + 1024;
	if(n <= 64) {
		if(n > 1) {
			wk_copy_range(tile, A, n, lane); __syncwarp();
			if(lane == 0) { rs_insertion<4>(tile, n); }
			__syncwarp(); wk_copy_range(A, tile, n, lane); __syncwarp();
		}
		return;
	}
	uint32_t sp = 1;
	if(lane == 0) { stack[0] = 0; stack[1] = n | (7u << 29); }						/* {beg, cnt | in B << 28 | shift / 8 << 29} */
	__syncwarp();
	while(sp > 0) {
		sp--;
		const uint32_t beg = stack[2 * sp], w1 = stack[2 * sp + 1];
		const uint32_t cnt = w1 & 0xffffu, inb = (w1 >> 28) & 1u;
		uint32_t s = (w1 >> 29) * 8;
		uint32_t *X = (inb ? Bf : A) + 4ull * beg, *Y = (inb ? A : Bf) + 4ull * beg, *Af = A + 4ull * beg;
		__syncwarp();
		uint64_t k0 = rs_key<4>(X), diff = 0;
		#pragma unroll 4
		for(uint32_t i = 1 + lane; i < cnt; i += 32) { diff |= rs_key<4>(X + 4 * i) ^ k0; }
		uint32_t dlo = __reduce_or_sync(0xffffffffu, (uint32_t)diff), dhi = __reduce_or_sync(0xffffffffu, (uint32_t)(diff >> 32));
		diff = (uint64_t)dhi << 32 | dlo;
		while(s > 0 && ((diff >> s) & 0xff) == 0) { s -= 8; }
		if(((diff >> s) & 0xff) == 0) {												/* nothing left to order: final where it is */
			if(inb) { wk_copy_range(Af, X, cnt, lane); }
			__syncwarp();
			continue;
		}
		for(int k = lane; k < 256; k += 32) { end[k] = 0; }
		__syncwarp();
		for(uint32_t i0 = 0; i0 < cnt; i0 += 128) {									/* digit histogram (leader of each digit group adds); the loads of four steps up front */
			uint32_t dd[4];
			#pragma unroll
			for(int u = 0; u < 4; u++) { uint32_t i = i0 + 32u * u + lane; dd[u] = i < cnt ? (uint32_t)((rs_key<4>(X + 4 * i) >> s) & 0xff) : 0x100u + (uint32_t)lane; }
			#pragma unroll
			for(int u = 0; u < 4; u++) {
				if(i0 + 32u * u >= cnt) { break; }
				uint32_t d = dd[u];
				uint32_t m = __match_any_sync(0xffffffffu, d);
				if(d < 0x100u && lane == __ffs((int)m) - 1) { end[d] += (uint32_t)__popc(m); }
				__syncwarp();
			}
		}
		{	/* bucket ends (inclusive prefix sum) and begins */
			uint32_t loc[8], sum = 0;
			for(int j = 0; j < 8; j++) { sum += end[8 * lane + j]; loc[j] = sum; }
			uint32_t inc = sum;
			for(int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, inc, d); if(lane >= d) { inc += y; } }
			uint32_t excl = inc - sum;
			__syncwarp();
			for(int j = 0; j < 8; j++) { uint32_t e = excl + loc[j]; end[8 * lane + j] = e; begin[8 * lane + j] = e - (loc[j] - (j ? loc[j - 1] : 0)); q[8 * lane + j] = 0; qe[8 * lane + j] = 0; tl[8 * lane + j] = 0; }
			__syncwarp();
		}
		/* classification: the region of a position is the first bucket whose end lies behind it; foreign elements are listed in
		 * position order (their destination digit and their position), each region notes where its part of the list starts and ends */
		uint32_t nf = 0;
		for(uint32_t i0 = 0; i0 < cnt; i0 += 128) {
			uint32_t dd[4];
			#pragma unroll
			for(int u = 0; u < 4; u++) { uint32_t i = i0 + 32u * u + lane; dd[u] = i < cnt ? (uint32_t)((rs_key<4>(X + 4 * i) >> s) & 0xff) : 0u; }
			#pragma unroll
			for(int u = 0; u < 4; u++) {
				if(i0 + 32u * u >= cnt) { break; }
				uint32_t i = i0 + 32u * u + lane, d = dd[u], r = 0;
				bool in = i < cnt;
				if(in) {
					uint32_t lo = 0, hi = 255;										/* first r with end[r] > i */
					while(lo < hi) { uint32_t mid = (lo + hi) >> 1; if(end[mid] > i) { hi = mid; } else { lo = mid + 1; } }
					r = lo;
				}
				bool foreign = in && d != r;
				uint32_t fm = __ballot_sync(0xffffffffu, foreign);
				uint32_t j = nf + (uint32_t)__popc(fm & ((1u << lane) - 1u));
				if(foreign) { fdig[j] = (uint8_t)d; fpos[j] = (uint16_t)i; }
				if(in && i == begin[r]) { q[r] = j; }
				if(in && i + 1 == end[r]) { qe[r] = (uint16_t)(j + (foreign ? 1u : 0u)); }
				nf += (uint32_t)__popc(fm);
			}
		}
		__syncwarp();
		if(lane == 0) {																/* the walk */
			for(uint32_t k = 0; k < 256; k++) {
				if(begin[k] == end[k]) { continue; }
				uint32_t wq = q[k]; const uint32_t qs_k = wq & 0xffffu, qe_k = qe[k];
				uint32_t qh_k = wq >> 16;
				tl[k] = (uint16_t)qh_k;
				while(qs_k + qh_k < qe_k) {
					const uint32_t js = qs_k + qh_k; qh_k++;
					uint32_t cur = js;
					for(;;) {
						const uint32_t l = fdig[cur];
						if(l == k) { where[cur] = (uint16_t)(0x8000u | js); break; }
						const uint32_t wl = q[l], j2 = (wl & 0xffffu) + (wl >> 16);
						q[l] = wl + 0x10000u;
						where[cur] = (uint16_t)j2;
						cur = j2;
					}
				}
			}
		}
		__syncwarp();
		/* tl[r] becomes the position in front of which the natives of region r move up by one (0: none do) */
		for(int j = 0; j < 8; j++) { int k = 8 * lane + j; uint32_t t = tl[k]; tl[k] = t ? fpos[(q[k] & 0xffffu) + t - 1] : 0; }
		__syncwarp();
		nf = 0;
		for(uint32_t i0 = 0; i0 < cnt; i0 += 128) {									/* placement */
			uint4 ee[4];
			#pragma unroll
			for(int u = 0; u < 4; u++) { uint32_t i = i0 + 32u * u + lane; if(i < cnt) { ee[u] = ((const uint4 *)X)[i]; } else { ee[u].x = 0; ee[u].y = 0; ee[u].z = 0; ee[u].w = 0; } }
			#pragma unroll
			for(int u = 0; u < 4; u++) {
				if(i0 + 32u * u >= cnt) { break; }
				uint32_t i = i0 + 32u * u + lane, d = 0, r = 0, dst = 0;
				bool in = i < cnt;
				const uint4 e = ee[u];
				if(in) {
					d = (uint32_t)((((uint64_t)e.y << 32 | e.x) >> s) & 0xff);
					uint32_t lo = 0, hi = 255;
					while(lo < hi) { uint32_t mid = (lo + hi) >> 1; if(end[mid] > i) { hi = mid; } else { lo = mid + 1; } }
					r = lo;
				}
				bool foreign = in && d != r;
				uint32_t fm = __ballot_sync(0xffffffffu, foreign);
				uint32_t j = nf + (uint32_t)__popc(fm & ((1u << lane) - 1u));
				nf += (uint32_t)__popc(fm);
				if(foreign) {
					uint32_t w = where[j];
					if(w & 0x8000u) { dst = fpos[w & 0x7fffu]; }
					else { dst = w == (q[d] & 0xffffu) ? begin[d] : (uint32_t)fpos[w - 1] + 1u; }
				} else if(in) { dst = i + (i < (uint32_t)tl[r] ? 1u : 0u); }
				if(in) { ((uint4 *)Y)[dst] = e; }
			}
		}
		__syncwarp();
		if(s == 0) {																	/* buckets of equal keys: final */
			if(!inb) { wk_copy_range(Af, Y, cnt, lane); __syncwarp(); }
			continue;
		}
		/* the buckets: large ones go on the stack (they now live in the other buffer), runs of small ones are insertion-sorted in
		 * the shared-memory tile, one bucket per lane, and written to A */
		const uint32_t ns = s - 8, y_in_b = inb ^ 1u;
		uint32_t k = 0;
		while(k < 256) {
			uint32_t sz = end[k] - begin[k];
			if(sz > 64) {
				if(sp < MAB_RS_STACK) { if(lane == 0) { stack[2 * sp] = beg + begin[k]; stack[2 * sp + 1] = sz | (y_in_b << 28) | ((ns >> 3) << 29); } sp++; }
				else { *err |= 1u; }
				k++;
				continue;
			}
			uint32_t k1 = k, tot = 0; bool work = false;
			while(k1 < 256 && end[k1] - begin[k1] <= 64 && tot + (end[k1] - begin[k1]) <= MAB_WK_TILE) { uint32_t z = end[k1] - begin[k1]; tot += z; work |= z > 1; k1++; }
			if(tot != 0 && (work || y_in_b)) {
				const uint32_t t0 = begin[k];
				if(work) {
					wk_copy_range(tile, Y + 4ull * t0, tot, lane); __syncwarp();
					for(uint32_t kk = k + lane; kk < k1; kk += 32) { uint32_t z = end[kk] - begin[kk]; if(z > 1) { rs_insertion<4>(tile + 4 * (begin[kk] - t0), z); } }
					__syncwarp();
					wk_copy_range(Af + 4ull * t0, tile, tot, lane);
				} else { wk_copy_range(Af + 4ull * t0, Y + 4ull * t0, tot, lane); }
				__syncwarp();
			}
			k = k1;
		}
		__syncwarp();
	}
}

/* ---------------------------------------------------------------- seeds and chaining (minialign.c:3340-3625) */
__device__ __forceinline__ uint32_t u_of(uint32_t x, uint32_t y) { return ((x << 1) - y) + MAB_OFS0; }
__device__ __forceinline__ uint32_t v_of(uint32_t x, uint32_t y) { return ((y << 1) - x) + MAB_OFS0; }
__device__ __forceinline__ int32_t as_of(const uint32_t *s) { return (int32_t)(((s[0] - MAB_OFS0) << 1) + (s[2] - MAB_OFS0)) / 3; }
__device__ __forceinline__ int32_t bs_of(const uint32_t *s) { return (int32_t)(((s[2] - MAB_OFS0) << 1) + (s[0] - MAB_OFS0)) / 3; }

/* mm_expand (minialign.c:3420-3446) for one occurrence */
__device__ __forceinline__ void make_seed(const DevParams &P, uint32_t *s, uint32_t rs, uint32_t rid, uint32_t qs)
{
	uint32_t rmask = 0u - (rid & 1);
	uint32_t _rs = rs + (P.k & rmask), _qs = qs ^ rmask;
	s[0] = u_of(_rs, _qs); s[1] = rid >> 1; s[2] = v_of(_rs, _qs); s[3] = 0x7fffffffu;
}

/* window vector lanes: (uub, rid, vub, vlb); position vector lanes: (upos, rid, vpos, vpos); signed compares
 * (minialign.c:3372-3402) */
struct V4 { int32_t l0, l1, l2, l3; };
__device__ __forceinline__ V4 load_pv(const uint32_t *s) { V4 r; r.l0 = (int32_t)s[0]; r.l1 = (int32_t)s[1]; r.l2 = (int32_t)s[2]; r.l3 = (int32_t)s[2]; return r; }
__device__ __forceinline__ V4 load_wv(const uint32_t *s, uint32_t len)
{
	V4 r = load_pv(s);
	r.l0 = (int32_t)((uint32_t)r.l0 + len); r.l2 = (int32_t)((uint32_t)r.l2 + len);
	return r;
}
__device__ __forceinline__ uint32_t inside_mask(const V4 &w, const V4 &d)
{
	return (d.l0 > w.l0 ? 0x000fu : 0) | (d.l1 > w.l1 ? 0x00f0u : 0) | (d.l2 > w.l2 ? 0x0f00u : 0) | (d.l3 > w.l3 ? 0xf000u : 0);
}
__device__ __forceinline__ bool inside_wv(const V4 &w, const V4 &d) { return inside_mask(w, d) == 0xf000u; }
__device__ __forceinline__ bool inside_uub(const V4 &w, const V4 &d) { return (inside_mask(w, d) & 0xff) == 0; }
__device__ __forceinline__ V4 update_wv(V4 w, const V4 &f)
{
	uint32_t d0 = (uint32_t)w.l0 - (uint32_t)f.l0, d2 = (uint32_t)w.l2 - (uint32_t)f.l2;
	w.l0 = (int32_t)((uint32_t)w.l0 - d2); w.l2 = (int32_t)((uint32_t)w.l2 - d0);
	return w;
}
__device__ __forceinline__ int32_t pdiff_wv(const V4 &w, const V4 &f)
{
	return (int32_t)(((uint32_t)w.l0 - (uint32_t)f.l0) + ((uint32_t)w.l2 - (uint32_t)f.l2));
}

/* mm_chain_seeds (minialign.c:3547-3625).  s = seed array ({upos, rid, vpos, lid} x n_seed, sentinel at n_seed), lf = the
 * array the leaves are appended to behind the sentinel (index n_seed + 1 onwards; the same array as s in the reference --
 * k_sortchain keeps the seeds in shared memory and the rarely touched leaves in global memory), c = root array
 * ({plen, lid}).  Returns #chains, *seed_n = seeds + sentinel + leaves. */
__device__ inline uint32_t chain_seeds(const DevParams &P, uint32_t *s, uint32_t *lfb, uint32_t n_seed, uint32_t *c, uint32_t *seed_n)
{
	uint32_t ncid = 0, nlid = n_seed + 1, nlsid = 0, tsid = n_seed;
	while(nlsid < tsid) {
		uint32_t lid = nlid++;
		uint32_t *lf = lfb + 4ull * lid;						/* leaf: {rsid, rid, lsid, cid} */
		uint32_t lf_lsid = nlsid;
		lf[0] = nlsid; lf[2] = nlsid; lf[1] = s[4ull * nlsid + 1]; lf[3] = 0xffffffffu;
		uint32_t plen = s[4ull * nlsid] + s[4ull * nlsid + 2], scnt = 1;
		uint64_t nrsid = nlsid; nlsid = 0xffffffffu;
		while(1) {
			uint32_t rsid = (uint32_t)nrsid; nrsid = 0;
			V4 wv = load_wv(s + 4ull * rsid, P.twlen);
			for(uint32_t sid = rsid + 1; ; sid++) {
				V4 fv = load_pv(s + 4ull * sid);
				if(!inside_wv(wv, fv)) {
					nlsid = nlsid < sid ? nlsid : sid;
					if(inside_uub(wv, fv)) { continue; }
					break;
				}
				wv = update_wv(wv, fv);
				int64_t di = (int64_t)(((uint64_t)(int64_t)pdiff_wv(wv, fv) << 32) | sid);
				nrsid = (uint64_t)((int64_t)nrsid > di ? (int64_t)nrsid : di);
			}
			if(nrsid == 0) { nrsid = rsid; break; }
			if(s[4ull * (uint32_t)nrsid + 3] != 0x7fffffffu) { nrsid = (uint32_t)nrsid; break; }
			s[4ull * (uint32_t)nrsid + 3] = lid; scnt++;
			if((uint64_t)nlsid <= nrsid) { nlsid = 0xffffffffu; }
		}
		if(nrsid == lf_lsid) { continue; }
		uint32_t cid = 0xffffffffu;
		if(s[4ull * nrsid + 3] < lid) {
			nrsid = lfb[4ull * s[4ull * nrsid + 3] + 0];
			cid = lfb[4ull * s[4ull * nrsid + 3] + 3];
		}
		if(cid == 0xffffffffu) { cid = ncid++; c[2ull * cid] = MAB_OFS0; c[2ull * cid + 1] = lid; }
		lf[3] = cid; lf[0] = (uint32_t)nrsid;
		uint32_t ps = s[4ull * nrsid] + s[4ull * nrsid + 2];
		double frac = __dsub_rn(1.0, __ddiv_rn(1.0, (double)scnt));
		double dl = __dmul_rn(frac, (double)(uint32_t)(ps - plen));
		plen = (uint32_t)((int32_t)MAB_OFS0 - (int32_t)(uint32_t)(int64_t)__double2ll_rz(dl));
		if(plen < c[2ull * cid]) { c[2ull * cid] = plen; c[2ull * cid + 1] = lid; }
	}
	*seed_n = nlid;
	return ncid;
}

/* The same, warp-wide.  What takes the time in mm_chain_seeds is the window scan: from every seed that joins a chain, the seeds
 * behind it are tested against a window that shrinks whenever one of them lies inside it, until one lies beyond its upper u
 * bound.  Here the 32 lanes test 32 candidates at a time against the current window; the events that change state are taken in
 * order out of the ballots (the first candidate inside the window narrows it and the lanes behind it are tested again; a
 * candidate beyond the bound ends the scan), so the result is the sequential scan's.  Everything else is computed redundantly
 * by all lanes (warp-uniform), lane 0 writes. */
__device__ inline uint32_t chain_seeds_warp(const DevParams &P, uint32_t *s, uint32_t *lfb, uint32_t n_seed, uint32_t *c, uint32_t *seed_n, int lane)
{
	uint32_t ncid = 0, nlid = n_seed + 1, nlsid = 0, tsid = n_seed;
	while(nlsid < tsid) {
		uint32_t lid = nlid++;
		uint32_t *lf = lfb + 4ull * lid;						/* leaf: {rsid, rid, lsid, cid} */
		uint32_t lf_lsid = nlsid;
		if(lane == 0) { lf[0] = nlsid; lf[2] = nlsid; lf[1] = s[4ull * nlsid + 1]; lf[3] = 0xffffffffu; }
		__syncwarp();
		uint32_t plen = s[4ull * nlsid] + s[4ull * nlsid + 2], scnt = 1;
		uint64_t nrsid = nlsid; nlsid = 0xffffffffu;
		while(1) {
			uint32_t rsid = (uint32_t)nrsid; nrsid = 0;
			V4 wv = load_wv(s + 4ull * rsid, P.twlen);
			uint32_t sid0 = rsid + 1;
			for(bool done = false; !done; sid0 += 32) {
				uint32_t sid = sid0 + (uint32_t)lane;
				V4 fv = load_pv(s + 4ull * (sid < n_seed ? sid : n_seed));		/* behind the sentinel: the sentinel again (it ends the scan) */
				uint32_t from = 0;
				while(from < 32) {
					uint32_t m = inside_mask(wv, fv);
					bool ins = m == 0xf000u, brk = !ins && (m & 0xffu) != 0;
					uint32_t lm = 0xffffffffu << from;
					uint32_t Im = __ballot_sync(0xffffffffu, ins) & lm, Bm = __ballot_sync(0xffffffffu, brk) & lm, Nm = __ballot_sync(0xffffffffu, !ins) & lm;
					uint32_t fi = Im ? (uint32_t)__ffs((int)Im) - 1u : 32u, fb = Bm ? (uint32_t)__ffs((int)Bm) - 1u : 32u;
					uint32_t upto = fb < fi ? fb + 1u : fi;								/* lanes [from, upto) are passed over in order */
					uint32_t Ns = Nm & (upto >= 32 ? 0xffffffffu : (1u << upto) - 1u);
					if(Ns) { uint32_t f = sid0 + (uint32_t)__ffs((int)Ns) - 1u; nlsid = nlsid < f ? nlsid : f; }
					if(fb < fi) { done = true; break; }
					if(fi == 32) { break; }
					V4 f;
					f.l0 = __shfl_sync(0xffffffffu, fv.l0, (int)fi); f.l1 = __shfl_sync(0xffffffffu, fv.l1, (int)fi); f.l2 = __shfl_sync(0xffffffffu, fv.l2, (int)fi); f.l3 = f.l2;
					wv = update_wv(wv, f);
					int64_t di = (int64_t)(((uint64_t)(int64_t)pdiff_wv(wv, f) << 32) | (sid0 + fi));
					nrsid = (uint64_t)((int64_t)nrsid > di ? (int64_t)nrsid : di);
					from = fi + 1;
				}
			}
			if(nrsid == 0) { nrsid = rsid; break; }
			if(s[4ull * (uint32_t)nrsid + 3] != 0x7fffffffu) { nrsid = (uint32_t)nrsid; break; }
			__syncwarp();
			if(lane == 0) { s[4ull * (uint32_t)nrsid + 3] = lid; }
			__syncwarp();
			scnt++;
			if((uint64_t)nlsid <= nrsid) { nlsid = 0xffffffffu; }
		}
		if(nrsid == lf_lsid) { continue; }
		uint32_t cid = 0xffffffffu;
		__syncwarp();
		if(s[4ull * nrsid + 3] < lid) {
			nrsid = lfb[4ull * s[4ull * nrsid + 3] + 0];
			cid = lfb[4ull * s[4ull * nrsid + 3] + 3];
		}
		uint32_t c0 = 0, c1 = 0;
		bool fresh = cid == 0xffffffffu;
		if(fresh) { cid = ncid++; c0 = MAB_OFS0; c1 = lid; } else { c0 = c[2ull * cid]; c1 = c[2ull * cid + 1]; }
		uint32_t ps = s[4ull * nrsid] + s[4ull * nrsid + 2];
		double frac = __dsub_rn(1.0, __ddiv_rn(1.0, (double)scnt));
		double dl = __dmul_rn(frac, (double)(uint32_t)(ps - plen));
		plen = (uint32_t)((int32_t)MAB_OFS0 - (int32_t)(uint32_t)(int64_t)__double2ll_rz(dl));
		if(plen < c0) { c0 = plen; c1 = lid; }
		__syncwarp();
		if(lane == 0) { lf[3] = cid; lf[0] = (uint32_t)nrsid; c[2ull * cid] = c0; c[2ull * cid + 1] = c1; }
		__syncwarp();
	}
	*seed_n = nlid;
	return ncid;
}

/* ---------------------------------------------------------------- position hash (kh_t, minialign.c:346-613) */
#define MAB_KH_EMPTY	0xffffffffffffffffull
#define MAB_KH_MOVED	0xfffffffffffffffeull
#define MAB_KH_INIT		0xffffffffffffffffull

__device__ inline void kh_reset(uint64_t *kh, ReadRec *r)								/* kh_clear (481-495) */
{
	r->kh_mask = 255; r->kh_cnt = 0; r->kh_ub = 102;									/* 256 * 0.4 */
	for(uint32_t i = 0; i < 256; i++) { kh[2 * i] = MAB_KH_EMPTY; kh[2 * i + 1] = MAB_KH_INIT; }
}

__device__ inline uint64_t kh_poll(const uint64_t *a, uint64_t *pi, uint64_t b0, uint64_t mask)
{
	int64_t b = (int64_t)b0; uint64_t i = *pi, k1;
	while(1) {
		k1 = a[2 * i];
		if(b <= (int64_t)(k1 & mask) + (int64_t)(k1 + 2 < 2)) { break; }
		b -= (int64_t)((i + 1) & (mask + 1));
		i = (i + 1) & mask;
	}
	*pi = i;
	return k1;
}

/* kh_allocate (503-536): returns slot index, *isnew = 1 when a slot was taken */
__device__ inline uint64_t kh_allocate(uint64_t *a, uint64_t k, uint64_t v, uint64_t mask, uint32_t *isnew)
{
	uint64_t i = k & mask, k0 = k, v0 = v;
	uint64_t k1 = kh_poll(a, &i, i, mask);
	if(k0 == k1) { *isnew = 0; return i; }
	uint64_t j = i;
	a[2 * i] = k0;
	while(k1 + 2 >= 2) {
		uint64_t v1 = a[2 * i + 1];
		a[2 * i + 1] = v0;
		k0 = k1; v0 = v1;
		i = (i + 1) & mask;
		k1 = kh_poll(a, &i, k0 & mask, mask);
		a[2 * i] = k0;
	}
	a[2 * i + 1] = v0;
	*isnew = 1;
	return j;
}

__device__ inline int kh_extend(uint64_t *kh, ReadRec *r)								/* 543-579 */
{
	uint64_t prev_size = (uint64_t)r->kh_mask + 1, size = 2 * prev_size, mask = size - 1;
	if(size > MAB_KH_CAP) { return -1; }
	r->kh_mask = (uint32_t)mask; r->kh_ub = (uint32_t)((double)size * 0.4);
	for(uint64_t i = 0; i < prev_size; i++) { kh[2 * (i + prev_size)] = MAB_KH_EMPTY; kh[2 * (i + prev_size) + 1] = MAB_KH_INIT; }
	for(uint64_t i = 0; i < size; i++) {
		uint64_t k = kh[2 * i];
		if(k + 2 < 2 || (k & mask) == i) { continue; }
		uint64_t v = kh[2 * i + 1];
		kh[2 * i] = MAB_KH_MOVED; kh[2 * i + 1] = MAB_KH_INIT;
		uint32_t dummy;
		kh_allocate(kh, k, v, mask, &dummy);
	}
	return 0;
}

/* kh_put_ptr (604-613): returns the slot index of the value */
__device__ inline uint64_t kh_put_ptr(uint64_t *kh, ReadRec *r, uint64_t key, int extend)
{
	if(extend && r->kh_cnt >= r->kh_ub) { if(kh_extend(kh, r) != 0) { r->err |= MAB_ERR_KH_OVF; r->kh_ub = 0xffffffffu; } }
	uint32_t isnew;
	uint64_t idx = kh_allocate(kh, key, MAB_KH_INIT, r->kh_mask, &isnew);
	r->kh_cnt += isnew;
	return idx;
}

__device__ __forceinline__ uint64_t bswap64(uint64_t x)
{
	uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
	return ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
}
__device__ __forceinline__ uint64_t pos_key(uint64_t x, uint64_t y) { return x ^ (x >> 29) ^ y ^ bswap64(y); }		/* minialign.c:3362 */

}  // namespace mab
