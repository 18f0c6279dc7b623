/*
 * mab_post.cuh -- what the reference does with a read's alignments after the extension (mm_align_seq's tail, minialign.c:4450-4473):
 * sort the results by score (radix_sort_64x, unstable), prune by score ratio (mm_prune_regs, 4185-4207), split them into
 * primary / supplementary and secondary by query-span cover (mm_collect_supp, 4214-4263), compute the mapping qualities
 * (mm_post_map, 4270-4325) and fix the output order (mm_pack_reg, 4364-4396) -- on the device, so that the SAM text can be
 * produced there too (mab_sam.cuh) and the host only ever sees text.
 *
 * Exactness.  Everything except one function is integer arithmetic or single IEEE double operations in the reference's order
 * (__dmul_rn / __dadd_rn / __ddiv_rn: no contraction).  The exception is the log10 inside the MAPQ formula
 * `(uint32_t)(-10.0 * 16 * log10(x))` (4175-4177): libm's log10 is not correctly rounded, so no device routine can be proven
 * equal to it.  The device never calls log10: mab_init tabulates, with the host's own libm, the 960 arguments at which the
 * reference's expression steps from one integer to the next (bisection over the double bit patterns, neighbourhoods checked for
 * monotonicity), and the device looks the argument up (binary search).  The result is the host's by construction.
 *
 * One warp per read.  The result record the extend kernel left in the pool ([n_res]{score, n_aln, plen, lb, ub, aln offsets})
 * is followed by scratch (4 words per result) and the plan this kernel writes:
 *   [count][n_uniq] then count x {alignment word offset (lo, hi), rank, mapq} in output order.
 */
#pragma once
#include "mab_pipe.cuh"

namespace mab {

#define MAB_MAPQ_STEPS 960

/* (uint32_t)(-160 * log10(v)) clipped to 960, from the threshold table: thr[k - 1] = largest v that still gives >= k */
__device__ __forceinline__ uint32_t mapq_lookup(const double *thr, double v)
{
	if(!(v > 0.0)) { return 0; }					/* log10 of 0, a negative number or NaN: the conversion of inf / NaN gives 0 on x86-64 */
	if(v > 1.0) { return v >= thr[MAB_MAPQ_STEPS] ? MAB_MAPQ_STEPS : 0; }	/* negative product: wraps to a huge unsigned, clipped (never reached: v <= 1) */
	uint32_t lo = 0, hi = MAB_MAPQ_STEPS;			/* answer in [lo, hi]: the largest k with v <= thr[k - 1] */
	while(lo < hi) {
		uint32_t mid = (lo + hi + 1) >> 1;
		if(v <= thr[mid - 1]) { lo = mid; } else { hi = mid - 1; }
	}
	return lo;
}

/* the scratch / plan area behind a read's result record, 8-byte aligned (the sort moves u32 pairs) */
__device__ __forceinline__ uint32_t *post_scratch(uint32_t *rec, uint32_t result_words)
{
	uint32_t *p = rec + result_words;
	return p + (((uintptr_t)p >> 2) & 1);
}

__device__ __forceinline__ int32_t post_sc(uint32_t x) { return (int32_t)0x40000000 - (int32_t)x; }

__device__ inline void post_read(const DevParams &P, uint32_t *pool, ReadRec *r, const double *thr, uint32_t *frames, uint32_t *sm, int lane)
{
	uint32_t *rec = pool + r->result_ofs;
	uint32_t n_res = rec[0];
	uint32_t *res = post_scratch(rec, r->result_words), *boff = res + 2ull * n_res, *mq = boff + n_res, *plan = mq + n_res;
	if(lane == 0) {
		uint32_t p = 1;
		for(uint32_t i = 0; i < n_res; i++) { boff[i] = p; res[2 * i] = rec[p]; res[2 * i + 1] = i; mq[i] = 0; p += 5 + 2 * rec[p + 1]; }
	}
	__syncwarp();
	uint32_t sort_err = 0;
	radix_sort_exact_warp<2>(res, n_res, frames, sm, lane, &sort_err);								/* minialign.c:4452 */
	__syncwarp();
	if(lane != 0) { return; }
	if(sort_err) { r->err |= MAB_ERR_SEED_OVF; }
	/* mm_prune_regs (4185-4207) */
	uint32_t q = n_res;
	uint32_t minv = (uint32_t)post_sc((uint32_t)__float2ll_rz(__fmul_rn(__ll2float_rn((long long)post_sc(res[0])), P.min_ratio)));
	while(res[2 * --q] > minv) {}
	n_res = q + 1;
	/* mm_collect_supp (4214-4263) */
	#define BIN_LB_(i) rec[boff[res[2 * (i) + 1]] + 3]
	#define BIN_UB_(i) rec[boff[res[2 * (i) + 1]] + 4]
	uint64_t pp, qq;
	for(pp = 1, qq = n_res; pp < qq; pp++) {
		uint64_t mx = 0;
		for(uint64_t i = pp; i < qq; i++) {
			int64_t lb = BIN_LB_(i), ub = BIN_UB_(i), span = ub - lb;
			bool covered = false;
			for(uint64_t j = 0; j < pp; j++) {
				int64_t tlb = BIN_LB_(j), tub = BIN_UB_(j);
				if(tub < ub) { lb = lb > tub ? lb : tub; } else { ub = ub < tlb ? ub : tlb; }
				if(__dmul_rn(1.2, __ll2double_rn(ub - lb)) < __ll2double_rn(span)) {
					qq--;
					uint32_t t0 = res[2 * i], t1 = res[2 * i + 1]; res[2 * i] = res[2 * qq]; res[2 * i + 1] = res[2 * qq + 1]; res[2 * qq] = t0; res[2 * qq + 1] = t1;
					i--; covered = true; break;
				}
			}
			if(covered) { continue; }
			uint64_t cand = ((uint64_t)(2 * (ub - lb) - span) << 32) | i;
			mx = mx > cand ? mx : cand;
		}
		if(mx & 0xffffffffu) {
			uint64_t y = mx & 0xffffffffu;
			uint32_t t0 = res[2 * pp], t1 = res[2 * pp + 1]; res[2 * pp] = res[2 * y]; res[2 * pp + 1] = res[2 * y + 1]; res[2 * y] = t0; res[2 * y + 1] = t1;
		}
	}
	uint64_t n_uniq = pp < qq ? pp : qq;
	/* mm_post_map (4270-4325) */
	int64_t usc = 0, lsc = 0x7fffffffffffffffll, tsc = 0;
	for(uint64_t i = n_uniq; i < n_res; i++) { int64_t s = post_sc(res[2 * i]); usc = usc > s ? usc : s; lsc = lsc < s ? lsc : s; tsc += s; }
	lsc = (lsc == 0x7fffffff) ? 0 : lsc;
	double tpc = 1.0, x = P.xcoef, mxc = __dadd_rn(P.mcoef, P.xcoef);
	for(uint64_t i = 0; i < n_uniq; i++) {
		uint32_t score = (uint32_t)post_sc(res[2 * i]);
		const uint32_t *h = rec + boff[res[2 * i + 1]];
		double pid = 0.0; uint64_t len = 0;
		for(uint32_t j = 0; j < h[1]; j++) {
			const uint32_t *a = pool + ((uint64_t)h[5 + 2 * j] | (uint64_t)h[6 + 2 * j] << 32);
			unsigned long long ib = (unsigned long long)a[2] | (unsigned long long)a[3] << 32;
			double identity; memcpy(&identity, &ib, 8);
			len += a[8]; pid = __dadd_rn(pid, __dmul_rn(__uint2double_rn(a[8]), identity));
		}
		pid = __ddiv_rn(pid, __ll2double_rn((long long)len));
		double ec = __ddiv_rn(2.0, __dsub_rn(__dmul_rn(pid, mxc), x));
		int64_t d = (int64_t)score - usc; d = d > 0 ? d : 0;
		double ulen = __dmul_rn(ec, __ll2double_rn(d)), pe = __ddiv_rn(1.0, __dadd_rn(__dmul_rn(ulen, ulen), 1.0));
		mq[res[2 * i + 1]] = mapq_lookup(thr, pe);
		tpc = __dmul_rn(tpc, __dsub_rn(1.0, pe));
	}
	double tpe = __dsub_rn(1.0, tpc); tpe = tpe < 1.0 ? tpe : 1.0;
	for(uint64_t i = n_uniq; i < n_res; i++) {
		double num = __dmul_rn(tpe, __ll2double_rn((int64_t)res[2 * i] - lsc + 1));
		mq[res[2 * i + 1]] = mapq_lookup(thr, __dsub_rn(1.0, __ddiv_rn(num, __ll2double_rn(tsc))));
	}
	/* mm_pack_reg (4364-4396): output order */
	uint32_t count = 0, nu = 0;
	for(uint64_t i = 0; i < n_res; i++) {
		const uint32_t *h = rec + boff[res[2 * i + 1]];
		for(uint32_t j = 0; j < h[1]; j++) {
			uint32_t *it = plan + 2 + 4ull * count;
			it[0] = h[5 + 2 * j]; it[1] = h[6 + 2 * j]; it[2] = (uint32_t)i; it[3] = mq[res[2 * i + 1]];
			count++;
		}
		if(i == n_uniq - 1) { nu = count; }
	}
	plan[0] = count; plan[1] = nu;
	#undef BIN_LB_
	#undef BIN_UB_
}

/* shared memory: 2 KB of sort scratch per warp */
__global__ void k_post(DevParams P, uint32_t *pool, ReadRec *reads, uint32_t n_reads, const double *thr, uint32_t *frames)
{
	MAB_DYN_SMEM(smem);
	int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	uint32_t *sm = (uint32_t *)smem + 512 * wib;
	uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
	for(uint32_t i = gw; i < n_reads; i += nw) {
		ReadRec *r = &reads[i];
		if(r->result_words == 0 || r->err != 0) { continue; }
		post_read(P, pool, r, thr, frames + (uint64_t)gw * 8 * MAB_RS_FRAME, sm, lane);
		__syncwarp();
	}
}

}  // namespace mab
