/*
 * mab_pipe.cuh -- the small kernels that keep a batch resident on the device between the stages of the mapping path, so that
 * the host never has to look at per-read state in the middle of a batch:
 *
 *   k_reads_init    builds the ReadRec array from (offset, length) arrays
 *   k_reads_reset   re-arms every read for a whole-batch re-run (after a buffer had to grow)
 *   k_order         work order of the persistent extend kernel: longest reads first (counting sort on the length)
 *   k_size          per-read workspace capacities and offsets (exclusive scan) from the totals k_seed_scan left behind
 *   k_rlen_predict  / k_rlen_verify: the reference's per-thread `rlen` carried from read to read (minialign.c:3865 runs before
 *                   3873, see ReadRec): predicted from the sorted chain lists before the extension, verified after it
 *
 * All single-CTA kernels: the arrays are a few thousand to a few million entries, the work per entry is a handful of loads.
 */
#pragma once
#include "mab_scalar.cuh"

namespace mab {

#define MAB_PIPE_THREADS 1024

/* block-wide exclusive sum / "latest non-zero" scans over one value per thread (blockDim.x = multiple of 32, <= 1024);
 * sm = 34 u64 of shared memory */
__device__ __forceinline__ uint64_t block_excl_sum(uint64_t x, uint64_t *sm, uint64_t *total)
{
	int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
	uint64_t s = x;
	for(int d = 1; d < 32; d <<= 1) { uint64_t y = __shfl_up_sync(0xffffffffu, s, d); if(lane >= d) { s += y; } }
	if(lane == 31) { sm[wid] = s; }
	__syncthreads();
	if(wid == 0) {
		uint64_t v = lane < nw ? sm[lane] : 0, t = v;
		for(int d = 1; d < 32; d <<= 1) { uint64_t y = __shfl_up_sync(0xffffffffu, t, d); if(lane >= d) { t += y; } }
		sm[lane] = t - v;
		if(lane == 31) { sm[32] = t; }
	}
	__syncthreads();
	uint64_t r = sm[wid] + s - x;
	*total = sm[32];
	__syncthreads();
	return r;
}
/* x = 0 for "nothing", otherwise (position + 1) << 32 | payload: the exclusive maximum is the latest entry before this thread */
__device__ __forceinline__ uint64_t block_excl_max(uint64_t x, uint64_t *sm, uint64_t *total)
{
	int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
	uint64_t s = x;
	for(int d = 1; d < 32; d <<= 1) { uint64_t y = __shfl_up_sync(0xffffffffu, s, d); if(lane >= d && y > s) { s = y; } }
	uint64_t e = __shfl_up_sync(0xffffffffu, s, 1);
	if(lane == 0) { e = 0; }
	if(lane == 31) { sm[wid] = s; }
	__syncthreads();
	if(wid == 0) {
		uint64_t t = lane < nw ? sm[lane] : 0;
		for(int d = 1; d < 32; d <<= 1) { uint64_t y = __shfl_up_sync(0xffffffffu, t, d); if(lane >= d && y > t) { t = y; } }
		uint64_t te = __shfl_up_sync(0xffffffffu, t, 1);
		sm[lane] = lane == 0 ? 0 : te;
		if(lane == 31) { sm[32] = t; }
	}
	__syncthreads();
	uint64_t r = sm[wid] > e ? sm[wid] : e;
	*total = sm[32];
	__syncthreads();
	return r;
}

__global__ void k_reads_init(ReadRec *reads, uint32_t n, const uint64_t *seq_ofs, const uint32_t *seq_len)
{
	for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		ReadRec r; memset(&r, 0, sizeof(r));
		r.seq_ofs = seq_ofs[i]; r.len = seq_len[i]; r.rlen_in = MAB_RLEN_OWN;
		reads[i] = r;
	}
}

__global__ void k_reads_reset(ReadRec *reads, uint32_t n)
{
	for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		ReadRec r; memset(&r, 0, sizeof(r));
		r.seq_ofs = reads[i].seq_ofs; r.len = reads[i].len; r.rlen_in = MAB_RLEN_OWN;
		reads[i] = r;
	}
}

/* order[] = read indices by descending length, 32-base granularity (the order only schedules the work: any order gives the same
 * results; longest-first shortens the tail of the persistent kernel) */
#define MAB_ORDER_BINS 2048
__global__ void k_order(const ReadRec *reads, uint32_t n, uint32_t *order)
{
	__shared__ uint32_t bin[MAB_ORDER_BINS];
	__shared__ uint64_t sm[34];
	for(uint32_t b = threadIdx.x; b < MAB_ORDER_BINS; b += blockDim.x) { bin[b] = 0; }
	__syncthreads();
	for(uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
		uint32_t k = reads[i].len >> 5; k = k < MAB_ORDER_BINS ? k : MAB_ORDER_BINS - 1;
		atomicAdd(&bin[MAB_ORDER_BINS - 1 - k], 1u);
	}
	__syncthreads();
	uint64_t carry = 0;
	for(uint32_t b0 = 0; b0 < MAB_ORDER_BINS; b0 += blockDim.x) {
		uint32_t b = b0 + threadIdx.x;
		uint64_t c = b < MAB_ORDER_BINS ? bin[b] : 0, tot;
		uint64_t e = block_excl_sum(c, sm, &tot);
		if(b < MAB_ORDER_BINS) { bin[b] = (uint32_t)(carry + e); }
		carry += tot;
	}
	__syncthreads();
	for(uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
		uint32_t k = reads[i].len >> 5; k = k < MAB_ORDER_BINS ? k : MAB_ORDER_BINS - 1;
		order[atomicAdd(&bin[MAB_ORDER_BINS - 1 - k], 1u)] = i;
	}
}

/* workspace capacities and offsets of the active reads from the count pass; reads that do not fit into ws_cap are finished
 * with MAB_ERR_WS_OVF (the host grows the buffer to ctr->ws_need and re-runs the batch) */
__global__ void k_size(ReadRec *reads, uint32_t n, uint64_t ws_cap, BatchCounters *ctr)
{
	__shared__ uint64_t sm[34];
	uint64_t carry = 0;
	for(uint32_t i0 = 0; i0 < n; i0 += blockDim.x) {
		uint32_t i = i0 + threadIdx.x;
		uint64_t sz = 0;
		ReadRec *r = i < n ? &reads[i] : nullptr;
		uint32_t seed_cap = 0, root_cap = 0, resc_cap = 0, bin_cap = 0;
		if(r != nullptr && r->state == 0) {
			seed_cap = 2 * (r->tot_seeds + 1) + 8; root_cap = r->tot_seeds + 8; resc_cap = r->tot_resc + 4; bin_cap = 2 * r->tot_seeds + 128;
			sz = ws_layout(seed_cap, root_cap, resc_cap, bin_cap).total;
		}
		uint64_t tot, e = carry + block_excl_sum(sz, sm, &tot);
		if(sz != 0) {
			r->seed_cap = seed_cap; r->root_cap = root_cap; r->resc_cap = resc_cap; r->bin_cap = bin_cap; r->ws_ofs = e;
			if(e + sz > ws_cap) { r->state = 1; r->err |= MAB_ERR_WS_OVF; r->result_words = 0; r->ws_ofs = 0; }
		}
		carry += tot;
	}
	if(threadIdx.x == 0) { ctr->ws_need = carry; if(carry > ws_cap) { atomicOr(&ctr->err_any, MAB_ERR_WS_OVF); } }
}

/* Before the round-0 extension: the value the reference thread's `rlen` will have when read i starts, assuming every read before
 * it is done after round 0.  A read loads the chains of its sorted root list one after the other until one is too short
 * (mm_search_load_root, minialign.c:3838-3848) and leaves the reference length of the last one it loaded behind; which chains
 * those are is known from the list alone.  Reads that only load a chain in a rescue round, exhausted chain budgets and the first
 * reads of the batch (their predecessor is in the previous batch) are what k_rlen_verify is for. */
__global__ void k_rlen_predict(DevParams P, ReadRec *reads, uint32_t n, const uint8_t *ws, uint32_t rlen_init, uint32_t init_known)
{
	__shared__ uint64_t sm[34];
	uint64_t carry = 0;
	for(uint32_t i0 = 0; i0 < n; i0 += blockDim.x) {
		uint32_t i = i0 + threadIdx.x;
		uint64_t x = 0;
		ReadRec *r = i < n ? &reads[i] : nullptr;
		bool active = r != nullptr && r->state == 0;
		if(active && r->n_root > 0 && r->seed_n != 0) {
			WsLayout L = ws_layout(r->seed_cap, r->root_cap, r->resc_cap, r->bin_cap);
			const uint32_t *seed = (const uint32_t *)(ws + r->ws_ofs + L.seed), *root = (const uint32_t *)(ws + r->ws_ofs + L.root);
			uint32_t m = 0;
			while(m < r->n_root) {
				uint32_t plen = (uint32_t)((int32_t)MAB_OFS0 - (int32_t)root[2ull * m]);
				if(__dmul_rn((double)plen, P.mcoef) < __dmul_rn(2.0, (double)P.min_score)) { break; }
				m++;
			}
			if(m > 0) {
				uint32_t lid = root[2ull * (m - 1) + 1], rsid = seed[4ull * lid], aid = seed[4ull * rsid + 1];
				x = ((uint64_t)(i + 1) << 32) | ref_seq(P, aid).l_seq;
			}
		}
		uint64_t tot, e = block_excl_max(x, sm, &tot);
		if(carry > e) { e = carry; }
		if(active) { uint32_t v = e != 0 ? (uint32_t)e : (init_known ? rlen_init : MAB_RLEN_OWN); r->rlen_in = v; r->rlen_cur = v; r->rlen_used = v; }
		if(tot > carry) { carry = tot; }
	}
}

/* After the last round: walk the reads in order with the `rlen` each one really left behind.  A read whose first root test
 * (recorded by load_root) would have gone the other way with the true value is re-armed with it (state = 0) for a redo pass;
 * the host repeats pass + verification until nothing is left (a redone read may leave a different value behind).
 * init_known = 0: the value left by the previous batch is not known yet; the first chain-loading read is reported in ctr->fd_*
 * and checked by a later call. */
__global__ void k_rlen_verify(ReadRec *reads, uint32_t n, uint32_t rlen_init, uint32_t init_known, BatchCounters *ctr)
{
	__shared__ uint64_t sm[34];
	uint64_t carry = 0;
	uint32_t n_redo = 0;
	if(threadIdx.x == 0) { ctr->fd_valid = 0; }
	__syncthreads();
	for(uint32_t i0 = 0; i0 < n; i0 += blockDim.x) {
		uint32_t i = i0 + threadIdx.x;
		ReadRec *r = i < n ? &reads[i] : nullptr;
		bool dep = r != nullptr && (r->dep_flags & 1) != 0;
		uint64_t x = dep ? ((uint64_t)(i + 1) << 32) | r->rlen_cur : 0;
		uint64_t tot, e = block_excl_max(x, sm, &tot);
		if(carry > e) { e = carry; }
		if(dep) {
			if(e == 0 && !init_known) {
				ctr->fd_valid = 1; ctr->fd_idx = i; ctr->fd_apos = r->dep_apos; ctr->fd_flags = r->dep_flags; ctr->fd_used = r->rlen_used;
			} else {
				uint32_t prev = e != 0 ? (uint32_t)e : rlen_init;
				bool used = (r->dep_apos >= r->rlen_used) || (r->dep_flags & 2), actual = (r->dep_apos >= prev) || (r->dep_flags & 2);
				if(used != actual) {
					r->state = 0; r->rlen_in = prev; r->err = 0; r->result_words = 0; r->n_res = 0; r->nbin = 0;
					n_redo++;
				}
			}
		}
		if(tot > carry) { carry = tot; }
	}
	if(n_redo) { atomicAdd(&ctr->n_redo, n_redo); }
	if(threadIdx.x == 0) { ctr->chain_valid = carry != 0; ctr->chain_rlen = (uint32_t)carry; }
}

}  // namespace mab
