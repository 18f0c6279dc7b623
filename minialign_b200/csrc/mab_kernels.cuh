/*
 * mab_kernels.cuh -- the kernels of the mapping path.
 *
 *   k_seed_scan     one warp per read: (w,k)-minimizer sketch (mm_sketch, minialign.c:2410-2435) computed position-parallel,
 *                   ordered ballot compaction into 16-byte minimizer records.
 *   k_seed_probe    one index probe per lane over those records (mm_idx_get, 2727-2748).
 *   k_seed_expand   occurrence expansion to (u,v) seeds and rescue entries (mm_collect_seed / mm_expand, 3420-3493).
 *   k_sort          one warp per read: rescue-round seeding (mm_seed, 3500-3541) and the reference's exact (unstable) radix sort
 *                   in its parallel form (radix_sort_walk_warp, mab_scalar.cuh).
 *   k_chain         one warp per read: array chaining (mm_chain, 3702-3721) and the root sort.
 *   k_sortchain     both in one kernel on a shared-memory copy, the sort by walking the permutation cycles (the earlier form; A/B).
 *   k_extend        persistent warps, one read at a time (longest reads first: `order`, so that the reads still running when the
 *                   work list is empty are the short ones): the mm_extend state machine (4118-4173) on lane 0, the GABA
 *                   fill / search / trace (mab_dp.cuh) on all 32 lanes.
 *   k_extend_pairs  stage-level test entry: one warp per explicit sequence pair.
 */
#pragma once
#include "mab_dp.cuh"
#include "mab_pipe.cuh"

namespace mab {

#ifdef MAB_EMU
#define MAB_DYN_SMEM(name) uint8_t *name = emu::g_blk->smem
#else
#define MAB_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#endif

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t x, int lane, uint32_t *total)
{
	uint32_t s = x;
	for(int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(MAB_FULL, s, d); if(lane >= d) { s += y; } }
	*total = __shfl_sync(MAB_FULL, s, 31);
	return s - x;
}

/* ---------------------------------------------------------------- k_seed_scan / k_seed_expand */
/* k_seed_scan: the sketch pass.  Every minimizer the read emits leaves a 16 B record {hash (lo, hi), qs, 0} in emission order at
 * recs[4 * (seq_ofs + j)] (a read emits at most one minimizer per base, so its slice of the block-sized record array cannot
 * overflow).  k_seed_probe turns the records into {qs, n, occurrence array offset} (one batched probe pass), k_seed_expand
 * replays them (mm_collect_seed / mm_expand, minialign.c:3420-3493): one sketch, one probe per minimizer.
 * shared memory per warp: five 64-entry rings of encoded minimizer candidates (u64): 2560 B */
__global__ void k_seed_scan(DevParams P, const uint8_t *base, ReadRec *reads, uint32_t n_reads, uint32_t *recs)
{
	MAB_DYN_SMEM(smem);
	int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	uint64_t *ring = (uint64_t *)smem + 320 * wib;
	uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
	uint64_t const k = P.k, w = P.w, kk = k - 1, shift1 = 2 * kk, mask = (1ull << (2 * k)) - 1;
	for(uint32_t rid = gw; rid < n_reads; rid += nw) {
		ReadRec *r = &reads[rid];
		uint32_t len = r->len;
		if(len < P.k || (double)len * P.mcoef < (double)P.min_score) {			/* minialign.c:4434 */
			if(lane == 0) { r->state = 1; r->tot_seeds = 0; r->tot_seeds0 = 0; r->tot_resc = 0; r->n_seed = 0; r->n_resc = 0; r->result_words = 0; r->n_rec = 0; }
			continue;
		}
		const uint8_t *seq = base + r->seq_ofs;
		uint32_t *rec = recs + 4ull * r->seq_ofs;
		uint32_t npos = len - (uint32_t)kk;
		uint32_t n_words = 0, n_rec = 0;
		uint64_t vcarry = 0;				/* u: previous window minimum */
		uint32_t idx_carry = (uint32_t)w;	/* decoder state: previous emitted in-block index (v = w), number of block starts */
		uint32_t nblk_carry = 0;
		/* Bit-plane k-mer construction (k <= 16, no N in reach): the two low bits of 64 consecutive base codes sit in ballot
		 * words; lane j cuts its k bits out with a funnel shift, reverses and spreads them to the even / odd bit positions, and
		 * gets the reverse-complement k-mer by complementing, bit-reversing and swapping the bits of each pair back.  Chunks that
		 * see an N (code 4 leaks into the neighbouring base, minialign.c:2383-2398) take the literal per-base loop below.
		 * The window minimum over w positions is min(m_p[j], m_p[j - (w - p)]) with p the largest power of two <= w and m_p built
		 * by doubling; the shifted reads go through per-level 64-entry rings so that they reach into the previous chunk. */
		uint32_t cur = seq[lane], p0c = __ballot_sync(MAB_FULL, cur & 1), p1c = __ballot_sync(MAB_FULL, cur & 2), nc = __ballot_sync(MAB_FULL, cur >= 4), nprev = 0;
		const bool fastk = k <= 16;
		uint32_t c0w = 0;					/* c0 mod w */
		const uint32_t lane_r = (uint32_t)lane % (uint32_t)w;
		int plog = 0; while((2u << plog) <= (uint32_t)w) { plog++; }		/* p = 1 << plog */
		const uint32_t pw = 1u << plog, tailofs = (uint32_t)w - pw;
		for(int t = lane; t < 64 * 5; t += 32) { ring[t] = 0xffffffffffffffffull; }
		__syncwarp();
		for(uint32_t c0 = 0; c0 < npos; c0 += 32) {
			uint32_t j = c0 + lane;
			uint32_t nxt = seq[c0 + 32 + lane];								/* inside the read or its 64 B margin */
			uint32_t p0n = __ballot_sync(MAB_FULL, nxt & 1), p1n = __ballot_sync(MAB_FULL, nxt & 2), nn = __ballot_sync(MAB_FULL, nxt >= 4);
			uint64_t enc = 0xffffffffffffffffull;
			uint32_t jm = c0w + lane_r; if(jm >= (uint32_t)w) { jm -= (uint32_t)w; } if(jm >= (uint32_t)w) { jm -= (uint32_t)w; }	/* j mod w */
			if(fastk && ((nprev >> 31) | nc | (nn & 0xffffu)) == 0) {
				if(j < npos) {
					uint32_t km1 = (1u << k) - 1;
					uint32_t x0 = __funnelshift_r(p0c, p0n, lane) & km1, x1 = __funnelshift_r(p1c, p1n, lane) & km1;		/* bit t = base j + t */
					uint32_t r0 = __brev(x0) >> (32 - (uint32_t)k), r1 = __brev(x1) >> (32 - (uint32_t)k);					/* bit i = base j + k - 1 - i */
					r0 = (r0 | (r0 << 8)) & 0x00ff00ffu; r0 = (r0 | (r0 << 4)) & 0x0f0f0f0fu; r0 = (r0 | (r0 << 2)) & 0x33333333u; r0 = (r0 | (r0 << 1)) & 0x55555555u;
					r1 = (r1 | (r1 << 8)) & 0x00ff00ffu; r1 = (r1 | (r1 << 4)) & 0x0f0f0f0fu; r1 = (r1 | (r1 << 2)) & 0x33333333u; r1 = (r1 | (r1 << 1)) & 0x55555555u;
					uint32_t m32 = (uint32_t)mask;
					uint32_t k0 = r0 | (r1 << 1);
					uint32_t z = __brev(~k0 & m32) >> (32 - 2 * (uint32_t)k);
					uint32_t k1 = ((z >> 1) & 0x55555555u) | ((z & 0x55555555u) << 1);
					uint32_t km = k0 < k1 ? k0 : k1, mm = k0 < k1 ? 0u : 0x80u;
					enc = (uint64_t)km << 8 | (uint64_t)(jm | mm);				/* k <= 16: the CRC term of hash64 is 0 */
				}
			} else if(j < npos) {
				/* k-mer seq[j .. j+kk]; N (code 4) leaks one bit into the neighbouring base exactly like the rolling update */
				uint64_t k0 = 0, k1 = 0;
				for(uint64_t t = 0; t < k; t++) { uint64_t c = seq[j + t]; k0 |= c << (2 * (kk - t)); k1 |= ((3ull ^ c) << shift1) >> (2 * (kk - t)); }
				k0 &= mask;
				if(j > 0) { k1 |= ((3ull ^ (uint64_t)seq[j - 1]) << shift1) >> (2 * k); }
				uint64_t km = k0 < k1 ? k0 : k1, kx = k0 < k1 ? k1 : k0, mm = k0 < k1 ? 0 : 0x80;
				uint64_t h = ((uint64_t)crc32c_u64((uint32_t)kx, kx) ^ km) & mask;
				enc = h << 8 | (uint64_t)jm | mm;
			}
			nprev = nc; p0c = p0n; p1c = p1n; nc = nn;
			c0w += 32 % (uint32_t)w; if(c0w >= (uint32_t)w) { c0w -= (uint32_t)w; }
			/* windowed minimum */
			uint64_t m = enc;
			for(int lv = 0; lv < plog; lv++) {
				uint64_t *rg = ring + 64 * lv;
				rg[j & 63] = m;
				__syncwarp();
				uint64_t o = rg[(j - (1u << lv)) & 63];
				m = o < m ? o : m;
			}
			uint64_t v = m;
			if(tailofs) {
				uint64_t *rg = ring + 64 * 4;
				rg[j & 63] = m;
				__syncwarp();
				uint64_t o = rg[(j - tailofs) & 63];
				v = o < m ? o : m;
			}
			if(j >= npos) { v = 0xffffffffffffffffull; }
			uint64_t u = __shfl_up_sync(MAB_FULL, v, 1);
			if(lane == 0) { u = vcarry; }
			vcarry = __shfl_sync(MAB_FULL, v, 31);
			int emit = j < npos && ((v == enc) || (v != u));
			uint32_t em = __ballot_sync(MAB_FULL, emit);
			n_words += (uint32_t)__popc(em);
			/* decoder of mm_collect_seed (3471-3475): base += (idx <= previous idx) ? w : 0 over the emitted stream */
			uint32_t myidx = (uint32_t)(v & 0x7f);
			uint32_t below = em & ((1u << lane) - 1);
			int pl = below ? 31 - __clz((int)below) : -1;
			uint32_t pidx = __shfl_sync(MAB_FULL, myidx, pl < 0 ? 0 : pl);
			if(pl < 0) { pidx = idx_carry; }
			int newblk = emit && (myidx <= pidx);
			uint32_t nbm = __ballot_sync(MAB_FULL, newblk);
			uint32_t blkcnt = nblk_carry + (uint32_t)__popc(nbm & ((2u << lane) - 1));		/* inclusive */
			if(em) { int last = 31 - __clz((int)em); idx_carry = __shfl_sync(MAB_FULL, myidx, last); }
			nblk_carry += (uint32_t)__popc(nbm);
			/* leave the minimizer for the probe pass: {hash (lo, hi), query position word, 0} in emission order */
			if(emit) {
				uint64_t fr = (v >> 7) & 1, h = v >> 8;
				uint64_t bpos = (uint64_t)(blkcnt - 1) * w + myidx;
				uint32_t qs = (uint32_t)((bpos + (k & (0 - fr))) ^ (0 - fr));
				uint4 e; e.x = (uint32_t)h; e.y = (uint32_t)(h >> 32); e.z = qs; e.w = 0;
				((uint4 *)rec)[n_rec + (uint32_t)__popc(em & ((1u << lane) - 1))] = e;
			}
			n_rec += (uint32_t)__popc(em);
			__syncwarp();
		}
		if(lane == 0) { r->n_words = n_words; r->n_rec = n_rec; }
	}
}

/* k_seed_probe: the index probes of a read, 32 minimizers at a time (mm_idx_get, minialign.c:2727-2748).  Sketching and probing
 * are separate passes on purpose: inside the sketch loop a probe is three dependent loads that only the ~6 emitting lanes of a
 * 32-position step issue and that the whole warp then waits for; here every lane has its own minimizer, a warp keeps 32
 * independent probe chains in flight and nothing else waits for them.  A record becomes {qs, n, byte offset of the occurrence
 * array in the index image (lo, hi)}; n = 0 drops the minimizer (absent, or more frequent than occ[n_occ - 1], 3479); the
 * per-read totals size the workspaces. */
__global__ void k_seed_probe(DevParams P, ReadRec *reads, uint32_t n_reads, uint32_t *recs)
{
	int lane = threadIdx.x & 31;
	uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
	uint32_t const max_occ = P.occ[P.n_occ - 1], resc_occ = P.occ[0];
	for(uint32_t rid = gw; rid < n_reads; rid += nw) {
		ReadRec *r = &reads[rid];
		if(r->state != 0) { continue; }
		uint4 *rec = (uint4 *)(recs + 4ull * r->seq_ofs);
		uint32_t n_rec = r->n_rec, tot_seeds = 0, tot_seeds0 = 0, tot_resc = 0;
		for(uint32_t j = lane; j < n_rec; j += 32) {
			uint4 e = rec[j];
			uint32_t n = 0;
			const uint8_t *occ = idx_get(P, (uint64_t)e.x | (uint64_t)e.y << 32, &n);
			if(n > max_occ) { n = 0; }
			uint64_t ofs = n ? (uint64_t)(occ - P.idx) : 0;
			uint4 o; o.x = e.z; o.y = n; o.z = (uint32_t)ofs; o.w = (uint32_t)(ofs >> 32);
			rec[j] = o;
			tot_seeds += n; tot_seeds0 += n > resc_occ ? 0 : n; tot_resc += n > resc_occ;
		}
		tot_seeds = __reduce_add_sync(MAB_FULL, tot_seeds); tot_seeds0 = __reduce_add_sync(MAB_FULL, tot_seeds0); tot_resc = __reduce_add_sync(MAB_FULL, tot_resc);
		if(lane == 0) { r->tot_seeds = tot_seeds; r->tot_seeds0 = tot_seeds0; r->tot_resc = tot_resc; }
	}
}

__global__ void k_seed_expand(DevParams P, ReadRec *reads, uint32_t n_reads, uint8_t *ws, const uint32_t *recs)
{
	int lane = threadIdx.x & 31;
	uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
	uint32_t const resc_occ = P.occ[0];
	for(uint32_t rid = gw; rid < n_reads; rid += nw) {
		ReadRec *r = &reads[rid];
		if(r->state != 0) { continue; }
		WsLayout L = ws_layout(r->seed_cap, r->root_cap, r->resc_cap, r->bin_cap);
		uint32_t *seeds = (uint32_t *)(ws + r->ws_ofs + L.seed), *resc = (uint32_t *)(ws + r->ws_ofs + L.resc);
		const uint32_t *rec = recs + 4ull * r->seq_ofs;
		uint32_t n_rec = r->n_rec, n_seed = 0, n_resc = 0;
		for(uint32_t c0 = 0; c0 < n_rec; c0 += 32) {
			uint32_t j = c0 + lane;
			uint32_t qs = 0, n = 0; uint64_t ofs = 0;
			if(j < n_rec) { uint4 e = ((const uint4 *)rec)[j]; qs = e.x; n = e.y; ofs = (uint64_t)e.z | (uint64_t)e.w << 32; }
			uint32_t is_resc = n > resc_occ, n_exp = is_resc ? 0 : n;
			uint32_t tot_e, tot_r;
			uint32_t eo = warp_excl_scan(n_exp, lane, &tot_e), ro = warp_excl_scan(is_resc, lane, &tot_r);
			if(n_seed + tot_e + 2 > r->seed_cap || n_resc + tot_r > r->resc_cap) { if(lane == 0) { r->err |= MAB_ERR_SEED_OVF; } break; }
			const uint8_t *occ = P.idx + ofs;
			for(uint32_t i = 0; i < n_exp; i++) {								/* mm_expand (3420-3446) */
				make_seed(P, seeds + 4ull * (n_seed + eo + i), ldg32(occ + 8ull * i), ldg32(occ + 8ull * i + 4), qs);
			}
			if(is_resc) {
				uint32_t *s = resc + 4ull * (n_resc + ro);
				s[0] = qs; s[1] = n; s[2] = (uint32_t)ofs; s[3] = (uint32_t)(ofs >> 32);
			}
			n_seed += tot_e; n_resc += tot_r;
		}
		if(lane == 0) {
			r->n_seed = n_seed; r->seed_n = n_seed; r->n_resc = n_resc; r->presc = 0; r->n_root = 0; r->n_next = 0; r->n_res = 0; r->nbin = 0;
			r->rlen_cur = r->rlen_in; r->rlen_used = r->rlen_in; r->dep_apos = 0; r->dep_flags = 0;
			kh_reset((uint64_t *)(ws + r->ws_ofs + L.kh), r);
		}
	}
}

/* emits the raw sketch words of one read (stage-level test entry for mab_sketch): same math as k_seed, single warp */
__global__ void k_sketch_words(DevParams P, const uint8_t *seq, uint32_t len, uint64_t *out, uint64_t cap, uint64_t *n_out)
{
	MAB_DYN_SMEM(smem);
	int lane = threadIdx.x & 31;
	uint64_t *ring = (uint64_t *)smem;
	uint64_t const k = P.k, w = P.w, kk = k - 1, shift1 = 2 * kk, mask = (1ull << (2 * k)) - 1;
	uint64_t n = 0, vcarry = 0;
	uint32_t npos = len > (uint32_t)kk ? len - (uint32_t)kk : 0;
	for(uint32_t c0 = 0; c0 < npos; c0 += 32) {
		uint32_t j = c0 + lane;
		uint64_t enc = 0xffffffffffffffffull;
		if(j < npos) {
			uint64_t k0 = 0, k1 = 0;
			for(uint64_t t = 0; t < k; t++) { uint64_t c = seq[j + t]; k0 |= c << (2 * (kk - t)); k1 |= ((3ull ^ c) << shift1) >> (2 * (kk - t)); }
			k0 &= mask;
			if(j > 0) { k1 |= ((3ull ^ (uint64_t)seq[j - 1]) << shift1) >> (2 * k); }
			uint64_t km = k0 < k1 ? k0 : k1, kx = k0 < k1 ? k1 : k0, mm = k0 < k1 ? 0 : 0x80;
			uint64_t h = ((uint64_t)crc32c_u64((uint32_t)kx, kx) ^ km) & mask;
			enc = h << 8 | (uint64_t)(j % (uint32_t)w) | mm;
		}
		ring[j & 63] = enc;
		__syncwarp();
		uint64_t v = 0xffffffffffffffffull;
		if(j < npos) {
			uint32_t lo = j + 1 >= (uint32_t)w ? j + 1 - (uint32_t)w : 0;
			for(uint32_t t = lo; t <= j; t++) { uint64_t e = ring[t & 63]; v = e < v ? e : v; }
		}
		uint64_t u = __shfl_up_sync(MAB_FULL, v, 1);
		if(lane == 0) { u = vcarry; }
		vcarry = __shfl_sync(MAB_FULL, v, 31);
		int emit = j < npos && ((v == enc) || (v != u));
		uint32_t em = __ballot_sync(MAB_FULL, emit);
		uint64_t o = n + (uint64_t)__popc(em & ((1u << lane) - 1));
		if(emit && o < cap) { out[o] = v; }
		n += (uint64_t)__popc(em);
		__syncwarp();
	}
	if(lane == 0) {
		/* the cap (minialign.c:2402-2408): only its marker word is ever read on the query path */
		if(n + 4 <= cap) { out[n] = 0xffffffffffff0000ull; out[n + 1] = 0; out[n + 2] = 0; out[n + 3] = 0; }
		*n_out = n + 4;
	}
}

/* ---------------------------------------------------------------- k_sortchain */
/* One WARP per read: the sorts are warp-cooperative (radix_sort_exact_warp), the data-dependent sequential parts (the
 * permutation cycles of the sort, rescue expansion, chaining) run on lane 0.  Those are chains of dependent loads, so the
 * read's seed array (16 B x (n + 1), sentinel included) is staged in shared memory for the duration (STAGED = true: ~30-cycle
 * instead of ~600-cycle steps, and the compiler knows the address space).  Leaves (appended behind the sentinel, touched once
 * each) and the small root / rescue arrays stay in global memory.  Shared memory per warp: 16 B x sc_cap + 2 KB sort scratch. */
template <bool STAGED>
__device__ __forceinline__ void sortchain_read(const DevParams &P, ReadRec *r, uint8_t *ws, uint32_t *fr, uint32_t round, uint32_t *sm, uint32_t *sseed, int lane)
{
	WsLayout L = ws_layout(r->seed_cap, r->root_cap, r->resc_cap, r->bin_cap);
	uint32_t *seed = (uint32_t *)(ws + r->ws_ofs + L.seed), *root = (uint32_t *)(ws + r->ws_ofs + L.root), *resc = (uint32_t *)(ws + r->ws_ofs + L.resc);
	uint32_t n = r->n_seed, sort_err = 0;
	uint32_t *sd = STAGED ? sseed : seed;
	if(STAGED) {
		for(uint32_t t = lane; t < n; t += 32) { ((uint4 *)sseed)[t] = ((const uint4 *)seed)[t]; }
		__syncwarp();
	}
	if(round > 0) {																/* mm_seed, cnt > 0 (3510-3526) */
		if(round == 1) { radix_sort_exact_warp<4>(resc, r->n_resc, fr, sm, lane, &sort_err); }
		for(uint32_t s = lane; s < n; s += 32) { sd[4ull * s + 3] = 0x7fffffffu; }
		__syncwarp();
		if(lane == 0) {
			uint32_t p = r->presc;
			while(p < r->n_resc && resc[4ull * p + 1] <= P.occ[round]) {
				const uint32_t *e = resc + 4ull * p;
				const uint8_t *occ = P.idx + ((uint64_t)e[2] | (uint64_t)e[3] << 32);
				if(n + e[1] + 2 > r->seed_cap / 2) { r->err |= MAB_ERR_SEED_OVF; break; }
				for(uint32_t t = 0; t < e[1]; t++) { make_seed(P, sd + 4ull * (n + t), ldg32(occ + 8ull * t), ldg32(occ + 8ull * t + 4), e[0]); }
				n += e[1]; p++;
			}
			r->presc = p;
		}
		n = __shfl_sync(0xffffffffu, n, 0);
	}
	if(lane == 0) {
		r->n_root = 0; r->n_next = 0; r->n_seed = n;
		if(n == 0) { r->seed_n = 0; }
		else { uint32_t *s = sd + 4ull * n; s[0] = 0x80000000u; s[1] = 0x7fffffffu; s[2] = 0x80000000u; s[3] = 0x7fffffffu; }	/* sentinel (3531) */
	}
	__syncwarp();
	if(n == 0) { return; }
	radix_sort_exact_warp<4>(sd, n + 1, fr, sm, lane, &sort_err);
	uint32_t nc = 0;
	if(lane == 0) {
		uint32_t seed_n = 0;
		nc = chain_seeds(P, sd, seed, n, root, &seed_n);							/* mm_chain (3702-3721); circular refs unsupported */
		r->seed_n = seed_n;
	}
	nc = __shfl_sync(0xffffffffu, nc, 0);
	if(STAGED) {
		__syncwarp();
		for(uint32_t t = lane; t < n + 1; t += 32) { ((uint4 *)seed)[t] = ((const uint4 *)sseed)[t]; }
	}
	if(nc != 0) {
		radix_sort_exact_warp<2>(root, nc, fr, sm, lane, &sort_err);
		if(lane == 0) { r->n_root = nc; }
	}
	if(__any_sync(0xffffffffu, sort_err != 0) && lane == 0) { r->err |= MAB_ERR_SEED_OVF; }
}

/* Handles the reads whose seed bound (round 0: its own seeds; later rounds: all seeds incl. rescued ones; + sentinel) lies in
 * (lo_cap, hi_cap]: the host launches it
 * once per size class so that the many ordinary reads run with a small shared-memory footprint (high occupancy) and the few
 * seed-rich ones with a large one.  hi_cap = UINT32_MAX in the last class; reads above sc_cap work in global memory. */
__global__ void k_sortchain(DevParams P, ReadRec *reads, uint32_t n_reads, uint8_t *ws, uint32_t *frames, uint32_t round, uint32_t sc_cap, uint32_t lo_cap, uint32_t hi_cap)
{
	MAB_DYN_SMEM(smem);
	int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	uint32_t per_warp = 4u * sc_cap + 512u;
	uint32_t *sm = (uint32_t *)smem + (uint64_t)per_warp * wib, *sseed = sm + 512;
	uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if(i >= n_reads) { return; }
	ReadRec *r = &reads[i];
	if(r->state != 0) { return; }
	uint32_t bound = (round == 0 ? r->tot_seeds0 : r->tot_seeds) + 2;
	if(bound <= lo_cap || bound > hi_cap) { return; }
	uint32_t *fr = frames + (uint64_t)i * 8 * MAB_RS_FRAME;
	if(bound <= sc_cap) { sortchain_read<true>(P, r, ws, fr, round, sm, sseed, lane); }
	else { sortchain_read<false>(P, r, ws, fr, round, sm, sseed, lane); }
}

/* ---------------------------------------------------------------- k_sort + k_chain */
/* The two halves of k_sortchain as kernels of their own.  k_sort: rescue-round seeding and the exact sort in its parallel form
 * (radix_sort_walk_warp: elements stay in global memory, shared memory holds a byte per element, so a few dozen reads are
 * resident per SM).  k_chain: chaining + root sort.  One warp per read in both; reads are taken longest first (order[]).
 * k_chain by default works on the sorted array where it lies (sc_cap = 0: 64 warps per SM, the scans are sequential and hit
 * L1): measured against staging it in shared memory (16 B per seed: 2-20 warps per SM) that is 7.4 vs 8.9 ms per chunk on the
 * E.coli-like workload and 23 vs 45 ms on the human-sized one; MAB_CHAIN_STAGED=1 keeps the staged classes for comparison. */
__global__ void k_sort(DevParams P, ReadRec *reads, const uint32_t *order, uint32_t n_reads, uint8_t *ws, uint32_t *frames, uint32_t round, uint32_t cap, uint32_t lo_cap, uint32_t hi_cap)
{
	MAB_DYN_SMEM(smem);
	int lane = threadIdx.x & 31;
	uint32_t *sm = (uint32_t *)smem;
	uint8_t *fdig = (uint8_t *)(sm + MAB_WK_SM_WORDS);
	uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if(slot >= n_reads) { return; }
	uint32_t i = order[slot];
	ReadRec *r = &reads[i];
	if(r->state != 0) { return; }
	uint32_t bound = (round == 0 ? r->tot_seeds0 : r->tot_seeds) + 2;
	if(bound <= lo_cap || bound > hi_cap) { return; }
	uint32_t *fr = frames + (uint64_t)i * 8 * MAB_RS_FRAME;
	WsLayout L = ws_layout(r->seed_cap, r->root_cap, r->resc_cap, r->bin_cap);
	uint32_t *seed = (uint32_t *)(ws + r->ws_ofs + L.seed), *root = (uint32_t *)(ws + r->ws_ofs + L.root), *resc = (uint32_t *)(ws + r->ws_ofs + L.resc);
	uint32_t n = r->n_seed, sort_err = 0;
	if(round > 0) {																/* mm_seed, cnt > 0 (3510-3526) */
		if(round == 1) { radix_sort_exact_warp<4>(resc, r->n_resc, fr, sm, lane, &sort_err); }
		for(uint32_t s = lane; s < n; s += 32) { seed[4ull * s + 3] = 0x7fffffffu; }
		__syncwarp();
		uint32_t p = r->presc; const uint32_t n_resc = r->n_resc;
		bool ovf = false;
		while(p < n_resc && resc[4ull * p + 1] <= P.occ[round]) {					/* warp-uniform; the occurrences of an entry spread over the lanes */
			const uint32_t *e = resc + 4ull * p;
			const uint32_t ne = e[1], qs = e[0];
			const uint8_t *occ = P.idx + ((uint64_t)e[2] | (uint64_t)e[3] << 32);
			if(n + ne + 2 > r->seed_cap / 2) { ovf = true; break; }
			for(uint32_t t = lane; t < ne; t += 32) { make_seed(P, seed + 4ull * (n + t), ldg32(occ + 8ull * t), ldg32(occ + 8ull * t + 4), qs); }
			n += ne; p++;
		}
		__syncwarp();
		if(lane == 0) { r->presc = p; if(ovf) { r->err |= MAB_ERR_SEED_OVF; } }
	}
	if(lane == 0) {
		r->n_root = 0; r->n_next = 0; r->n_seed = n;
		if(n == 0) { r->seed_n = 0; }
		else { uint32_t *s = seed + 4ull * n; s[0] = 0x80000000u; s[1] = 0x7fffffffu; s[2] = 0x80000000u; s[3] = 0x7fffffffu; }	/* sentinel (3531) */
	}
	__syncwarp();
	if(n == 0) { return; }
	if(n + 1 <= 32767u && n + 1 <= cap) {
		uint16_t *fpos = (uint16_t *)root, *where = fpos + ((n + 8u) & ~7u);		/* the root / next arrays are free until the chaining */
		radix_sort_walk_warp(seed, seed + 4ull * (n + 1), n + 1, fr, sm, fdig, fpos, where, lane, &sort_err);
	} else { radix_sort_exact_warp<4>(seed, n + 1, fr, sm, lane, &sort_err); }
	if(__any_sync(0xffffffffu, sort_err != 0) && lane == 0) { r->err |= MAB_ERR_SEED_OVF; }
}

template <bool STAGED>
__device__ __forceinline__ void chain_read(const DevParams &P, ReadRec *r, uint8_t *ws, uint32_t *fr, uint32_t *sm, uint32_t *sseed, int lane, uint32_t wide)
{
	WsLayout L = ws_layout(r->seed_cap, r->root_cap, r->resc_cap, r->bin_cap);
	uint32_t *seed = (uint32_t *)(ws + r->ws_ofs + L.seed), *root = (uint32_t *)(ws + r->ws_ofs + L.root);
	const uint32_t n = r->n_seed;
	uint32_t sort_err = 0;
	if(n == 0) { return; }
	uint32_t *sd = STAGED ? sseed : seed;
	if(STAGED) {
		for(uint32_t t = lane; t < n + 1; t += 32) { ((uint4 *)sseed)[t] = ((const uint4 *)seed)[t]; }
		__syncwarp();
	}
	uint32_t seed_n = 0, nc = 0;													/* mm_chain (3702-3721); circular refs unsupported */
	if(wide) { nc = chain_seeds_warp(P, sd, seed, n, root, &seed_n, lane); }
	else {
		if(lane == 0) { nc = chain_seeds(P, sd, seed, n, root, &seed_n); }
		nc = __shfl_sync(0xffffffffu, nc, 0);
	}
	if(lane == 0) { r->seed_n = seed_n; }
	if(STAGED) {
		__syncwarp();
		for(uint32_t t = lane; t < n + 1; t += 32) { ((uint4 *)seed)[t] = ((const uint4 *)sseed)[t]; }
	}
	__syncwarp();
	if(nc != 0) {
		radix_sort_exact_warp<2>(root, nc, fr, sm, lane, &sort_err);
		if(lane == 0) { r->n_root = nc; }
	}
	if(__any_sync(0xffffffffu, sort_err != 0) && lane == 0) { r->err |= MAB_ERR_SEED_OVF; }
}

__global__ void k_chain(DevParams P, ReadRec *reads, const uint32_t *order, uint32_t n_reads, uint8_t *ws, uint32_t *frames, uint32_t sc_cap, uint32_t lo_cap, uint32_t hi_cap, uint32_t wide)
{
	MAB_DYN_SMEM(smem);
	int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	uint32_t *sm = (uint32_t *)smem + (uint64_t)(512u + 4u * sc_cap) * wib, *sseed = sm + 512;
	uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if(slot >= n_reads) { return; }
	uint32_t i = order[slot];
	ReadRec *r = &reads[i];
	if(r->state != 0) { return; }
	uint32_t bound = r->n_seed + 2;												/* the sorted array, sentinel included */
	if(bound <= lo_cap || bound > hi_cap) { return; }
	uint32_t *fr = frames + (uint64_t)i * 8 * MAB_RS_FRAME;
	if(bound <= sc_cap) { chain_read<true>(P, r, ws, fr, sm, sseed, lane, wide); }
	else { chain_read<false>(P, r, ws, fr, sm, sseed, lane, wide); }
}

/* test kernel: the same array through the cycle-walking sort (a) and its parallel form (b; b holds 2 n elements) */
__global__ void k_sort_check(uint32_t *a, uint32_t *b, uint32_t n, uint32_t *frames, uint16_t *scratch, uint32_t *err_out)
{
	MAB_DYN_SMEM(smem);
	int lane = threadIdx.x & 31;
	uint32_t *sm = (uint32_t *)smem;
	uint32_t err = 0;
	radix_sort_exact_warp<4>(a, n, frames, sm, lane, &err);
	__syncwarp();
	radix_sort_walk_warp(b, b + 4ull * n, n, frames, sm, (uint8_t *)(sm + MAB_WK_SM_WORDS), scratch, scratch + ((n + 8u) & ~7u), lane, &err);
	if(lane == 0) { *err_out = err; }
}

/* ---------------------------------------------------------------- mm_extend state machine, lane-0 routines */
#define MAB_CREM 50000u
#define MAB_SREM 8u
struct Search {
	uint32_t cp[2], tp[2];
	uint32_t aid, bid, iid, eid, sid, rev;
	int64_t prem; uint32_t pacc, crem, srem, narrow, min_score;
	uint32_t rlen, qlen;
};

struct RCtx {				/* per-read pointers (lane 0) */
	ReadRec *r;
	uint32_t *seed, *root, *next;
	uint64_t *kh, *bin;
	uint32_t *pool;
};

#define BIN_NALN(_x, _iid)	( ((uint32_t *)&(_x).bin[_iid])[0] )
#define BIN_PLEN(_x, _iid)	( ((uint32_t *)&(_x).bin[_iid])[1] )
#define BIN_LB(_x, _iid)	( ((uint32_t *)&(_x).bin[(_iid) + 1])[0] )
#define BIN_UB(_x, _iid)	( ((uint32_t *)&(_x).bin[(_iid) + 1])[1] )

__device__ inline uint64_t bin_push(RCtx &x, uint64_t v)
{
	ReadRec *r = x.r;
	if(r->nbin >= r->bin_cap) { r->err |= MAB_ERR_BIN_OVF; return r->nbin - 1; }
	x.bin[r->nbin] = v;
	return r->nbin++;
}

__device__ inline void load_pos(const DevParams &P, const Search &st, const uint32_t *p, uint32_t *rev, uint32_t *cp)	/* 3817-3833 */
{
	int32_t bs = bs_of(p);
	*rev = bs < 0;
	cp[0] = (uint32_t)as_of(p);
	cp[1] = (uint32_t)bs + ((uint32_t)(bs >> 31) & st.qlen);
	if(cp[0] >= st.rlen || cp[1] >= st.qlen) {
		cp[0] -= cp[0] < P.k ? cp[0] : P.k;
		cp[1] -= cp[1] < P.k ? cp[1] : P.k;
	}
}

__device__ inline int load_root(const DevParams &P, RCtx &x, Search &st, uint32_t cid)									/* 3838-3881 */
{
	ReadRec *r = x.r;
	uint32_t lid = x.root[2ull * cid + 1];
	uint32_t plen = (uint32_t)((int32_t)MAB_OFS0 - (int32_t)x.root[2ull * cid]);
	if(__dmul_rn((double)plen, P.mcoef) < __dmul_rn(2.0, (double)P.min_score)) { return 1; }
	r->n_next = 0;
	uint32_t iid = (uint32_t)bin_push(x, 0); bin_push(x, 0);			/* header starts all-zero: see oracle/mm_oracle.c load_root */
	uint32_t eid = r->n_res++;
	x.root[2ull * eid] = MAB_OFS0; x.root[2ull * eid + 1] = iid;
	uint32_t rsid = x.seed[4ull * lid + 0];
	const uint32_t *p = x.seed + 4ull * rsid;
	st.aid = p[1]; st.bid = 0;
	if(!(r->dep_flags & 1)) {											/* first root of this read: rlen is the previous READ's (see ReadRec) */
		if(st.rlen == MAB_RLEN_OWN) { st.rlen = ref_seq(P, st.aid).l_seq; }
		int32_t bs = bs_of(p);
		uint32_t bpos = (uint32_t)bs + ((uint32_t)(bs >> 31) & st.qlen);
		r->dep_apos = (uint32_t)as_of(p); r->dep_flags = 1u | (bpos >= st.qlen ? 2u : 0u); r->rlen_used = st.rlen;
	}
	load_pos(P, st, p, &st.rev, st.cp);									/* reads rlen of the PREVIOUS chain (3865 before 3873) */
	st.tp[0] = st.cp[0]; st.tp[1] = st.cp[1];
	st.iid = iid; st.eid = eid; st.sid = rsid;
	st.prem = plen; st.pacc = 0; st.srem = MAB_SREM; st.narrow = 0;
	st.rlen = ref_seq(P, st.aid).l_seq;
	return 0;
}

/* mm_search_load_next (3887-3944) in two lane-0 halves around the sort of the candidate list, which runs warp-cooperatively
 * (and keeps the single-thread radix sort, whose control flow defeats the compiler's convergence analysis for the whole
 * kernel, out of k_extend).  collect returns the number of candidates to sort (0 = no next seed). */
__device__ inline uint32_t load_next_collect(const DevParams &P, RCtx &x, Search &st)
{
	ReadRec *r = x.r;
	if(st.srem == 0) { return 0; }
	st.srem--;
	const uint32_t *s = x.seed;
	uint32_t *n = x.next;
	uint64_t ncnt = r->n_next, ofs = 2ull * P.tglen;
	uint32_t fa = st.cp[0], fb = st.cp[1] - (st.rev ? st.qlen : 0);
	V4 fv; fv.l0 = (int32_t)u_of(fa, fb); fv.l1 = (int32_t)st.aid; fv.l2 = (int32_t)v_of(fa, fb); fv.l3 = fv.l2;
	uint64_t plim = ofs - st.pacc;
	if(st.pacc > ofs) { ncnt = 0; }
	for(uint64_t i = 0; i < ncnt; i++) {
		if(n[2 * i] >= plim) { ncnt = i; break; }
		n[2 * i] += st.pacc;
	}
	uint64_t sid = st.sid;
	for(uint64_t rcnt = 2ull * st.srem; sid > 0 && rcnt > 0; sid--) {
		V4 wv = load_wv(s + 4ull * (sid - 1), P.tglen), zv = load_wv(s + 4ull * (sid - 1), 128);
		if(!inside_uub(wv, fv)) { break; }
		if(!inside_wv(wv, fv) || inside_wv(zv, fv)) { continue; }
		n[2 * ncnt] = (uint32_t)pdiff_wv(wv, fv); n[2 * ncnt + 1] = (uint32_t)(sid - 1); ncnt++; rcnt--;
	}
	st.sid = (uint32_t)sid;
	r->n_next = (uint32_t)ncnt;
	if(ncnt == 0) { st.pacc = 0; st.srem = 0; }
	return (uint32_t)ncnt;
}
__device__ inline void load_next_pick(const DevParams &P, RCtx &x, Search &st)
{
	ReadRec *r = x.r;
	const uint32_t *n = x.next;
	r->n_next--;
	uint32_t nsid = n[2ull * r->n_next + 1];
	st.pacc = (uint32_t)(2ull * P.tglen - n[2ull * r->n_next]);
	load_pos(P, st, x.seed + 4ull * nsid, &st.rev, st.cp);
}

__device__ inline int test_dup(RCtx &x, Search &st, const PosPair &cp)													/* 3952-3981 */
{
	ReadRec *r = x.r;
	uint64_t k = pos_key((uint64_t)cp.apos | (uint64_t)cp.bpos << 32, (uint64_t)st.aid | (uint64_t)st.bid << 32);
	uint64_t t = kh_put_ptr(x.kh, r, k, 1);
	uint64_t prev = x.kh[2 * t + 1];
	int32_t pa = (int32_t)cp.apos < (int32_t)st.rlen ? (int32_t)cp.apos : (int32_t)st.rlen; pa = pa > 1 ? pa : 1;
	int32_t pb = (int32_t)cp.bpos < (int32_t)st.qlen ? (int32_t)cp.bpos : (int32_t)st.qlen; pb = pb > 1 ? pb : 1;
	st.tp[0] = (uint32_t)pa; st.tp[1] = (uint32_t)pb;
	x.kh[2 * t + 1] = (uint64_t)st.eid | 0xffffffff00000000ull;
	if(prev == MAB_KH_INIT) { return 0; }
	uint32_t eid = (uint32_t)x.kh[2 * t + 1];									/* reads back what was just stored */
	if(eid != st.eid && cp.plen < BIN_PLEN(x, x.root[2ull * eid + 1])) { st.srem = 0; }
	else { st.narrow = st.narrow + 1 < 2 ? st.narrow + 1 : 2; }
	return 1;
}

__device__ inline int record(const DevParams &P, RCtx &x, Search &st, uint64_t aofs)										/* 3986-4067 */
{
	ReadRec *r = x.r;
	const uint32_t *a = x.pool + aofs;
	int64_t score = (int64_t)((uint64_t)a[0] | (uint64_t)a[1] << 32);
	unsigned long long ib = (unsigned long long)a[2] | (unsigned long long)a[3] << 32;
	double identity; memcpy(&identity, &ib, 8);
	uint32_t slen = a[7], plen = a[8], sn = a[10];
	const uint32_t *s0 = a + MAB_ALN_HDR + 8ull * (sn - slen), *sl = a + MAB_ALN_HDR + 8ull * (sn - 1);
	uint32_t p[4] = { st.rlen - (sl[2] + sl[4]), st.qlen - (sl[3] + sl[5]), st.rlen - s0[2], st.qlen - s0[3] };
	st.cp[0] = p[0]; st.cp[1] = p[1];
	st.prem -= plen; st.pacc = plen;
	uint64_t id = (uint64_t)st.aid | (uint64_t)st.bid << 32;
	uint64_t hk = pos_key((uint64_t)p[0] | (uint64_t)p[1] << 32, id), tk = pos_key((uint64_t)p[2] | (uint64_t)p[3] << 32, id);
	uint64_t h = kh_put_ptr(x.kh, r, hk, 1);
	uint64_t t = kh_put_ptr(x.kh, r, tk, 0);
	int isnew = (uint32_t)(x.kh[2 * h + 1] >> 32) == 0xffffffffu;
	uint32_t nid = isnew ? (uint32_t)bin_push(x, aofs) : (uint32_t)(x.kh[2 * h + 1] >> 32);
	uint32_t iid = st.iid;
	uint32_t lb = BIN_LB(x, iid), ub = BIN_UB(x, iid);
	uint32_t ovl = (lb > p[1] ? lb : p[1]) - (ub < p[3] ? ub : p[3]) - p[1] + p[3];
	uint32_t pen = (uint32_t)(int64_t)__double2ll_rz(__dmul_rn((double)(ovl * 2), identity));
	x.root[2ull * st.eid] = (uint32_t)((int64_t)x.root[2ull * st.eid] - (score + (int64_t)pen));
	BIN_NALN(x, iid) += (uint32_t)isnew;
	BIN_PLEN(x, iid) += plen;
	BIN_LB(x, iid) = lb < p[1] ? lb : p[1];
	BIN_UB(x, iid) = ub > p[3] ? ub : p[3];
	if(nid >= r->bin_cap) { r->err |= MAB_ERR_BIN_OVF; nid = iid + 2; }
	const uint32_t *bo = x.pool + x.bin[nid];
	int64_t bscore = (int64_t)((uint64_t)bo[0] | (uint64_t)bo[1] << 32);
	if(bscore > score) {
		x.kh[2 * t + 1] = (uint64_t)st.eid | 0xffffffff00000000ull;
	} else {
		if(x.bin[nid] != aofs) { x.bin[nid] = aofs; }
		x.kh[2 * h + 1] = x.kh[2 * t + 1] = (uint64_t)st.eid | (uint64_t)nid << 32;
	}
	st.srem = MAB_SREM; st.narrow = 0;
	float ms = __fmul_rn(__ll2float_rn(score), P.min_ratio), cur = __uint2float_rn(st.min_score);
	st.min_score = (uint32_t)(int64_t)__float2ll_rz(cur > ms ? cur : ms);
	return (isnew && st.prem > 0) ? 0 : 1;
}

__device__ inline int finish_root(const DevParams &P, RCtx &x, Search &st)													/* 3794-3811 */
{
	ReadRec *r = x.r;
	if(BIN_NALN(x, st.iid) == 0 || x.root[2ull * st.eid] > (uint32_t)((int32_t)MAB_OFS0 - (int32_t)P.min_score)) {
		r->nbin = st.iid; r->n_res--; st.crem--;
	} else {
		st.crem = st.crem != 0 ? MAB_CREM : 0;
	}
	return st.crem == 0;
}

/* write the read's result record into the pool: [n_res] then per result {score, n_aln, plen, lb, ub, n_aln x aln offset (lo,hi)} */
__device__ inline void finalize_read(RCtx &x, BatchCounters *ctr, uint64_t pool_cap)
{
	ReadRec *r = x.r;
	r->state = 1;
	if(r->n_res == 0) { r->result_words = 0; return; }
	uint64_t words = 1, n_aln = 0;
	for(uint32_t i = 0; i < r->n_res; i++) { uint32_t na = BIN_NALN(x, x.root[2ull * i + 1]); words += 5 + 2ull * na; n_aln += na; }
	/* behind the record: room for the post-processing kernel (mab_post.cuh): 4 words of scratch per result and the output plan */
	uint64_t room = words + 1 + 4ull * r->n_res + 2 + 4 * n_aln;			/* + 1: the scratch is 8-byte aligned (sorted as u32 pairs) */
	unsigned long long ofs = atomicAdd(&ctr->pool_top, (unsigned long long)room);
	if(ofs + room > pool_cap) { r->err |= MAB_ERR_POOL_OVF; r->result_words = 0; return; }
	uint32_t *o = x.pool + ofs;
	*o++ = r->n_res;
	for(uint32_t i = 0; i < r->n_res; i++) {
		uint32_t iid = x.root[2ull * i + 1], na = BIN_NALN(x, iid);
		*o++ = x.root[2ull * i]; *o++ = na; *o++ = BIN_PLEN(x, iid); *o++ = BIN_LB(x, iid); *o++ = BIN_UB(x, iid);
		for(uint32_t j = 0; j < na; j++) { uint64_t ao = x.bin[iid + 2 + j]; *o++ = (uint32_t)ao; *o++ = (uint32_t)(ao >> 32); }
	}
	r->result_ofs = ofs; r->result_words = (uint32_t)words;
}

/* per-warp DP arena: [SlotHdr pad 64 B][TailRec x MAB_MAX_TAILS][BlkEntry x blk_cap][masks 2 KB x blk_cap][frames] */
struct ArenaLayout { uint64_t tails, blk, masks, frames, total; };
static inline __host__ __device__ ArenaLayout arena_layout(uint32_t blk_cap)
{
	ArenaLayout a;
	a.tails = 64;
	a.blk = a.tails + sizeof(TailRec) * MAB_MAX_TAILS;
	a.blk = (a.blk + 127) & ~127ull;
	a.masks = a.blk + sizeof(BlkEntry) * (uint64_t)blk_cap;
	a.masks = (a.masks + 127) & ~127ull;
	a.frames = a.masks + 2048ull * blk_cap;
	a.total = (a.frames + 4ull * 8 * MAB_RS_FRAME + 255) & ~255ull;
	return a;
}

__device__ __forceinline__ void dp_ctx_init(DpCtx &c, const DevParams *P, uint8_t *arena, uint32_t blk_cap, const uint32_t *lut, int lane)
{
	ArenaLayout A = arena_layout(blk_cap);
	c.P = P; c.tails = (TailRec *)(arena + A.tails); c.blk = (BlkEntry *)(arena + A.blk); c.masks = (uint32_t *)(arena + A.masks);
	c.blk_cap = blk_cap; c.lut = lut; c.lane = lane; c.nblk = 1; c.ntail = 1; c.W = 64; c.nl = 32; c.widx = 0; c.err = 0; c.n_vectors = 0;
}

__device__ __forceinline__ SecDesc make_sec(const uint8_t *base, uint32_t len, uint32_t id, uint32_t rev)
{
	SecDesc s; s.base = (uint64_t)(uintptr_t)base; s.len = len; s.id = id; s.rev = rev; s._pad = 0;
	return s;
}

/* ---------------------------------------------------------------- k_extend */
/* shared memory per CTA: 1 KB score LUT + per warp two 2.1 KB tiles (traceback mask blocks, double-buffered; the first doubles as sort scratch) */
/* CTAS = resident CTAs per SM the register budget is cut for: 6 (80 registers) when the kernel has the GPU to itself, 4 (128
 * registers, no spills) for the pipelined contexts, which launch 4 per SM anyway and leave the rest of the SM to the other chunks */
template <int CTAS>
__global__ void __launch_bounds__(32 * MAB_WARPS_PER_CTA, CTAS) k_extend(DevParams P, const uint8_t *base, const uint8_t *ntail, ReadRec *reads, const uint32_t *order, uint32_t n_reads, uint8_t *ws,
	uint8_t *arenas, uint64_t arena_stride, uint32_t blk_cap, uint32_t *pool, uint64_t pool_cap, BatchCounters *ctr, uint32_t round, uint32_t last_round)
{
	MAB_DYN_SMEM(smem);
	uint32_t *lut = (uint32_t *)smem;
	/* warp-uniform indices go through a lane-0 broadcast: the compiler then knows that every pointer derived from them (arena,
	 * tile) and every branch on data loaded through them is warp-uniform, and drops the divergence guards around the shuffles */
	int lane = threadIdx.x & 31, wib = __shfl_sync(MAB_FULL, (int)(threadIdx.x >> 5), 0);
	uint32_t *tile = (uint32_t *)smem + 256 + MAB_TILE_WORDS * wib;
	build_lut(P, lut, threadIdx.x, blockDim.x);
	__syncthreads();
	uint32_t gw = __shfl_sync(MAB_FULL, (blockIdx.x * blockDim.x + threadIdx.x) >> 5, 0);
	uint8_t *arena = arenas + arena_stride * gw;
	DpCtx c; dp_ctx_init(c, &P, arena, blk_cap, lut, lane);
	uint32_t *frames = (uint32_t *)(arena + arena_layout(blk_cap).frames);
	SecDesc tsec = make_sec(ntail, 96, 0xfffffffeu, 0);								/* minialign.c:4512-4518 */
	uint64_t n_fill = 0, n_trace = 0;
	while(1) {
		uint32_t rid = 0;
		if(lane == 0) { rid = atomicAdd(&ctr->work_next, 1u); rid = rid < n_reads ? order[rid] : n_reads; }
		rid = __shfl_sync(MAB_FULL, rid, 0);
		if(rid >= n_reads) { break; }
		ReadRec *r = &reads[rid];
		if(r->state != 0) { continue; }
		WsLayout L = ws_layout(r->seed_cap, r->root_cap, r->resc_cap, r->bin_cap);
		RCtx x; x.r = r; x.seed = (uint32_t *)(ws + r->ws_ofs + L.seed); x.root = (uint32_t *)(ws + r->ws_ofs + L.root);
		x.next = (uint32_t *)(ws + r->ws_ofs + L.next); x.kh = (uint64_t *)(ws + r->ws_ofs + L.kh); x.bin = (uint64_t *)(ws + r->ws_ofs + L.bin); x.pool = pool;
		uint32_t n_root = r->n_root, qlen = r->len;
		const uint8_t *qseq = base + r->seq_ofs;
		Search st; memset(&st, 0, sizeof(st));
		st.crem = MAB_CREM; st.min_score = P.min_score; st.qlen = qlen; st.rlen = r->rlen_cur;
		c.err = 0;
		for(uint32_t k = 0; k < n_root && r->seed_n != 0; k++) {
			int stop = 0;
			if(lane == 0) { stop = load_root(P, x, st, k); }
			stop = __shfl_sync(MAB_FULL, stop, 0);
			if(stop) { break; }
			uint32_t aid = __shfl_sync(MAB_FULL, st.aid, 0);
			RefSeq ref = ref_seq(P, aid);
			SecDesc rsec[2] = { make_sec(ref.seq, ref.l_seq, aid << 1, 0), make_sec(ref.seq, ref.l_seq, (aid << 1) + 1, 1) };
			SecDesc qsec[2] = { make_sec(qseq, qlen, 0, 0), make_sec(qseq, qlen, 1, 1) };
			while(1) {
				int go = lane == 0 ? (st.srem > 0 && st.prem > 0) : 0;
				go = __shfl_sync(MAB_FULL, go, 0);
				if(!go) { break; }
				uint32_t rev = __shfl_sync(MAB_FULL, st.rev, 0), narrow = __shfl_sync(MAB_FULL, st.narrow, 0);
				uint32_t cpa = __shfl_sync(MAB_FULL, st.cp[0], 0), cpb = __shfl_sync(MAB_FULL, st.cp[1], 0);
				int adv = 1, brk = 0;
				dp_flush(c, (int)narrow);
				int32_t f = extend_core<false>(c, rsec[0], tsec, qsec[rev], tsec, cpa, cpb);	/* downward */
				n_fill++;
				if(c.err == 0 && c.tails[f].max != 0) {
					PosPair cp = dp_search_max(c, f);
					int dup = 0;
					if(lane == 0) { dup = test_dup(x, st, cp); }
					dup = __shfl_sync(MAB_FULL, dup, 0);
					if(!dup) {
						uint32_t tpa = __shfl_sync(MAB_FULL, st.tp[0], 0), tpb = __shfl_sync(MAB_FULL, st.tp[1], 0);
						dp_flush(c, (int)narrow);												/* the downward blocks are dead: reuse the arena */
						f = extend_core<true>(c, rsec[1], tsec, qsec[1 - rev], tsec, ref.l_seq - tpa, qlen - tpb);	/* upward, traced */
						n_fill++;
						if(c.err == 0 && c.tails[f].max >= (int64_t)P.min_score) {
							uint64_t aofs = dp_trace(c, f, pool, pool_cap, ctr, tile);
							n_trace++;
							if(aofs != 0xffffffffffffffffull) {
								if(lane == 0) { brk = record(P, x, st, aofs); }
								brk = __shfl_sync(MAB_FULL, brk, 0);
								if(brk) { adv = 0; }
							}
						}
					}
				}
				if(c.err) { break; }
				if(brk) { break; }
				if(adv) {
					uint32_t ncnt = 0;
					if(lane == 0) { ncnt = load_next_collect(P, x, st); }
					ncnt = __shfl_sync(MAB_FULL, ncnt, 0);
					if(ncnt > 1) { radix_sort_exact_warp<2>(x.next, ncnt, frames, tile, lane, &c.err); }	/* the trace tile is idle here */
					if(ncnt != 0 && lane == 0) { load_next_pick(P, x, st); }
				}
				__syncwarp();
			}
			if(c.err) { break; }
			int fin = 0;
			if(lane == 0) { fin = finish_root(P, x, st); }
			fin = __shfl_sync(MAB_FULL, fin, 0);
			if(fin) { break; }
		}
		if(lane == 0) {
			r->rlen_cur = st.rlen;
			if(c.err) { r->err |= c.err; }
			if(r->n_res > 0 || round == last_round || r->err) { finalize_read(x, ctr, pool_cap); }
			if(r->err) { atomicOr(&ctr->err_any, r->err); }
		}
		__syncwarp();
	}
	if(lane == 0) {
		atomicAdd(&ctr->n_vectors, (unsigned long long)c.n_vectors);
		atomicAdd(&ctr->n_fill, (unsigned long long)n_fill);
		atomicAdd(&ctr->n_trace, (unsigned long long)n_trace);
	}
}

/* ---------------------------------------------------------------- k_extend_pairs (stage-level parity entry) */
struct PairIn { uint64_t a_ofs, b_ofs; uint32_t alen, blen, apos, bpos, brev, narrow; int64_t min_score; };

__global__ void k_extend_pairs(DevParams P, const uint8_t *base, const uint8_t *ntail, const PairIn *pairs, uint32_t n, uint32_t *res, uint64_t *aln_ofs,
	uint8_t *arenas, uint64_t arena_stride, uint32_t blk_cap, uint32_t *pool, uint64_t pool_cap, BatchCounters *ctr)
{
	MAB_DYN_SMEM(smem);
	uint32_t *lut = (uint32_t *)smem;
	int lane = threadIdx.x & 31, wib = __shfl_sync(MAB_FULL, (int)(threadIdx.x >> 5), 0);
	uint32_t *tile = (uint32_t *)smem + 256 + MAB_TILE_WORDS * wib;
	build_lut(P, lut, threadIdx.x, blockDim.x);
	__syncthreads();
	uint32_t gw = __shfl_sync(MAB_FULL, (blockIdx.x * blockDim.x + threadIdx.x) >> 5, 0), nw = (gridDim.x * blockDim.x) >> 5;
	DpCtx c; dp_ctx_init(c, &P, arenas + arena_stride * gw, blk_cap, lut, lane);
	SecDesc tsec = make_sec(ntail, 96, 0xfffffffeu, 0);
	for(uint32_t i = gw; i < n; i += nw) {
		PairIn pr = pairs[i];
		const uint8_t *a = base + pr.a_ofs, *b = base + pr.b_ofs;
		SecDesc rsec[2] = { make_sec(a, pr.alen, 0, 0), make_sec(a, pr.alen, 1, 1) };
		SecDesc qsec[2] = { make_sec(b, pr.blen, 0, 0), make_sec(b, pr.blen, 1, 1) };
		uint32_t *o = res + 16ull * i;
		uint64_t aofs = 0xffffffffffffffffull;
		c.err = 0;
		if(lane < 16) { o[lane] = 0; }
		__syncwarp();
		dp_flush(c, (int)pr.narrow);
		int32_t f = extend_core<false>(c, rsec[0], tsec, qsec[pr.brev], tsec, pr.apos, pr.bpos);
		const TailRec *t = &c.tails[f];
		if(lane == 0) { o[0] = (uint32_t)t->max; o[1] = (uint32_t)((uint64_t)t->max >> 32); o[2] = t->status; o[3] = (uint32_t)t->apos; o[4] = (uint32_t)t->bpos; }
		if(c.err == 0 && t->max != 0) {
			PosPair cp = dp_search_max(c, f);
			int32_t ta = (int32_t)cp.apos < (int32_t)pr.alen ? (int32_t)cp.apos : (int32_t)pr.alen; ta = ta > 1 ? ta : 1;
			int32_t tb = (int32_t)cp.bpos < (int32_t)pr.blen ? (int32_t)cp.bpos : (int32_t)pr.blen; tb = tb > 1 ? tb : 1;
			if(lane == 0) { o[5] = cp.aid; o[6] = cp.bid; o[7] = cp.apos; o[8] = cp.bpos; o[9] = (uint32_t)cp.plen; o[14] = (uint32_t)ta; o[15] = (uint32_t)tb; }
			dp_flush(c, (int)pr.narrow);
			f = extend_core<true>(c, rsec[1], tsec, qsec[1 - pr.brev], tsec, pr.alen - (uint32_t)ta, pr.blen - (uint32_t)tb);
			t = &c.tails[f];
			if(lane == 0) { o[10] = (uint32_t)t->max; o[11] = (uint32_t)((uint64_t)t->max >> 32); o[12] = t->status; }
			if(c.err == 0 && t->max >= pr.min_score) {
				aofs = dp_trace(c, f, pool, pool_cap, ctr, tile);
				if(lane == 0 && aofs != 0xffffffffffffffffull) { o[13] = 1; }
			}
		}
		if(lane == 0) { aln_ofs[i] = aofs; if(c.err) { atomicOr(&ctr->err_any, c.err); } }
		__syncwarp();
	}
	if(lane == 0) { atomicAdd(&ctr->n_vectors, (unsigned long long)c.n_vectors); }
}


/* ---------------------------------------------------------------- k_fill_peak (integer roofline of the DP step) */
/* The bulk block loop of fill_blocks<MASKS> (the very same bulk_step code) on register-resident synthetic band state, with the
 * same launch shape as k_extend: no sequence fetches, no block bookkeeping, no search / trace, mask rows written to a small
 * per-warp ring that stays in L2.  Its vectors/s is the ceiling the instruction mix of the DP step allows on this GPU; bench.py
 * reports k_extend's vectors/s against it (SURVEY.md section 8d asks for exactly this denominator). */
template <bool MASKS>
__global__ void __launch_bounds__(32 * MAB_WARPS_PER_CTA, MAB_EXT_CTAS_PER_SM) k_fill_peak(DevParams P, uint32_t *ring, uint32_t n_blocks, uint32_t *sink)
{
	MAB_DYN_SMEM(smem);
	uint32_t *lut = (uint32_t *)smem;
	int lane = threadIdx.x & 31;
	build_lut(P, lut, threadIdx.x, blockDim.x);
	__syncthreads();
	uint32_t gw = __shfl_sync(MAB_FULL, (blockIdx.x * blockDim.x + threadIdx.x) >> 5, 0);
	DpCtx c; c.P = &P; c.lut = lut; c.lane = lane; c.W = 64; c.nl = 32; c.widx = 0;
	const StepK k = make_stepk(c);
	const RootTpl &R = P.root[0];
	Vec v;
	v.A = vneg2(unpack8h((uint32_t)(uint8_t)R.dh[2 * lane] | ((uint32_t)(uint8_t)R.dh[2 * lane + 1] << 8)));
	v.V = unpack8h((uint32_t)(uint8_t)R.dv[2 * lane] | ((uint32_t)(uint8_t)R.dv[2 * lane + 1] << 8));
	v.E = unpack8h((uint32_t)(uint8_t)R.de[2 * lane] | ((uint32_t)(uint8_t)R.de[2 * lane + 1] << 8));
	v.F = unpack8h((uint32_t)(uint8_t)R.df[2 * lane] | ((uint32_t)(uint8_t)R.df[2 * lane + 1] << 8));
	v.delta = 0; v.ndrop = 0; v.md = 0; v.wa = 0; v.wb = 0; v.dir = 0;
	v.acc = __reduce_add_sync(MAB_FULL, 0);
	uint32_t x = 0x9e3779b9u * (gw * 32 + lane + 1);
	uint32_t *mrow = ring + 512ull * 4 * gw + lane;
	uint32_t acc = 0;
	for(uint32_t b = 0; b < n_blocks; b++) {
		x = x * 1664525u + 1013904223u;
		uint32_t an = ((x >> 8) & 3u) << 2, bn = ((x >> 16) & 3u) << 20;			/* random bases, pre-scaled like fill_blocks */
		BulkCnt n; n.dir = 0; n.acnt = 0; n.bcnt = 0;
#ifndef MAB_EMU
		asm volatile("" : "+r"(n.dir), "+r"(n.acnt), "+r"(n.bcnt));
#endif
		uint32_t *m = mrow + 512 * (b & 3);
		#pragma unroll 1
		for(int g = 0; g < MAB_BLK / 4; g++) {
			uint32_t b0 = bulk_step<MASKS>(P, k, v, an, bn, n), b1 = bulk_step<MASKS>(P, k, v, an, bn, n);
			if(MASKS) { m[64 * g] = __byte_perm(b0, b1, 0x6420); }
			uint32_t b2 = bulk_step<MASKS>(P, k, v, an, bn, n), b3 = bulk_step<MASKS>(P, k, v, an, bn, n);
			if(MASKS) { m[64 * g + 32] = __byte_perm(b2, b3, 0x6420); }
		}
		acc += __shfl_sync(MAB_FULL, n.dir, 0) + v.delta;
		v.delta = 0; v.acc = (int32_t)(int8_t)v.acc;
	}
	if(lane == 0) { sink[gw] = acc + v.A + v.V + v.E + v.F + v.ndrop; }
}

/* ---------------------------------------------------------------- k_selftest */
/* evaluates every packed-SIMD / permute / warp primitive the DP relies on, on lane-dependent inputs; tests compare the
 * device output word for word with the CUDA-on-CPU shim's, which pins the intrinsics' semantics on real hardware */
__global__ void k_selftest(uint32_t *out)
{
	int lane = threadIdx.x & 31;
	uint32_t x = 0x9e3779b9u * (uint32_t)(lane + 1) ^ 0x7f4a7c15u, y = 0x85ebca6bu * (uint32_t)(lane + 3) ^ 0xc2b2ae35u, z = x * 31u + y;
	uint32_t xs = x & 0x00ff00ffu, ys = y & 0x00ff00ffu;
	int f = 0;
	#define PUT(v) out[32 * (f++) + lane] = (uint32_t)(v)
	PUT(sext8x2(x)); PUT(unpack8(x & 0xffff)); PUT(pack8(x)); PUT(unpack8h(x & 0xffff)); PUT(pack8h(x)); PUT(h8_to_s16(x));
	PUT(pack2(-5)); PUT(pack2h(-5)); PUT(clamp8x2(x)); PUT(__vadd2(x, y)); PUT(__vsub2(x, y)); PUT(__vmaxs2(x, y)); PUT(__vmins2(x, y));
	PUT(__vminu2(x, y)); PUT(__vimax3_s16x2(x, y, z)); PUT(__viaddmax_s16x2(x, y, z)); PUT(__byte_perm(x, y, 0x5432)); PUT(__byte_perm(x, 0, 0x0123));
	PUT(__shfl_up_sync(MAB_FULL, x, 1)); PUT(__shfl_down_sync(MAB_FULL, x, 1)); PUT(__shfl_sync(MAB_FULL, x, (lane * 7 + 3) & 31));
	PUT(__ballot_sync(MAB_FULL, (x >> 5) & 1)); PUT(__reduce_add_sync(MAB_FULL, (int)(x & 0xffff) - 30000)); PUT(__reduce_max_sync(MAB_FULL, (int)x));
	PUT(__popc(x)); PUT(__clz((int)x)); PUT(__ffs((int)x)); PUT(mask_tz(x & 0xff00, y & 0xf0f0));
	PUT(lo16(x)); PUT(hi16(x)); PUT(xs | ys);
	{ unsigned long long w = ((unsigned long long)x << 32) | y; unsigned long long r = __shfl_sync(MAB_FULL, w, (lane + 5) & 31); PUT(r); PUT(r >> 32); }
	{ double d = __dsub_rn(__dmul_rn(__ddiv_rn((double)(int)x, (double)((y & 0xffff) + 1)), 1.0 / 6.0), -4.0 / 6.0); unsigned long long b; memcpy(&b, &d, 8); PUT(b); PUT(b >> 32); }
	{ float fl = __fmul_rn(__ll2float_rn((long long)(int)x), 0.3f); PUT((uint32_t)(int64_t)__float2ll_rz(fl)); }
	PUT((uint32_t)(int64_t)__double2ll_rz((double)(x & 0xffff) * 0.9371));
	{ uint64_t k = pos_key(((uint64_t)x << 32) | y, ((uint64_t)z << 32) | x); PUT(k); PUT(k >> 32); }
	PUT(crc32c_u64(x, ((uint64_t)y << 32) | z));
	PUT(__viaddmin_u16x2(x, y, 0x01000100u)); PUT(__viaddmin_s16x2(x, y, 0x00800080u)); PUT(dp4a_ss(x, y, (int)z)); PUT(dp4a_ss(x, 0xff000000u, 0));
	PUT(prmt(x, y, 0x5444)); PUT(prmt(x, y, 0x0032)); PUT(vneg2(x)); PUT(win_store(win_load(x & 0xffff)));
	PUT(__match_any_sync(MAB_FULL, (x >> 7) & 3)); PUT(__ldg(&out[32 * 0 + ((lane + 1) & 31)]) * 0 + h8_to_s16(x));
	#undef PUT
	if(lane == 0) { out[32 * 63] = (uint32_t)f; }
}

}  // namespace mab

#include "mab_text.cuh"
#include "mab_sam.cuh"
