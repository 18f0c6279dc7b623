/*
 * mab_text_host.inl -- host driver of the text path (mab_text_begin / commit / finish, include/minialign_b200.h): FASTA / FASTQ
 * bytes in, SAM bytes out.  Included at the end of mab_host.inl.  The host's part is buffer management and three waits per
 * chunk: after the parse (it has to know the number of reads, their total and longest length to size the buffers), after the
 * mapping (counters, rlen chain) and after the SAM size pass (bytes to copy back).
 */

/* ---- MAPQ step table (mab_post.cuh): the reference's expression evaluated with this host's libm ---- */
namespace {
inline uint32_t host_mapq(double v) { double x = -10.0 * 16 * log10(v); uint32_t r = (uint32_t)x; return std::min(r, 60u * 16); }	/* _clip(-10.0 * MAPQ_COEF * log10(v)), minialign.c:4175-4177 */
inline double bits_d(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }
inline uint64_t d_bits(double d) { uint64_t b; memcpy(&b, &d, 8); return b; }

/* thr[k - 1] = the largest double v in (0, 1] with host_mapq(v) >= k, k = 1..960; thr[960] = the smallest v > 1 with
 * host_mapq(v) != 0 (negative product wrapping around; unreachable, kept for completeness).  Returns false when the step
 * function is not monotone around a threshold (then no table can stand in for log10 and the text path is refused). */
bool build_mapq_table(double *thr)
{
	for(uint32_t k = 1; k <= MAB_MAPQ_STEPS; k++) {
		uint64_t lo = d_bits(4.9e-324), hi = d_bits(1.0);						/* positive doubles order like their bit patterns */
		if(host_mapq(bits_d(lo)) < k) { return false; }
		while(lo < hi) {															/* largest pattern with mapq >= k */
			uint64_t mid = lo + (hi - lo + 1) / 2;
			if(host_mapq(bits_d(mid)) >= k) { lo = mid; } else { hi = mid - 1; }
		}
		for(int d = -16; d <= 16; d++) {											/* the step must be clean in the neighbourhood */
			uint64_t b = lo + (uint64_t)(int64_t)d;
			if(b > d_bits(1.0) || b < 1) { continue; }
			bool ge = host_mapq(bits_d(b)) >= k;
			if(ge != (d <= 0)) { return false; }
		}
		thr[k - 1] = bits_d(lo);
	}
	{
		uint64_t lo = d_bits(1.0), hi = d_bits(2.0);								/* smallest pattern > 1.0 whose product truncates to <= -1 */
		while(lo < hi) { uint64_t mid = lo + (hi - lo) / 2; if(host_mapq(bits_d(mid)) != 0) { hi = mid; } else { lo = mid + 1; } }
		thr[MAB_MAPQ_STEPS] = bits_d(lo);
	}
	return true;
}
}  // namespace

static int text_init(mab_ctx *ctx)
{
	std::vector<double> thr(MAB_MAPQ_STEPS + 1);
	ctx->thr_ok = build_mapq_table(thr.data());
	CK(RT_MALLOC(&ctx->d_thr, 8 * (MAB_MAPQ_STEPS + 1)));
	CK(RT_MEMCPY_H2D(ctx->d_thr, thr.data(), 8 * (MAB_MAPQ_STEPS + 1)));
	CK(RT_MALLOC(&ctx->d_tc, sizeof(TextCounters)));
	return MAB_OK;
}

static void text_destroy(mab_ctx *ctx)
{
	RT_FREE(ctx->d_text); RT_FREE(ctx->d_base); RT_FREE(ctx->d_marks); RT_FREE(ctx->d_tiles); RT_FREE(ctx->d_trec); RT_FREE(ctx->d_tc); RT_FREE(ctx->d_thr); RT_FREE(ctx->d_sam);
	RT_HOST_FREE(ctx->h_sam);
}

/* page-locks memory the caller allocated (and, better, already touched: populating pages is the slow part of pinning and can be
 * done on several threads without holding up the CUDA calls of the rest of the process, which cudaHostAlloc does) */
extern "C" int mab_host_register(int device, void *p, uint64_t bytes)
{
	if(!RT_OK(RT_USE_DEVICE(device))) { g_err = std::string("no usable CUDA device: ") + RT_ERRSTR(); return MAB_ENODEV; }
	if(!RT_OK(RT_HOST_REGISTER(p, bytes))) { g_err = std::string("page-locking host memory failed: ") + RT_ERRSTR(); return MAB_ENOMEM; }
	return MAB_OK;
}
extern "C" void mab_host_unregister(void *p) { RT_HOST_UNREGISTER(p); }
extern "C" void *mab_host_alloc_on(int device, uint64_t bytes)
{
	if(!RT_OK(RT_USE_DEVICE(device))) { g_err = std::string("no usable CUDA device: ") + RT_ERRSTR(); return nullptr; }
	return mab_host_alloc(bytes);
}
extern "C" void *mab_host_alloc(uint64_t bytes) { void *p = nullptr; if(!RT_OK(RT_HOST_ALLOC(&p, bytes))) { g_err = std::string("pinned host allocation failed: ") + RT_ERRSTR(); return nullptr; } return p; }
extern "C" void mab_host_free(void *p) { RT_HOST_FREE(p); }

extern "C" uint64_t mab_sam_header_text(const mab_ctx *ctx, const char *version, const char *cmdline, char *out, uint64_t cap)
{
	std::string h = "@HD\tVN:1.0\tSO:unsorted\n";
	for(uint32_t i = 0; i < ctx->P.n_ref; i++) {
		const char *name; uint32_t l_name, l_seq;
		mab_ref_info(ctx, i, &name, &l_name, &l_seq, nullptr);
		h += "@SQ\tSN:"; h.append(name, l_name); h += "\tLN:" + std::to_string(l_seq) + "\n";
	}
	h += std::string("@PG\tID:minialign\tPN:minialign\tVN:") + version + "\tCL:" + cmdline + "\n";
	if(out != nullptr && cap > 0) { memcpy(out, h.data(), std::min<uint64_t>(cap, h.size())); }
	return h.size();
}

static void text_fill_info(const mab_ctx *ctx, mab_text_info_t *info)
{
	if(info == nullptr) { return; }
	info->n_reads = ctx->tx.n_kept; info->n_bases = ctx->tx.tc.tot_len; info->sam_bytes = ctx->tx.sam_total;
	info->rlen_valid = ctx->hc.chain_valid; info->rlen_next = ctx->hc.chain_rlen;
}

extern "C" int mab_text_reserve(mab_ctx *ctx, uint64_t max_chunk_bytes)
{
	if(ctx == nullptr || max_chunk_bytes >= 0xfffffff0ull) { g_err = "mab_text_reserve: bad arguments (chunks are limited to 4 GiB)"; return MAB_EINVAL; }
	std::lock_guard<std::mutex> lk(ctx->cal->mu);
	ctx->cal->reserve_bytes = max_chunk_bytes;
	return MAB_OK;
}

extern "C" int mab_text_begin(mab_ctx *ctx, const char *text, uint64_t n_bytes, uint32_t flags, uint32_t rlen_prev, int rlen_known, mab_text_info_t *info)
{
	mab_stats_t &S = ctx->stats;
	memset(&S, 0, sizeof(S));
	const double t_call = RT_WALL_MS();
	ctx->tx.stage = 0; ctx->tx.n_rec = 0; ctx->tx.n_kept = 0; ctx->tx.sam_total = 0; ctx->tx.flags = flags; ctx->tx.rlen_known = rlen_known != 0;
	memset(&ctx->tx.tc, 0, sizeof(TextCounters)); memset(&ctx->hc, 0, sizeof(BatchCounters));
	if(!ctx->thr_ok) { g_err = "text path unavailable: this host's log10 is not monotone around a MAPQ step"; return MAB_EINVAL; }
	if(n_bytes >= 0xfffffff0ull) { g_err = "mab_text_begin: chunks are limited to 4 GiB"; return MAB_EINVAL; }
	CK(RT_USE_DEVICE(ctx->device));
	trace_line(ctx, "begin: enter, MB", n_bytes / 1048576.0);
	/* buffer sizes follow the largest chunk the caller announced (mab_text_reserve) when this one is a fair sample of it; a much
	 * shorter chunk (the tail of a file) is sized as it is */
	double reserve = 1.0; uint64_t mark_hw, rec_hw;
	{
		std::lock_guard<std::mutex> lk(ctx->cal->mu);
		const uint64_t rb = ctx->cal->reserve_bytes;
		if(n_bytes != 0 && rb > n_bytes && rb <= 2 * n_bytes) { reserve = (double)rb / (double)n_bytes; }
		mark_hw = ctx->cal->mark_hw; rec_hw = ctx->cal->rec_hw;
	}
	ctx->tx.reserve = reserve;
	RT_EVENT_RECORD(ctx->ev[0], ctx->stream);
	if(n_bytes == 0) { ctx->tx.stage = 2; text_fill_info(ctx, info); return MAB_OK; }
	/* the chunk: on the device, closed by a newline, padded with newlines for the 16-byte loads of the scanner */
	const uint8_t *d_text;
	uint8_t first, last;
	if(ctx->device_input) {
		/* caller contract in this mode: the device buffer ends with '\n' and has 64 readable bytes behind it */
		d_text = (const uint8_t *)text;
		if(((uintptr_t)text & 15) != 0) { g_err = "mab_text_begin: a device-resident chunk must be 16-byte aligned"; return MAB_EINVAL; }
		uint8_t fl[2];
		CK(RT_MEMCPY_D2H_ASYNC(&fl[0], d_text, 1, ctx->stream)); CK(RT_MEMCPY_D2H_ASYNC(&fl[1], d_text + n_bytes - 1, 1, ctx->stream));
		CK(ctx_sync(ctx));
		first = fl[0]; last = fl[1];
		if(last != '\n') { g_err = "mab_text_begin: a device-resident chunk must end with a newline"; return MAB_EINVAL; }
	} else {
		{ int rc = grow(&ctx->d_text, &ctx->text_cap, (uint64_t)((double)n_bytes * reserve) + 128); if(rc) { return rc; } }
		CK(RT_MEMCPY_H2D_ASYNC(ctx->d_text, text, n_bytes, ctx->stream));
		CK(RT_MEMSET_ASYNC(ctx->d_text + n_bytes, '\n', 64, ctx->stream));
		S.h2d_bytes += n_bytes;
		d_text = ctx->d_text;
		first = (uint8_t)text[0]; last = (uint8_t)text[n_bytes - 1];
		if(last != '\n') { n_bytes++; }
	}
	if(first != '>' && first != '@') { g_err = "mab_text_begin: the chunk does not start with a FASTA / FASTQ record"; return MAB_EFORMAT; }
	const uint32_t fastq = first == '@';
	ctx->tx.d_text = d_text; ctx->tx.n_text = n_bytes;
	const uint32_t n_tiles = (uint32_t)((n_bytes + MAB_TXT_TILE - 1) / MAB_TXT_TILE);
	{ int rc = grow(&ctx->d_tiles, &ctx->tiles_cap, (uint64_t)(4.0 * n_tiles * reserve) + 64); if(rc) { return rc; } }
	const uint64_t z_bytes = (uint64_t)((double)n_bytes * reserve);
	uint64_t mark_cap = std::max<uint64_t>(mark_hw + mark_hw / 2, z_bytes / 64 + 1024);
	uint64_t rec_cap = std::max<uint64_t>(rec_hw + rec_hw / 2, z_bytes / 256 + 1024);
	{ int rc = pin_reserve(ctx, sizeof(TextCounters) + 2 * sizeof(BatchCounters) + 512); if(rc) { return rc; } }
	TextCounters *pin_tc = (TextCounters *)ctx->pin;
	TextCounters tc;
	for(int attempt = 0; ; attempt++) {
		{ int rc = grow(&ctx->d_marks, &ctx->marks_cap, 4 * mark_cap + 64); if(rc) { return rc; } }
		{ int rc = grow(&ctx->d_trec, &ctx->trec_cap, sizeof(TextRec) * rec_cap + 64); if(rc) { return rc; } }
		{ int rc = grow(&ctx->d_reads, &ctx->reads_cap, sizeof(ReadRec) * rec_cap + 64); if(rc) { return rc; } }
		memset(pin_tc, 0, sizeof(TextCounters)); pin_tc->fastq = fastq;
		CK(RT_MEMCPY_H2D_ASYNC(ctx->d_tc, pin_tc, sizeof(TextCounters), ctx->stream));
		RT_LAUNCH(k_text_count, n_tiles, 256, 0, ctx->stream, d_text, n_bytes, fastq, ctx->d_tiles);
		RT_LAUNCH(k_scan_u32, 1, MAB_PIPE_THREADS, 0, ctx->stream, ctx->d_tiles, n_tiles, &ctx->d_tc->n_mark);
		RT_LAUNCH(k_text_mark, n_tiles, 256, 0, ctx->stream, d_text, n_bytes, fastq, (const uint32_t *)ctx->d_tiles, ctx->d_marks, mark_cap, ctx->d_tc);
		RT_LAUNCH(k_text_index, ctx->n_sm * 8, 32 * MAB_WARPS_PER_CTA, 0, ctx->stream, d_text, n_bytes, (const uint32_t *)ctx->d_marks, ctx->d_tc, ctx->d_trec, rec_cap);
		RT_LAUNCH(k_text_layout, 1, MAB_PIPE_THREADS, 0, ctx->stream, (const TextRec *)ctx->d_trec, ctx->d_reads, ctx->d_tc);
		S.n_launches += 5;
		CK(RT_MEMCPY_D2H_ASYNC(pin_tc, ctx->d_tc, sizeof(TextCounters), ctx->stream));
		{ double tw = RT_WALL_MS(); CK(ctx_sync(ctx)); S.ms_wall_wait += (float)(RT_WALL_MS() - tw); }
		tc = *pin_tc;
		trace_line(ctx, "begin: parsed, reads", tc.n_rec);
		S.d2h_bytes += sizeof(TextCounters);
		if(tc.err & (MAB_TXT_EMARKS | MAB_TXT_ERECS)) {
			if(attempt >= 2) { g_err = "text index arrays overflowed after retries"; return MAB_EOVERFLOW; }
			mark_cap = std::max<uint64_t>(mark_cap, tc.n_mark + tc.n_mark / 8 + 1024);
			rec_cap = std::max<uint64_t>(rec_cap, (fastq ? tc.n_mark / 4 : tc.n_mark) + tc.n_mark / 8 + 1024);
			S.n_retry++;
			continue;
		}
		break;
	}
	if(tc.err & MAB_TXT_EFORMAT) { g_err = "mab_text_begin: record layout not handled by the device reader (wrapped FASTQ, blank lines, text before the first record)"; return MAB_EFORMAT; }
	{
		std::lock_guard<std::mutex> lk(ctx->cal->mu);
		ctx->cal->mark_hw = std::max<uint64_t>(ctx->cal->mark_hw, (uint64_t)((double)tc.n_mark * reserve));
		ctx->cal->rec_hw = std::max<uint64_t>(ctx->cal->rec_hw, (uint64_t)((double)tc.n_rec * reserve));
	}
	ctx->tx.tc = tc; ctx->tx.n_rec = tc.n_rec;
	/* the read block */
	{ int rc = grow(&ctx->d_base, &ctx->base_cap, (uint64_t)((double)tc.span * reserve) + 256); if(rc) { return rc; } }
	CK(RT_MEMSET_ASYNC(ctx->d_base, 0, tc.span + 256, ctx->stream));
	RT_LAUNCH(k_text_pack, ctx->n_sm * 8, 32 * MAB_WARPS_PER_CTA, 0, ctx->stream, d_text, (const TextRec *)ctx->d_trec, (const ReadRec *)ctx->d_reads, tc.n_rec, ctx->d_base);
	S.n_launches++;
	PipeShape sh; sh.n_seq = tc.n_rec; sh.maxlen = tc.maxlen; sh.tot_len = tc.tot_len; sh.span = tc.span;
	ctx->pin_user = sizeof(TextCounters) + 256;
	if(tc.n_rec != 0) {
		int rc = pipeline_run(ctx, ctx->d_base, sh, rlen_prev, rlen_known ? 1u : 0u, true, reserve);
		if(rc) { return rc; }
	}
	ctx->tx.n_kept = tc.n_rec;						/* dropped (empty) records are counted out in finish, where the records come back */
	ctx->tx.stage = rlen_known ? 2 : 1; ctx->tx.rlen_committed = rlen_prev;
	if(rlen_known && ctx->hc.chain_valid) { ctx->rlen_last = ctx->hc.chain_rlen; }
	text_fill_info(ctx, info);
	S.ms_wall = (float)(RT_WALL_MS() - t_call);
	trace_line(ctx, "begin: mapped, ms", S.ms_wall);
	return MAB_OK;
}

extern "C" int mab_text_commit(mab_ctx *ctx, uint32_t rlen_prev, mab_text_info_t *info)
{
	if(ctx->tx.stage != 1 && ctx->tx.stage != 2) { g_err = "mab_text_commit: no chunk in flight"; return MAB_EINVAL; }
	CK(RT_USE_DEVICE(ctx->device));
	bool check = false;
	if(ctx->tx.n_rec != 0) {
		if(ctx->tx.stage == 1) {
			/* the first chain-loading read was mapped assuming its own reference's length; only if the true value flips its first
			 * seed test does anything have to be redone (the verification pass finds that out, and what follows from it) */
			const BatchCounters &hc = ctx->hc;
			if(hc.fd_valid) {
				bool used = (hc.fd_apos >= hc.fd_used) || (hc.fd_flags & 2), actual = (hc.fd_apos >= rlen_prev) || (hc.fd_flags & 2);
				check = used != actual;
			}
		} else { check = rlen_prev != ctx->tx.rlen_committed; }				/* committed before with another value (a predecessor was corrected): verify again */
	}
	if(check) {
		int rc = pipe_verify(ctx, rlen_prev, 1u); if(rc) { return rc; }
		if((ctx->hc.err_any & (MAB_ERR_POOL_OVF | MAB_ERR_WS_OVF)) || ctx->hc.pool_top > ctx->pool_cap / 4) { g_err = "result pool overflow while re-mapping a read"; return MAB_EOVERFLOW; }
	}
	ctx->tx.stage = 2; ctx->tx.rlen_committed = rlen_prev;
	text_fill_info(ctx, info);
	return MAB_OK;
}

extern "C" int mab_text_finish(mab_ctx *ctx, char *sam_out, uint64_t sam_cap, const char **sam_ptr, mab_text_info_t *info)
{
	mab_stats_t &S = ctx->stats;
	if(ctx->tx.stage != 2) { g_err = "mab_text_finish: call mab_text_begin (and mab_text_commit) first"; return MAB_EINVAL; }
	if(sam_ptr) { *sam_ptr = nullptr; }
	const uint32_t n = ctx->tx.n_rec;
	if(n == 0) { ctx->tx.sam_total = 0; ctx->tx.stage = 0; text_fill_info(ctx, info); return MAB_OK; }
	CK(RT_USE_DEVICE(ctx->device));
	const double t_call = RT_WALL_MS();
	trace_line(ctx, "finish: enter");
	const uint32_t flags = ctx->tx.flags, tags = flags & ~(MAB_TEXT_KEEP_QUAL | MAB_TEXT_DEVICE_OUT), keep_qual = (flags & MAB_TEXT_KEEP_QUAL) != 0;
	const uint32_t ctas = std::max<uint32_t>(1, std::min<uint32_t>((n + MAB_WARPS_PER_CTA - 1) / MAB_WARPS_PER_CTA, ctx->n_sm * 8));
	const uint64_t rr_bytes = sizeof(ReadRec) * (uint64_t)n;
	{ int rc = pin_reserve(ctx, sizeof(TextCounters) + 256 + (uint64_t)((double)rr_bytes * ctx->tx.reserve) + 2 * sizeof(BatchCounters) + 512); if(rc) { return rc; } }
	TextCounters *pin_tc = (TextCounters *)ctx->pin;
	ReadRec *pin_rr = (ReadRec *)(ctx->pin + sizeof(TextCounters) + 256);
	RT_EVENT_RECORD(ctx->ev[6], ctx->stream);
	RT_LAUNCH(k_post, ctas, 32 * MAB_WARPS_PER_CTA, 2048 * MAB_WARPS_PER_CTA, ctx->stream, ctx->P, ctx->d_pool, ctx->d_reads, n, (const double *)ctx->d_thr, ctx->d_frames);
	RT_LAUNCH((k_sam<false>), ctas, 32 * MAB_WARPS_PER_CTA, 0, ctx->stream, ctx->P, (const uint32_t *)ctx->d_pool, (const ReadRec *)ctx->d_reads, ctx->d_trec, n, ctx->tx.d_text, (const uint8_t *)ctx->d_base, tags, keep_qual, (uint8_t *)nullptr);
	RT_LAUNCH(k_sam_offsets, 1, MAB_PIPE_THREADS, 0, ctx->stream, ctx->d_trec, n, ctx->d_tc);
	S.n_launches += 3;
	CK(RT_MEMCPY_D2H_ASYNC(pin_tc, ctx->d_tc, sizeof(TextCounters), ctx->stream));
	/* the output buffer is sized from earlier chunks; the size pass usually confirms it while the write pass is already queued */
	double sam_per_byte;
	{ std::lock_guard<std::mutex> lk(ctx->cal->mu); sam_per_byte = ctx->cal->sam_per_byte; }
	uint64_t est = (uint64_t)((sam_per_byte * (double)ctx->tx.n_text + 512.0 * n) * ctx->tx.reserve) + 4096;
	{ int rc = grow(&ctx->d_sam, &ctx->sam_cap, est); if(rc) { return rc; } }
	{ double tw = RT_WALL_MS(); CK(ctx_sync(ctx)); S.ms_wall_wait += (float)(RT_WALL_MS() - tw); }
	const uint64_t total = pin_tc->sam_total;
	trace_line(ctx, "finish: sized, MB", total / 1048576.0);
	S.d2h_bytes += sizeof(TextCounters);
	{ int rc = grow(&ctx->d_sam, &ctx->sam_cap, total + 64); if(rc) { return rc; } }
	{ double r = (double)total / (double)ctx->tx.n_text; std::lock_guard<std::mutex> lk(ctx->cal->mu); if(r * 1.1 > ctx->cal->sam_per_byte) { ctx->cal->sam_per_byte = r * 1.1; } }
	RT_LAUNCH((k_sam<true>), ctas, 32 * MAB_WARPS_PER_CTA, 0, ctx->stream, ctx->P, (const uint32_t *)ctx->d_pool, (const ReadRec *)ctx->d_reads, ctx->d_trec, n, ctx->tx.d_text, (const uint8_t *)ctx->d_base, tags, keep_qual, ctx->d_sam);
	S.n_launches++;
	RT_EVENT_RECORD(ctx->ev[7], ctx->stream);
	ctx->tx.sam_total = total;
	/* per-read records: error flags, statistics for the next chunk's launch shapes */
	CK(RT_MEMCPY_D2H_ASYNC(pin_rr, ctx->d_reads, rr_bytes, ctx->stream));
	S.d2h_bytes += rr_bytes;
	if(!(flags & MAB_TEXT_DEVICE_OUT)) {
		char *dst = sam_out;
		if(dst == nullptr) {
			if(total + 1 > ctx->h_sam_cap) {
				RT_HOST_FREE(ctx->h_sam); ctx->h_sam = nullptr; ctx->h_sam_cap = 0;
				uint64_t nb = total + total / 4 + 4096;
				if(!RT_OK(RT_HOST_ALLOC(&ctx->h_sam, nb))) { g_err = std::string("pinned host allocation failed: ") + RT_ERRSTR(); return MAB_ENOMEM; }
				ctx->h_sam_cap = nb;
			}
			dst = (char *)ctx->h_sam;
		} else if(total > sam_cap) {										/* the chunk stays in flight: call again with info->sam_bytes of room */
			g_err = "mab_text_finish: output buffer too small (" + std::to_string(total) + " bytes needed)";
			CK(ctx_sync(ctx)); text_fill_info(ctx, info); return MAB_ENOMEM;
		}
		if(total) { CK(RT_MEMCPY_D2H_ASYNC(dst, ctx->d_sam, total, ctx->stream)); }
		S.d2h_bytes += total;
		if(sam_ptr) { *sam_ptr = dst; }
	}
	RT_EVENT_RECORD(ctx->ev[5], ctx->stream);
	{ double tw = RT_WALL_MS(); CK(ctx_sync(ctx)); S.ms_wall_wait += (float)(RT_WALL_MS() - tw); }
	update_sc_caps(ctx, pin_rr, n);
	uint32_t n_failed = 0; uint64_t kept = 0;
	for(uint32_t i = 0; i < n; i++) { n_failed += pin_rr[i].err != 0; kept += pin_rr[i].len != 0; }
	S.n_failed = n_failed; ctx->tx.n_kept = kept;
	S.ms_post = RT_EVENT_MS(ctx->ev[6], ctx->ev[7]);										/* post-processing + SAM kernels */
	S.ms_h2d = RT_EVENT_MS(ctx->ev[0], ctx->ev[1]);
	S.ms_seed = RT_EVENT_MS(ctx->ev[1], ctx->ev[3]);
	S.ms_sortchain = 0.f; S.ms_extend = 0.f; S.ms_extend_r0 = 0.f;
	for(uint32_t r = 0; r < ctx->P.n_occ && r < 8; r++) {
		S.ms_sortchain += RT_EVENT_MS(ctx->rev[3 * r], ctx->rev[3 * r + 1]);
		float e = RT_EVENT_MS(ctx->rev[3 * r + 1], ctx->rev[3 * r + 2]);
		S.ms_extend += e; if(r == 0) { S.ms_extend_r0 = e; }
	}
	S.ms_d2h = RT_EVENT_MS(ctx->ev[7], ctx->ev[5]);
	S.ms_total = RT_EVENT_MS(ctx->ev[0], ctx->ev[5]);
	S.ms_wall += (float)(RT_WALL_MS() - t_call);
	trace_line(ctx, "finish: done; device ms sortchain, extend", S.ms_sortchain, S.ms_extend);
	trace_line(ctx, "finish: done; device ms seed, post", S.ms_seed, S.ms_post);
	trace_line(ctx, "finish: done; device ms h2d, d2h", S.ms_h2d, S.ms_d2h);
	ctx->tx.stage = 0;
	text_fill_info(ctx, info);
	return MAB_OK;
}

extern "C" int mab_map_text(mab_ctx *ctx, const char *text, uint64_t n_bytes, uint32_t flags, char *sam_out, uint64_t sam_cap, const char **sam_ptr, mab_text_info_t *info)
{
	int rc = mab_text_begin(ctx, text, n_bytes, flags, ctx->rlen_last, 1, nullptr);
	if(rc) { return rc; }
	return mab_text_finish(ctx, sam_out, sam_cap, sam_ptr, info);
}
