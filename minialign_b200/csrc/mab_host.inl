/*
 * mab_host.inl -- host driver behind the C ABI (include/minialign_b200.h): context setup, batch orchestration, D2H and the
 * floating-point post-processing (pruning, supplementary/secondary split, MAPQ: minialign.c:4175-4396) that the reference
 * also runs on the CPU.  Included by mab_cuda.cu (real CUDA runtime) and by tests/emu/mab_emu.cpp (CUDA-on-CPU shim used by
 * the "not gpu" tests); the RT_* macros are the only difference between the two builds.
 */
#include "../../include/minialign_b200.h"
#include "mab_kernels.cuh"
#include <new>
#include <memory>
#include <vector>
#include <string>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <thread>
#include <mutex>

using namespace mab;

static thread_local std::string g_err;
extern "C" const char *mab_last_error(void) { return g_err.c_str(); }

struct PipeShape { uint32_t n_seq, maxlen; uint64_t tot_len, span; };		/* what the host knows about a batch: counts, longest read, bases, size of the base block */
/* launch geometry and buffer shapes of the batch in flight (kept in the context: redo passes may be issued by a later call) */
struct RunState {
	PipeShape sh; const uint8_t *d_base;
	uint32_t blk_cap, ext_ctas, seed_ctas, init_ctas, sc_cls[2][MAB_SC_CLASSES], sc_ncls[2], rlen_init, init_known;
	uint64_t arena_stride;
};

/* What the contexts of one index have learnt about the workload: output and workspace sizes per input byte, seed counts, the
 * parser's array sizes.  Shared between a context and its clones, so that a clone's first chunk is sized by what the others
 * have already seen instead of finding everything out again (a buffer found too small costs a re-run, and re-allocating one
 * waits for the whole device). */
struct Calib {
	std::mutex mu;
	uint64_t mark_hw = 0, rec_hw = 0;	/* high-water marks of the parser's arrays */
	double sam_per_byte = 1.3;			/* SAM bytes per input byte (high-water mark, sizes d_sam) */
	uint32_t sc_c0[2] = { 1280, 1280 };	/* k_sortchain: staging capacity of the smallest size class (median seed bound of the previous batch), per round kind */
	double ws_per_base = 6.0;			/* workspace estimate: bytes per read base beyond the fixed 20 KB per read (high-water mark) */
	bool ws_seen = false;				/* ws_per_base is a measurement (a batch has gone through), not the built-in guess */
	uint64_t reserve_bytes = 0;			/* mab_text_reserve: the largest chunk the caller will pass; buffers are sized for it from the first chunk on */
};

/* streams the size classes of k_sort are spread over.  Few on purpose: a process has 8 hardware queues by default
 * (CUDA_DEVICE_MAX_CONNECTIONS; the hosts here raise it to 32) and streams that share one wait for each other's kernels. */
#define MAB_SIDE_STREAMS 3

struct mab_ctx {
	int device;
	DevParams P;
	std::shared_ptr<std::vector<uint8_t>> blob;	/* host copy of the index image (reference names / sequences for the printer); shared by clones */
	const uint8_t *blob_ptr = nullptr;	/* = blob->data(), or the caller's image when it lends it (MAB_FLAG_BORROW_INDEX) */
	mab_ctx *parent = nullptr;			/* clone: d_idx / d_ntail / d_thr belong to the parent */
	uint8_t *d_idx = nullptr, *d_ntail = nullptr;
	uint32_t n_sm = 0, n_slots = 0;
	mab_params_t prm;
	double xcoef;
	/* batch buffers (grown on demand) */
	uint8_t *d_seq = nullptr; uint64_t seq_cap = 0;
	ReadRec *d_reads = nullptr; uint64_t reads_cap = 0;
	uint8_t *d_ws = nullptr; uint64_t ws_cap = 0;
	uint32_t *d_frames = nullptr; uint64_t frames_cap = 0;
	uint32_t *d_order = nullptr; uint64_t order_cap = 0;	/* work order of k_extend: read indices by descending length */
	uint32_t *d_recs = nullptr; uint64_t recs_cap = 0;		/* minimizer records of k_seed_scan: 16 B per base position of the read block */
	uint8_t *d_arenas = nullptr; uint64_t arenas_cap = 0;
	uint32_t *d_pool = nullptr; uint64_t pool_cap = 0;
	BatchCounters *d_ctr = nullptr;
	RT_STREAM stream; bool have_stream = false; int n_ev = 0;
	RT_STREAM side[MAB_SIDE_STREAMS]; RT_EVENT fork_ev, join_ev[MAB_SIDE_STREAMS]; int n_side = 0; bool have_fork_ev = false;	/* the size classes of k_sort / k_chain run side by side */
	RT_EVENT sync_ev; bool have_sync_ev = false;	/* the host waits on this one asleep (blocking-sync event) instead of spinning in a stream synchronize */
	RT_EVENT ev[8];
	RT_EVENT rev[24];					/* per-round kernel boundaries: [3r] sortchain start, [3r+1] extend start, [3r+2] extend end */
	int device_input = 0;
	uint32_t rlen_last = 0;				/* the reference's self->rlen carried from read to read and batch to batch */
	/* results of the last batch */
	uint32_t *res_words = nullptr; uint64_t res_cap = 0;	/* flat results of the last batch (uninitialised storage, filled in parallel) */
	std::vector<uint64_t> res_ofs;
	uint32_t *h_pool = nullptr; uint64_t h_pool_cap = 0;	/* pinned host copy of the device result pool */
	uint8_t *pin = nullptr; uint64_t pin_cap = 0;			/* pinned bounce buffer: [ReadRec x n][order u32 x n][BatchCounters x 2] (every async copy has a
															 * pinned host side: a pageable one blocks inside the runtime and stalls the other contexts' submissions) */
	uint64_t pin_user = 0;				/* bytes at the head of `pin` the caller of pipeline_run uses itself */
	uint8_t *d_io = nullptr; uint64_t io_cap = 0;			/* [ofs u64 x n][len u32 x n] of the record-level entry point */
	BatchCounters hc;					/* counters of the last batch */
	RunState rs;
	uint64_t arena_budget = 40ull << 30;	/* HBM the DP arenas of this context may take (mab_set_arena_budget) */
	bool chain_warp = false;				/* MAB_CHAIN_WARP=1: the chaining's window scan 32 candidates at a time (pays off with long scans only: 15.8 vs 9.9 ms per chunk on the E.coli-like workload) */
	bool chain_staged = false;			/* MAB_CHAIN_STAGED=1: k_chain on a shared-memory copy of the seed array, one launch per size class (A/B switch) */
	bool class_streams = true;			/* MAB_CLASS_STREAMS=0: the size classes one after the other on the context's stream (A/B switch) */
	bool sort_walk = true;				/* MAB_SORT_WALK=0: the fused k_sortchain (sort by cycle-walking in shared memory) instead of k_sort + k_chain (A/B switch) */
	bool ext_wide = true;				/* MAB_EXT_WIDE=0: always the 80-register build of k_extend (A/B switch) */
	/* text path (mab_text_*): the chunk, its index, the packed read block, the SAM text */
	uint8_t *d_text = nullptr; uint64_t text_cap = 0;
	uint8_t *d_base = nullptr; uint64_t base_cap = 0;
	uint32_t *d_marks = nullptr; uint64_t marks_cap = 0;
	uint32_t *d_tiles = nullptr; uint64_t tiles_cap = 0;
	TextRec *d_trec = nullptr; uint64_t trec_cap = 0;
	TextCounters *d_tc = nullptr;
	double *d_thr = nullptr;			/* MAPQ step table (mab_post.cuh) */
	bool thr_ok = false;
	uint8_t *d_sam = nullptr; uint64_t sam_cap = 0;
	uint8_t *h_sam = nullptr; uint64_t h_sam_cap = 0;		/* pinned; used when the caller passes no output buffer */
	std::shared_ptr<Calib> cal;			/* shared with the clones */
	struct TextState { const uint8_t *d_text; uint64_t n_text, n_kept, sam_total; uint32_t n_rec, flags, stage, rlen_committed; bool rlen_known; double reserve = 1.0; TextCounters tc; } tx;
	mab_stats_t stats;
};

/* ---------------------------------------------------------------- GABA constants (gaba.c:3613-3842) */
namespace {
struct GapModel {
	int M, gi, ge, gfa, gfb;
	int gap_h(int l) const { return std::max(-1 * (l > 0) * gi - ge * l, -1 * gfb * l); }
	int gap_v(int l) const { return std::max(-1 * (l > 0) * gi - ge * l, -1 * gfa * l); }
	int gap_e(int l) const { return -1 * (l > 0) * gi - ge * l; }
};

int check_params(const mab_params_t *p)
{
	int M = -128, X = 127;
	for(int i = 0; i < 16; i++) { M = std::max(M, (int)p->score_matrix[i]); X = std::min(X, (int)p->score_matrix[i]); }
	if(p->gi == 0 || p->gfa == 0 || p->gfb == 0) { return -1; }				/* combined model only */
	if(M <= 0 || M > 6 || X >= 0 || X < -7) { return -1; }
	if(X < -2 * (p->gi + p->ge)) { return -1; }
	if(X <= -1 * (p->gfa + p->gfb)) { return -1; }
	if(p->ge <= 0 || p->gi < 0) { return -1; }
	if(p->gfa <= p->ge || p->gfb <= p->ge) { return -1; }
	GapModel g = { M, p->gi, p->ge, p->gfa, p->gfb };
	int ofs = p->gi + p->ge;
	for(int i = 0; i < 8; i++) {											/* the wrapper validates with the 16-cell object */
		int t1 = ofs + g.gap_h(i * 2 + 1) - g.gap_h(i * 2);
		int t2 = ofs + (M + g.gap_v(i * 2 + 1)) - g.gap_v((i + 1) * 2);
		int t3 = ofs + (M + g.gap_h(i * 2 + 1)) - g.gap_h((i + 1) * 2);
		if(std::max(std::max(t1, t2), t3) > 127) { return -1; }
		if(std::min(std::min(t2, t3), t1) < 0) { return -1; }
	}
	return 0;
}

void init_gaba_consts(DevParams &P, const mab_params_t *p)
{
	int M = -128;
	for(int i = 0; i < 16; i++) { M = std::max(M, (int)p->score_matrix[i]); }
	GapModel g = { M, p->gi, p->ge, p->gfa, p->gfb };
	int ofs = p->gi + p->ge;
	for(int i = 0; i < 16; i++) { P.sb[i] = (int8_t)(p->score_matrix[i] + 2 * ofs); }
	P.adjh = P.adjv = p->gi; P.ofsh = P.ofsv = -ofs; P.gfh = ofs - p->gfb; P.gfv = ofs - p->gfa;
	P.tx = (int8_t)(p->xdrop - 128);
	{
		auto h8 = [](int v) { uint32_t x = ((uint32_t)(uint8_t)(int8_t)v) << 8; return x | (x << 16); };
		const uint32_t ulp = 0x00010001u;
		P.K_GFH1 = h8(P.gfh) + ulp; P.K_GFV1 = h8(P.gfv) + ulp; P.K_ADJH1 = h8(P.adjh) + ulp; P.K_ADJV1 = h8(P.adjv) + ulp; P.K_OFS = h8(P.ofsh);
		P.K_M1 = 0xffffffffu;
	}
	P.gi = p->gi; P.ge = p->ge; P.gfa = p->gfa; P.gfb = p->gfb;
	long long diag = 0, off = 0;
	for(int i = 0; i < 16; i++) { if((i & 3) == (i >> 2)) { diag += p->score_matrix[i]; } else { off += p->score_matrix[i]; } }
	double m = (double)diag / 4.0, x = (double)off / 12.0;
	P.imx = 1 / (m - x); P.xmx = x / (m - x);
	for(int wi = 0; wi < 3; wi++) {
		int W = 64 >> wi;
		RootTpl &R = P.root[wi];
		memset(&R, 0, sizeof(R));
		for(int i = 0; i < W / 2; i++) {
			int lo = W / 2 - 1 - i, hi = W / 2 + i;
			uint8_t dh_lo = (uint8_t)(ofs + g.gap_h(i * 2 + 1) - g.gap_h(i * 2));
			uint8_t dh_hi = (uint8_t)(ofs + M + g.gap_v(i * 2 + 1) - g.gap_v((i + 1) * 2));
			uint8_t dv_lo = (uint8_t)(ofs + M + g.gap_h(i * 2 + 1) - g.gap_h((i + 1) * 2));
			uint8_t dv_hi = (uint8_t)(ofs + g.gap_v(i * 2 + 1) - g.gap_v(i * 2));
			R.dh[lo] = (int8_t)(uint8_t)(0 - dh_lo); R.dh[hi] = (int8_t)(uint8_t)(0 - dh_hi);		/* dh is kept negated */
			R.dv[lo] = (int8_t)dv_lo; R.dv[hi] = (int8_t)dv_hi;
			R.de[lo] = (int8_t)(uint8_t)(p->gi + dv_lo + g.gap_e(i * 2 + 1) - g.gap_h(i * 2 + 1));
			R.de[hi] = (int8_t)(uint8_t)(p->gi + dv_hi - p->gi);
			R.df[lo] = (int8_t)(uint8_t)(p->gi + dh_lo - p->gi);
			R.df[hi] = (int8_t)(uint8_t)(p->gi + dh_hi + g.gap_e(i * 2 + 1) - g.gap_v(i * 2 + 1));
			R.md[lo] = (int16_t)(-(i + 1) * M + g.gap_h(i * 2 + 1));
			R.md[hi] = (int16_t)(-(i + 1) * M + g.gap_v(i * 2 + 1));
		}
		R.init_max = -(M + g.gap_h(1));
		R.mdrop = (int16_t)(R.init_max - 128);
	}
}

inline uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }
}  // namespace

/* MAB_TRACE=1: one stderr line per stage boundary / reallocation (wall clock in ms since the first line; for reading how the
 * chunks of several contexts interleave on one GPU) */
static bool trace_on() { static int on = -1; if(on < 0) { const char *e = getenv("MAB_TRACE"); on = e && atoi(e) > 0; } return on != 0; }
static void trace_line(const void *ctx, const char *what, double a = 0, double b = 0)
{
	if(!trace_on()) { return; }
	static double t0 = RT_WALL_MS();
	fprintf(stderr, "[T %10.2f ctx %p] %s %.3f %.3f\n", RT_WALL_MS() - t0, ctx, what, a, b);
}

static int text_init(struct mab_ctx *ctx);
static void text_destroy(struct mab_ctx *ctx);
#define CK(call) do { if(!RT_OK(call)) { g_err = std::string(#call) + ": " + RT_ERRSTR(); return MAB_ENODEV; } } while(0)
#define CKP(call) do { if(!RT_OK(call)) { g_err = std::string(#call) + ": " + RT_ERRSTR(); mab_destroy(ctx); return nullptr; } } while(0)

/* wait for everything queued on the context's stream with the host thread asleep: several contexts per GPU and several GPUs
 * per process each have a thread waiting most of the time, and spinning waits would eat the cores the reader needs */
static inline auto ctx_sync(mab_ctx *ctx) -> decltype(RT_STREAM_SYNC(ctx->stream))
{
	if(!ctx->have_sync_ev) { return RT_STREAM_SYNC(ctx->stream); }
	RT_EVENT_RECORD(ctx->sync_ev, ctx->stream);
	return RT_EVENT_SYNC(ctx->sync_ev);
}

/* the per-context part of the set-up: counters, stream, events, launch shape */
static int ctx_private_init(mab_ctx *ctx)
{
	CK(RT_MALLOC(&ctx->d_ctr, sizeof(BatchCounters)));
	CK(RT_STREAM_CREATE(&ctx->stream)); ctx->have_stream = true;
	for(int i = 0; i < 8; i++) { CK(RT_EVENT_CREATE(&ctx->ev[i])); ctx->n_ev++; }
	for(int i = 0; i < 24; i++) { CK(RT_EVENT_CREATE(&ctx->rev[i])); ctx->n_ev++; }
	CK(RT_SYNC_EVENT_CREATE(&ctx->sync_ev)); ctx->have_sync_ev = true;
	CK(RT_LIGHT_EVENT_CREATE(&ctx->fork_ev)); ctx->have_fork_ev = true;
	for(int i = 0; i < MAB_SIDE_STREAMS; i++) { CK(RT_STREAM_CREATE(&ctx->side[i])); CK(RT_LIGHT_EVENT_CREATE(&ctx->join_ev[i])); ctx->n_side++; }
	if(const char *e = getenv("MAB_CHAIN_WARP")) { ctx->chain_warp = atoi(e) != 0; }
	if(const char *e = getenv("MAB_CLASS_STREAMS")) { ctx->class_streams = atoi(e) != 0; }
	if(const char *e = getenv("MAB_CHAIN_STAGED")) { ctx->chain_staged = atoi(e) != 0; }
	RT_FUNC_MAX_SMEM(k_sortchain, 16 * MAB_SC_MAX + 2048);
	RT_FUNC_MAX_SMEM(k_chain, 16 * MAB_SC_MAX + 2048);
	RT_FUNC_MAX_SMEM(k_sort, 4 * MAB_WK_SM_WORDS + 32768);
	RT_FUNC_MAX_SMEM(k_sort_check, 4 * MAB_WK_SM_WORDS + 32768);
	if(const char *e = getenv("MAB_SORT_WALK")) { ctx->sort_walk = atoi(e) != 0; }
	ctx->n_slots = RT_EXTEND_SLOTS(ctx->n_sm);
	if(const char *e = getenv("MAB_EXT_WIDE")) { ctx->ext_wide = atoi(e) != 0; }
	if(const char *e = getenv("MAB_EXT_CTAS")) {										/* resident k_extend CTAs per SM actually launched (<= MAB_EXT_CTAS_PER_SM) */
		int v = atoi(e);
		if(v >= 1 && v <= MAB_EXT_CTAS_PER_SM) { ctx->n_slots = ctx->n_sm * MAB_WARPS_PER_CTA * (uint32_t)v; }
	}
	return MAB_OK;
}

/* the host part of a context's set-up: parameters, the header of the index image, the GABA constants */
static mab_ctx *ctx_new(const uint8_t *b, uint64_t size, const mab_params_t *params, int device)
{
	if(b == nullptr || size < 64 || params == nullptr) { g_err = "mab_init: bad arguments"; return nullptr; }
	if(check_params(params) != 0) { g_err = "mab_init: unsupported scoring parameters (combined gap model with validated ranges only)"; return nullptr; }
	mab_ctx *ctx = new (std::nothrow) mab_ctx();
	if(ctx == nullptr) { g_err = "host allocation failed"; return nullptr; }
	ctx->device = device; ctx->prm = *params;
	try { ctx->cal = std::make_shared<Calib>(); } catch(const std::bad_alloc &) { g_err = "host allocation failed"; delete ctx; return nullptr; }
	memset(&ctx->stats, 0, sizeof(ctx->stats));
	if(!RT_OK(RT_SET_DEVICE(device))) { g_err = std::string("no usable CUDA device: ") + RT_ERRSTR(); delete ctx; return nullptr; }
	ctx->n_sm = RT_SM_COUNT(device);
	if(params->flags & MAB_FLAG_BORROW_INDEX) { ctx->blob_ptr = b; }
	else {
		try { ctx->blob = std::make_shared<std::vector<uint8_t>>(b, b + size); } catch(const std::bad_alloc &) { g_err = "host allocation failed"; delete ctx; return nullptr; }
		ctx->blob_ptr = ctx->blob->data();
	}
	DevParams &P = ctx->P;
	memset(&P, 0, sizeof(P));
	P.bkt_ofs = rd64(b); P.bkt_mask = rd64(b + 8);
	P.b = b[16]; P.w = b[17]; P.k = b[18]; P.n_occ = b[19];
	for(int i = 0; i < 7; i++) { P.occ[i] = rd32(b + 20 + 4 * i); }
	P.n_ref = rd32(b + 48); P.seq_ofs = rd64(b + 56);
	if(P.n_occ == 0 || P.n_occ > 7 || P.w == 0 || P.w >= 32 || P.k == 0 || P.k > 31 || P.bkt_ofs >= size || P.seq_ofs >= size) { g_err = "mab_init: malformed index image"; delete ctx; return nullptr; }
	P.twlen = (uint32_t)((params->wlen << 1) - params->wlen); P.tglen = (uint32_t)((params->glen << 1) - params->glen);
	P.min_score = params->min_score; P.min_ratio = params->min_ratio;
	double mc = 0.0, xc = 0.0;																/* minialign.c:4676-4681 */
	for(int i = 0; i < 16; i++) { if((i & 3) == (i >> 3)) { mc += params->score_matrix[0]; } else { xc += params->score_matrix[0]; } }
	P.mcoef = mc / 4.0; ctx->xcoef = xc / 12.0; P.xcoef = ctx->xcoef;
	init_gaba_consts(P, params);
	return ctx;
}

/* the device part after the index image is in place (ctx->d_idx) */
static mab_ctx *ctx_finish(mab_ctx *ctx)
{
	ctx->P.idx = ctx->d_idx;
	uint8_t nt[128]; memset(nt, 4, sizeof(nt));
	CKP(RT_MALLOC(&ctx->d_ntail, 256));
	CKP(RT_MEMCPY_H2D(ctx->d_ntail, nt, 128));
	if(text_init(ctx) != MAB_OK) { mab_destroy(ctx); return nullptr; }
	CKP(RT_DEVICE_SYNC());								/* the uploads above ran on the legacy stream: nothing on the context's own (non-blocking) stream may overtake them */
	if(ctx_private_init(ctx) != MAB_OK) { mab_destroy(ctx); return nullptr; }
	trace_line(ctx, "init: done");
	return ctx;
}

extern "C" mab_ctx *mab_init(const void *mai_blob, uint64_t size, const mab_params_t *params, int device)
{
	mab_ctx *ctx = ctx_new((const uint8_t *)mai_blob, size, params, device);
	if(ctx == nullptr) { return nullptr; }
	trace_line(ctx, "init: enter, index MB", size / 1048576.0);
	CKP(RT_MALLOC(&ctx->d_idx, size + 256));
	CKP(RT_MEMCPY_H2D(ctx->d_idx, ctx->blob_ptr, size));
	trace_line(ctx, "init: index on the device");
	return ctx_finish(ctx);
}

/* ---- staged set-up: the index image goes to the device(s) in pieces while the caller is still producing it ----
 * A .mai file is a sequence of deflated 1 MiB frames (minialign.c:1137-1290); inflating a human-sized index takes seconds on all
 * host cores and copying 15-17 GB from pageable memory takes as long again.  Here every piece is handed over as soon as it exists:
 * mab_load_put stages it in a ring of page-locked slots and queues one asynchronous copy per device, so the upload runs under the
 * inflation and every GPU of the process gets its copy from the same staging slot. */
struct mab_loader {
	std::vector<int> devices;
	std::vector<uint8_t *> d_idx;
	std::vector<RT_STREAM> streams;
	mab_params_t prm;
	uint64_t cap = 0;
	uint8_t *ring = nullptr;
	static constexpr uint32_t SLOTS = 64;
	static constexpr uint64_t SLOT_BYTES = 1ull << 20;
	std::mutex slot_mu[SLOTS];
	std::vector<RT_EVENT> slot_ev;			/* [slot][device] */
	std::vector<uint8_t> slot_used;			/* [slot] */
	std::mutex mu; uint32_t next = 0;
	bool failed = false;
};

extern "C" void mab_load_abort(mab_loader *ld)
{
	if(ld == nullptr) { return; }
	for(size_t d = 0; d < ld->devices.size(); d++) {
		RT_USE_DEVICE(ld->devices[d]);
		if(d < ld->streams.size()) { RT_STREAM_SYNC(ld->streams[d]); RT_STREAM_DESTROY(ld->streams[d]); }
		for(uint32_t sl = 0; sl < mab_loader::SLOTS; sl++) { size_t e = sl * ld->devices.size() + d; if(e < ld->slot_ev.size()) { RT_EVENT_DESTROY(ld->slot_ev[e]); } }
		if(d < ld->d_idx.size()) { RT_FREE(ld->d_idx[d]); }
	}
	RT_HOST_FREE(ld->ring);
	delete ld;
}

extern "C" mab_loader *mab_load_begin(uint64_t max_size, const mab_params_t *params, const int *devices, int n_devices)
{
	if(params == nullptr || devices == nullptr || n_devices <= 0 || max_size < 64) { g_err = "mab_load_begin: bad arguments"; return nullptr; }
	if(check_params(params) != 0) { g_err = "mab_init: unsupported scoring parameters (combined gap model with validated ranges only)"; return nullptr; }
	mab_loader *ld = new (std::nothrow) mab_loader();
	if(ld == nullptr) { g_err = "host allocation failed"; return nullptr; }
	trace_line(ld, "load: enter");
	try {
	ld->prm = *params; ld->cap = max_size; ld->devices.assign(devices, devices + n_devices);
	ld->slot_used.assign(mab_loader::SLOTS, 0);
	/* the devices in parallel: creating a CUDA context takes a few hundred milliseconds each */
	ld->d_idx.assign(n_devices, nullptr); ld->streams.resize(n_devices);
	ld->slot_ev.resize((size_t)mab_loader::SLOTS * n_devices);
	std::vector<std::string> errs(n_devices); std::vector<int> n_ev(n_devices, 0), have_stream(n_devices, 0);
	{
		std::vector<std::thread> th;
		for(int d = 0; d < n_devices; d++) {
			th.emplace_back([&, d]() {
				if(!RT_OK(RT_SET_DEVICE(devices[d]))) { errs[d] = std::string("no usable CUDA device: ") + RT_ERRSTR(); return; }
				if(!RT_OK(RT_MALLOC(&ld->d_idx[d], max_size + 256))) { errs[d] = std::string("device allocation failed: ") + RT_ERRSTR(); ld->d_idx[d] = nullptr; return; }
				if(!RT_OK(RT_STREAM_CREATE(&ld->streams[d]))) { errs[d] = std::string("stream creation failed: ") + RT_ERRSTR(); return; }
				have_stream[d] = 1;
				for(uint32_t sl = 0; sl < mab_loader::SLOTS; sl++) {
					if(!RT_OK(RT_EVENT_CREATE(&ld->slot_ev[(size_t)sl * n_devices + d]))) { errs[d] = std::string("event creation failed: ") + RT_ERRSTR(); return; }
					n_ev[d]++;
				}
			});
		}
		for(auto &t : th) { t.join(); }
	}
	for(int d = 0; d < n_devices; d++) {
		if(errs[d].empty()) { continue; }
		g_err = errs[d];
		for(int e = 0; e < n_devices; e++) {										/* partial set-up: undo by hand */
			RT_USE_DEVICE(devices[e]);
			for(int sl = 0; sl < n_ev[e]; sl++) { RT_EVENT_DESTROY(ld->slot_ev[(size_t)sl * n_devices + e]); }
			if(have_stream[e]) { RT_STREAM_DESTROY(ld->streams[e]); }
			RT_FREE(ld->d_idx[e]);
		}
		delete ld;
		return nullptr;
	}
	trace_line(ld, "load: devices ready");
	RT_USE_DEVICE(devices[0]);
	if(!RT_OK(RT_HOST_ALLOC(&ld->ring, mab_loader::SLOTS * mab_loader::SLOT_BYTES))) { g_err = std::string("pinned host allocation failed: ") + RT_ERRSTR(); mab_load_abort(ld); return nullptr; }
	trace_line(ld, "load: begin, index MB", max_size / 1048576.0);
	return ld;
	} catch(const std::exception &e) {									/* vectors, threads: nothing may be thrown across the C ABI */
		g_err = std::string("mab_load_begin: ") + e.what();
		delete ld;
		return nullptr;
	}
}

/* bytes [offset, offset + n) of the image; any thread, any order.  Returns once the bytes are staged (src may be reused). */
extern "C" int mab_load_put(mab_loader *ld, uint64_t offset, const void *src, uint64_t n)
{
	if(ld == nullptr || src == nullptr || offset + n > ld->cap) { g_err = "mab_load_put: bad arguments"; return MAB_EINVAL; }
	const uint8_t *s = (const uint8_t *)src;
	const size_t nd = ld->devices.size();
	while(n > 0) {
		const uint64_t m = std::min<uint64_t>(n, mab_loader::SLOT_BYTES);
		uint32_t sl;
		{ std::lock_guard<std::mutex> lk(ld->mu); sl = ld->next++ % mab_loader::SLOTS; }
		std::lock_guard<std::mutex> lk(ld->slot_mu[sl]);
		uint8_t *stage = ld->ring + (uint64_t)sl * mab_loader::SLOT_BYTES;
		if(ld->slot_used[sl]) { for(size_t d = 0; d < nd; d++) { CK(RT_EVENT_SYNC(ld->slot_ev[sl * nd + d])); } }	/* the slot's previous copies */
		memcpy(stage, s, m);
		for(size_t d = 0; d < nd; d++) {
			CK(RT_USE_DEVICE(ld->devices[d]));
			CK(RT_MEMCPY_H2D_ASYNC(ld->d_idx[d] + offset, stage, m, ld->streams[d]));
			RT_EVENT_RECORD(ld->slot_ev[sl * nd + d], ld->streams[d]);
		}
		ld->slot_used[sl] = 1;
		s += m; offset += m; n -= m;
	}
	return MAB_OK;
}

/* all pieces are in: waits for the copies and builds one context per device around the uploaded image.  blob = the complete
 * host image of `size` bytes (kept or borrowed like in mab_init).  The loader is gone afterwards, whatever the outcome. */
extern "C" int mab_load_end(mab_loader *ld, const void *blob, uint64_t size, mab_ctx **ctx_out)
{
	if(ld == nullptr || blob == nullptr || ctx_out == nullptr || size > ld->cap) { g_err = "mab_load_end: bad arguments"; mab_load_abort(ld); return MAB_EINVAL; }
	const size_t nd = ld->devices.size();
	int rc = MAB_OK;
	for(size_t d = 0; d < nd; d++) { ctx_out[d] = nullptr; }
	for(size_t d = 0; d < nd && rc == MAB_OK; d++) {
		RT_USE_DEVICE(ld->devices[d]);
		if(!RT_OK(RT_STREAM_SYNC(ld->streams[d]))) { g_err = std::string("index upload failed: ") + RT_ERRSTR(); rc = MAB_ENODEV; }
	}
	trace_line(ld, "load: index on the device(s)");
	for(size_t d = 0; d < nd && rc == MAB_OK; d++) {
		mab_ctx *ctx = ctx_new((const uint8_t *)blob, size, &ld->prm, ld->devices[d]);
		if(ctx == nullptr) { rc = MAB_EINVAL; break; }
		ctx->d_idx = ld->d_idx[d]; ld->d_idx[d] = nullptr;			/* the context owns it from here */
		ctx = ctx_finish(ctx);
		if(ctx == nullptr) { rc = MAB_ENODEV; break; }
		ctx_out[d] = ctx;
	}
	if(rc != MAB_OK) { for(size_t d = 0; d < nd; d++) { if(ctx_out[d]) { mab_destroy(ctx_out[d]); ctx_out[d] = nullptr; } } }
	mab_load_abort(ld);
	return rc;
}

/* a second (third, ...) context on the parent's device that shares its index image: batches in flight at the same time cost one
 * copy of the index, not one each.  The parent must outlive its clones. */
extern "C" mab_ctx *mab_clone(mab_ctx *parent)
{
	if(parent == nullptr) { g_err = "mab_clone: bad arguments"; return nullptr; }
	while(parent->parent != nullptr) { parent = parent->parent; }
	mab_ctx *ctx = new (std::nothrow) mab_ctx();
	if(ctx == nullptr) { g_err = "host allocation failed"; return nullptr; }
	ctx->parent = parent; ctx->device = parent->device; ctx->prm = parent->prm; ctx->P = parent->P; ctx->xcoef = parent->xcoef; ctx->n_sm = parent->n_sm;
	ctx->cal = parent->cal;
	ctx->blob = parent->blob; ctx->blob_ptr = parent->blob_ptr; ctx->d_idx = parent->d_idx; ctx->d_ntail = parent->d_ntail; ctx->d_thr = parent->d_thr; ctx->thr_ok = parent->thr_ok;
	memset(&ctx->stats, 0, sizeof(ctx->stats));
	if(!RT_OK(RT_USE_DEVICE(ctx->device))) { g_err = std::string("no usable CUDA device: ") + RT_ERRSTR(); delete ctx; return nullptr; }
	CKP(RT_MALLOC(&ctx->d_tc, sizeof(TextCounters)));
	if(ctx_private_init(ctx) != MAB_OK) { mab_destroy(ctx); return nullptr; }
	trace_line(ctx, "clone: done");
	return ctx;
}

extern "C" void mab_destroy(mab_ctx *ctx)
{
	if(ctx == nullptr) { return; }
	RT_USE_DEVICE(ctx->device);
	if(ctx->parent != nullptr) { ctx->d_idx = nullptr; ctx->d_ntail = nullptr; ctx->d_thr = nullptr; }	/* the parent's */
	if(ctx->have_sync_ev) { RT_EVENT_DESTROY(ctx->sync_ev); }
	if(ctx->have_fork_ev) { RT_EVENT_DESTROY(ctx->fork_ev); }
	for(int i = 0; i < ctx->n_side; i++) { RT_STREAM_DESTROY(ctx->side[i]); RT_EVENT_DESTROY(ctx->join_ev[i]); }
	RT_FREE(ctx->d_idx); RT_FREE(ctx->d_ntail); RT_FREE(ctx->d_ctr); RT_FREE(ctx->d_seq); RT_FREE(ctx->d_reads); RT_FREE(ctx->d_ws);
	RT_FREE(ctx->d_frames); RT_FREE(ctx->d_order); RT_FREE(ctx->d_recs); RT_FREE(ctx->d_arenas); RT_FREE(ctx->d_pool); RT_FREE(ctx->d_io);
	RT_HOST_FREE(ctx->h_pool); RT_HOST_FREE(ctx->pin); delete[] ctx->res_words;
	text_destroy(ctx);
	if(ctx->have_stream) { RT_STREAM_DESTROY(ctx->stream); }
	for(int i = 0; i < ctx->n_ev; i++) { RT_EVENT_DESTROY(i < 8 ? ctx->ev[i] : ctx->rev[i - 8]); }
	delete ctx;
}

extern "C" uint32_t mab_n_ref(const mab_ctx *ctx) { return ctx->P.n_ref; }
extern "C" int mab_ref_info(const mab_ctx *ctx, uint32_t rid, const char **name, uint32_t *l_name, uint32_t *l_seq, const uint8_t **seq)
{
	if(rid >= ctx->P.n_ref) { return MAB_EINVAL; }
	const uint8_t *b = ctx->blob_ptr, *s = b + ctx->P.seq_ofs + 24ull * rid;
	if(seq) { *seq = b + rd64(s); }
	if(name) { *name = (const char *)(b + rd64(s + 8)); }
	if(l_seq) { *l_seq = rd32(s + 16); }
	if(l_name) { *l_name = rd16(s + 20); }
	return MAB_OK;
}
extern "C" int mab_index_params(const mab_ctx *ctx, uint32_t *k, uint32_t *w, uint32_t *b, uint32_t *n_occ, uint32_t *occ)
{
	if(k) { *k = ctx->P.k; } if(w) { *w = ctx->P.w; } if(b) { *b = ctx->P.b; } if(n_occ) { *n_occ = ctx->P.n_occ; }
	if(occ) { for(int i = 0; i < 7; i++) { occ[i] = ctx->P.occ[i]; } }
	return MAB_OK;
}
extern "C" uint32_t mab_get_rlen(const mab_ctx *ctx) { return ctx->rlen_last; }
extern "C" void mab_set_rlen(mab_ctx *ctx, uint32_t rlen) { ctx->rlen_last = rlen; }
extern "C" int mab_device_memory(const mab_ctx *ctx, uint64_t *free_bytes, uint64_t *total_bytes)
{
	CK(RT_USE_DEVICE(ctx->device));
	size_t f = 0, t = 0;
	CK(RT_MEM_INFO(&f, &t));
	if(free_bytes) { *free_bytes = f; } if(total_bytes) { *total_bytes = t; }
	return MAB_OK;
}
extern "C" void mab_set_arena_budget(mab_ctx *ctx, uint64_t bytes) { ctx->arena_budget = bytes < (256ull << 20) ? (256ull << 20) : bytes; }
extern "C" int mab_set_device_input(mab_ctx *ctx, int on) { ctx->device_input = on; return MAB_OK; }
extern "C" int mab_last_stats(const mab_ctx *ctx, mab_stats_t *out) { *out = ctx->stats; return MAB_OK; }

template <class T> static int grow(T **p, uint64_t *cap, uint64_t need_bytes)
{
	if(need_bytes <= *cap && *p != nullptr) { return MAB_OK; }
	double tg = RT_WALL_MS();
	RT_FREE(*p); *p = nullptr;
	uint64_t nb = need_bytes + need_bytes / 4 + 4096;
	if(!RT_OK(RT_MALLOC(p, nb))) { g_err = std::string("device allocation failed: ") + RT_ERRSTR(); *cap = 0; return MAB_ENOMEM; }
	*cap = nb;
	trace_line(p, "grow MB, ms", nb / 1048576.0, RT_WALL_MS() - tg);
	return MAB_OK;
}

/* ---------------------------------------------------------------- host post-processing (minialign.c:4175-4396) */
namespace {
/* the reference's radix_sort_64x on {score, idx} pairs: same unstable algorithm as on the device (ksort.h:82-131) */
void rs_insertion64(uint32_t *a, uint32_t n)
{
	for(uint32_t i = 1; i < n; i++) {
		if(a[2 * i] < a[2 * (i - 1)]) {
			uint32_t t0 = a[2 * i], t1 = a[2 * i + 1], j;
			for(j = i; j > 0 && t0 < a[2 * (j - 1)]; j--) { a[2 * j] = a[2 * (j - 1)]; a[2 * j + 1] = a[2 * (j - 1) + 1]; }
			a[2 * j] = t0; a[2 * j + 1] = t1;
		}
	}
}
void rs_sort64(uint32_t *a, uint32_t n, int s)
{
	uint32_t head[256], end[256];
	memset(end, 0, sizeof(end));
	for(uint32_t i = 0; i < n; i++) { end[(a[2 * i] >> s) & 0xff]++; }
	head[0] = 0;
	for(int k = 1; k < 256; k++) { end[k] += end[k - 1]; head[k] = end[k - 1]; }
	for(int k = 0; k < 256;) {
		if(head[k] != end[k]) {
			int l = (a[2 * head[k]] >> s) & 0xff;
			if(l != k) {
				uint32_t t0 = a[2 * head[k]], t1 = a[2 * head[k] + 1];
				do {
					uint32_t s0 = t0, s1 = t1; t0 = a[2 * head[l]]; t1 = a[2 * head[l] + 1]; a[2 * head[l]] = s0; a[2 * head[l] + 1] = s1; head[l]++;
					l = (t0 >> s) & 0xff;
				} while(l != k);
				a[2 * head[k]] = t0; a[2 * head[k] + 1] = t1; head[k]++;
			} else { head[k]++; }
		} else { k++; }
	}
	if(s) {
		s = s > 8 ? s - 8 : 0;
		uint32_t beg = 0;
		for(int k = 0; k < 256; k++) {
			uint32_t sz = end[k] - beg;
			if(sz > 64) { rs_sort64(a + 2 * beg, sz, s); } else if(sz > 1) { rs_insertion64(a + 2 * beg, sz); }
			beg = end[k];
		}
	}
}
void radix_sort_64x(uint32_t *a, uint32_t n) { if(n <= 64) { rs_insertion64(a, n); } else { rs_sort64(a, n, 24); } }

struct HRes { uint32_t score, n_aln, plen, lb, ub; const uint32_t *alns; };		/* alns: n_aln x (lo, hi) pool offsets */

inline int32_t SC(uint32_t x) { return (int32_t)0x40000000 - (int32_t)x; }
inline uint32_t clip_mapq(double x) { uint32_t v = (uint32_t)x; return std::min(v, 60u * 16); }
}  // namespace

/* One alignment of a read's result in output order: where its record sits in the pool, its rank and MAPQ. */
struct PlanItem { const uint32_t *a; uint32_t rank, mapq; };
struct ReadPlan { size_t first; uint32_t count, n_uniq; uint64_t words; };

/* sort, prune, supplementary/secondary split, MAPQ (minialign.c:4175-4396): decides the output of one read without copying
 * anything; appends its items to `items` and returns the plan (words = size of the flat record post_emit will write) */
static ReadPlan post_plan(mab_ctx *ctx, const uint32_t *pool, const uint32_t *rec, std::vector<PlanItem> &items)
{
	uint32_t n_res = rec[0];
	std::vector<HRes> bins(n_res);
	std::vector<uint32_t> res(2 * (size_t)n_res);
	const uint32_t *p = rec + 1;
	for(uint32_t i = 0; i < n_res; i++) {
		bins[i].score = p[0]; bins[i].n_aln = p[1]; bins[i].plen = p[2]; bins[i].lb = p[3]; bins[i].ub = p[4]; bins[i].alns = p + 5;
		res[2 * i] = p[0]; res[2 * i + 1] = i;
		p += 5 + 2 * (size_t)p[1];
	}
	radix_sort_64x(res.data(), n_res);															/* minialign.c:4452 */
	/* mm_prune_regs (4185-4207) */
	uint32_t q = n_res;
	uint32_t minv = (uint32_t)SC((uint32_t)(SC(res[0]) * ctx->prm.min_ratio));
	while(res[2 * --q] > minv) {}
	n_res = q + 1;
	/* mm_collect_supp (4214-4263) */
	auto swap_res = [&](uint64_t x, uint64_t y) { std::swap(res[2 * x], res[2 * y]); std::swap(res[2 * x + 1], res[2 * y + 1]); };
	uint64_t pp, qq;
	for(pp = 1, qq = n_res; pp < qq; pp++) {
		uint64_t mx = 0;
		for(uint64_t i = pp; i < qq; i++) {
			const HRes &s = bins[res[2 * i + 1]];
			int64_t lb = s.lb, ub = s.ub, span = ub - lb;
			bool covered = false;
			for(uint64_t j = 0; j < pp; j++) {
				const HRes &t = bins[res[2 * j + 1]];
				if(t.ub < ub) { lb = std::max(lb, (int64_t)t.ub); } else { ub = std::min(ub, (int64_t)t.lb); }
				if(1.2 * (ub - lb) < span) { qq--; swap_res(i, qq); i--; covered = true; break; }
			}
			if(covered) { continue; }
			mx = std::max(mx, ((uint64_t)(2 * (ub - lb) - span) << 32) | i);
		}
		if(mx & 0xffffffff) { swap_res(pp, mx & 0xffffffff); }
	}
	uint64_t n_uniq = std::min(pp, qq);
	/* mm_post_map (4270-4325) */
	auto aln_at = [&](const HRes &h, uint32_t j) { return pool + ((uint64_t)h.alns[2 * j] | (uint64_t)h.alns[2 * j + 1] << 32); };
	int64_t usc = 0, lsc = INT64_MAX, tsc = 0;
	for(uint64_t i = n_uniq; i < n_res; i++) { usc = std::max(usc, (int64_t)SC(res[2 * i])); lsc = std::min(lsc, (int64_t)SC(res[2 * i])); tsc += SC(res[2 * i]); }
	lsc = (lsc == INT32_MAX) ? 0 : lsc;
	double tpc = 1.0, x = ctx->xcoef, mxc = ctx->P.mcoef + ctx->xcoef;
	std::vector<uint32_t> mapq(bins.size(), 0);
	for(uint64_t i = 0; i < n_uniq; i++) {
		uint32_t score = (uint32_t)SC(res[2 * i]); const HRes &h = bins[res[2 * i + 1]];
		double pid = 0.0; uint64_t len = 0;
		for(uint32_t j = 0; j < h.n_aln; j++) {
			const uint32_t *a = aln_at(h, j);
			double identity; uint64_t ib = (uint64_t)a[2] | (uint64_t)a[3] << 32; memcpy(&identity, &ib, 8);
			len += a[8]; pid += (double)a[8] * identity;
		}
		pid /= (double)len;
		double ec = 2.0 / (pid * mxc - x);
		double ulen = ec * std::max((int64_t)score - usc, (int64_t)0), pe = 1.0 / (ulen * ulen + 1);
		mapq[res[2 * i + 1]] = clip_mapq(-10.0 * 16 * log10(pe));
		tpc *= 1.0 - pe;
	}
	double tpe = std::min(1.0 - tpc, 1.0);
	for(uint64_t i = n_uniq; i < n_res; i++) {
		mapq[res[2 * i + 1]] = clip_mapq(-10.0 * 16 * log10(1.0 - tpe * (double)(res[2 * i] - lsc + 1) / (double)tsc));
	}
	/* mm_pack_reg (4364-4396): output order */
	ReadPlan pl; pl.first = items.size(); pl.count = 0; pl.n_uniq = 0; pl.words = 2;
	for(uint64_t i = 0; i < n_res; i++) {
		const HRes &h = bins[res[2 * i + 1]];
		for(uint32_t j = 0; j < h.n_aln; j++) {
			const uint32_t *a = aln_at(h, j);
			PlanItem it; it.a = a; it.rank = (uint32_t)i; it.mapq = mapq[res[2 * i + 1]];
			items.push_back(it);
			pl.words += 16 + 8ull * a[7] + a[9];
			pl.count++;
		}
		if(i == n_uniq - 1) { pl.n_uniq = pl.count; }
	}
	return pl;
}

/* writes the flat record of one read (layout in include/minialign_b200.h) */
static void post_emit(const ReadPlan &pl, const PlanItem *items, uint32_t *out)
{
	out[0] = pl.count; out[1] = pl.n_uniq;
	uint32_t *w = out + 2;
	for(uint32_t k = 0; k < pl.count; k++) {
		const uint32_t *a = items[k].a;
		uint32_t slen = a[7], npw = a[9], sn = a[10];
		memcpy(w, a, 40);
		w[10] = items[k].rank; w[11] = items[k].mapq; w[12] = w[13] = w[14] = w[15] = 0;
		memcpy(w + 16, a + MAB_ALN_HDR + 8ull * (sn - slen), 32ull * slen);
		memcpy(w + 16 + 8ull * slen, a + MAB_ALN_HDR + 8ull * sn, 4ull * npw);
		w += 16 + 8ull * slen + npw;
	}
}

/* ---------------------------------------------------------------- batch driver */
static uint32_t dp_blk_cap(uint32_t maxlen) { return (uint32_t)((4ull * ((uint64_t)maxlen + 512)) / 32 + 64); }


static int pin_reserve(mab_ctx *ctx, uint64_t need)
{
	if(need <= ctx->pin_cap) { return MAB_OK; }
	RT_HOST_FREE(ctx->pin); ctx->pin = nullptr; ctx->pin_cap = 0;
	if(!RT_OK(RT_HOST_ALLOC(&ctx->pin, need + need / 4))) { g_err = std::string("pinned host allocation failed: ") + RT_ERRSTR(); return MAB_ENOMEM; }
	ctx->pin_cap = need + need / 4;
	return MAB_OK;
}

static void pipe_rounds(mab_ctx *ctx, bool first, bool timed)
{
	const DevParams &P = ctx->P; const RunState &R = ctx->rs; mab_stats_t &S = ctx->stats;
	const uint32_t n_seq = R.sh.n_seq;
	if(timed && first) { RT_EVENT_RECORD(ctx->ev[2], ctx->stream); }
	RT_LAUNCH(k_seed_expand, R.seed_ctas, 32 * MAB_WARPS_PER_CTA, 0, ctx->stream, P, ctx->d_reads, n_seq, ctx->d_ws, (const uint32_t *)ctx->d_recs);
	S.n_launches++;
	if(timed && first) { RT_EVENT_RECORD(ctx->ev[3], ctx->stream); }
	for(uint32_t round = 0; round < P.n_occ; round++) {
		bool ev = timed && first && round < 8;
		if(ev) { RT_EVENT_RECORD(ctx->rev[3 * round], ctx->stream); }
		int kind = round == 0 ? 0 : 1;
		if(ctx->sort_walk) {
			/* sort (elements in global memory, a byte of shared memory per element: classes of x2), then chaining on the sorted array
			 * staged in shared memory (16 B per seed: the classes sized from the previous batch).  A class is as long as its slowest
			 * read (one warp per read, latency-bound), so the classes run side by side on streams of their own, forked off and
			 * joined back into the context's stream. */
			const bool par = ctx->class_streams;
			uint32_t used = 0;
			if(par) { RT_EVENT_RECORD(ctx->fork_ev, ctx->stream); }
			for(uint32_t cap = 32768, ci = 0; ; cap /= 2, ci++) {						/* the seed-rich classes first: they take longest */
				const bool first = cap >= 32768, last = cap <= 1024;
				RT_STREAM st = par ? ctx->side[ci % MAB_SIDE_STREAMS] : ctx->stream;
				if(par && ci < MAB_SIDE_STREAMS) { RT_STREAM_WAIT(st, ctx->fork_ev); }
				RT_LAUNCH(k_sort, n_seq, 32, 4 * MAB_WK_SM_WORDS + cap, st, P, ctx->d_reads, (const uint32_t *)ctx->d_order, n_seq, ctx->d_ws, ctx->d_frames, round, cap, last ? 0u : cap / 2, first ? 0xffffffffu : cap);
				S.n_launches++; used = ci + 1;
				if(last) { break; }
			}
			if(par) { for(uint32_t k = 0; k < std::min<uint32_t>(used, MAB_SIDE_STREAMS); k++) { RT_EVENT_RECORD(ctx->join_ev[k], ctx->side[k]); RT_STREAM_WAIT(ctx->stream, ctx->join_ev[k]); } }
			if(!ctx->chain_staged) {
				RT_LAUNCH(k_chain, (n_seq + MAB_WARPS_PER_CTA - 1) / MAB_WARPS_PER_CTA, 32 * MAB_WARPS_PER_CTA, 2048 * MAB_WARPS_PER_CTA, ctx->stream, P, ctx->d_reads, (const uint32_t *)ctx->d_order, n_seq, ctx->d_ws, ctx->d_frames, 0u, 0u, 0xffffffffu, ctx->chain_warp ? 1u : 0u);
				S.n_launches++;
			} else {
				if(par) { RT_EVENT_RECORD(ctx->fork_ev, ctx->stream); }
				for(uint32_t k = 0; k < R.sc_ncls[kind]; k++) {
					const uint32_t ci = R.sc_ncls[kind] - 1 - k;
					const uint32_t cap = R.sc_cls[kind][ci], lo = ci == 0 ? 0u : R.sc_cls[kind][ci - 1], hi = ci + 1 == R.sc_ncls[kind] ? 0xffffffffu : cap;
					RT_STREAM st = par ? ctx->side[k % MAB_SIDE_STREAMS] : ctx->stream;
					if(par && k < MAB_SIDE_STREAMS) { RT_STREAM_WAIT(st, ctx->fork_ev); }
					RT_LAUNCH(k_chain, n_seq, 32, 16 * cap + 2048, st, P, ctx->d_reads, (const uint32_t *)ctx->d_order, n_seq, ctx->d_ws, ctx->d_frames, cap, lo, hi, ctx->chain_warp ? 1u : 0u);
					S.n_launches++;
				}
				if(par) { for(uint32_t k = 0; k < std::min<uint32_t>(R.sc_ncls[kind], MAB_SIDE_STREAMS); k++) { RT_EVENT_RECORD(ctx->join_ev[k], ctx->side[k]); RT_STREAM_WAIT(ctx->stream, ctx->join_ev[k]); } }
			}
		} else {
			for(uint32_t ci = 0, lo = 0; ci < R.sc_ncls[kind]; ci++) {				/* one launch per size class: shared memory cut to the class */
				const uint32_t cap = R.sc_cls[kind][ci], hi = ci + 1 == R.sc_ncls[kind] ? 0xffffffffu : cap;
				RT_LAUNCH(k_sortchain, n_seq, 32, 16 * cap + 2048, ctx->stream, P, ctx->d_reads, n_seq, ctx->d_ws, ctx->d_frames, round, cap, lo, hi);
				lo = cap; S.n_launches++;
			}
		}
		if(first && round == 0) { RT_LAUNCH(k_rlen_predict, 1, MAB_PIPE_THREADS, 0, ctx->stream, P, ctx->d_reads, n_seq, (const uint8_t *)ctx->d_ws, R.rlen_init, R.init_known); S.n_launches++; }
		RT_MEMSET_ASYNC(&ctx->d_ctr->work_next, 0, sizeof(unsigned int), ctx->stream);
		if(ev) { RT_EVENT_RECORD(ctx->rev[3 * round + 1], ctx->stream); }
		if(ctx->n_slots <= ctx->n_sm * MAB_WARPS_PER_CTA * 4 && ctx->ext_wide) {
			RT_LAUNCH((k_extend<4>), R.ext_ctas, 32 * MAB_WARPS_PER_CTA, 1024 + 4 * MAB_TILE_WORDS * MAB_WARPS_PER_CTA, ctx->stream, P, R.d_base, (const uint8_t *)ctx->d_ntail, ctx->d_reads, (const uint32_t *)ctx->d_order, n_seq, ctx->d_ws,
				ctx->d_arenas, R.arena_stride, R.blk_cap, ctx->d_pool, ctx->pool_cap / 4, ctx->d_ctr, round, P.n_occ - 1);
		} else {
			RT_LAUNCH((k_extend<MAB_EXT_CTAS_PER_SM>), R.ext_ctas, 32 * MAB_WARPS_PER_CTA, 1024 + 4 * MAB_TILE_WORDS * MAB_WARPS_PER_CTA, ctx->stream, P, R.d_base, (const uint8_t *)ctx->d_ntail, ctx->d_reads, (const uint32_t *)ctx->d_order, n_seq, ctx->d_ws,
				ctx->d_arenas, R.arena_stride, R.blk_cap, ctx->d_pool, ctx->pool_cap / 4, ctx->d_ctr, round, P.n_occ - 1);
		}
		S.n_launches += 1;
		if(ev) { RT_EVENT_RECORD(ctx->rev[3 * round + 2], ctx->stream); }
	}
	if(timed && first) { RT_EVENT_RECORD(ctx->ev[4], ctx->stream); }
}

/* verification of the rlen speculation (+ redo passes until it holds) against the value the previous batch left behind;
 * init_known = 0: that value is not known yet, the first chain-loading read is only reported (ctx->hc.fd_*).  Ends with the
 * batch counters in ctx->hc; a buffer overflow is left for the caller to see there. */
static int pipe_verify(mab_ctx *ctx, uint32_t rlen_init, uint32_t init_known)
{
	mab_stats_t &S = ctx->stats;
	BatchCounters *pin_ctr = (BatchCounters *)(ctx->pin + ((ctx->pin_user + 127) & ~127ull));
	for(;;) {
		double t_sub = RT_WALL_MS();
		CK(RT_MEMSET_ASYNC(&ctx->d_ctr->n_redo, 0, sizeof(unsigned int), ctx->stream));
		RT_LAUNCH(k_rlen_verify, 1, MAB_PIPE_THREADS, 0, ctx->stream, ctx->d_reads, ctx->rs.sh.n_seq, rlen_init, init_known, ctx->d_ctr);
		S.n_launches++;
		CK(RT_MEMCPY_D2H_ASYNC(pin_ctr, ctx->d_ctr, sizeof(BatchCounters), ctx->stream));
		S.ms_wall_submit += (float)(RT_WALL_MS() - t_sub);
		{ double tw = RT_WALL_MS(); CK(ctx_sync(ctx)); S.ms_wall_wait += (float)(RT_WALL_MS() - tw); }
		ctx->hc = *pin_ctr;
		S.d2h_bytes += sizeof(BatchCounters);
		const BatchCounters &hc = ctx->hc;
		if((hc.err_any & (MAB_ERR_POOL_OVF | MAB_ERR_WS_OVF)) || hc.pool_top > ctx->pool_cap / 4 || hc.n_redo == 0) { break; }
		S.n_retry += hc.n_redo;
		trace_line(ctx, "rlen redo, reads", hc.n_redo);
		pipe_rounds(ctx, false, false);
	}
	return MAB_OK;
}

/* One batch through the device pipeline: scan -> size -> expand -> rounds of sort/chain + extend -> rlen verification (+ redo
 * passes).  On entry ctx->d_reads holds n_seq records with seq_ofs / len set (k_reads_init or the text parser); nothing between
 * the first launch and the verification needs the host: workspace offsets, the work order and the `rlen` speculation are
 * computed on the device, buffers are sized from the shape and from high-water marks of earlier batches, and a buffer that
 * turns out too small (result pool, workspace) is grown and the batch re-run.  On return ctx->hc holds the batch counters. */
static int pipeline_run(mab_ctx *ctx, const uint8_t *d_base, const PipeShape &sh, uint32_t rlen_init, uint32_t init_known, bool timed, double reserve = 1.0)
{
	const DevParams &P = ctx->P;
	mab_stats_t &S = ctx->stats;
	RunState &R = ctx->rs;
	const uint32_t n_seq = sh.n_seq;
	double t_sub = RT_WALL_MS();
	R.sh = sh; R.d_base = d_base; R.rlen_init = rlen_init; R.init_known = init_known;
	/* buffers are sized for `reserve` (>= 1) times this batch: the largest chunk the caller has announced (mab_text_reserve), so
	 * that a context whose first chunk happens to be a short one does not re-allocate everything on its second */
	reserve = std::max(1.0, reserve);
	const uint64_t z_seq = reserve > 1.0 ? (uint64_t)((double)n_seq * reserve * 1.05) + 64 : n_seq;
	const uint64_t z_span = (uint64_t)((double)sh.span * reserve), z_len = (uint64_t)((double)sh.tot_len * reserve);
	double ws_per_base; bool ws_seen; uint32_t sc_c0[2];
	{ std::lock_guard<std::mutex> lk(ctx->cal->mu); ws_per_base = ctx->cal->ws_per_base; ws_seen = ctx->cal->ws_seen; sc_c0[0] = ctx->cal->sc_c0[0]; sc_c0[1] = ctx->cal->sc_c0[1]; }
	{ int rc = pin_reserve(ctx, ctx->pin_user + 2 * sizeof(BatchCounters) + 256); if(rc) { return rc; } }
	{ int rc = grow(&ctx->d_recs, &ctx->recs_cap, 16ull * z_span + 256); if(rc) { return rc; } }
	{ int rc = grow(&ctx->d_frames, &ctx->frames_cap, 4ull * 8 * MAB_RS_FRAME * (z_seq + 8)); if(rc) { return rc; } }
	{ int rc = grow(&ctx->d_order, &ctx->order_cap, 4ull * z_seq); if(rc) { return rc; } }
	R.blk_cap = dp_blk_cap(sh.maxlen);
	ArenaLayout AL = arena_layout(R.blk_cap);
	R.arena_stride = AL.total;
	R.ext_ctas = std::max<uint32_t>(1, std::min<uint32_t>(ctx->n_slots / MAB_WARPS_PER_CTA, (n_seq + MAB_WARPS_PER_CTA - 1) / MAB_WARPS_PER_CTA));
	{	/* the DP arenas are sized by the longest read of the batch (312 B per base per resident warp): with very long reads fewer
		 * warps stay resident rather than asking for more than MAB_ARENA_BUDGET bytes of HBM per context */
		uint64_t budget = ctx->arena_budget;
		if(const char *e = getenv("MAB_ARENA_BUDGET_MB")) { long v = atol(e); if(v > 0) { budget = (uint64_t)v << 20; } }
		uint64_t fit = budget / (AL.total * MAB_WARPS_PER_CTA);
		if(fit < R.ext_ctas) { R.ext_ctas = (uint32_t)std::max<uint64_t>(1, fit); }
	}
	{ int rc = grow(&ctx->d_arenas, &ctx->arenas_cap, AL.total * (uint64_t)R.ext_ctas * MAB_WARPS_PER_CTA); if(rc) { return rc; } }
	uint64_t pool_need = z_len / 4 + 64ull * z_seq + (1u << 16);					/* ~2 bits per base and alignment, x4 head room */
	uint64_t ws_need = (uint64_t)(ws_per_base * (double)z_len) + 20480ull * z_seq + 4096;
	R.seed_ctas = std::max<uint32_t>(1, std::min<uint32_t>((n_seq + MAB_WARPS_PER_CTA - 1) / MAB_WARPS_PER_CTA, ctx->n_sm * 8));
	R.init_ctas = std::max<uint32_t>(1, std::min<uint32_t>((n_seq + 255) / 256, ctx->n_sm * 4));
	/* k_sortchain size classes, per round kind (round 0 stages a read's own seeds, later rounds all of them).  A read's seed array
	 * is staged in shared memory, so the reads resident per SM are what 228 KB holds: every class is launched with the footprint of
	 * its largest member.  Classes: the median seed bound of the previous batch, then steps of x1.4 up to MAB_SC_MAX seeds (96 KB);
	 * beyond that the read works in global memory. */
	for(int k = 0; k < 2; k++) {
		uint32_t c = std::max<uint32_t>(64u, std::min<uint32_t>(sc_c0[k], MAB_SC_MAX)), n = 0;
		while(n + 1 < MAB_SC_CLASSES && c < MAB_SC_MAX) { R.sc_cls[k][n++] = c; c = std::min<uint32_t>(MAB_SC_MAX, ((c * 7 / 5) + 63u) & ~63u); }
		R.sc_cls[k][n++] = MAB_SC_MAX;
		R.sc_ncls[k] = n;
	}
	if(const char *e = getenv("MAB_SC_CAP")) {										/* test hook: tiny caps push reads through the unstaged (global memory) path */
		uint32_t v = (uint32_t)atoi(e);
		if(v >= 64) { for(int k = 0; k < 2; k++) { R.sc_cls[k][0] = v; R.sc_ncls[k] = 1; } }
	}
	S.ms_wall_sizing += (float)(RT_WALL_MS() - t_sub);
	RT_LAUNCH(k_order, 1, MAB_PIPE_THREADS, 0, ctx->stream, (const ReadRec *)ctx->d_reads, n_seq, ctx->d_order);
	S.n_launches++;
	for(int attempt = 0; ; attempt++) {
		t_sub = RT_WALL_MS();
		{ int rc = grow(&ctx->d_pool, &ctx->pool_cap, 4 * pool_need); if(rc) { return rc; } }
		{ int rc = grow(&ctx->d_ws, &ctx->ws_cap, ws_need); if(rc) { return rc; } }
		if(attempt > 0) { RT_LAUNCH(k_reads_reset, R.init_ctas, 256, 0, ctx->stream, ctx->d_reads, n_seq); S.n_launches++; }
		CK(RT_MEMSET_ASYNC(ctx->d_ctr, 0, sizeof(BatchCounters), ctx->stream));
		if(timed) { RT_EVENT_RECORD(ctx->ev[1], ctx->stream); }
		RT_LAUNCH(k_seed_scan, R.seed_ctas, 32 * MAB_WARPS_PER_CTA, 2560 * MAB_WARPS_PER_CTA, ctx->stream, P, d_base, ctx->d_reads, n_seq, ctx->d_recs);
		RT_LAUNCH(k_seed_probe, R.seed_ctas, 32 * MAB_WARPS_PER_CTA, 0, ctx->stream, P, ctx->d_reads, n_seq, ctx->d_recs);
		RT_LAUNCH(k_size, 1, MAB_PIPE_THREADS, 0, ctx->stream, ctx->d_reads, n_seq, ctx->ws_cap, ctx->d_ctr);
		S.n_launches += 3;
		if(!ws_seen) {
			/* first batch on this index: how many seeds a read base yields is not known yet (4 to 40 workspace bytes per base from a
			 * bacterial to a human index), so the host looks at k_size's total here, before the expensive stages, instead of finding the
			 * workspace too small after them.  Later batches go through without this wait on the high-water mark. */
			BatchCounters *pin_ctr = (BatchCounters *)(ctx->pin + ((ctx->pin_user + 127) & ~127ull));
			CK(RT_MEMCPY_D2H_ASYNC(pin_ctr, ctx->d_ctr, sizeof(BatchCounters), ctx->stream));
			{ double tw = RT_WALL_MS(); CK(ctx_sync(ctx)); S.ms_wall_wait += (float)(RT_WALL_MS() - tw); }
			S.d2h_bytes += sizeof(BatchCounters);
			ws_seen = true;
			if(pin_ctr->ws_need > ctx->ws_cap) {
				ws_need = (uint64_t)((double)pin_ctr->ws_need * reserve); ws_need += ws_need / 8 + 4096;
				trace_line(ctx, "first batch: workspace sized after the seed scan, MB", ws_need / 1048576.0);
				continue;
			}
		}
		pipe_rounds(ctx, true, timed);
		S.ms_wall_submit += (float)(RT_WALL_MS() - t_sub);
		{ int rc = pipe_verify(ctx, rlen_init, init_known); if(rc) { return rc; } }
		const BatchCounters &hc = ctx->hc;
		if(hc.ws_need > ctx->ws_cap || (hc.err_any & MAB_ERR_WS_OVF)) {					/* workspace estimate too small: now it is known exactly */
			ws_need = hc.ws_need + hc.ws_need / 8 + 4096; S.n_retry++;
			trace_line(ctx, "workspace overflow: re-run, MB", ws_need / 1048576.0);
			if(attempt >= 3) { g_err = "workspace overflow after retries"; return MAB_EOVERFLOW; }
			continue;
		}
		if((hc.err_any & MAB_ERR_POOL_OVF) || hc.pool_top > ctx->pool_cap / 4) {
			pool_need = std::max<uint64_t>(4 * pool_need, hc.pool_top + hc.pool_top / 2); S.n_retry++;
			trace_line(ctx, "pool overflow: re-run, MB", 4.0 * pool_need / 1048576.0);
			if(attempt >= 3) { g_err = "result pool overflow after retries"; return MAB_EOVERFLOW; }
			continue;
		}
		S.n_vectors += hc.n_vectors; S.n_fill_calls += hc.n_fill; S.n_trace += hc.n_trace;
		if(sh.tot_len) {																	/* high-water mark for the next batch's estimate */
			double per_base = ((double)hc.ws_need - 20480.0 * n_seq) / (double)sh.tot_len;
			std::lock_guard<std::mutex> lk(ctx->cal->mu);
			if(per_base * 1.25 > ctx->cal->ws_per_base) { ctx->cal->ws_per_base = per_base * 1.25; }
			ctx->cal->ws_seen = true;
		}
		break;
	}
	return MAB_OK;
}

/* size classes of k_sortchain for the next batch from this batch's per-read seed bounds */
static void update_sc_caps(mab_ctx *ctx, const ReadRec *hr, uint32_t n_seq)
{
	for(int kind = 0; kind < 2; kind++) {
		std::vector<uint32_t> bnd;
		bnd.reserve(n_seq);
		for(uint32_t i = 0; i < n_seq; i++) { if(hr[i].len >= ctx->P.k && hr[i].seed_cap != 0) { bnd.push_back((kind == 0 ? hr[i].tot_seeds0 : hr[i].tot_seeds) + 2); } }
		if(bnd.size() < 16) { continue; }
		size_t k50 = bnd.size() / 2;
		std::nth_element(bnd.begin(), bnd.begin() + k50, bnd.end());
		std::lock_guard<std::mutex> lk(ctx->cal->mu);
		ctx->cal->sc_c0[kind] = std::max<uint32_t>(64u, std::min<uint32_t>((bnd[k50] + 63u) & ~63u, MAB_SC_SMALL));
	}
}

extern "C" int mab_map_batch(mab_ctx *ctx, const uint8_t *seq_block, uint64_t block_size, const uint64_t *seq_ofs, const uint32_t *seq_len, uint32_t n_seq)
{
	mab_stats_t &S = ctx->stats;
	memset(&S, 0, sizeof(S));
	const double t_call = RT_WALL_MS();
	try { ctx->res_ofs.assign((size_t)n_seq + 1, 0); } catch(const std::bad_alloc &) { g_err = "host allocation failed"; return MAB_ENOMEM; }
	if(n_seq == 0) { return MAB_OK; }
	CK(RT_USE_DEVICE(ctx->device));
	PipeShape sh; sh.n_seq = n_seq; sh.maxlen = 0; sh.tot_len = 0; sh.span = 0;
	for(uint32_t i = 0; i < n_seq; i++) {
		sh.maxlen = std::max(sh.maxlen, seq_len[i]); sh.tot_len += seq_len[i]; sh.span = std::max<uint64_t>(sh.span, seq_ofs[i] + seq_len[i]);
		/* contract (header): ascending, non-overlapping, 64 readable bytes behind the last read inside the block */
		if(seq_ofs[i] + seq_len[i] + 64 > block_size || (i > 0 && seq_ofs[i] < seq_ofs[i - 1] + seq_len[i - 1])) { g_err = "mab_map_batch: reads must be ascending and non-overlapping, with 64 bytes of margin behind the last one inside the block"; return MAB_EINVAL; }
	}
	RT_EVENT_RECORD(ctx->ev[0], ctx->stream);
	const uint8_t *d_base;
	if(ctx->device_input) { d_base = seq_block; }
	else {
		int rc = grow(&ctx->d_seq, &ctx->seq_cap, block_size + 256); if(rc) { return rc; }
		CK(RT_MEMCPY_H2D_ASYNC(ctx->d_seq, seq_block, block_size, ctx->stream));
		d_base = ctx->d_seq; S.h2d_bytes += block_size;
	}
	const uint64_t rr_bytes = sizeof(ReadRec) * (uint64_t)n_seq;
	ctx->pin_user = rr_bytes;																/* [ReadRec x n] doubles as [ofs u64 x n][len u32 x n] on the way in */
	{ int rc = pin_reserve(ctx, ctx->pin_user + 2 * sizeof(BatchCounters) + 256); if(rc) { return rc; } }
	{ int rc = grow(&ctx->d_reads, &ctx->reads_cap, rr_bytes); if(rc) { return rc; } }
	{ int rc = grow(&ctx->d_io, &ctx->io_cap, 12ull * n_seq + 64); if(rc) { return rc; } }
	{
		uint64_t *po = (uint64_t *)ctx->pin; uint32_t *pl = (uint32_t *)(ctx->pin + 8ull * n_seq);
		memcpy(po, seq_ofs, 8ull * n_seq); memcpy(pl, seq_len, 4ull * n_seq);
		CK(RT_MEMCPY_H2D_ASYNC(ctx->d_io, ctx->pin, 12ull * n_seq, ctx->stream));
		S.h2d_bytes += 12ull * n_seq;
		uint32_t init_ctas = std::max<uint32_t>(1, std::min<uint32_t>((n_seq + 255) / 256, ctx->n_sm * 4));
		RT_LAUNCH(k_reads_init, init_ctas, 256, 0, ctx->stream, ctx->d_reads, n_seq, (const uint64_t *)ctx->d_io, (const uint32_t *)(ctx->d_io + 8ull * n_seq));
		S.n_launches++;
	}
	{ int rc = pipeline_run(ctx, d_base, sh, ctx->rlen_last, 1u, true); if(rc) { return rc; } }
	const BatchCounters &hc = ctx->hc;
	if(hc.chain_valid) { ctx->rlen_last = hc.chain_rlen; }
	/* results: per-read records and the pool */
	ReadRec *pin_rr = (ReadRec *)ctx->pin;
	uint64_t top = hc.pool_top;
	if(4 * (top + 4) > ctx->h_pool_cap) {
		RT_HOST_FREE(ctx->h_pool); ctx->h_pool = nullptr; ctx->h_pool_cap = 0;
		uint64_t nb = 4 * (top + 4) + (top + 4);											/* 25 % head room */
		if(!RT_OK(RT_HOST_ALLOC(&ctx->h_pool, nb))) { g_err = std::string("pinned host allocation failed: ") + RT_ERRSTR(); return MAB_ENOMEM; }
		ctx->h_pool_cap = nb;
	}
	CK(RT_MEMCPY_D2H_ASYNC(pin_rr, ctx->d_reads, rr_bytes, ctx->stream));
	if(top) { CK(RT_MEMCPY_D2H_ASYNC(ctx->h_pool, ctx->d_pool, 4 * top, ctx->stream)); }
	RT_EVENT_RECORD(ctx->ev[5], ctx->stream);
	{ double tw = RT_WALL_MS(); CK(ctx_sync(ctx)); S.ms_wall_wait += (float)(RT_WALL_MS() - tw); }
	S.d2h_bytes += 4 * top + rr_bytes;
	const ReadRec *hr = pin_rr;
	update_sc_caps(ctx, hr, n_seq);
	uint32_t n_failed = 0;
	for(uint32_t i = 0; i < n_seq; i++) { n_failed += hr[i].err != 0; }
	S.n_failed = n_failed;																	/* reads given up on (a per-read device structure overflowed): reported unmapped */
	double t0 = RT_WALL_MS();
	/* host post-processing, fanned out over the host cores: plan every read (no copying), prefix-sum the sizes, then write the
	 * flat records straight into their final place */
	uint32_t hw = std::thread::hardware_concurrency();
	if(const char *e = getenv("MAB_HOST_THREADS")) { int v = atoi(e); if(v > 0) { hw = (uint32_t)v; } }
	uint32_t nth = std::max(1u, std::min<uint32_t>(std::min<uint32_t>(hw, 32u), n_seq / 64 + 1));
	try {
		std::vector<std::vector<PlanItem>> items(nth);
		std::vector<ReadPlan> plan(n_seq);
		auto run_par = [&](auto &&fn) {
			std::vector<std::thread> th;
			for(uint32_t t = 1; t < nth; t++) { th.emplace_back(fn, t); }
			fn(0u);
			for(auto &x : th) { x.join(); }
		};
		run_par([&](uint32_t t) {
			for(uint32_t i = t; i < n_seq; i += nth) {
				if(hr[i].result_words != 0 && hr[i].err == 0) { plan[i] = post_plan(ctx, ctx->h_pool, ctx->h_pool + hr[i].result_ofs, items[t]); }
				else { plan[i].first = 0; plan[i].count = 0; plan[i].n_uniq = 0; plan[i].words = 0; }
			}
		});
		uint64_t total = 0;
		for(uint32_t i = 0; i < n_seq; i++) { ctx->res_ofs[i] = total; total += plan[i].words; }
		ctx->res_ofs[n_seq] = total;
		if(total > ctx->res_cap) { delete[] ctx->res_words; ctx->res_words = nullptr; ctx->res_cap = 0; ctx->res_words = new uint32_t[total + total / 4 + 1024]; ctx->res_cap = total + total / 4 + 1024; }
		run_par([&](uint32_t t) {
			for(uint32_t i = t; i < n_seq; i += nth) {
				if(plan[i].words != 0) { post_emit(plan[i], items[t].data() + plan[i].first, ctx->res_words + ctx->res_ofs[i]); }
			}
		});
	} catch(const std::bad_alloc &) { g_err = "host allocation failed"; return MAB_ENOMEM; }
	S.ms_post = (float)(RT_WALL_MS() - t0);
	S.ms_h2d = RT_EVENT_MS(ctx->ev[0], ctx->ev[1]);
	S.ms_seed = RT_EVENT_MS(ctx->ev[1], ctx->ev[3]);
	S.ms_sortchain = 0.f; S.ms_extend = 0.f; S.ms_extend_r0 = 0.f;
	for(uint32_t r = 0; r < ctx->P.n_occ && r < 8; r++) {
		S.ms_sortchain += RT_EVENT_MS(ctx->rev[3 * r], ctx->rev[3 * r + 1]);
		float e = RT_EVENT_MS(ctx->rev[3 * r + 1], ctx->rev[3 * r + 2]);
		S.ms_extend += e; if(r == 0) { S.ms_extend_r0 = e; }
	}
	S.ms_d2h = RT_EVENT_MS(ctx->ev[4], ctx->ev[5]);
	S.ms_total = RT_EVENT_MS(ctx->ev[0], ctx->ev[5]);
	S.ms_wall = (float)(RT_WALL_MS() - t_call);
	return MAB_OK;
}

extern "C" uint64_t mab_result(const mab_ctx *ctx, uint32_t i, const uint32_t **words)
{
	if((size_t)i + 1 >= ctx->res_ofs.size()) { return 0; }
	uint64_t n = ctx->res_ofs[i + 1] - ctx->res_ofs[i];
	if(words) { *words = n ? ctx->res_words + ctx->res_ofs[i] : nullptr; }
	return n;
}
extern "C" void mab_release_batch(mab_ctx *ctx) { ctx->res_ofs.clear(); }

struct mab_results { uint32_t *words; std::vector<uint64_t> ofs; };
extern "C" mab_results *mab_detach_batch(mab_ctx *ctx)
{
	mab_results *r = new mab_results();
	r->words = ctx->res_words; r->ofs.swap(ctx->res_ofs);
	ctx->res_words = nullptr; ctx->res_cap = 0; ctx->res_ofs.clear();
	return r;
}
extern "C" uint64_t mab_results_get(const mab_results *r, uint32_t i, const uint32_t **words)
{
	if((size_t)i + 1 >= r->ofs.size()) { return 0; }
	uint64_t n = r->ofs[i + 1] - r->ofs[i];
	if(words) { *words = n ? r->words + r->ofs[i] : nullptr; }
	return n;
}
extern "C" void mab_results_free(mab_results *r) { if(r) { delete[] r->words; delete r; } }

/* ---------------------------------------------------------------- stage-level entry points */
extern "C" uint64_t mab_sketch(mab_ctx *ctx, const uint8_t *seq, uint32_t len, uint64_t *out, uint64_t cap)
{
	uint8_t *d_seq = nullptr; uint64_t *d_out = nullptr, *d_n = nullptr;
	uint64_t n = 0, dcap = (uint64_t)len + 8;
	if(!RT_OK(RT_USE_DEVICE(ctx->device))) { return 0; }
	if(!RT_OK(RT_MALLOC(&d_seq, (uint64_t)len + 64)) || !RT_OK(RT_MALLOC(&d_out, 8 * dcap)) || !RT_OK(RT_MALLOC(&d_n, 8))) { return 0; }
	RT_MEMCPY_H2D_ASYNC(d_seq, seq, len, ctx->stream);
	RT_LAUNCH(k_sketch_words, 1, 32, 512, ctx->stream, ctx->P, (const uint8_t *)d_seq, len, d_out, dcap, d_n);
	ctx_sync(ctx);
	RT_MEMCPY_D2H_ASYNC(&n, d_n, 8, ctx->stream);
	if(n <= cap) { RT_MEMCPY_D2H_ASYNC(out, d_out, 8 * n, ctx->stream); }
	RT_FREE(d_seq); RT_FREE(d_out); RT_FREE(d_n);
	return n;
}

extern "C" uint64_t mab_seed_chain(mab_ctx *ctx, const uint8_t *seq, uint32_t len, uint32_t round,
	uint32_t *seeds, uint64_t seed_cap, uint64_t *n_total, uint32_t *roots, uint64_t root_cap, uint64_t *n_root)
{
	const DevParams &P = ctx->P;
	*n_total = 0; *n_root = 0;
	if(!RT_OK(RT_USE_DEVICE(ctx->device))) { return 0; }
	uint8_t *d_seq = nullptr; ReadRec *d_r = nullptr; uint8_t *d_ws = nullptr; uint32_t *d_fr = nullptr;
	if(!RT_OK(RT_MALLOC(&d_seq, (uint64_t)len + 256)) || !RT_OK(RT_MALLOC(&d_r, sizeof(ReadRec))) || !RT_OK(RT_MALLOC(&d_fr, 4ull * 8 * MAB_RS_FRAME))) { return 0; }
	RT_MEMCPY_H2D_ASYNC(d_seq, seq, len, ctx->stream);
	ReadRec r; memset(&r, 0, sizeof(r)); r.len = len;
	RT_MEMCPY_H2D_ASYNC(d_r, &r, sizeof(r), ctx->stream);
	uint32_t *d_rec = nullptr;
	if(!RT_OK(RT_MALLOC(&d_rec, 16ull * len + 256))) { RT_FREE(d_seq); RT_FREE(d_r); RT_FREE(d_fr); return 0; }
	RT_LAUNCH(k_seed_scan, 1, 32, 2560, ctx->stream, P, (const uint8_t *)d_seq, d_r, 1u, d_rec);
	RT_LAUNCH(k_seed_probe, 1, 32, 0, ctx->stream, P, d_r, 1u, d_rec);
	RT_MEMCPY_D2H_ASYNC(&r, d_r, sizeof(r), ctx->stream);
	ctx_sync(ctx);
	uint64_t ns = 0;
	if(r.state == 0) {
		r.seed_cap = 2 * (r.tot_seeds + 1) + 8; r.root_cap = r.tot_seeds + 8; r.resc_cap = r.tot_resc + 4; r.bin_cap = 2 * r.tot_seeds + 128; r.ws_ofs = 0;
		WsLayout L = ws_layout(r.seed_cap, r.root_cap, r.resc_cap, r.bin_cap);
		RT_MALLOC(&d_ws, L.total + 256);
		RT_MEMCPY_H2D_ASYNC(d_r, &r, sizeof(r), ctx->stream);
		RT_LAUNCH(k_seed_expand, 1, 32, 0, ctx->stream, P, d_r, 1u, d_ws, (const uint32_t *)d_rec);
		{
			uint32_t sc_cap = std::max<uint32_t>(64u, std::min<uint32_t>(r.tot_seeds + 2, MAB_SC_SMALL));
			uint32_t *d_ord = nullptr; RT_MALLOC(&d_ord, 64); RT_MEMSET_ASYNC(d_ord, 0, 64, ctx->stream);
			for(uint32_t i = 0; i <= round && i < P.n_occ; i++) {
				if(ctx->sort_walk) {
					uint32_t cap = 1024; while(cap < r.tot_seeds + 2 && cap < 32768) { cap *= 2; }
					RT_LAUNCH(k_sort, 1, 32, 4 * MAB_WK_SM_WORDS + cap, ctx->stream, P, d_r, (const uint32_t *)d_ord, 1u, d_ws, d_fr, i, cap, 0u, 0xffffffffu);
					RT_LAUNCH(k_chain, 1, 32, 16 * sc_cap + 2048, ctx->stream, P, d_r, (const uint32_t *)d_ord, 1u, d_ws, d_fr, sc_cap, 0u, 0xffffffffu, ctx->chain_warp ? 1u : 0u);
				} else { RT_LAUNCH(k_sortchain, 1, 32, 16 * sc_cap + 2048, ctx->stream, P, d_r, 1u, d_ws, d_fr, i, sc_cap, 0u, 0xffffffffu); }
			}
			ctx_sync(ctx); RT_FREE(d_ord);
		}
		RT_MEMCPY_D2H_ASYNC(&r, d_r, sizeof(r), ctx->stream);
		ctx_sync(ctx);
		if(r.n_seed) {
			ns = r.n_seed; *n_total = r.seed_n; *n_root = r.n_root;
			if(r.seed_n <= seed_cap) { RT_MEMCPY_D2H_ASYNC(seeds, d_ws + L.seed, 16ull * r.seed_n, ctx->stream); }
			if(r.n_root && r.n_root <= root_cap) { RT_MEMCPY_D2H_ASYNC(roots, d_ws + L.root, 8ull * r.n_root, ctx->stream); }
			ctx_sync(ctx);
		}
		RT_FREE(d_ws);
	}
	RT_FREE(d_seq); RT_FREE(d_r); RT_FREE(d_fr); RT_FREE(d_rec);
	return ns;
}

extern "C" int mab_extend_pairs(mab_ctx *ctx, const uint8_t *seq_block, uint64_t block_size, const mab_pair_t *pairs, uint32_t n,
	uint32_t *res, uint32_t *aln_out, uint64_t aln_cap, uint64_t *aln_ofs)
{
	const DevParams &P = ctx->P;
	if(n == 0) { aln_ofs[0] = 0; return MAB_OK; }
	CK(RT_USE_DEVICE(ctx->device));
	uint32_t maxlen = 0; uint64_t tot = 0;
	for(uint32_t i = 0; i < n; i++) { maxlen = std::max(maxlen, std::max(pairs[i].alen, pairs[i].blen)); tot += pairs[i].alen + pairs[i].blen; }
	uint32_t blk_cap = dp_blk_cap(maxlen);
	ArenaLayout AL = arena_layout(blk_cap);
	uint32_t ctas = std::max<uint32_t>(1, std::min<uint32_t>(ctx->n_slots / MAB_WARPS_PER_CTA, (n + MAB_WARPS_PER_CTA - 1) / MAB_WARPS_PER_CTA));
	uint64_t pool_words = tot / 2 + 256ull * n + 4096;
	uint8_t *d_seq = nullptr, *d_ar = nullptr; PairIn *d_p = nullptr; uint32_t *d_res = nullptr, *d_pool = nullptr; uint64_t *d_ao = nullptr;
	static_assert(sizeof(PairIn) == sizeof(mab_pair_t), "pair layout");
	CK(RT_MALLOC(&d_seq, block_size + 256)); CK(RT_MALLOC(&d_ar, AL.total * (uint64_t)ctas * MAB_WARPS_PER_CTA)); CK(RT_MALLOC(&d_p, sizeof(PairIn) * (uint64_t)n));
	CK(RT_MALLOC(&d_res, 64ull * n)); CK(RT_MALLOC(&d_pool, 4 * pool_words)); CK(RT_MALLOC(&d_ao, 8ull * n));
	CK(RT_MEMCPY_H2D_ASYNC(d_seq, seq_block, block_size, ctx->stream)); CK(RT_MEMCPY_H2D_ASYNC(d_p, pairs, sizeof(PairIn) * (uint64_t)n, ctx->stream));
	BatchCounters zero; memset(&zero, 0, sizeof(zero));
	CK(RT_MEMCPY_H2D_ASYNC(ctx->d_ctr, &zero, sizeof(zero), ctx->stream));
	RT_LAUNCH(k_extend_pairs, ctas, 32 * MAB_WARPS_PER_CTA, 1024 + 4 * MAB_TILE_WORDS * MAB_WARPS_PER_CTA, ctx->stream, P, (const uint8_t *)d_seq, (const uint8_t *)ctx->d_ntail, (const PairIn *)d_p, n, d_res, d_ao,
		d_ar, AL.total, blk_cap, d_pool, pool_words, ctx->d_ctr);
	CK(ctx_sync(ctx));
	BatchCounters hc; CK(RT_MEMCPY_D2H_ASYNC(&hc, ctx->d_ctr, sizeof(hc), ctx->stream));
	std::vector<uint32_t> pool((size_t)std::min<uint64_t>(hc.pool_top, pool_words) + 4);
	std::vector<uint64_t> ao(n);
	CK(RT_MEMCPY_D2H_ASYNC(res, d_res, 64ull * n, ctx->stream)); CK(RT_MEMCPY_D2H_ASYNC(ao.data(), d_ao, 8ull * n, ctx->stream));
	if(hc.pool_top) { CK(RT_MEMCPY_D2H_ASYNC(pool.data(), d_pool, 4 * std::min<uint64_t>(hc.pool_top, pool_words), ctx->stream)); }
	ctx->stats.n_vectors = hc.n_vectors;
	uint64_t o = 0;
	int rc = (hc.err_any || hc.pool_top > pool_words) ? MAB_EOVERFLOW : MAB_OK;
	for(uint32_t i = 0; i < n; i++) {
		aln_ofs[i] = o;
		if(ao[i] == 0xffffffffffffffffull || rc != MAB_OK) { continue; }
		const uint32_t *a = pool.data() + ao[i];
		uint32_t slen = a[7], npw = a[9], sn = a[10];
		uint64_t need = 16 + 8ull * slen + npw;
		if(o + need > aln_cap) { rc = MAB_ENOMEM; break; }
		uint32_t *w = aln_out + o;
		for(int t = 0; t < 10; t++) { w[t] = a[t]; }
		for(int t = 10; t < 16; t++) { w[t] = 0; }
		memcpy(w + 16, a + MAB_ALN_HDR + 8ull * (sn - slen), 32ull * slen);
		memcpy(w + 16 + 8ull * slen, a + MAB_ALN_HDR + 8ull * sn, 4ull * npw);
		o += need;
	}
	aln_ofs[n] = o;
	RT_FREE(d_seq); RT_FREE(d_ar); RT_FREE(d_p); RT_FREE(d_res); RT_FREE(d_pool); RT_FREE(d_ao);
	if(rc == MAB_EOVERFLOW) { g_err = "extend_pairs: device workspace overflow"; }
	return rc;
}

extern "C" int mab_fill_peak(mab_ctx *ctx, int masks, uint32_t n_blocks, double *vectors_per_s)
{
	CK(RT_USE_DEVICE(ctx->device));
	uint32_t ctas = std::max<uint32_t>(1, ctx->n_slots / MAB_WARPS_PER_CTA), warps = ctas * MAB_WARPS_PER_CTA;
	uint32_t *d_ring = nullptr, *d_sink = nullptr;
	CK(RT_MALLOC(&d_ring, 4ull * 512 * 4 * warps)); CK(RT_MALLOC(&d_sink, 4ull * warps));
	for(int rep = 0; rep < 2; rep++) {												/* first launch warms up */
		RT_EVENT_RECORD(ctx->ev[6], ctx->stream);
		if(masks) { RT_LAUNCH((k_fill_peak<true>), ctas, 32 * MAB_WARPS_PER_CTA, 1024, ctx->stream, ctx->P, d_ring, n_blocks, d_sink); }
		else { RT_LAUNCH((k_fill_peak<false>), ctas, 32 * MAB_WARPS_PER_CTA, 1024, ctx->stream, ctx->P, d_ring, n_blocks, d_sink); }
		RT_EVENT_RECORD(ctx->ev[7], ctx->stream);
		CK(ctx_sync(ctx));
	}
	float ms = RT_EVENT_MS(ctx->ev[6], ctx->ev[7]);
	RT_FREE(d_ring); RT_FREE(d_sink);
	*vectors_per_s = ms > 0.f ? (double)warps * n_blocks * MAB_BLK / (ms * 1e-3) : 0.0;
	return MAB_OK;
}

/* debugging / test entry: runs k_selftest and copies its 64 x 32 words out */
extern "C" int mab_selftest(mab_ctx *ctx, uint32_t *out)
{
	uint32_t *d = nullptr;
	CK(RT_USE_DEVICE(ctx->device));
	CK(RT_MALLOC(&d, 4ull * 64 * 32));
	CK(RT_MEMSET_ASYNC(d, 0, 4ull * 64 * 32, ctx->stream));
	RT_LAUNCH(k_selftest, 1, 32, 0, ctx->stream, d);
	CK(ctx_sync(ctx));
	CK(RT_MEMCPY_D2H_ASYNC(out, d, 4ull * 64 * 32, ctx->stream));
	CK(ctx_sync(ctx));
	RT_FREE(d);
	return MAB_OK;
}

/* test entry: n 16-byte elements (key = words 1:0) sorted by the reference's sort as the device walks it (out_exact) and by its
 * parallel form (out_walk); n <= 32767 */
extern "C" int mab_sort_check(mab_ctx *ctx, const uint32_t *elems, uint32_t n, uint32_t *out_exact, uint32_t *out_walk)
{
	if(n == 0 || n > 32767) { g_err = "mab_sort_check: 1..32767 elements"; return MAB_EINVAL; }
	CK(RT_USE_DEVICE(ctx->device));
	uint32_t *a = nullptr, *b = nullptr, *fr = nullptr, *e = nullptr; uint16_t *sc = nullptr;
	CK(RT_MALLOC(&a, 16ull * n)); CK(RT_MALLOC(&b, 32ull * n)); CK(RT_MALLOC(&fr, 4ull * 8 * MAB_RS_FRAME)); CK(RT_MALLOC(&sc, 4ull * (n + 16))); CK(RT_MALLOC(&e, 64));
	CK(RT_MEMCPY_H2D(a, elems, 16ull * n)); CK(RT_MEMCPY_H2D(b, elems, 16ull * n));
	CK(RT_DEVICE_SYNC());
	RT_LAUNCH(k_sort_check, 1, 32, 4 * MAB_WK_SM_WORDS + 32768, ctx->stream, a, b, n, fr, sc, e);
	CK(ctx_sync(ctx));
	uint32_t err = 0;
	CK(RT_MEMCPY_D2H(out_exact, a, 16ull * n)); CK(RT_MEMCPY_D2H(out_walk, b, 16ull * n)); CK(RT_MEMCPY_D2H(&err, e, 4));
	RT_FREE(a); RT_FREE(b); RT_FREE(fr); RT_FREE(sc); RT_FREE(e);
	if(err) { g_err = "mab_sort_check: frame stack overflow"; return MAB_EOVERFLOW; }
	return MAB_OK;
}

#include "mab_text_host.inl"
