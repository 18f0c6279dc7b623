/*
 * mab_dp.cuh -- warp-parallel GABA engine (adaptive-banded, difference-recurrence, piecewise-affine SW-Gotoh) for sm_100a.
 *
 * One warp owns one band.  A band of W cells (W = 64, or 32 / 16 for the reference's "narrow" retries) is spread over the
 * warp two cells per lane, each cell a signed 16-bit half of a 32-bit register, so the whole recurrence runs on
 * Blackwell's native packed-16x2 integer SIMD (VIADD.16x2 / VIMNMX.S16x2 / VIMNMX3.S16x2 / VIADDMNMX.S16x2); the
 * reference's int8 wrap-around is re-imposed only where it is observable (the per-block delta accumulator and the
 * saturating drop vector).  Band shifts are one SHFL + one PRMT per shifted vector, the direction accumulator is a single
 * REDUX per anti-diagonal.  Semantics follow gaba.c (COMBINED model) and are checked against oracle/gaba_oracle.c;
 * citations are into /root/reference/gaba.c.
 *
 * Everything here is executed by all 32 lanes with warp-uniform control flow.
 */
#pragma once
#include "mab_types.h"
#include "mab_scalar.cuh"

namespace mab {

#define MAB_FULL 0xffffffffu
#ifdef MAB_DEBUG_DEV
#define MAB_DBG(...) do { if(c.lane == 0) { printf(__VA_ARGS__); } } while(0)
#else
#define MAB_DBG(...) do {} while(0)
#endif

/* PTX prmt.b32 in default mode: selector nibble bit 3 replicates the sign bit of the selected byte.  (__byte_perm masks the
 * selector with 0x7777, so the sign-extending permutes below have to be issued as PTX.) */
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
#ifdef MAB_EMU
	return emu_prmt(a, b, sel);
#else
	uint32_t d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
	return d;
#endif
}
__device__ __forceinline__ uint32_t pack2(int v) { uint32_t x = (uint32_t)(uint16_t)(int16_t)v; return x | (x << 16); }
__device__ __forceinline__ uint32_t sext8x2(uint32_t x) { return prmt(x, 0, 0xA280); }		/* int8 wrap of both halves */
__device__ __forceinline__ uint32_t unpack8(uint32_t v) { return prmt(v, 0, 0x9180); }		/* {b0,b1} -> s16x2 */
__device__ __forceinline__ uint32_t pack8(uint32_t x) { return __byte_perm(x, 0, 0x4420); }			/* s16x2 -> {b0,b1} */
/* "H8" form: the int8 value sits in the HIGH byte of its 16-bit half (low byte zero), so VIADD.16x2 wraps exactly like the
 * reference's int8 lanes (observable at the band edges, where dv/de run away) while signed 16-bit max/compare keep order */
__device__ __forceinline__ uint32_t unpack8h(uint32_t v) { return __byte_perm(v, 0, 0x1404); }		/* {b0,b1} -> H8 */
__device__ __forceinline__ uint32_t pack8h(uint32_t x) { return __byte_perm(x, 0, 0x4431); }			/* H8 -> {b0,b1} */
__device__ __forceinline__ uint32_t h8_to_s16(uint32_t x) { return prmt(x, 0, 0xB391); }		/* H8 -> sign-extended s16x2 */
__device__ __forceinline__ uint32_t pack2h(int v) { uint32_t x = ((uint32_t)(uint8_t)(int8_t)v) << 8; return x | (x << 16); }
__device__ __forceinline__ int lo16(uint32_t x) { return (int)(int16_t)(x & 0xffff); }
__device__ __forceinline__ int hi16(uint32_t x) { return (int)(int16_t)(x >> 16); }
__device__ __forceinline__ uint32_t clamp8x2(uint32_t x) { return __vmaxs2(__vmins2(x, 0x007f007fu), 0xff80ff80u); }

/* warp-uniform DP context: arena pointers + band geometry */
struct DpCtx {
	const DevParams *P;
	BlkEntry *blk; uint32_t blk_cap;
	uint32_t *masks;				/* 512 u32 (2 KB) per entry, only written by traced fills: row j (128 B) holds vectors 2j and 2j+1,
									 * lane l's word = bytes {2j: cell 2l, cell 2l+1, 2j+1: cell 2l, cell 2l+1}, one flag nibble per byte */
	TailRec *tails;
	const uint32_t *lut;			/* 256-entry packed score LUT in shared memory */
	uint32_t nblk, ntail;
	int W, nl, lane, widx;			/* band width, active lanes (W/2), lane id, root template index */
	uint32_t err;
	uint64_t n_vectors;
};

/* Band registers of one warp (lane l holds cells 2l and 2l+1 as the two 16-bit halves of each register).
 * Representation (chosen so that one step needs no packed subtraction, which sm_100 has no instruction for):
 *   A = -dh, V = dv, E = de, F = df    exact H8 values (int8 in the high byte, low byte 0)
 *   every candidate of the max (score, dfh, dfv, de, df) and T = max, TE, TF carry +1 ulp in the low byte, so that
 *   ~x (= -x - 1 ulp) of an exact value added to a T-space value is exact again:  de' = TE + ~A, dv' = ~A + T,
 *   df' = TF + ~V, A' = T + ~V.  The ulp never reaches the high byte, so int8 wrap-around and signed order are unchanged.
 *   ndrop = -drop as sign-extended s16 (saturating, gaba.c:1650), delta = H8 (wrapping), md = s16.
 *   wa / wb = base codes of the two cells, pre-scaled by 4 (a << 2, b << 4) so their OR is a byte offset into the LUT. */
struct Vec {
	uint32_t A, V, E, F, delta, ndrop, md, wa, wb;
	int32_t acc; uint32_t dir;
};
/* Pipe balance.  On sm_100 the ALU pipe (LOP3 / PRMT / SHF / VIADD / VIMNMX ...) and the FMA pipe (IMAD ...) each accept one warp
 * instruction every two cycles per scheduler; the DP step is almost pure ALU work and saturates that pipe at ~0.5 IPC while the
 * FMA pipe idles (k_fill_peak: ~106 cycles per anti-diagonal for ~50 ALU instructions).  The helpers below express a few
 * operations as integer multiply-adds so that they issue on the FMA pipe; their constant operands come from kernel parameters
 * because the assembler folds a literal multiplier back into the ALU form. */
__device__ __forceinline__ uint32_t not_fma(uint32_t x, uint32_t m1)					/* ~x = x * -1 + -1 */
{
#if defined(MAB_EMU) || defined(MAB_NO_FMA_NOT)
	(void)m1; return ~x;
#else
	uint32_t d;
	asm("mad.lo.u32 %0, %1, %2, %2;" : "=r"(d) : "r"(x), "r"(m1));
	return d;
#endif
}
__device__ __forceinline__ uint32_t vneg2(uint32_t x) { return __vadd2(~x, 0x00010001u); }
/* window registers <-> the packed {cell 2l, cell 2l+1} byte pairs kept in block / tail records */
__device__ __forceinline__ uint32_t win_load(uint32_t c16) { return ((c16 & 0xffu) | ((c16 >> 8) << 16)) << 2; }
__device__ __forceinline__ uint32_t win_store(uint32_t w) { return ((w >> 2) & 0xffu) | (((w >> 18) & 0xffu) << 8); }

/* reader work (gaba_reader_work_s, gaba.c:400-423), warp-uniform */
struct FillWork {
	SecDesc sec[2];
	uint32_t rem[2], sridx[2], pridx;
	int32_t ofsd;
	int32_t wtail;
};

/* sequence fetch with the reference's code tables (gaba.c:864-879, 957-1118); `rev` replaces the mirrored pointers */
__device__ __forceinline__ uint32_t fetch_a(const SecDesc &s, uint32_t i)
{
	const uint8_t *b = (const uint8_t *)s.base;
	if(!s.rev) { return b[i]; }
	uint32_t c = b[s.len - 1 - i];
	return c < 4 ? 3 - c : 4;
}
__device__ __forceinline__ uint32_t fetch_b(const SecDesc &s, uint32_t i)
{
	const uint8_t *b = (const uint8_t *)s.base;
	uint32_t c = s.rev ? b[s.len - 1 - i] : b[i];
	if(c >= 4) { return 2; }
	return (s.rev ? 3 - c : c) << 2;
}

/* build the 256-entry LUT: index = idx(cell 2l) | idx(cell 2l+1) << 4, value = packed sb[] pair (gaba.c:1612, 3657) */
__device__ __forceinline__ void build_lut(const DevParams &P, uint32_t *lut, int tid, int nthreads)
{
	for(int i = tid; i < 256; i += nthreads) {
		uint32_t lo = ((uint32_t)(uint8_t)P.sb[i & 15] << 8) | 1u, hi = ((uint32_t)(uint8_t)P.sb[i >> 4] << 8) | 1u;	/* T-space: +1 ulp */
		lut[i] = lo | (hi << 16);
	}
}

/* ---------------------------------------------------------------- one anti-diagonal (gaba.c:1604-1699) */
/* per-lane operands of a step (everything else is a kernel-parameter constant) */
struct StepK {
	uint32_t selR, selD;		/* PRMT selectors of the two band shifts: the edge lane pulls in a zero byte instead of its neighbour */
	uint32_t insA, insB;		/* bit-select masks that drop the new base into cell 0 (lane 0, low half) / cell W-1 (last lane, high half) */
	uint32_t accw;				/* DP4A weights: lane 0 contributes +delta[0], the last lane -delta[W-1] to the direction accumulator */
	/* the packed step constants as plain 32-bit vector registers.  The assembler keeps a warp-uniform 16x2 operand as two 16-bit
	 * halves in uniform registers (LDCU.U16) and rebuilds the pair with two moves and a PRMT in front of every use: 5 PRMT + 7
	 * moves per anti-diagonal.  OR-ing in a per-lane word that is zero on every lane, but not provably so (lane (lane + 1) & 1), makes the
	 * value thread-varying in its eyes: one ordinary register, usable as it is. */
	uint32_t gfh1, gfv1, adjh1, adjv1, ofs;
	uint32_t one, two;			/* 1 and 2 as run-time values (multipliers of the FMA-pipe increments / packs) */
	const uint8_t *lut;
};
__device__ __forceinline__ StepK make_stepk(const DpCtx &c)
{
	StepK k;
	bool is0 = c.lane == 0, isL = c.lane == c.nl - 1;
	k.lut = (const uint8_t *)c.lut;
	k.selR = is0 ? 0x5444u : 0x5432u; k.selD = isL ? 0x0032u : 0x5432u;
	k.insA = is0 ? 0x0000ffffu : 0u; k.insB = isL ? 0xffff0000u : 0u;
	k.accw = is0 ? 0x00000100u : (isL ? 0xff000000u : 0u);
	const DevParams &P = *c.P;
#ifndef MAB_EMU
	/* keep the per-lane words in registers: without the barrier the compiler re-derives them from the lane id every step */
	asm volatile("" : "+r"(k.selR), "+r"(k.selD), "+r"(k.insA), "+r"(k.insB), "+r"(k.accw));
#endif
	const uint32_t lane_zero = ((uint32_t)c.lane * ((uint32_t)c.lane + 1u)) & 1u;	/* a product of consecutive numbers is even */
	k.one = 0u - P.K_M1; k.two = k.one + k.one;
	k.gfh1 = P.K_GFH1 | lane_zero; k.gfv1 = P.K_GFV1 | lane_zero; k.adjh1 = P.K_ADJH1 | lane_zero; k.adjv1 = P.K_ADJV1 | lane_zero; k.ofs = P.K_OFS | lane_zero;
#ifndef MAB_EMU
	asm volatile("" : "+r"(k.gfh1), "+r"(k.gfv1), "+r"(k.adjh1), "+r"(k.adjv1), "+r"(k.ofs));
#endif
	return k;
}
__device__ __forceinline__ int dp4a_ss(uint32_t a, uint32_t b, int c)
{
#ifdef MAB_EMU
	for(int i = 0; i < 4; i++) { c += (int)(int8_t)(a >> (8 * i)) * (int)(int8_t)(b >> (8 * i)); }
	return c;
#else
	return __dp4a((int)a, (int)b, c);
#endif
}

/* byte offset into the score LUT from the OR of the two windows: (x | x >> 12) & 0x3fc as one shift and one three-input
 * logic operation (the compiler ORs the two windows in again and masks separately: one ALU instruction more per step) */
__device__ __forceinline__ uint32_t lut_index(uint32_t x)
{
#ifdef MAB_EMU
	return (x | (x >> 12)) & 0x3fcu;
#else
	uint32_t d;
	asm("lop3.b32 %0, %1, %2, 0x3fc, 0xA8;" : "=r"(d) : "r"(x >> 12), "r"(x));		/* (a | b) & c */
	return d;
#endif
}
/* x + 1 on the FMA pipe (the step is bound by the ALU pipe; `one` is a run-time 1 the assembler cannot fold into an add) */
__device__ __forceinline__ uint32_t inc_fma(uint32_t x, uint32_t one)
{
#ifdef MAB_EMU
	(void)one; return x + 1;
#else
	uint32_t d;
	asm("mad.lo.u32 %0, %1, %2, %2;" : "=r"(d) : "r"(x), "r"(one));
	return d;
#endif
}

__device__ __forceinline__ uint32_t mad_fma(uint32_t a, uint32_t m, uint32_t b)		/* a * m + b with a run-time multiplier: stays an IMAD */
{
#ifdef MAB_EMU
	return a * m + b;
#else
	uint32_t d;
	asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(m), "r"(b));
	return d;
#endif
}

/* the two band shifts (gaba.c:1673-1699); na = new a-base code in the low half, nb = new b-base code in the HIGH half */
__device__ __forceinline__ void shift_right(const StepK &k, Vec &v, uint32_t na)		/* _fill_right: bsl dh, df; a new a-base enters at cell 0 */
{
	uint32_t ua = __shfl_up_sync(MAB_FULL, v.A, 1), uf = __shfl_up_sync(MAB_FULL, v.F, 1), uw = __shfl_up_sync(MAB_FULL, v.wa, 1);
	v.A = prmt(ua, v.A, k.selR); v.F = prmt(uf, v.F, k.selR);
	uint32_t w = __byte_perm(uw, v.wa, 0x5432);
	v.wa = (w & ~k.insA) | (na & k.insA);
}
__device__ __forceinline__ void shift_down(const StepK &k, Vec &v, uint32_t nb)			/* _fill_down: bsr dv, de; a new b-base enters at cell W-1 */
{
	uint32_t nv = __shfl_down_sync(MAB_FULL, v.V, 1), ne = __shfl_down_sync(MAB_FULL, v.E, 1), nw = __shfl_down_sync(MAB_FULL, v.wb, 1);
	v.V = prmt(v.V, nv, k.selD); v.E = prmt(v.E, ne, k.selD);
	uint32_t w = __byte_perm(v.wb, nw, 0x5432);
	v.wb = (w & ~k.insB) | (nb & k.insB);
}

/* The cell update of one anti-diagonal after the shift (gaba.c:1604-1655).  Returns (when MASKS) the traceback nibble
 * {bit0 = ~h, bit1 = ~v, bit2 = ~e, bit3 = ~f} of the two cells in the low bits of each 16-bit half. */
template <bool DOWN, bool MASKS>
__device__ __forceinline__ uint32_t vec_core(const DevParams &P, const StepK &k, Vec &v)
{
	const uint32_t EPS = 0x00010001u, ONE = 0x00010001u;
	uint32_t x = v.wa | v.wb;
	uint32_t S = *(const uint32_t *)(k.lut + lut_index(x));
	uint32_t dfh = __vadd2(v.V, k.gfh1), dfv = __vadd2(v.A, k.gfv1);
	uint32_t T = __vimax3_s16x2(S, dfh, dfv);
	T = __viaddmax_s16x2(v.E, EPS, T);
	T = __viaddmax_s16x2(v.F, EPS, T);
	uint32_t TE = __viaddmax_s16x2(v.E, k.adjh1, T), TF = __viaddmax_s16x2(v.F, k.adjv1, T);
	uint32_t bits = 0;
	if(MASKS) {
		/* T is the maximum, so T - x is 0..255 in the high byte and "min.u16 against ONE" leaves ONE set <=> not equal */
		uint32_t n_e = __viaddmin_u16x2(T, not_fma(v.E, P.K_M1), ONE), n_f = __viaddmin_u16x2(T, not_fma(v.F, P.K_M1), ONE);
		uint32_t n_fh = __vminu2(T ^ dfh, ONE), n_fv = __vminu2(T ^ dfv, ONE);
		uint32_t g_e = __vminu2(TE ^ T, ONE), g_f = __vminu2(TF ^ T, ONE);				/* set <=> te != t */
		uint32_t NH = n_fh & n_e, NV = n_fv & n_f;										/* ~h, ~v */
		uint32_t NE = (n_e | ~n_fh) & g_e, NF = (n_f | ~n_fv) & g_f;						/* ~e, ~f */
		bits = (mad_fma(NF, k.two, NE)) * 4 + (NV * 2 + NH);								/* disjoint bits: three IMADs pack the nibble */
	}
	uint32_t Pp = not_fma(v.A, P.K_M1), N = not_fma(v.V, P.K_M1);
	v.E = __vadd2(TE, Pp); v.F = __vadd2(TF, N);
	v.V = __vadd2(Pp, T); v.A = __vadd2(T, N);
	uint32_t dH = __vadd2(k.ofs, DOWN ? v.V : v.A);										/* _fill_update_delta (ofsh == ofsv) */
	v.delta = __vadd2(v.delta, dH);														/* wraps like int8 */
	v.ndrop = __vmaxs2(__viaddmin_s16x2(v.ndrop, h8_to_s16(dH), 0x00800080u), 0xff81ff81u);	/* -(drop subs delta), saturating */
	v.acc += __reduce_add_sync(MAB_FULL, dp4a_ss(dH, k.accw, 0));							/* _dir_update */
	return bits;
}
/* One unchecked step of a bulk block: direction from the accumulator (_dir_fetch), next base of that side, shift, update.
 * The direction word and the two base counters live in ordinary (per-lane, identical) registers inside the block and are
 * made warp-uniform again by one broadcast at the block end: the per-step uniform-datapath bookkeeping would cost more. */
struct BulkCnt { uint32_t dir, acnt, bcnt; };
template <bool MASKS>
__device__ __forceinline__ uint32_t bulk_step(const DevParams &P, const StepK &k, Vec &v, uint32_t an, uint32_t bn, BulkCnt &n)
{
	if(v.acc < 0) {																		/* warp-uniform */
		n.dir = n.dir * 2 + 1; shift_down(k, v, __shfl_sync(MAB_FULL, bn, n.bcnt)); n.bcnt = inc_fma(n.bcnt, k.one);
		return vec_core<true, MASKS>(P, k, v);
	}
	n.dir = n.dir * 2; shift_right(k, v, __shfl_sync(MAB_FULL, an, n.acnt)); n.acnt = inc_fma(n.acnt, k.one);
	return vec_core<false, MASKS>(P, k, v);
}
template <bool MASKS>
__device__ __forceinline__ uint32_t vec_step(const DevParams &P, const StepK &k, Vec &v, bool down, uint32_t newch)
{
	if(down) { shift_down(k, v, newch); return vec_core<true, MASKS>(P, k, v); }
	shift_right(k, v, newch); return vec_core<false, MASKS>(P, k, v);
}
template <bool MASKS>
__device__ __forceinline__ uint32_t vec_step_rt(const DevParams &P, const StepK &k, Vec &v, int down, uint32_t newch)
{
	return vec_step<MASKS>(P, k, v, down != 0, newch);
}

/* load the vector registers from the entry physically before a block (_fill_load_context, gaba.c:1527-1550) */
__device__ __forceinline__ void vec_load(const DpCtx &c, Vec &v, const BlkEntry *prev, uint32_t xd)
{
	int l = c.lane;
	v.A = vneg2(unpack8h(prev->dh[l])); v.V = unpack8h(prev->dv[l]); v.E = unpack8h(prev->de[l]); v.F = unpack8h(prev->df[l]);
	v.delta = 0; v.ndrop = vneg2(xd);
	v.acc = __reduce_add_sync(MAB_FULL, l == 0 ? (int)prev->acc : 0); v.dir = 0;		/* warp-uniform by construction */
}

/* ---------------------------------------------------------------- tails, sections */
__device__ __forceinline__ int32_t push_tail(DpCtx &c)
{
	if(c.ntail >= MAB_MAX_TAILS) { c.err |= MAB_ERR_TAIL_OVF; return (int32_t)c.ntail - 1; }
	return (int32_t)c.ntail++;
}
__device__ __forceinline__ int32_t push_blk(DpCtx &c)
{
	if(c.nblk >= c.blk_cap) { c.err |= MAB_ERR_DP_OVF; return (int32_t)c.nblk - 1; }
	return (int32_t)c.nblk++;
}

/* gaba_dp_flush (gaba.c:3969-4002): entry 0 / tail 0 are the root templates of the current band width */
__device__ inline void dp_flush(DpCtx &c, int widx)
{
	const DevParams &P = *c.P;
	const RootTpl &R = P.root[widx];
	__syncwarp();				/* nobody is still reading the previous extension's records */
	c.widx = widx; c.W = 64 >> widx; c.nl = c.W / 2;
	int l = c.lane;
	BlkEntry *b = &c.blk[0]; TailRec *t = &c.tails[0];
	if(l < c.nl) {
		b->dh[l] = (uint16_t)((uint8_t)R.dh[2 * l] | ((uint8_t)R.dh[2 * l + 1] << 8));
		b->dv[l] = (uint16_t)((uint8_t)R.dv[2 * l] | ((uint8_t)R.dv[2 * l + 1] << 8));
		b->de[l] = (uint16_t)((uint8_t)R.de[2 * l] | ((uint8_t)R.de[2 * l + 1] << 8));
		b->df[l] = (uint16_t)((uint8_t)R.df[2 * l] | ((uint8_t)R.df[2 * l + 1] << 8));
		t->cha[l] = (l == 0) ? 0x000c : 0;								/* ch.w[0] = 0x0c (gaba.c:3780) */
		t->chb[l] = (l == c.nl - 1) ? 0x0300 : 0;						/* ch.w[W-1] = 0x03 << 4 */
		t->xd[l] = 0x8080;												/* -128 */
		t->md[2 * l] = R.md[2 * l]; t->md[2 * l + 1] = R.md[2 * l + 1];
	}
	if(l == 0) {
		b->acc = 0; b->xstat = MAB_X_ROOT; b->acnt = 0; b->bcnt = 0; b->link = -1; b->dir_mask = 0; b->mm_lo = b->mm_hi = 0;
		t->mdrop = R.mdrop; t->istat = 0; t->pridx = 0;
		t->ridx[0] = t->ridx[1] = 0; t->adv[0] = t->adv[1] = 0; t->tail = -1; t->last_blk = 0;
		t->aid = t->bid = 0; t->ascnt = t->bscnt = 0;
		t->apos = -(c.W / 2); t->bpos = -(c.W / 2); t->max = R.init_max; t->status = MAB_UPDATE_A | MAB_UPDATE_B;
	}
	c.nblk = 1; c.ntail = 1;
	__syncwarp();
}

/* fill_load_section (gaba.c:1269-1308); breakpoint masks are always zero on minialign's path */
__device__ __forceinline__ void load_section(DpCtx &c, FillWork &w, int32_t tail, const SecDesc &a, const SecDesc &b, uint32_t pridx)
{
	const TailRec *t = &c.tails[tail];
	w.sec[0] = a; w.sec[1] = b;
	uint32_t ra = t->ridx[0], rb = t->ridx[1];
	w.rem[0] = w.sridx[0] = ra == 0 ? a.len : ra;
	w.rem[1] = w.sridx[1] = rb == 0 ? b.len : rb;
	w.pridx = pridx; w.ofsd = 0; w.wtail = tail;
}

/* fill_load_vectors + fill_create_phantom (gaba.c:1315-1332, 1376-1399); returns the head entry index.
 * The char windows, xd and md of the tail are returned in registers. */
__device__ __forceinline__ int32_t load_vectors(DpCtx &c, int32_t tail, Vec &v, uint32_t &xd)
{
	const TailRec *t = &c.tails[tail];
	int l = c.lane, ls = l < c.nl ? l : 0;
	v.wa = win_load(t->cha[ls]); v.wb = win_load(t->chb[ls]);
	xd = unpack8(t->xd[ls]);
	v.md = (uint32_t)(uint16_t)t->md[2 * ls] | ((uint32_t)(uint16_t)t->md[2 * ls + 1] << 16);
	int32_t prev = t->last_blk;
	int32_t ph = push_blk(c);
	BlkEntry *p = &c.blk[ph]; const BlkEntry *pb = &c.blk[prev];
	if(l < c.nl) { p->dh[l] = pb->dh[l]; p->dv[l] = pb->dv[l]; p->de[l] = pb->de[l]; p->df[l] = pb->df[l]; }
	if(l == 0) {
		p->acc = pb->acc; p->xstat = (int8_t)((pb->xstat & MAB_X_ROOT) | MAB_X_HEAD);
		p->acnt = 0; p->bcnt = 0; p->link = prev; p->dir_mask = 0; p->mm_lo = p->mm_hi = 0; p->tail = (uint32_t)tail;
	}
	__syncwarp();
	return ph;
}

/* shift `n` freshly fetched bases of one side into the window registers without a DP step (init fetch) */
__device__ __forceinline__ void window_consume(const DpCtx &c, Vec &v, const FillWork &w, int side, uint32_t n)
{
	const SecDesc &s = w.sec[side];
	for(uint32_t k = 0; k < n; k++) {
		uint32_t pos = s.len - w.rem[side] + k;
		if(side == 0) {
			uint32_t ch = fetch_a(s, pos) << 2;
			uint32_t ua = __shfl_up_sync(MAB_FULL, v.wa, 1);
			if(c.lane == 0) { ua = ch << 16; }
			v.wa = __byte_perm(ua, v.wa, 0x5432);
		} else {
			uint32_t ch = fetch_b(s, pos) << 2;
			uint32_t nb = __shfl_down_sync(MAB_FULL, v.wb, 1);
			if(c.lane == c.nl - 1) { nb = ch; }
			v.wb = __byte_perm(v.wb, nb, 0x5432);
		}
	}
}

/* fill_init_fetch (gaba.c:1168-1210) */
__device__ __forceinline__ int64_t init_fetch(DpCtx &c, FillWork &w, Vec &v, int32_t ph, int64_t apos, int64_t bpos)
{
	int32_t irem[2] = { (int32_t)(-1 - (int32_t)apos), (int32_t)(-1 - (int32_t)bpos) };
	int32_t srem[2] = { (int32_t)w.rem[0], (int32_t)w.rem[1] };
	int32_t len[2];
	for(int i = 0; i < 2; i++) {
		int32_t x = irem[i] < srem[i] ? irem[i] : srem[i];
		int32_t y = (srem[1 - i] - irem[1 - i]) + ((i == 0 ? 1 : 0) + irem[i]);
		len[i] = x < y ? x : y;
	}
	MAB_DBG("init_fetch apos %lld bpos %lld irem %d %d srem %d %d len %d %d\n", (long long)apos, (long long)bpos, irem[0], irem[1], srem[0], srem[1], len[0], len[1]);
	if(len[0] < 0 || len[1] < 0 || len[0] > MAB_WMAX || len[1] > MAB_WMAX) { c.err |= MAB_ERR_DP_OVF; len[0] = len[1] = 0; }	/* cannot happen for non-empty sections */
	window_consume(c, v, w, 0, (uint32_t)len[0]); window_consume(c, v, w, 1, (uint32_t)len[1]);
	if(c.lane == 0) { c.blk[ph].acnt = (int8_t)len[0]; c.blk[ph].bcnt = (int8_t)len[1]; }
	w.rem[0] = (uint32_t)(srem[0] - len[0]); w.rem[1] = (uint32_t)(srem[1] - len[1]);
	__syncwarp();
	return bpos + len[1];
}

/* fill_create_tail (gaba.c:1405-1499); `last` = last processed entry */
__device__ __forceinline__ int32_t create_tail(DpCtx &c, FillWork &w, const Vec &v, uint32_t xd, int32_t last)
{
	const BlkEntry *b = &c.blk[last];
	int xstat = b->xstat;
	int cnt = ((uint8_t)b->acnt) | (((uint8_t)b->bcnt) << 8);
	if(cnt == 0 && !(xstat & MAB_X_HEAD)) { c.nblk = (uint32_t)last; last--; }		/* squash the empty block */
	else if(cnt == 0) { last--; }
	int32_t ti = push_tail(c);
	TailRec *t = &c.tails[ti];
	const TailRec *prev = &c.tails[w.wtail];
	int l = c.lane;
	/* fill_save_vectors: mdrop = hmax(md + xd) */
	int m0 = (int)(int16_t)(lo16(v.md) + lo16(xd)), m1 = (int)(int16_t)(hi16(v.md) + hi16(xd));
	int mx = l < c.nl ? (m0 > m1 ? m0 : m1) : -32768;
	int mdrop = __reduce_max_sync(MAB_FULL, mx);
	if(l < c.nl) {
		t->cha[l] = (uint16_t)win_store(v.wa); t->chb[l] = (uint16_t)win_store(v.wb);
		t->xd[l] = (uint16_t)pack8(xd);
		t->md[2 * l] = (int16_t)lo16(v.md); t->md[2 * l + 1] = (int16_t)hi16(v.md);
	}
	if(l == 0) {
		t->mdrop = mdrop; t->istat = 0; t->pridx = w.pridx;
		uint32_t upd = 0;
		for(int i = 0; i < 2; i++) {
			t->ridx[i] = w.rem[i]; t->adv[i] = w.sridx[i] - w.rem[i];
			if(w.rem[i] == 0) { upd |= i == 0 ? MAB_UPDATE_A : MAB_UPDATE_B; }
		}
		t->tail = w.wtail; t->last_blk = last;
		t->aid = w.sec[0].id; t->bid = w.sec[1].id;
		t->ascnt = prev->ascnt + (w.rem[0] == 0); t->bscnt = prev->bscnt + (w.rem[1] == 0);
		t->apos = prev->apos + (w.sridx[0] - w.rem[0]); t->bpos = prev->bpos + (w.sridx[1] - w.rem[1]);
		t->max = (prev->max - prev->mdrop) + w.ofsd + mdrop;
		t->status = ((uint32_t)(xstat & MAB_X_TERM) << 8) | upd;
		t->sec[0] = w.sec[0]; t->sec[1] = w.sec[1];
	}
	__syncwarp();
	MAB_DBG("create_tail ti %d last %d xstat %d cnt %d rem %u %u sridx %u %u mdrop %d ofsd %d max %lld status %x nblk %u\n", ti, last, xstat, cnt, w.rem[0], w.rem[1], w.sridx[0], w.sridx[1], mdrop, w.ofsd, (long long)c.tails[ti].max, c.tails[ti].status, c.nblk);
	return ti;
}

/* ---------------------------------------------------------------- block loop (gaba.c:1821-2103) */
/* Bulk blocks (>= 32 bases left on both sides) run 32 steps without bound checks, unrolled by four so the traceback
 * nibbles of four anti-diagonals land in one 32-bit word per lane (one coalesced 128 B row per four vectors); cap blocks
 * test the section bounds before every step.  The diff vectors stay in registers from block to block; only the direction
 * accumulator is re-truncated to int8 at each block boundary like the reference's `blk->acc` store. */
template <bool MASKS>
__device__ __forceinline__ int32_t fill_blocks(DpCtx &c, FillWork &w, Vec &v, uint32_t &xd, int32_t cur)
{
	const DevParams &P = *c.P;
	const StepK k = make_stepk(c);
	int cap = 0, first = 1;
	const int l = c.lane;
	int xstat_prev = c.blk[cur].xstat;
	while(1) {
		if(xstat_prev < 0) { break; }												/* TERM */
		if(!cap && (w.rem[0] < MAB_BLK || w.rem[1] < MAB_BLK || w.pridx < MAB_BLK)) { cap = 1; }
		int32_t bi = push_blk(c);
		if(c.err) { break; }
		BlkEntry *b = &c.blk[bi];
		/* prefetch up to 32 bases of each side, one per lane (fill_fetch_core, gaba.c:1125-1144), pre-scaled like the windows */
		uint32_t an = 0, bn = 0;
		if((uint32_t)l < w.rem[0]) { an = fetch_a(w.sec[0], w.sec[0].len - w.rem[0] + l) << 2; }
		if((uint32_t)l < w.rem[1]) { bn = fetch_b(w.sec[1], w.sec[1].len - w.rem[1] + l) << 18; }		/* enters a high half */
		if(l < c.nl) { b->cha[l] = (uint16_t)win_store(v.wa); b->chb[l] = (uint16_t)win_store(v.wb); }
		if(first) { vec_load(c, v, &c.blk[bi - 1], xd); first = 0; }
		else { v.delta = 0; v.acc = (int32_t)(int8_t)v.acc; v.dir = 0; }				/* ndrop carries over: xd == drop of the previous block */
		uint32_t *mrow = c.masks + 512ull * bi + l;
		int acnt = 0, bcnt = 0, i = 0;
		if(!cap) {
			BulkCnt n; n.dir = 0; n.acnt = 0; n.bcnt = 0;
#ifndef MAB_EMU
			asm volatile("" : "+r"(n.dir), "+r"(n.acnt), "+r"(n.bcnt));					/* plain registers, see bulk_step */
#endif
			#pragma unroll 1
			for(int g = 0; g < MAB_BLK / 4; g++) {
				uint32_t b0 = bulk_step<MASKS>(P, k, v, an, bn, n), b1 = bulk_step<MASKS>(P, k, v, an, bn, n);
				if(MASKS) { mrow[64 * g] = __byte_perm(b0, b1, 0x6420); }				/* one coalesced 128 B row per two vectors */
				uint32_t b2 = bulk_step<MASKS>(P, k, v, an, bn, n), b3 = bulk_step<MASKS>(P, k, v, an, bn, n);
				if(MASKS) { mrow[64 * g + 32] = __byte_perm(b2, b3, 0x6420); }
			}
			v.dir = __shfl_sync(MAB_FULL, n.dir, 0);
			bcnt = __popc(v.dir); acnt = MAB_BLK - bcnt;
			i = MAB_BLK;
		} else {
			for(; i < MAB_BLK; i++) {
				int down = v.acc < 0;
				{																/* _fill_cap_test_idx */
					int64_t ar = (int64_t)w.rem[0] - (acnt + !down), br = (int64_t)w.rem[1] - (bcnt + down);
					int64_t pr = ar + br + (int64_t)w.pridx;
					if((ar | br | pr) < 0) { break; }							/* direction not consumed (the reference winds it back) */
				}
				v.dir = (v.dir << 1) | (uint32_t)down;
				uint32_t newch = down ? __shfl_sync(MAB_FULL, bn, bcnt) : __shfl_sync(MAB_FULL, an, acnt);
				bcnt += down; acnt += 1 - down;
				uint32_t bits = vec_step_rt<MASKS>(P, k, v, down, newch);
				if(MASKS) { ((uint16_t *)(mrow + 32 * (i >> 1)))[i & 1] = (uint16_t)__byte_perm(bits, 0, 0x4420); }
			}
		}
		c.n_vectors += (uint64_t)i;
		w.pridx -= (uint32_t)i;
		if(i < MAB_BLK && i != 0) { v.dir <<= (MAB_BLK - i); }						/* _dir_adjust_remainder */
		/* _fill_store_context (gaba.c:1734-1778) */
		int ctr = c.W / 4;															/* lane holding cell W/2 (low half) */
		uint32_t dl = h8_to_s16(v.delta);										/* delta as sign-extended int8 */
		uint32_t drop = vneg2(v.ndrop);
		int dropc = lo16(__shfl_sync(MAB_FULL, drop, ctr)), cofs = lo16(__shfl_sync(MAB_FULL, dl, ctr));
		int xstat = (P.tx - dropc) & MAB_X_TERM;
		uint32_t sum = sext8x2(__vadd2(drop, dl));
		uint32_t mlo = __ballot_sync(MAB_FULL, l < c.nl && lo16(sum) > lo16(xd));
		uint32_t mhi = __ballot_sync(MAB_FULL, l < c.nl && hi16(sum) > hi16(xd));
		/* middle delta with the overflow / underflow rescue terms */
		uint32_t md = __vadd2(v.md, dl);
		uint32_t ov = sext8x2(~sum & (drop & dl));
		md = __vadd2(md, ov & 0x01000100u);
		uint32_t uv = sext8x2(clamp8x2(__vsub2(dl, 0x00400040u)) | drop);
		md = __vadd2(md, uv & 0x01000100u);
		md = __vsub2(md, pack2(cofs + 0x0100));
		v.md = md; xd = drop;
		w.ofsd += cofs; w.rem[0] -= (uint32_t)acnt; w.rem[1] -= (uint32_t)bcnt;
		if(l < c.nl) {
			b->dh[l] = (uint16_t)pack8h(vneg2(v.A)); b->dv[l] = (uint16_t)pack8h(v.V); b->de[l] = (uint16_t)pack8h(v.E); b->df[l] = (uint16_t)pack8h(v.F);
		}
		if(l == 0) {
			b->acc = (int8_t)v.acc; b->xstat = (int8_t)xstat; b->acnt = (int8_t)acnt; b->bcnt = (int8_t)bcnt;
			b->dir_mask = v.dir; b->mm_lo = mlo; b->mm_hi = mhi; b->link = -1;
			b->arem = w.rem[0] + (uint32_t)acnt; b->brem = w.rem[1] + (uint32_t)bcnt; b->tail = (uint32_t)c.ntail;	/* the tail created next */
		}
		xstat_prev = (int)(int8_t)xstat;
		cur = bi;
		if(i != MAB_BLK) { break; }
	}
	__syncwarp();
	return cur;
}

/* gaba_dp_fill_root (gaba.c:2110-2154) */
template <bool MASKS>
__device__ inline int32_t dp_fill_root(DpCtx &c, const SecDesc &a, uint32_t apos, const SecDesc &b, uint32_t bpos)
{
	/* fill_create_bridge (gaba.c:1339-1369) */
	int32_t bi = push_tail(c);
	TailRec *brg = &c.tails[bi]; const TailRec *rt = &c.tails[0];
	int l = c.lane;
	if(l < c.nl) { brg->cha[l] = rt->cha[l]; brg->chb[l] = rt->chb[l]; brg->xd[l] = rt->xd[l]; brg->md[2 * l] = rt->md[2 * l]; brg->md[2 * l + 1] = rt->md[2 * l + 1]; }
	if(l == 0) {
		brg->mdrop = rt->mdrop; brg->istat = 1; brg->pridx = rt->pridx;
		brg->ridx[0] = a.len - apos; brg->ridx[1] = b.len - bpos; brg->adv[0] = apos; brg->adv[1] = bpos;
		brg->tail = 0; brg->last_blk = 0;
		brg->aid = a.id; brg->bid = b.id; brg->ascnt = rt->ascnt; brg->bscnt = rt->bscnt;
		brg->apos = rt->apos; brg->bpos = rt->bpos; brg->max = rt->max; brg->status = rt->status;
		brg->sec[0] = a; brg->sec[1] = b;
	}
	__syncwarp();
	FillWork w; Vec v; uint32_t xd;
	load_section(c, w, bi, a, b, 0xffffffffu);
	int32_t ph = load_vectors(c, 0, v, xd);
	int64_t rapos = c.tails[0].apos, rbpos = c.tails[0].bpos;
	if(init_fetch(c, w, v, ph, rapos, rbpos) < -1) { return create_tail(c, w, v, xd, ph); }
	{ int32_t last = fill_blocks<MASKS>(c, w, v, xd, ph); return create_tail(c, w, v, xd, last); }	/* sequenced: xd is updated by fill_blocks */
}

/* gaba_dp_fill (gaba.c:2161-2203) */
template <bool MASKS>
__device__ inline int32_t dp_fill(DpCtx &c, int32_t prev, const SecDesc &a, const SecDesc &b)
{
	FillWork w; Vec v; uint32_t xd;
	load_section(c, w, prev, a, b, c.tails[prev].pridx);
	int32_t ph = load_vectors(c, prev, v, xd);
	int64_t papos = c.tails[prev].apos, pbpos = c.tails[prev].bpos;
	if(pbpos < -1) {
		if(init_fetch(c, w, v, ph, papos, pbpos) < -1) { return create_tail(c, w, v, xd, ph); }
	}
	{ int32_t last = fill_blocks<MASKS>(c, w, v, xd, ph); return create_tail(c, w, v, xd, last); }	/* sequenced: xd is updated by fill_blocks */
}

/* mm_extend_core (minialign.c:4075-4112): fill_root, then keep filling over the N-tail sections until X-drop or a second
 * section end; returns the tail with the largest max */
template <bool MASKS>
__device__ inline int32_t extend_core(DpCtx &c, SecDesc a, const SecDesc &at, SecDesc b, const SecDesc &bt, uint32_t apos, uint32_t bpos)
{
	int32_t f = dp_fill_root<MASKS>(c, a, apos, b, bpos);
	int32_t m = f;
	uint32_t flag = MAB_TERM;
	while(c.err == 0) {
		uint32_t st = c.tails[f].status;
		if((flag & st) != 0) { break; }
		if(st & MAB_UPDATE_A) { a = at; }
		if(st & MAB_UPDATE_B) { b = bt; }
		flag |= st & (MAB_UPDATE_A | MAB_UPDATE_B);
		f = dp_fill<MASKS>(c, f, a, b);
		m = c.tails[f].max > c.tails[m].max ? f : m;
	}
	return m;
}

/* ---------------------------------------------------------------- max search (gaba.c:2604-2817) */
struct Leaf {
	int32_t blk; uint32_t p, q;
	int32_t gidx[2], sgidx[2];
	uint64_t plen;
};

/* lowest set cell index of a two-plane mask (cell 2l = plane lo bit l, cell 2l+1 = plane hi bit l); 64 when empty */
__device__ __forceinline__ uint32_t mask_tz(uint32_t lo, uint32_t hi)
{
	uint32_t a = lo ? 2u * (uint32_t)(__ffs((int)lo) - 1) : 64u, b = hi ? 2u * (uint32_t)(__ffs((int)hi) - 1) + 1u : 64u;
	return a < b ? a : b;
}

__device__ inline void leaf_search(DpCtx &c, int32_t ti, Leaf &lf)
{
	const TailRec *t = &c.tails[ti];
	int l = c.lane, ls = l < c.nl ? l : 0;
	/* leaf_load_max_mask */
	uint32_t xd = unpack8(t->xd[ls]);
	int m0 = (int)(int16_t)(t->md[2 * ls] + lo16(xd)), m1 = (int)(int16_t)(t->md[2 * ls + 1] + hi16(xd));
	int mdrop = (int)(int16_t)t->mdrop;
	uint32_t mlo = __ballot_sync(MAB_FULL, l < c.nl && m0 == mdrop), mhi = __ballot_sync(MAB_FULL, l < c.nl && m1 == mdrop);
	int32_t b = t->last_blk + 1;
	int32_t ridx[2] = { (int32_t)t->ridx[0], (int32_t)t->ridx[1] };
	lf.plen = 0; lf.blk = 0; lf.p = 0; lf.q = 0; lf.gidx[0] = lf.gidx[1] = lf.sgidx[0] = lf.sgidx[1] = 0;
	while(1) {
		--b;
		if((c.blk[b].xstat & MAB_X_ROOT) == MAB_X_ROOT) { return; }
		while(c.blk[b].xstat & MAB_X_HEAD) { b = c.blk[b].link; }
		ridx[0] += c.blk[b].acnt; ridx[1] += c.blk[b].bcnt;
		uint32_t blo = c.blk[b].mm_lo, bhi = c.blk[b].mm_hi;
		if(((mlo & ~blo) | (mhi & ~bhi)) == 0) { break; }
		mlo &= ~blo; mhi &= ~bhi;
	}
	/* leaf_detect_pos: replay the block, lane i keeps the update mask of vector i */
	const BlkEntry *blk = &c.blk[b];
	int n = blk->acnt + blk->bcnt;
	uint32_t my_lo = 0, my_hi = 0;
	{
		const TailRec *bt = &c.tails[blk->tail];
		SecDesc sa = bt->sec[0], sb = bt->sec[1];
		uint32_t arem = blk->arem, brem = blk->brem;
		Vec v;
		v.wa = win_load(blk->cha[ls]); v.wb = win_load(blk->chb[ls]);
		v.md = 0;
		vec_load(c, v, &c.blk[b - 1], 0);
		uint32_t an = 0, bn = 0;
		if((uint32_t)l < arem) { an = fetch_a(sa, sa.len - arem + l) << 2; }
		if((uint32_t)l < brem) { bn = fetch_b(sb, sb.len - brem + l) << 18; }
		uint32_t mx = 0;
		const StepK kk = make_stepk(c);
		int acnt = 0, bcnt = 0;
		for(int i = 0; i < n; i++) {
			v.dir = (v.dir << 1) | (uint32_t)(v.acc < 0);
			int down = (int)(v.dir & 1);
			uint32_t newch = down ? __shfl_sync(MAB_FULL, bn, bcnt) : __shfl_sync(MAB_FULL, an, acnt);
			if(down) { bcnt++; } else { acnt++; }
			vec_step_rt<false>(*c.P, kk, v, down, newch);
			uint32_t ulo = __ballot_sync(MAB_FULL, l < c.nl && lo16(v.delta) > lo16(mx));
			uint32_t uhi = __ballot_sync(MAB_FULL, l < c.nl && hi16(v.delta) > hi16(mx));
			mx = __vmaxs2(mx, v.delta);
			if(l == i) { my_lo = ulo; my_hi = uhi; }
		}
		c.n_vectors += (uint64_t)n;
	}
	/* leaf_search_pos: while(m > mask_arr && (max_mask & ~(--m)->all) != 0) { max_mask &= ~m->all; } */
	int m = n;
	uint32_t clo = 0, chi = 0;
	for(;;) {
		if(!(m > 0)) { break; }
		m--;
		clo = __shfl_sync(MAB_FULL, my_lo, m); chi = __shfl_sync(MAB_FULL, my_hi, m);
		if(((mlo & ~clo) | (mhi & ~chi)) == 0) { break; }
		mlo &= ~clo; mhi &= ~chi;
	}
	if(n == 0) { clo = chi = 0; }
	uint32_t p = (uint32_t)m, q = mask_tz(clo & mlo, chi & mhi);
	lf.blk = b; lf.p = p & 0xff; lf.q = q & 0xff;
	int32_t fcnt = (int32_t)p + 1;
	uint32_t dir_mask = blk->dir_mask >> (MAB_BLK - fcnt);
	ridx[0] -= (int32_t)((uint32_t)(fcnt - __popc(dir_mask)) - (1 + q));
	ridx[1] -= (int32_t)((uint32_t)(0 + __popc(dir_mask)) - ((uint32_t)c.W - q));
	for(int i = 0; i < 2; i++) { lf.gidx[i] = lf.sgidx[i] = 1 - ridx[i] + (int32_t)t->ridx[i]; }
	int32_t rem0 = ridx[0] - (int32_t)t->ridx[0], rem1 = ridx[1] - (int32_t)t->ridx[1];
	lf.plen = (uint64_t)t->apos + (uint64_t)t->bpos + 2 + (uint64_t)c.W - (uint64_t)(int64_t)rem1 - (uint64_t)(int64_t)rem0;
}

struct PosPair { uint32_t aid, bid, apos, bpos; uint64_t plen; };

/* gaba_dp_search_max (gaba.c:2776-2817) */
__device__ inline PosPair dp_search_max(DpCtx &c, int32_t ti)
{
	Leaf lf; leaf_search(c, ti, lf);
	PosPair pos; pos.plen = lf.plen;
	int32_t gidx[2] = { lf.gidx[0], lf.gidx[1] }, acc[2] = { 0, 0 };
	const TailRec *t = &c.tails[ti];
	uint32_t id[2] = { t->aid, t->bid };
	while(t->tail >= 0) {
		int upd0 = 1 > gidx[0], upd1 = 1 > gidx[1];
		if(!upd0 && !upd1) { break; }
		uint32_t nid0 = t->aid, nid1 = t->bid;
		acc[0] += (int32_t)t->adv[0]; acc[1] += (int32_t)t->adv[1];
		t = &c.tails[t->tail];
		if(upd0 && t->ridx[0] == 0) { gidx[0] += acc[0]; id[0] = nid0; acc[0] = 0; }
		if(upd1 && t->ridx[1] == 0) { gidx[1] += acc[1]; id[1] = nid1; acc[1] = 0; }
	}
	pos.aid = id[0]; pos.bid = id[1]; pos.apos = (uint32_t)gidx[0]; pos.bpos = (uint32_t)gidx[1];
	__syncwarp();
	return pos;
}

/* ---------------------------------------------------------------- traceback (gaba.c:2820-3393) */
#define MAB_TS_H 1
#define MAB_TS_V 2
#define MAB_TS_S 4
enum { mab_ts_d = 3, mab_ts_v0 = 2, mab_ts_v1 = 6, mab_ts_h0 = 1, mab_ts_h1 = 5 };
enum { MP_D_HEAD, MP_D_MID, MP_D_TAIL, MP_H_HEAD, MP_H_BODY, MP_H_TAIL, MP_V_HEAD, MP_V_BODY, MP_V_TAIL };

struct Trace {
	int32_t blk, mi; uint32_t q, state;
	int32_t gidx[2], sgidx[2]; uint32_t ofs[2], id[2];
	int32_t tail[2];
	uint32_t gi[2], ge[2], gf[2];
	/* Path bits are produced from the end of the alignment towards its start.  pidx = bits still to be produced; the 64-bit
	 * accumulator holds the nacc produced-but-unstored bits [pidx, 32 (widx + 1)), newest at bit 0; every path word above widx
	 * is already in memory.  Whole words are stored at block boundaries only (<= 32 pops apart). */
	uint32_t pidx, nacc; int32_t widx;
	uint64_t pacc;
	uint32_t *path;				/* path words in the result pool */
	int32_t sblk, pblk; uint32_t cur;	/* block held by tile buffer cur, block being prefetched into the other buffer (-1: none) */
};

/* Start the copy of the 2 KB mask block of entry b into a tile buffer: 16 asynchronous 4-byte copies per lane (coalesced
 * 128 B rows in global memory, pitch MAB_TPITCH in shared memory), no registers and no waiting: the block the walk will need
 * next is fetched while the current one is walked. */
__device__ __forceinline__ void stage_issue(const DpCtx &c, int32_t b, uint8_t *buf)
{
	const uint32_t *src = c.masks + 512ull * b + c.lane;
#ifdef MAB_EMU
	for(int j = 0; j < 16; j++) { ((uint32_t *)buf)[33 * j + c.lane] = src[32 * j]; }
#else
	const uint32_t dst = (uint32_t)__cvta_generic_to_shared(buf) + 4u * c.lane;
	const unsigned long long g = (unsigned long long)__cvta_generic_to_global(src);
	#pragma unroll
	for(int j = 0; j < 16; j++) {
		asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst + MAB_TPITCH * j), "l"(g + 128ull * j) : "memory");
	}
#endif
}

/* Wait for this lane's copies and publish them to the warp.  For W = 16 the cells 16..31 read as "all flags clear" like the
 * zero-extended 16-bit mask words of the reference. */
__device__ __forceinline__ void stage_wait(const DpCtx &c, uint8_t *buf)
{
#ifndef MAB_EMU
	asm volatile("cp.async.wait_all;" ::: "memory");
#endif
	if(c.W == 16 && c.lane >= 8) {
		for(int j = 0; j < 16; j++) { ((uint32_t *)buf)[33 * j + c.lane] = 0x0f0f0f0fu; }
	}
	__syncwarp();
}

/* no copy may still be in flight when the tile is reused as sort scratch */
__device__ __forceinline__ void stage_drain()
{
#ifndef MAB_EMU
	asm volatile("cp.async.wait_all;" ::: "memory");
#endif
	__syncwarp();
}

/* Make block b the one held by the current tile buffer and start fetching b - 1, the usual successor, into the other. */
__device__ __forceinline__ void stage_block(const DpCtx &c, Trace &w, int32_t b, uint8_t *tile8)
{
	if(w.sblk != b) {
		if(w.pblk == b) { w.cur ^= 1; w.pblk = -1; }
		else { __syncwarp(); stage_issue(c, b, tile8 + MAB_TBUF * w.cur); }		/* every lane is done reading the old contents */
		w.sblk = b;
		stage_wait(c, tile8 + MAB_TBUF * w.cur);
	}
	if(b > 0 && w.pblk != b - 1) {
		__syncwarp();
		stage_issue(c, b - 1, tile8 + MAB_TBUF * (w.cur ^ 1)); w.pblk = b - 1;
	}
}

/* trace_reload_section (gaba.c:2826-2859) */
__device__ __forceinline__ void trace_reload_section(const DpCtx &c, Trace &w, int i)
{
	int32_t tail = w.tail[i], prev = tail;
	int32_t gidx = w.gidx[i];
	while(gidx <= 0 && tail > 0) {
		do {
			gidx += c.tails[tail].istat ? 0 : (int32_t)c.tails[tail].adv[i];
			prev = tail; tail = c.tails[tail].tail;
		} while(tail > 0 && c.tails[tail].ridx[i] != 0);
	}
	w.tail[i] = tail;
	w.id[i] = i == 0 ? c.tails[prev].aid : c.tails[prev].bid;
	w.ofs[i] = c.tails[prev].istat ? c.tails[prev].adv[i] : 0;
	w.gidx[i] = gidx; w.sgidx[i] = gidx;
}

/* one byte of this warp's tile, addressed in the shared window (a generic-pointer load makes the compiler rebuild the
 * shared base from SR_CgaCtaId on every pop) */
__device__ __forceinline__ uint32_t tile_ld8(const uint8_t *tile8, uint32_t tile_s, int32_t idx)
{
#ifdef MAB_EMU
	(void)tile_s; return (uint32_t)tile8[idx];
#else
	(void)tile8; uint32_t v;
	asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(tile_s + (uint32_t)idx) : "memory");
	return v;
#endif
}

/* trace_core (gaba.c:3111-3232): the reference's diag / h-gap / v-gap loops with its bulk and tail modes, as straight
 * gotos.  Per popped vector: one byte load gives the cell's flag nibble (bit0 = ~h, bit1 = ~v, bit2 = ~e, bit3 = ~f; q is
 * taken modulo the mask word width like the reference's `mask->x.all >> q` on x86, gaba.c:2952-2973) and the path bit is
 * shifted into a 64-bit accumulator (the reference's `path_array << 1 | isV`, gaba.c:2978-2991); whole path words are stored
 * at block boundaries.  The rare block-boundary work sits in one out-of-line `reload` section that returns to the pop site
 * through `ret`.  In bulk mode the section counters were advanced by whole blocks beforehand (dec = 0) and stay >= W, so the
 * "counter exhausted" tests need no mode check. */
__device__ __forceinline__ void trace_core(DpCtx &c, Trace &w, uint8_t *tile8)
{
	const int W = c.W;
	const uint32_t HEAD_CNT = (uint32_t)(W / MAB_BLK + (W == 16));
	const uint32_t qmask = W == 64 ? 63u : 31u;
	int32_t b = w.blk, mi = w.mi; uint32_t q = w.q, save = HEAD_CNT;
	uint32_t dir = c.blk[b].dir_mask >> (MAB_BLK - (mi + 1));
	int bulk = 0, ret = 0;
	int32_t g0 = w.gidx[0], g1 = w.gidx[1], dec = 1;
	uint64_t pacc = w.pacc; uint32_t nacc = w.nacc; int32_t widx = w.widx;
	uint32_t nb;
	/* at most two whole words are pending: nacc < 32 after a flush, <= 32 pops per block */
	#define FLUSH_PATH() { \
		if(nacc >= 32) { nacc -= 32; if(c.lane == 0) { w.path[widx] = (uint32_t)(pacc >> nacc); } widx--; } \
		if(nacc >= 32) { nacc -= 32; if(c.lane == 0) { w.path[widx] = (uint32_t)(pacc >> nacc); } widx--; } \
	}
#ifdef MAB_EMU
	const uint32_t tile_s = 0;
#else
	const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile8);
#endif
	/* acc, xstat, acnt, bcnt of a block entry as one word */
	#define BLK_META(_b) (*(const uint32_t *)&c.blk[_b].acc)
	#define META_XSTAT(_m) ((_m) >> 8)
	#define META_ACNT(_m) ((int32_t)(int8_t)((_m) >> 16))
	#define META_BCNT(_m) ((int32_t)(_m) >> 24)
	#define PREFETCH_META() { if(b > 0) { pf_meta = BLK_META(b - 1); pf_dir = c.blk[b - 1].dir_mask; } }
	uint32_t pf_meta = 0, pf_dir = 0;
	stage_block(c, w, b, tile8);
	PREFETCH_META();
	uint32_t cofs = MAB_TBUF * w.cur;
	/* byte of (vector m, cell x) in the raw block layout */
	#define TIDX(_m, _x) (int32_t)(cofs + (uint32_t)((_m) >> 1) * MAB_TPITCH + (((uint32_t)(_m) & 1u) << 1) + (((_x) >> 1) << 2) + ((_x) & 1u))
	#define NIB() tile_ld8(tile8, tile_s, TIDX(mi, q & qmask))
	nb = NIB();
	#define POP(_v, _id) { \
		if(_v) { g1 -= dec; } else { g0 -= dec; } \
		pacc = (pacc << 1) | (uint64_t)(_v); nacc++; \
		q += (dir & 1) - (uint32_t)(_v); dir >>= 1; mi--; \
		if(mi < 0) { ret = _id; goto reload; } \
		R##_id: nb = NIB(); \
	}
	switch(w.state) {
		case mab_ts_d:  goto L_D_HEAD;
		case mab_ts_v0: goto L_V_HEAD;
		case mab_ts_v1: goto L_V_TAIL;
		case mab_ts_h0: goto L_H_HEAD;
		case mab_ts_h1: goto L_H_TAIL;
		default: return;
	}
L_D_HEAD:
	if(!(nb & 1)) { goto L_H_HEAD; }											/* h bit set */
	{
		/* Run of diagonal steps inside this block, looked up by all lanes at once: along a pure diagonal run the column after
		 * s steps is known from the direction bits alone (q + popc(low 2s bits) - s), so lane s reads the flags the walk would
		 * see after s steps, a ballot finds the first step that leaves the diagonal (v flag at D_TAIL, h flag at D_HEAD) and
		 * the state jumps there: two path bits "01" per step, both section counters down by one in tail mode.  The run stops
		 * short of the block end and of the section ends; those steps take the scalar path below. */
		int32_t lim = mi >> 1;
		if(dec) { const int32_t gl = g0 < g1 ? g0 : g1; lim = gl < lim ? gl : lim; }
		if(lim >= 1) {
			const uint32_t run_mask = (1u << (2 * (c.lane & 15))) - 1u, run_need = c.lane == 0 ? 1u : 3u;
			const int32_t ms = mi - 2 * c.lane;
			const uint32_t qs = q + (uint32_t)__popc(dir & run_mask) - (uint32_t)c.lane;
			uint32_t ns = 0;
			if(ms >= 0) { ns = tile_ld8(tile8, tile_s, TIDX(ms, qs & qmask)); }
			const uint32_t bal = __ballot_sync(MAB_FULL, (ns & run_need) == run_need);
			const int32_t nf = __ffs((int)~bal) - 1, n = nf < lim ? nf : lim;						/* >= 1: lane 0 re-reads nb; lanes past the block read 0 */
			nb = __shfl_sync(MAB_FULL, ns, n); q = __shfl_sync(MAB_FULL, qs, n);
			dir >>= 2 * n; mi -= 2 * n; g0 -= dec * n; g1 -= dec * n;
			pacc = (pacc << (2 * n)) | (0x5555555555555555ull >> (64 - 2 * n)); nacc += 2 * n;
			goto L_D_TAIL;
		}
	}
	if(g0 == 0 || g1 == 0) { w.state = mab_ts_d; goto term; }
	POP(0, 1);
	POP(1, 2);
L_D_TAIL:
	if(!(nb & 2)) { goto L_V_HEAD; }
	goto L_D_HEAD;
L_H_HEAD:
	if(nb & 4) {																/* e bit clear: short gap */
		if(g0 == 0) { w.state = mab_ts_h0; goto term; }
		w.gf[0]++; POP(0, 3);
		goto L_D_HEAD;
	}
	w.gi[0]++;
L_H_BODY:
	if(g0 == 0) { w.state = mab_ts_h1; goto term; }
	w.ge[0]++; POP(0, 4);
L_H_TAIL:
	if(!((nb & 1) && !(nb & 4))) { goto L_H_BODY; }								/* (~h & e) bit clear */
	goto L_D_HEAD;
L_V_HEAD:
	if(nb & 8) {
		if(g1 == 0) { w.state = mab_ts_v0; goto term; }
		w.gf[1]++; POP(1, 5);
		goto L_D_TAIL;
	}
	w.gi[1]++;
L_V_BODY:
	if(g1 == 0) { w.state = mab_ts_v1; goto term; }
	w.ge[1]++; POP(1, 6);
L_V_TAIL:
	if(!((nb & 2) && !(nb & 8))) { goto L_V_BODY; }
	goto L_D_TAIL;

reload:
	{
		/* _trace_test_bulk (3052-3060), _trace_reload_block (3032-3043), _trace_reload_tail (3009-3026); the header words of
		 * entry b - 1 were fetched when block b was entered (pf_meta, pf_dir) */
		uint32_t meta;
		#define TEST_BULK(_ok) { \
			int32_t _ga = g0 - META_ACNT(meta), _gb = g1 - META_BCNT(meta); \
			_ok = !(W > _ga) && !(W > _gb); \
			if(_ok) { g0 = _ga; g1 = _gb; } \
		}
		#define RELOAD_HEAD() { \
			do { b = c.blk[b].link; } while(c.blk[b].xstat & MAB_X_HEAD); \
			meta = BLK_META(b); \
			int _cnt = META_ACNT(meta) + META_BCNT(meta); \
			mi = _cnt - 1; dir = c.blk[b].dir_mask >> (MAB_BLK - _cnt); \
		}
		#define RELOAD_BLOCK() { \
			b--; mi = MAB_BLK - 1; dir = pf_dir; meta = pf_meta; \
			if(META_XSTAT(meta) & MAB_X_HEAD) { RELOAD_HEAD(); } \
		}
		if(bulk) {																/* _trace_bulk_load_n (3070-3082) */
			RELOAD_BLOCK();
			int ok; TEST_BULK(ok);
			if(!ok) {
				if(q >= (uint32_t)W) { goto term; }
				g1 += (int32_t)(q - save); g0 += (int32_t)(save - q); save = HEAD_CNT; bulk = 0; dec = 1;
			}
		} else {																/* _trace_tail_load_n (3083-3099) */
			if(META_XSTAT(pf_meta) & MAB_X_HEAD) {
				b--; RELOAD_HEAD();
			} else {
				RELOAD_BLOCK();
				if(--save >= HEAD_CNT) { int ok; TEST_BULK(ok); if(ok) { save = q; bulk = 1; dec = 0; } }
			}
		}
		#undef RELOAD_HEAD
		#undef TEST_BULK
		#undef RELOAD_BLOCK
		FLUSH_PATH();
		stage_block(c, w, b, tile8); cofs = MAB_TBUF * w.cur;
		PREFETCH_META();
		switch(ret) { case 1: goto R1; case 2: goto R2; case 3: goto R3; case 4: goto R4; case 5: goto R5; default: goto R6; }
	}
term:
	FLUSH_PATH();
	w.blk = b; w.mi = mi; w.q = q & 0xff; w.gidx[0] = g0; w.gidx[1] = g1;
	w.pacc = pacc; w.nacc = nacc; w.widx = widx; w.pidx = (uint32_t)(widx + 1) * 32u - nacc;
	__syncwarp();
	#undef POP
	#undef NIB
	#undef TIDX
	#undef FLUSH_PATH
	#undef BLK_META
	#undef META_XSTAT
	#undef META_ACNT
	#undef META_BCNT
	#undef PREFETCH_META
}

/* gaba_dp_trace (gaba.c:3244-3393).  Allocates the alignment record from the result pool (lane 0 bumps the pointer),
 * returns its word offset or UINT64_MAX when the path left the band / the pool is full (err set in that case). */
__device__ inline uint64_t dp_trace(DpCtx &c, int32_t ti, uint32_t *pool, uint64_t pool_cap, BatchCounters *ctr, uint32_t *tile32)
{
	const DevParams &P = *c.P;
	const TailRec *t = &c.tails[ti];
	Leaf lf; lf.plen = 0; lf.blk = 0; lf.p = 0; lf.q = 0; lf.gidx[0] = lf.gidx[1] = lf.sgidx[0] = lf.sgidx[1] = 0;
	if(!(t->bpos < -1)) { leaf_search(c, ti, lf); }
	uint64_t plen = lf.plen;
	uint32_t sn = t->ascnt + t->bscnt + 2, npw = (uint32_t)((plen + 31) / 32 + 1);
	uint64_t words = MAB_ALN_HDR + 8ull * sn + npw + 1;
	unsigned long long ofs = 0;
	if(c.lane == 0) { ofs = atomicAdd(&ctr->pool_top, (unsigned long long)words); }
	ofs = __shfl_sync(MAB_FULL, ofs, 0);
	if(ofs + words > pool_cap) { c.err |= MAB_ERR_POOL_OVF; return 0xffffffffffffffffull; }
	uint32_t *rec = pool + ofs;
	Trace w;
	w.blk = lf.blk; w.mi = (int32_t)lf.p; w.q = lf.q; w.state = mab_ts_d;
	for(int i = 0; i < 2; i++) { w.gidx[i] = lf.gidx[i]; w.sgidx[i] = lf.sgidx[i]; w.tail[i] = ti; w.ofs[i] = 0; w.id[i] = 0; w.gi[i] = w.ge[i] = w.gf[i] = 0; }
	if(plen >= 0x80000000ull) { c.err |= MAB_ERR_DP_OVF; return 0xffffffffffffffffull; }	/* 2^31 path bits: not a read */
	w.pidx = (uint32_t)plen; w.path = rec + MAB_ALN_HDR + 8ull * sn;
	w.sblk = w.pblk = -1; w.cur = 0;
	w.pacc = 1; w.widx = (int32_t)(plen >> 5); w.nacc = 32u - ((uint32_t)plen & 31u);	/* sentinel bit at plen (gaba.c:3287), zeros above it */
	if(c.lane == 0) {
		w.path[(plen >> 5) + 1] = 0;
		if(plen == 0) { w.path[0] = 1u; }												/* nothing will be popped */
	}
	uint32_t nseg = 0;
	uint32_t fuel = sn + 8;
	/* the remaining-bits counter comes out of the walk, whose data-dependent branches the compiler cannot prove warp-uniform:
	 * re-broadcast it so that this loop (and with it everything after the trace) counts as convergent code */
	while(__shfl_sync(MAB_FULL, w.pidx, 0) != 0) {
		if(fuel-- == 0) { c.err |= MAB_ERR_DP_OVF; stage_drain(); return 0xffffffffffffffffull; }		/* more segments than sections: cannot happen */
		if(w.gidx[0] < (int32_t)((w.state & MAB_TS_H) != 0)) { trace_reload_section(c, w, 0); }
		if(w.gidx[1] < (int32_t)((w.state & MAB_TS_V) != 0)) { trace_reload_section(c, w, 1); }
		trace_core(c, w, (uint8_t *)tile32);
		if(w.q >= (uint32_t)c.W) { stage_drain(); return 0xffffffffffffffffull; }					/* out of band: abort (gaba.c:3324-3328) */
		/* trace_push_segment (gaba.c:2865-2895): slots fill from the back */
		if(c.lane == 0 && nseg < sn) {
			uint32_t *s = rec + MAB_ALN_HDR + 8ull * (sn - 1 - nseg);
			uint64_t ppos = w.pidx;
			s[0] = w.id[0]; s[1] = w.id[1];
			s[2] = w.ofs[0] + (uint32_t)w.gidx[0]; s[3] = w.ofs[1] + (uint32_t)w.gidx[1];
			s[4] = (uint32_t)(w.sgidx[0] - w.gidx[0]); s[5] = (uint32_t)(w.sgidx[1] - w.gidx[1]);
			s[6] = (uint32_t)ppos; s[7] = (uint32_t)(ppos >> 32);
		}
		nseg++;
		w.sgidx[0] = w.gidx[0]; w.sgidx[1] = w.gidx[1];
	}
	if(c.lane == 0) {
		/* identity (gaba.c:3334-3355): only the a-side counters enter the sum (_mm_mul_epi32 multiplies the low lane) */
		uint32_t gc0 = w.ge[0] + w.gf[0], gc1 = w.ge[1] + w.gf[1];
		int32_t g0 = (int32_t)((uint32_t)P.gi * w.gi[0] + (uint32_t)P.ge * w.ge[0] + (uint32_t)P.gfa * w.gf[0]);
		uint64_t dlen = ((uint32_t)plen - gc1 - gc0) >> 1;
		int64_t score = t->max, dsc = score + g0;
		double identity = dlen == 0 ? 0.0 : __dsub_rn(__dmul_rn(__ddiv_rn((double)dsc, (double)dlen), P.imx), P.xmx);
		unsigned long long ib; memcpy(&ib, &identity, 8);
		rec[0] = (uint32_t)score; rec[1] = (uint32_t)((uint64_t)score >> 32); rec[2] = (uint32_t)ib; rec[3] = (uint32_t)(ib >> 32);
		rec[4] = gc0; rec[5] = gc1; rec[6] = (uint32_t)dlen; rec[7] = nseg < sn ? nseg : sn; rec[8] = (uint32_t)plen; rec[9] = npw; rec[10] = sn;
		rec[11] = rec[12] = rec[13] = rec[14] = rec[15] = 0;
	}
	if(nseg > sn) { c.err |= MAB_ERR_POOL_OVF; }
	stage_drain();
	return (uint64_t)ofs;
}

}  // namespace mab
