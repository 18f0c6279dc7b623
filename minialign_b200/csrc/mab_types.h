/*
 * mab_types.h -- plain structs shared by the host driver (mab_host.cpp) and the device code (mab_device.cuh).
 * Layouts in HBM are described in DESIGN.md section 3.
 */
#pragma once
#include <stdint.h>

namespace mab {

#define MAB_WMAX		64
#define MAB_BLK			32
#define MAB_KH_CAP		1024u		/* slots of the per-read dedup hash (reference starts at 256 and doubles) */
#define MAB_MAX_TAILS	24
/* per-warp traceback tiles in shared memory: two buffers (the block being walked and the prefetched next one), each a raw
 * copy of a block's mask stream: 16 rows of 32 words (row j = vectors 2j and 2j+1, one flag byte per cell, word of lane l =
 * cells 2l, 2l+1 of both vectors).  The row pitch is 33 words so that the 16 lanes of the diagonal-run lookahead, which read
 * consecutive rows at nearly the same column, hit different banks */
#define MAB_TPITCH 132
#define MAB_TBUF (16 * MAB_TPITCH)
#define MAB_TILE_WORDS (2 * MAB_TBUF / 4)
#define MAB_SC_SMALL		2944u		/* largest shared-memory staging of the ordinary size class of k_sortchain (16 B per seed + 2 KB = 48 KB) */
#define MAB_SC_MAX		6016u		/* ... and of the seed-rich class (96 KB, opt-in dynamic shared memory) */
#define MAB_SC_CLASSES		8u			/* size classes of k_sortchain (one launch each) */
#define MAB_WARPS_PER_CTA 4
#ifndef MAB_EXT_CTAS_PER_SM
#define MAB_EXT_CTAS_PER_SM 6		/* resident CTAs of the persistent extend kernel per SM (register budget = 65536 / (128 x this)) */
#endif

/* GABA status bits (gaba.h:45-51) and block status (gaba.c:670-680) */
#define MAB_UPDATE_A	0x000fu
#define MAB_UPDATE_B	0x00f0u
#define MAB_TERM		0x8000u
#define MAB_X_TERM		0x80
#define MAB_X_HEAD		0x20
#define MAB_X_MERGE		0x40
#define MAB_X_ROOT		0x60

/* error bits per read */
#define MAB_ERR_SEED_OVF	0x01u
#define MAB_ERR_KH_OVF		0x02u
#define MAB_ERR_BIN_OVF		0x04u
#define MAB_ERR_POOL_OVF	0x08u
#define MAB_ERR_DP_OVF		0x10u
#define MAB_ERR_TAIL_OVF	0x20u
#define MAB_ERR_WS_OVF		0x40u		/* the batch's workspace buffer was too small for this read (host grows it and re-runs) */

/* root (phantom) vectors of one band width: gaba_init_diff_vectors / gaba_init_middle_delta / gaba_init_phantom
 * (gaba.c:3684-3791) evaluated on the host once per context */
struct RootTpl {
	int8_t dh[MAB_WMAX], dv[MAB_WMAX], de[MAB_WMAX], df[MAB_WMAX];
	int16_t md[MAB_WMAX];
	int32_t mdrop, init_max;
};

/* uniform constants, passed to every kernel by value */
struct DevParams {
	const uint8_t *idx;				/* relocatable mm_idx_t image in HBM */
	uint64_t bkt_ofs, bkt_mask, seq_ofs;
	uint32_t b, w, k, n_occ, occ[8], n_ref;
	uint32_t twlen, tglen, min_score;
	float min_ratio;
	double mcoef, xcoef, imx, xmx;
	int8_t sb[16];
	int32_t adjh, adjv, ofsh, ofsv, gfh, gfv, tx;
	/* the same constants as packed H8 pairs (value in the high byte of both 16-bit halves; "+1" = plus one ulp, see mab_dp.cuh)
	 * ready to be used as constant-bank operands: K_OFS = ofsh = ofsv */
	uint32_t K_GFH1, K_GFV1, K_ADJH1, K_ADJV1, K_OFS;
	uint32_t K_M1;					/* 0xffffffff as a run-time value (a multiplier operand the assembler must not fold, see mab_dp.cuh) */
	int32_t gi, ge, gfa, gfb;
	RootTpl root[3];				/* W = 64, 32, 16 */
};

/* one read of the batch */
struct ReadRec {
	uint64_t seq_ofs;				/* into the batch's base block */
	uint32_t len;
	uint32_t state;					/* 0 = active, 1 = finished */
	uint32_t tot_seeds, tot_resc;	/* count pass: upper bounds over all rescue rounds */
	uint64_t ws_ofs;				/* byte offset of this read's workspace */
	uint32_t seed_cap, root_cap, resc_cap, bin_cap;
	uint32_t n_seed, seed_n, n_resc, presc, n_root, n_next, n_res, nbin;
	uint32_t kh_mask, kh_cnt, kh_ub, n_words;
	uint64_t result_ofs;			/* word offset into the result pool, valid when result_words != 0 */
	uint32_t result_words, err;
	/* The reference's per-thread buffer keeps `rlen` of the LAST chain it loaded, and mm_search_load_root tests the first
	 * seed of the next read against that stale value (minialign.c:3865 runs before 3873).  With -t1 this is a sequential
	 * dependency between consecutive reads; it is reproduced by speculating (rlen_in) and verifying on the host. */
	uint32_t rlen_in;				/* value of self->rlen when this read starts; MAB_RLEN_OWN = assume the first chain's own reference */
	uint32_t rlen_cur;				/* running value (carried across rescue rounds) */
	uint32_t rlen_used;				/* the value the first load_root actually compared against */
	uint32_t dep_apos, dep_flags;	/* first load_root: raw a-position; bit0 = valid, bit1 = (bpos >= qlen) */
	uint32_t n_rec;					/* minimizer records left by k_seed_scan for k_seed_expand */
	uint32_t tot_seeds0, _pad3;		/* seeds of round 0 alone (occurrence count <= occ[0]): what k_sortchain has to stage in round 0 */
};
#define MAB_RLEN_OWN 0xffffffffu

/* per-read workspace = [seeds 16B x seed_cap][root 8B x root_cap][next 8B x root_cap][resc 16B x resc_cap]
 *                      [kh 16B x MAB_KH_CAP][bin 8B x bin_cap] */
struct WsLayout { uint64_t seed, root, next, resc, kh, bin, total; };
static inline
#if defined(__CUDACC__) || defined(MAB_EMU)
__host__ __device__
#endif
WsLayout ws_layout(uint32_t seed_cap, uint32_t root_cap, uint32_t resc_cap, uint32_t bin_cap)
{
	WsLayout l;
	l.seed = 0;
	l.root = l.seed + 16ull * seed_cap;
	l.next = l.root + 8ull * root_cap;
	l.resc = l.next + 8ull * root_cap;
	l.kh = l.resc + 16ull * resc_cap;
	l.bin = l.kh + 16ull * MAB_KH_CAP;
	l.total = (l.bin + 8ull * bin_cap + 255) & ~255ull;
	return l;
}

/* DP block-chain entry: the reference's gaba_block_s / gaba_phantom_s (gaba.c:308-322) minus the mask array, which
 * lives in a separate stream (1 KB per real block of the traced pass), plus the char windows and remaining lengths at
 * the START of the block so a block can be replayed without re-deriving the fetch state */
struct BlkEntry {
	uint16_t dh[32], dv[32], de[32], df[32];	/* lane l holds cells 2l (low byte) and 2l+1 (high byte), int8 each */
	uint16_t cha[32], chb[32];					/* char windows at block start, same packing */
	int8_t acc, xstat, acnt, bcnt;
	uint32_t dir_mask;
	uint32_t mm_lo, mm_hi;						/* max_mask planes: bit l of lo/hi = cell 2l / 2l+1 */
	int32_t link;								/* head only: previous entry (-1 at the root) */
	uint32_t arem, brem;						/* remaining section lengths at block start */
	uint32_t tail;								/* tail record this block belongs to (gives the sections) */
	uint32_t _pad;
};

struct SecDesc { uint64_t base; uint32_t len, id, rev, _pad; };	/* base = absolute device address */

/* tail record: gaba_joint_tail_s (gaba.c:351-366) */
struct TailRec {
	uint16_t cha[32], chb[32];
	uint16_t xd[32];							/* int8 pairs */
	int16_t md[MAB_WMAX];
	int32_t mdrop; uint32_t istat, pridx;
	uint32_t ridx[2], adv[2];
	int32_t tail, last_blk;
	uint32_t aid, bid, ascnt, bscnt;
	int64_t apos, bpos, max;
	uint32_t status, _pad;
	SecDesc sec[2];
};

/* per-warp DP arena header (lane 0 <-> warp communication) */
struct SlotHdr {
	uint32_t nblk, ntail;
	uint64_t n_vectors, n_fill, n_trace;
};

/* alignment record in the result pool (u32 words):
 *   [0..1] score  [2..3] identity bits  [4] agcnt [5] bgcnt [6] dcnt [7] slen [8] plen [9] npw [10] sn [11..15] 0
 *   then sn x 8 words of segment slots (the LAST slen are valid, forward order), then npw (+1 spare) path words */
#define MAB_ALN_HDR 16

/* read result record in the pool: [0] n_res, then per result {score, n_aln, plen, lb, ub} followed by n_aln x {aln word
 * offset (lo, hi)} */

struct BatchCounters {
	unsigned long long pool_top;	/* result pool bump pointer (u32 words) */
	unsigned int work_next;			/* persistent-warp work counter */
	unsigned int err_any;
	unsigned long long n_vectors, n_fill, n_trace;
	unsigned long long ws_need;		/* k_size: bytes of per-read workspace the batch asks for */
	unsigned int n_redo;			/* k_rlen_verify: reads whose first-root test was mis-speculated (re-activated for a redo pass) */
	unsigned int chain_valid, chain_rlen;	/* the reference thread's `rlen` after the last chain-loading read of the batch */
	unsigned int fd_valid, fd_idx, fd_apos, fd_flags, fd_used;	/* first chain-loading read: its predecessor lives in the previous batch */
	unsigned int _pad;
};

/* ---- text path (FASTA/FASTQ bytes in, SAM bytes out) ---- */
/* one record of a text chunk (bseq_seq_t, minialign.c:1589-1596, as offsets into the chunk) */
struct TextRec {
	uint64_t name_ofs;				/* first byte of the name */
	uint64_t seq_beg, seq_end;		/* text range of the sequence lines (newlines inside are skipped) */
	uint64_t qual_ofs;				/* FASTQ: first byte of the quality line */
	uint32_t name_len, len;			/* len = number of bases */
	uint32_t flags, _pad;
	uint64_t sam_ofs;				/* where this read's SAM lines start in the output */
	uint64_t sam_len;
};
#define MAB_TR_DROPPED	1u			/* shorter than the reader's min_len (1): the record vanishes (minialign.c:2077) */
#define MAB_TR_HASQUAL	2u

#define MAB_TXT_EFORMAT	1u			/* record layout the device parser does not take (see mab_text.cuh) */
#define MAB_TXT_EMARKS	2u			/* mark array too small (host grows it and parses again) */
#define MAB_TXT_ERECS	4u			/* record array too small (ditto) */
struct TextCounters {
	unsigned long long n_mark;		/* record starts (FASTA) or line ends (FASTQ) found */
	unsigned long long tot_len, span, sam_total;
	unsigned int n_rec, maxlen, err, fastq;
	unsigned int mapq_flag, _pad[3];
};

}  // namespace mab
