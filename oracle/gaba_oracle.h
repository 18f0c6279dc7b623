/*
 * gaba_oracle.h -- TEST INFRASTRUCTURE ONLY.  CPU restatement (plain scalar C, one lane at a time) of the GABA
 * adaptive-banded DP that minialign vendors as gaba.c.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may use it; the product path (minialign_b200/csrc) never includes or links this.
 *
 * Pinned against: the reference's own objects gaba.{combined}.{16,32,64}.o driven through oracle/ref_harness.c
 * (tests/test_oracle_vs_ref.py) and the golden vectors under tests/golden/ that script generated.
 */
#ifndef GABA_ORACLE_H
#define GABA_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#define ORA_WMAX   64
#define ORA_BLK    32

/* status flags, gaba.h:45-51 */
#define ORA_UPDATE_A 0x000f
#define ORA_UPDATE_B 0x00f0
#define ORA_TERM     0x8000

/* block status, gaba.c:670-680 */
#define ORA_X_TERM   0x80
#define ORA_X_HEAD   0x20
#define ORA_X_MERGE  0x40
#define ORA_X_ROOT   0x60

typedef struct {
	int8_t score_matrix[16];
	int8_t gi, ge, gfa, gfb;
	int8_t xdrop;
} ora_params_t;

/* a sequence section (gaba.h:131-135); `rev` replaces the reference's mirrored ("phantom") pointer trick
 * (gaba.h:151-155): position i of a rev section reads base[len-1-i] complemented */
typedef struct {
	uint32_t id, len;
	uint8_t const *base;
	uint32_t rev;
} ora_section_t;

typedef struct { uint64_t h, v, e, f; } ora_mask_t;

/* one entry of the block chain: either a real block (<=32 vectors) or a head (the reference's phantom block) */
typedef struct {
	int8_t dh[ORA_WMAX], dv[ORA_WMAX], de[ORA_WMAX], df[ORA_WMAX];	/* diff vectors after the last vector */
	int8_t acc, xstat, acnt, bcnt;
	uint32_t dir_mask;
	uint64_t max_mask;
	int64_t link;				/* head only: index of the previous entry (-1 for the root) */
	uint64_t na, nb;			/* stream positions (consumed base counts) at the START of this block */
	ora_mask_t mask[ORA_BLK];
} ora_block_t;

typedef struct {
	uint32_t aid, bid, ascnt, bscnt;
	int64_t apos, bpos;
	int64_t max;
	uint32_t status;
} ora_fill_t;

typedef struct {
	uint8_t cha[ORA_WMAX], chb[ORA_WMAX];	/* char window: a codes (0..4), b codes (0,4,8,12,2) per lane */
	int8_t xd[ORA_WMAX];
	int16_t md[ORA_WMAX];
	int16_t mdrop; uint16_t istat; uint32_t pridx;
	uint32_t ridx[2], adv[2];			/* [0]=a, [1]=b */
	int64_t tail;						/* previous tail index, -1 for none */
	int64_t last_blk;					/* index of the last block entry belonging to this tail */
	ora_fill_t f;
} ora_tail_t;

typedef struct { uint32_t aid, bid, apos, bpos; uint64_t plen; } ora_pos_t;

typedef struct { uint32_t aid, bid, apos, bpos, alen, blen; uint64_t ppos; } ora_seg_t;

typedef struct {
	int64_t score;
	double identity;
	uint32_t agcnt, bgcnt, dcnt, slen, plen;
	ora_seg_t *seg;		/* slen entries, forward order */
	uint32_t *path;		/* (plen + 31) / 32 + 1 words (bit plen is the sentinel) */
	uint32_t npath;
} ora_aln_t;

typedef struct {
	/* constants */
	int W;
	int8_t sb[16], adjh, adjv, ofsh, ofsv, gfh, gfv, tx;
	int8_t gi, ge, gfa, gfb;
	double imx, xmx;
	ora_tail_t root;		/* template root tail */
	ora_block_t rootblk;	/* template root phantom */

	/* stack (reset by ora_dp_flush) */
	ora_block_t *blk; size_t nblk, mblk;
	ora_tail_t *tl; size_t ntl, mtl;
	uint8_t *sa, *sb_; size_t na, nb, msa, msb;	/* consumed-base streams (a codes / b codes) */

	/* reader work (gaba.c:400-423) */
	uint32_t rem[2], sridx[2], pridx, ids[2];
	int32_t ofsd;
	ora_section_t sec[2];
	int64_t wtail;
	int8_t xd[ORA_WMAX]; int16_t md[ORA_WMAX];
	uint64_t vec_count;		/* statistics: #vectors filled since init */
} ora_dp_t;

int  ora_dp_init(ora_dp_t *dp, ora_params_t const *p, int W);		/* returns nonzero on unsupported params */
void ora_dp_clean(ora_dp_t *dp);
void ora_dp_flush(ora_dp_t *dp);
int64_t ora_dp_fill_root(ora_dp_t *dp, ora_section_t const *a, uint32_t apos, ora_section_t const *b, uint32_t bpos, uint32_t pridx);
int64_t ora_dp_fill(ora_dp_t *dp, int64_t prev, ora_section_t const *a, ora_section_t const *b, uint32_t pridx);
ora_pos_t ora_dp_search_max(ora_dp_t *dp, int64_t tail);
ora_aln_t *ora_dp_trace(ora_dp_t *dp, int64_t tail);				/* NULL when the path leaves the band */
void ora_aln_free(ora_aln_t *a);
uint64_t ora_dump_cigar_reverse(char *buf, uint32_t const *path, uint64_t offset, uint64_t len);
uint64_t ora_dump_cigar_forward(char *buf, uint32_t const *path, uint64_t offset, uint64_t len);

static inline ora_fill_t const *ora_fill(ora_dp_t const *dp, int64_t t) { return(&dp->tl[t].f); }
#endif
