/*
 * mm_oracle.h -- TEST INFRASTRUCTURE ONLY.  CPU restatement of minialign's per-read mapper (minialign.c:2349-4474:
 * sketch, index probe, seed collection, exact radix sort, array chaining, the mm_extend state machine, pruning,
 * supplementary/secondary split and MAPQ).  Uses gaba_oracle.c for the DP.  See gaba_oracle.h for who may call this.
 */
#ifndef MM_ORACLE_H
#define MM_ORACLE_H
#include "gaba_oracle.h"

typedef struct {
	uint32_t k, w, b, n_occ;
	uint32_t occ[8];
	int32_t wlen, glen;
	uint32_t min_score;
	float min_ratio;
	ora_params_t gp;
} mmo_params_t;

typedef struct mmo_s mmo_t;

/* flat result layout shared by ref_harness.c (refh_dump_aln), this oracle and the C-ABI:
 *   out[0] = n_all, out[1] = n_uniq, then n_all alignments, each
 *   16 x u32 header: score(lo,hi) identity(lo,hi) agcnt bgcnt dcnt slen plen npathwords rank mapq 0 0 0 0
 *   slen x 8 x u32 segments: aid bid apos bpos alen blen ppos(lo,hi)
 *   npathwords path words */
mmo_t *mmo_init(void const *mai_blob, uint64_t size, mmo_params_t const *p);		/* blob = raw (inflated) .mai payload after the 12-byte header */
void mmo_destroy(mmo_t *m);
uint64_t mmo_sketch(mmo_t *m, uint8_t const *seq, uint32_t len, uint64_t *out, uint64_t cap);
uint32_t mmo_get(mmo_t *m, uint64_t minier, uint64_t *out, uint32_t cap);
uint64_t mmo_seed_chain(mmo_t *m, uint8_t const *seq, uint32_t len, uint32_t round,
	uint32_t *seeds, uint64_t seed_cap, uint64_t *n_total, uint32_t *roots, uint64_t root_cap, uint64_t *n_root);
uint64_t mmo_align(mmo_t *m, uint8_t const *seq, uint32_t len, uint32_t qid, uint32_t *out, uint64_t cap);
uint64_t mmo_extend(mmo_params_t const *p, uint8_t const *a, uint32_t alen, uint8_t const *b, uint32_t blen,
	uint32_t apos, uint32_t bpos, uint32_t brev, uint32_t narrow, int64_t min_score, uint32_t *res, uint32_t *aln_out, uint64_t cap);
uint64_t mmo_vec_count(mmo_t *m);
#endif
