/*
 * mab_sam.h -- host-side SAM writer (plain C++, no CUDA): the text formatting the reference does in its printer
 * (minialign.c:5095-5426) and path parser (gaba_parse.h:107-263), fed with the flat per-read results of mab_result().
 */
#pragma once
#include "../../../include/minialign_b200.h"
#include <cstdint>
#include <string>
#include <vector>

struct MabSamRef { const char *name; uint32_t l_name; uint32_t l_seq; const uint8_t *seq; };		/* mm_idx_seq_t view */
struct MabSamRead { const char *name; uint32_t l_name; const uint8_t *seq; uint32_t l_seq; const char *qual; };	/* bseq_seq_t view; qual may be null */

/* tag bits (MAB_TAG_*, MAB_OMIT_REP): include/minialign_b200.h */

extern "C" {
/* header: @HD, @SQ per reference, @PG with the command line (minialign.c:5095-5120) */
void mab_sam_header(std::string &out, const MabSamRef *refs, uint32_t n_ref, const char *version, const char *cmdline);
/* one read: `words`/`n_words` as returned by mab_result (0 words = unmapped) (minialign.c:5126-5426) */
void mab_sam_record(std::string &out, const MabSamRef *refs, const MabSamRead *read, const uint32_t *words, uint64_t n_words, uint32_t tags);
/* C entry used by the tests: formats into a malloc'ed buffer */
char *mab_sam_format_c(const MabSamRef *refs, uint32_t n_ref, const MabSamRead *read, const uint32_t *words, uint64_t n_words, uint32_t tags, uint64_t *len);
void mab_sam_free(char *p);
uint32_t mab_sam_parse_tags(const char *list);		/* "AS,XS,NM" -> bit set (minialign.c:5925-5938) */
}
