/*
 * fmx_oracle.h -- CPU oracle for the FM-index query path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference
 * crate's (ajalab/fm-index 0.3.1) count / locate / extract path and of the
 * rank/select/access semantics of its un-vendored dependency
 * vers-vecs =1.10.1 (Cargo.toml:16).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The
 * product (fm-index_b200/) never links, imports or calls anything in here.
 *
 * Parity pinning: vers-vecs' source is absent from /root/reference and no
 * Rust toolchain exists in this image, so the crate itself cannot run here.
 * The oracle is pinned against every known-answer vector the reference's
 * own tests hold for this path (see tests/test_oracle_golden.py, which cites
 * each file:line) and against a naive scan following tests/testutil/mod.rs.
 *
 * Every function cites the reference file:line it restates
 * (paths relative to /root/reference).
 */
#ifndef FMX_ORACLE_H
#define FMX_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_index orc_index;

enum { ORC_FM = 0, ORC_RLFM = 1, ORC_MULTI = 2 };
/* wrapper.rs:37-42, 61-82 */
enum { ORC_SEARCH = 0, ORC_SEARCH_PREFIX = 1, ORC_SEARCH_SUFFIX = 2, ORC_SEARCH_EXACT = 3 };

#define ORC_NONE UINT64_MAX /* Option::None */

/* Construction (frontend.rs:195-267).  level < 0 => count-only (DiscardedSuffixArray).
 * Returns NULL and fills err (the reference's Error::InvalidText message,
 * sais.rs:128-139) when the text is rejected. */
orc_index *orc_build(const uint8_t *text, uint64_t n, uint64_t max_character, int kind,
                     int level, char *err, size_t errlen);
/* Same, but the suffix array is supplied by the caller (bench-only speed path for GB-scale
 * texts).  The array is VERIFIED first (orc_check_suffix_array, linear time): the suffix array
 * of a text is unique, so a verified one gives exactly the index orc_build would. */
orc_index *orc_build_from_sa(const uint8_t *text, uint64_t n, uint64_t max_character,
                             int kind, int level, const uint64_t *sa, char *err,
                             size_t errlen);
void orc_free(orc_index *idx);

/* 0 iff sa is the suffix array of text (independent linear-time checker). */
int orc_check_suffix_array(const uint8_t *text, uint64_t n, const uint64_t *sa);

/* Suffix array alone (sais.rs:115-144 incl. validation). returns 0 ok, -1 invalid text. */
int orc_suffix_array(const uint8_t *text, uint64_t n, uint64_t *sa, char *err, size_t errlen);

uint64_t orc_len(const orc_index *idx);           /* backend.rs:25 */
uint64_t orc_pieces_count(const orc_index *idx);  /* multi_pieces.rs:220-222 */
uint64_t orc_heap_bits(const orc_index *idx);

/* backend primitives (backend.rs:5-40) */
uint64_t orc_get_l(const orc_index *idx, uint64_t i);
uint64_t orc_lf_map(const orc_index *idx, uint64_t i);
uint64_t orc_lf_map2(const orc_index *idx, uint64_t c, uint64_t i);
uint64_t orc_get_f(const orc_index *idx, uint64_t i);
uint64_t orc_fl_map(const orc_index *idx, uint64_t i); /* ORC_NONE for None */
uint64_t orc_get_sa(const orc_index *idx, uint64_t i, uint64_t *steps_out);
uint64_t orc_piece_id(const orc_index *idx, uint64_t i);
uint64_t orc_sample_get(const orc_index *idx, uint64_t i); /* sample.rs:46-60; ORC_NONE */

/* introspection used by the golden tests */
uint64_t orc_cs(const orc_index *idx, uint64_t c);
uint64_t orc_cs_len(const orc_index *idx);
uint64_t orc_rlfm_runs(const orc_index *idx);
uint64_t orc_rlfm_s(const orc_index *idx, uint64_t i);
int orc_rlfm_b(const orc_index *idx, uint64_t i);
int orc_rlfm_bp(const orc_index *idx, uint64_t i);
uint64_t orc_doc(const orc_index *idx, uint64_t k);
uint64_t orc_first_row(const orc_index *idx);
uint32_t orc_sample_level(const orc_index *idx);
uint32_t orc_sample_word_size(const orc_index *idx);

/* wrapper.rs:103-124: one backward search.  use_init!=0 refines (init_s,init_e).
 * returns the number of loop iterations executed (for the roofline numerator);
 * -1 if a pattern character exceeds max_character (the reference panics there). */
int64_t orc_search(const orc_index *idx, int mode, const uint8_t *pat, uint64_t m,
                   int use_init, uint64_t init_s, uint64_t init_e, uint64_t *s, uint64_t *e);

/* batched + OpenMP (the stand-in for a rayon pool over &index).  pat_off has npat+1
 * entries.  steps (nullable) receives executed iterations per pattern.
 * returns 0, or -1 if any pattern hit a character > max_character. */
int orc_search_batch(const orc_index *idx, int mode, const uint8_t *pat,
                     const uint64_t *pat_off, uint64_t npat, const uint64_t *init_s,
                     const uint64_t *init_e, uint64_t *s, uint64_t *e, uint32_t *steps,
                     int nthreads);

/* wrapper.rs:203-217 + 238-248: enumerate matches of row ranges in SA-row order.
 * pass 1 (positions==NULL): fills hit_off[npat+1] (exclusive prefix of hit counts).
 * pass 2: fills positions / piece_ids (nullable) / lf_steps (nullable, total LF steps). */
int orc_locate_batch(const orc_index *idx, int prefix_only, const uint64_t *s,
                     const uint64_t *e, uint64_t npat, uint64_t *hit_off,
                     uint64_t *positions, uint64_t *piece_ids, uint64_t *lf_steps,
                     int nthreads);

/* wrapper.rs:143-183: k chars per row. backward always yields k chars;
 * forward may stop early on MultiPieces (fl_map None). out is nrows*k bytes,
 * out_len (nullable) per row. */
void orc_extract_batch(const orc_index *idx, const uint64_t *rows, uint64_t nrows,
                       uint32_t k, int forward, uint8_t *out, uint32_t *out_len,
                       int nthreads);

int orc_max_threads(void);

/* ---- texts and patterns of wider characters (character.rs:24-42: u16, u32, u64, usize) ----
 * char_width = bytes per character (2, 4 or 8; 1 gives the u8 entry points above); every character and
 * max_character must be below 2^32 - 1 (the reference allocates max_character + 1 counters, sais.rs:9-19).
 * Offsets count characters.  Everything else (locate, primitives, accessors) is shared with the u8 index. */
orc_index *orc_build_w(const void *text, uint32_t char_width, uint64_t n, uint64_t max_character, int kind, int level,
                       char *err, size_t errlen);
int orc_suffix_array_w(const void *text, uint32_t char_width, uint64_t n, uint64_t max_character, uint64_t *sa, char *err,
                       size_t errlen);
int orc_search_batch_w(const orc_index *idx, int mode, const void *pat, uint32_t char_width, const uint64_t *off,
                       uint64_t npat, const uint64_t *init_s, const uint64_t *init_e, uint64_t *s, uint64_t *e,
                       uint32_t *steps, int nthreads);
void orc_extract_batch_w(const orc_index *idx, const uint64_t *rows, uint64_t nrows, uint32_t k, int forward, void *out,
                         uint32_t char_width, uint32_t *out_len, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
