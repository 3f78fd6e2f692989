/*
 * fmx.h -- C ABI of the B200-native batched FM-index query engine.
 *
 * This is the drop-in boundary for the one hot path of ajalab/fm-index 0.3.1:
 * backward search (count), LF walks to sampled suffix-array entries (locate) and the
 * character-extraction walks, for FMIndex / RLFMIndex / FMIndexMultiPieces with and
 * without locate support.  The reference has no FFI of its own;
 * each entry point below names the reference interface it replaces (paths relative to
 * the reference crate root).  A Rust facade binds these with `extern "C"` (see
 * INTEGRATION.md); include/fmx.hpp is the same facade in C++.
 *
 * Conventions
 *   - every function returns FMX_OK (0) or a negative fmx_status; fmx_last_error()
 *     returns a thread-local message for the last failure.  Nothing unwinds across the ABI.
 *   - positions, rows and counts are uint64_t (Rust usize).  Characters are bytes (u8).
 *   - input buffers are caller-owned; buffers returned through `T**` are library-owned
 *     and released with fmx_free().
 *   - "_device" entry points take raw device pointers resident in the index's GPU and an
 *     optional cudaStream_t (as void*, NULL = the index's own stream); they are
 *     asynchronous unless stated.  The host-buffer entry points copy H2D/D2H themselves
 *     and are synchronous.
 *   - there is no CPU fallback: without a usable CUDA device every query entry point
 *     fails with FMX_ERR_CUDA.
 */
#ifndef FMX_H
#define FMX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fmx_index fmx_index; /* opaque; owns the device-resident index */

typedef enum fmx_status {
    FMX_OK = 0,
    FMX_ERR_INVALID_TEXT = -1,  /* Error::InvalidText, src/error.rs:3-6 (same messages as sais.rs:128-139) */
    FMX_ERR_INVALID_ARG = -2,
    FMX_ERR_CUDA = -3,
    FMX_ERR_NO_LOCATE = -4,     /* index built without a sampled SA (DiscardedSuffixArray, discard.rs) */
    FMX_ERR_PATTERN_CHAR = -5,  /* a processed pattern char > max_character: the reference panics (fm_index.rs:94) */
    FMX_ERR_OOM = -6,
    FMX_ERR_UNSUPPORTED = -7,
    FMX_ERR_IO = -8,
    FMX_ERR_CAPACITY = -9       /* caller-provided output buffer too small */
} fmx_status;

/* src/frontend.rs:110-193: the three backends; locate support is chosen by `level` */
typedef enum fmx_kind {
    FMX_KIND_FM = 0,    /* FMIndex / FMIndexWithLocate            (src/fm_index.rs) */
    FMX_KIND_RLFM = 1,  /* RLFMIndex / RLFMIndexWithLocate        (src/rlfmi.rs) */
    FMX_KIND_MULTI = 2  /* FMIndexMultiPieces(/WithLocate)        (src/multi_pieces.rs) */
} fmx_kind;

/* src/wrapper.rs:37-42, 61-82: initial SA range and the prefix filter */
typedef enum fmx_mode {
    FMX_SEARCH = 0,         /* (0, n),            all rows */
    FMX_SEARCH_PREFIX = 1,  /* (0, n),            rows with L == 0 only */
    FMX_SEARCH_SUFFIX = 2,  /* (0, pieces_count), all rows */
    FMX_SEARCH_EXACT = 3    /* (0, pieces_count), rows with L == 0 only */
} fmx_mode;

#define FMX_LEVEL_COUNT_ONLY (-1)

/* How much HBM an index may spend beyond the reference's own structures (fmx_index_build_ex):
 *   COMPACT  rank structure + the caller's suffix-array samples + an L2-resident k-mer table: the reference's
 *            space (about n/2 + 4n/2^level bytes for DNA); locate walks LF to the samples.
 *   RICH     additionally the text, the FULL suffix array and its inverse (9 n bytes) and a k-mer table sized to
 *            end the search of an absent pattern: every query step that the reference answers with a chain of
 *            dependent rank probes becomes one or two memory requests, locate is SA[row].  Bit-identical results.
 *   AUTO     RICH when the rank structure does not fit the 126 MB L2 (RLFM with locate: when the text has 2^27 symbols
 *            or more) and the budgets allow, else COMPACT (FMX_MODE=compact|rich in the environment overrides AUTO). */
typedef enum fmx_index_mode { FMX_MODE_AUTO = 0, FMX_MODE_COMPACT = 1, FMX_MODE_RICH = 2 } fmx_index_mode;

const char *fmx_last_error(void);
int fmx_version(void);
int fmx_device_count(void);
void fmx_free(void *p);

/* ---------------------------------------------------------------- construction
 * Replaces Text::{new, with_max_character} (src/text.rs:28-49) + the six ::new
 * constructors (src/frontend.rs:195-267).
 * char_width = bytes per character of `text`: 1, 2, 4 or 8 for Text<u8>, <u16>, <u32>, <u64 / usize>
 * (src/character.rs:38-42).  Every character must be <= max_character, and max_character < 2^32 - 1 (the reference
 * allocates max_character + 1 counters, sais.rs:9-19, so it cannot go further either): Text::new's
 * C::max_value() is therefore available for u8 and u16 texts, wider ones use Text::with_max_character.
 * An index over wide characters takes patterns (pat, fixed_len, pat_off: all counted in CHARACTERS) and returns
 * extracted characters in that same width; it is a compact index (no k-mer tables, no packed patterns), and with
 * max_character > 255 its rank structure is the WIDE layout (csrc/fmx_layout.h).  Multi-GPU groups take u8 texts.
 * max_character = 255 is Text::new for u8; anything else is Text::with_max_character.
 * level = FMX_LEVEL_COUNT_ONLY builds the count-only variant; level >= 0 the
 * ...WithLocate variant with that sampling level (sample.rs:21-44).
 * Text validation and its messages follow sais.rs:128-139. */
int fmx_index_build(const void *text, uint64_t n, uint32_t char_width, uint64_t max_character,
                    int kind, int level, int device, fmx_index **out);

int fmx_index_build_ex(const void *text, uint64_t n, uint32_t char_width, uint64_t max_character,
                       int kind, int level, int device, int mode, fmx_index **out);
/* the mode an index was built with, resolved (FMX_MODE_COMPACT or FMX_MODE_RICH) */
int fmx_index_mode_of(const fmx_index *idx);

/* Host-only half of construction: text -> device-layout blob (no GPU needed). */
int fmx_blob_build(const void *text, uint64_t n, uint32_t char_width, uint64_t max_character,
                   int kind, int level, void **blob, uint64_t *blob_bytes);
int fmx_blob_build_ex(const void *text, uint64_t n, uint32_t char_width, uint64_t max_character,
                      int kind, int level, int mode, void **blob, uint64_t *blob_bytes);
/* Upload a blob (serialised once; the reference only has un-exposed serde derives,
 * fm_index.rs:13, rlfmi.rs:15, multi_pieces.rs:16, sample.rs:12). */
int fmx_index_from_blob(const void *blob, uint64_t blob_bytes, int device, fmx_index **out);
/* A copy of a device-resident index on another device of this process (blob and k-mer tables copied device to
 * device; nothing is rebuilt).  What FMX_GROUP_REPLICATE uses. */
int fmx_index_clone(const fmx_index *src, int device, fmx_index **out);
int fmx_index_save(const fmx_index *idx, const char *path);
int fmx_index_load(const char *path, int device, fmx_index **out);
void fmx_index_free(fmx_index *idx);

/* Suffix array of a text exactly as sais::build_suffix_array (sais.rs:115-144) returns it. */
int fmx_build_suffix_array(const void *text, uint64_t n, uint32_t char_width, uint64_t *sa_out);
/* The same suffix array built on the GPU (prefix doubling over radix sorts; what fmx_index_build
 * uses for texts of 2^16 symbols and more, FMX_HOST_SA=1 disables).  No text validation.
 * *rounds (nullable) receives the number of doubling rounds. */
int fmx_build_suffix_array_device(const void *text, uint64_t n, uint32_t char_width, uint64_t max_character,
                                  int device, uint64_t *sa_out, int *rounds);

/* Tuning knobs (A/B measurement; results never change): "search_persistent" 0|1 (persistent
 * per-lane-refill search kernels instead of one pattern per thread), "kmer" 0|1 (memoised first
 * search iterations), "kmer_big" 0|1 (the large HBM-resident table), "kmer_budget_mb" (rebuild the
 * large table with that many MiB of HBM; default = twice the index size, FMX_KMER_BUDGET_MB at
 * build time), "bucket" -1|0|1 (visit the batch in k-mer bucket order: auto / never / always),
 * "l2_fetch_granularity" 32|64|128, "persist_blocks_per_sm" 1..32, "pipeline_chunk" (patterns per chunk of
 * fmx_search_locate_batch's copy/compute pipeline, 0 = automatic), "locate_refill" 0|1 (per-lane refill
 * locate kernel; 2 = stable warp-level compaction), "verify" 0|1 (seed-and-verify tail of the search),
 * "stage_patterns" 0|1 (fixed-length pattern bytes through shared memory), "locate_expand" 0|1|2 (hit rows: auto / always expanded by scans / always found by
 * binary search inside the locate kernel), "locate_ranges" -1|0|1 (walk whole SA sub-ranges instead of
 * single rows: auto = RLFM indexes whose patterns average >= 8 matches / never / always), "extract_text" 0|1
 * (extraction from the resident text), "order_by_length" 0|1 (ragged batches through the fused kernel in order of
 * pattern length; off: it measured slower), "fused_defer" 0|1|2 (the fused query kernel parks the patterns its table lookup does not finish
 * in a shared-memory queue and runs them with full warps: never / with the 16-byte table entries / always; 3..7 = as 1 with
 * (resident blocks per SM, rounds between two looks at the queue) = (6,1) (5,1) (5,2) (6,2) (5,4): tuning A/B), "query_blocks" (> 0: at most this many blocks for the fused query kernels; tests), "emit_fused" 0|1 (rich locate: hit offsets and
 * positions in one pass over the ranges -- the scan's last phase and the emit kernel as one; FMX_EMIT_FUSED=0|1 at construction time), "table_ctx" 0|1 (16-byte entries of the large k-mer table that carry the 16 text
 * characters in front of one-row ranges: a pattern with <= 16 characters left after the table is finished by ONE
 * request; rebuilds the table; FMX_NO_TABLE_CTX=1 at construction time; built while afterwards at least as much HBM stays
 * free as the table takes, FMX_TABLE_CTX_FORCE=1 lifts that).  Environment at construction time: FMX_FORCE_WAVELET=1
 * keeps the binary wavelet matrix; FMX_SYM_BUDGET_MB caps the per-symbol bit-vector layout (default
 * 49152; 0 = use the quaternary wavelet matrix instead); FMX_VERIFY_BUDGET_MB caps the dense
 * seed-and-verify structures (text + full suffix array + inverse, 9 bytes per symbol; default 32768),
 * FMX_VERIFY_MIN_RANK_MB is the smallest rank structure they are built for (default 192: smaller ones
 * sit in L2), FMX_NO_VERIFY=1 never builds them.  csrc/fmx_layout.h describes every structure. */
int fmx_index_set_option(fmx_index *idx, const char *key, int64_t value);

/* SearchIndex::len (frontend.rs:35-39), heap_size (:41-44, here: device bytes),
 * HasMultiPieces::pieces_count (multi_pieces.rs:220-222) */
uint64_t fmx_index_len(const fmx_index *idx);
uint64_t fmx_index_device_bytes(const fmx_index *idx);
uint64_t fmx_index_pieces_count(const fmx_index *idx);
int fmx_index_kind(const fmx_index *idx);
int fmx_index_has_locate(const fmx_index *idx);
int fmx_index_device(const fmx_index *idx);
uint32_t fmx_index_wavelet_levels(const fmx_index *idx); /* Text::max_bits, text.rs:61-63 */
uint32_t fmx_index_sample_level(const fmx_index *idx);
/* 32-byte sectors one rank/access probe of the BWT touches in this index's device layout:
 * 1 for Q4 (max_character <= 4) and SYM (per-symbol bit vectors), ceil(L/2) for the quaternary
 * wavelet matrix, L for the binary one. */
uint32_t fmx_index_sectors_per_rank(const fmx_index *idx);
/* the device layout the builder chose: 0 binary wavelet matrix, 1 Q4, 2 WM4, 3 SYM, 4 WIDE (csrc/fmx_layout.h) */
uint32_t fmx_index_layout(const fmx_index *idx);
/* 1 when the index holds the text and the full suffix array (HBM-rich mode): extraction then reads the characters
 * from the text at SA[row] instead of walking LF / FL steps (same characters; option "extract_text" 0|1 for A/B) */
int fmx_index_has_text(const fmx_index *idx);
/* bytes per character of the text the index was built from = width of pattern and extracted characters (1, 2, 4, 8) */
uint32_t fmx_index_char_width(const fmx_index *idx);
/* characters memoised by the small (big = 0) / large (big = 1) k-mer table of fresh searches; 0 = none */
uint32_t fmx_index_kmer_k(const fmx_index *idx, int big);
/* bytes per entry of the large table: 8 ((s, e) or (s, position)), 16 (+ the text in front of one-row ranges), 0 = no such table */
uint32_t fmx_index_kmer_entry_bytes(const fmx_index *idx);

/* ---------------------------------------------------------------- search / count
 * Batched SearchIndex::search / search_prefix / search_suffix / search_exact and
 * Search::search (refinement) -- src/wrapper.rs:37-42, 61-82, 99-124.
 * Patterns are concatenated in `pat`; pattern p is pat[pat_off[p] .. pat_off[p+1]).
 * If pat_off is NULL every pattern has `fixed_len` characters.
 * init_s/init_e: NULL for a fresh search, else the ranges being refined (the new
 * pattern is prepended, wrapper.rs:115).  out_s/out_e receive the SA range; count
 * (wrapper.rs:132-134) is out_e - out_s.  The loop breaks as soon as s == e exactly
 * as the reference does, so (s, e) are bit-identical, not just the count. */
int fmx_search_batch(const fmx_index *idx, int mode, const uint8_t *pat, const uint64_t *pat_off,
                     uint64_t fixed_len, uint64_t npat, const uint64_t *init_s,
                     const uint64_t *init_e, uint64_t *out_s, uint64_t *out_e);
int fmx_search_batch_device(const fmx_index *idx, int mode, const uint8_t *d_pat,
                            const uint64_t *d_pat_off, uint64_t fixed_len, uint64_t npat,
                            const uint64_t *d_init_s, const uint64_t *d_init_e, uint64_t *d_out_s,
                            uint64_t *d_out_e, void *stream);
/* after a *_device search on `stream` has completed: FMX_ERR_PATTERN_CHAR if any processed
 * pattern character exceeded max_character.  Synchronises the stream. */
int fmx_search_check(const fmx_index *idx, void *stream);

/* ---------------------------------------------------------------- locate
 * Batched Search::iter_matches + MatchWithLocate::locate (+ MatchWithPieceId::piece_id)
 * -- src/wrapper.rs:137-139, 203-217, 238-248; get_sa: fm_index.rs:127-140,
 * rlfmi.rs:176-189, multi_pieces.rs:188-201; piece_id: multi_pieces.rs:208-218.
 * For each range [s[p], e[p]) the matches are emitted in ascending SA-row order (the
 * reference's iteration order); with prefix_only != 0 only rows whose L is 0 are kept
 * (wrapper.rs:208).  hit_off (npat+1 entries, caller-owned) receives the exclusive
 * prefix sum of hit counts; *positions (and *piece_ids if non-NULL; MultiPieces only)
 * receive hit_off[npat] values each and are released with fmx_free(). */
int fmx_locate_batch(const fmx_index *idx, int prefix_only, const uint64_t *s, const uint64_t *e,
                     uint64_t npat, uint64_t *hit_off, uint64_t **positions, uint64_t **piece_ids);
/* Fused search + locate for host buffers: the batched form of
 *     index.search(p).iter_matches().map(|m| m.locate())           (README.md:49-64)
 * in ONE call.  The batch is cut into chunks that flow through a four-lane pipeline (H2D copy of
 * chunk k+1 overlaps the kernels of chunk k and the D2H copy of chunk k-1); SA ranges never make
 * the round trip to the host in between.  out_s / out_e (nullable) receive the SA ranges,
 * hit_off (npat+1) the exclusive prefix sum of hit counts, positions / piece_ids (nullable,
 * caller-owned, `capacity` entries each) the matches in the reference's iteration order.
 * If the batch has more hits than `capacity`, returns FMX_ERR_CAPACITY with hit_off and
 * *total_hits filled so the caller can size the buffers and use fmx_locate_batch.
 * Pinned (page-locked) host buffers make the copies asynchronous; pageable ones work too. */
int fmx_search_locate_batch(const fmx_index *idx, int mode, const uint8_t *pat, const uint64_t *pat_off,
                            uint64_t fixed_len, uint64_t npat, uint64_t *out_s, uint64_t *out_e,
                            uint64_t *hit_off, uint64_t *positions, uint64_t *piece_ids,
                            uint64_t capacity, uint64_t *total_hits);
/* two-phase device form: count (synchronous, returns the total), then fill (async). */
int fmx_locate_count_device(const fmx_index *idx, int prefix_only, const uint64_t *d_s,
                            const uint64_t *d_e, uint64_t npat, uint64_t *d_hit_off,
                            uint64_t *total_hits, void *stream);
int fmx_locate_fill_device(const fmx_index *idx, int prefix_only, const uint64_t *d_s,
                           const uint64_t *d_e, uint64_t npat, const uint64_t *d_hit_off,
                           uint64_t total_hits, uint64_t *d_positions, uint64_t *d_piece_ids,
                           void *stream);

/* Paged hit list -- the bounded-memory form of the lazy iter_matches (wrapper.rs:137-139, 203-217)
 * for match sets too large to hold at once (BASELINE config 3: ~6e8 hits).  Hits are numbered in the
 * reference's iteration order (pattern by pattern, rows ascending); a page is hits
 * [first_hit, first_hit + nhits).  Unfiltered modes only (search / search_suffix).
 * Host form: SA ranges in, one page out (fewer if the list ends), *total_hits = size of the whole list.
 * Device form: d_hit_off is what fmx_locate_count_device(prefix_only = 0) produced; asynchronous. */
int fmx_locate_page(const fmx_index *idx, const uint64_t *s, const uint64_t *e, uint64_t npat,
                    uint64_t first_hit, uint64_t nhits, uint64_t *positions, uint64_t *piece_ids,
                    uint64_t *total_hits);
int fmx_locate_page_device(const fmx_index *idx, const uint64_t *d_s, const uint64_t *d_hit_off,
                           uint64_t npat, uint64_t first_hit, uint64_t nhits, uint64_t *d_positions,
                           uint64_t *d_piece_ids, void *stream);

/* Fully asynchronous form (no host synchronisation, CUDA-graph capturable when prefix_only == 0):
 * launches are sized by `capacity` (entries of d_positions / d_piece_ids); the real number of hits
 * stays on the device in d_hit_off[npat]; if it exceeds `capacity` only the first `capacity` hits are
 * written.  Scratch must have been sized by an earlier identical call before a graph capture. */
int fmx_locate_batch_device(const fmx_index *idx, int prefix_only, const uint64_t *d_s,
                            const uint64_t *d_e, uint64_t npat, uint64_t *d_hit_off,
                            uint64_t *d_positions, uint64_t *d_piece_ids, uint64_t capacity,
                            void *stream);

/* ---------------------------------------------------------------- fused query (count + locate), any input form
 * One descriptor for the batched form of
 *     index.search(p).count()  /  index.search(p).iter_matches().map(|m| m.locate())     (README.md:49-64)
 * over byte patterns or PACKED patterns, with 64- or 32-bit outputs.
 *
 * Packed patterns (packed_bits = 2 or 4; needs max_character <= 2^packed_bits): every pattern has `fixed_len`
 * characters and occupies ceil(fixed_len * packed_bits / 64) consecutive 64-bit little-endian words; character k
 * of a pattern sits in bits [k * packed_bits, (k + 1) * packed_bits) of its words (bit 0 = least significant bit
 * of the first word) and is stored as (character - 1), so DNA coded 1..4 packs a 32-mer into ONE word -- 8 bytes
 * over PCIe instead of 32.  A packed pattern cannot hold the character 0.
 *
 * out_width = 8: hit_off / positions / piece_ids / counts are uint64_t (Rust usize); out_width = 4: uint32_t
 * (every index has n < 2^32; hit_off then needs the batch to have fewer than 2^32 hits, else FMX_ERR_CAPACITY).
 * out_s / out_e (nullable, always uint64_t): the exact SA ranges.  Leaving them NULL lets the HBM-rich path skip
 * the inverse-suffix-array request that only serves to name the final row.
 * counts (nullable): e - s per pattern (Search::count, wrapper.rs:132-134).  hit_off (nullable, npat + 1): the
 * exclusive prefix sum of the hit counts.  positions / piece_ids (nullable, `capacity` entries): the matches in
 * the reference's iteration order.  Unfiltered modes need nothing else; FMX_SEARCH_PREFIX / _EXACT apply the
 * L == 0 filter to hit_off / positions (counts stay unfiltered, as in the reference). */
typedef struct fmx_query {
    int mode;                /* fmx_mode */
    uint32_t packed_bits;    /* 0: one byte per character; 2 or 4: packed (see above) */
    const void *patterns;    /* bytes (pattern p = [pat_off[p], pat_off[p+1]) or p * fixed_len ..) or packed words */
    const uint64_t *pat_off; /* NULL: every pattern has fixed_len characters (required for packed input) */
    uint64_t fixed_len;
    uint64_t npat;
    uint32_t out_width;      /* 8 or 4 */
    uint32_t reserved;
    uint64_t *out_s, *out_e; /* nullable */
    void *counts;            /* nullable */
    void *hit_off;           /* nullable unless positions / piece_ids are given */
    void *positions;         /* nullable */
    void *piece_ids;         /* nullable */
    uint64_t capacity;
} fmx_query;
/* Device buffers, asynchronous on `stream` (NULL = the index's stream), CUDA-graph capturable for the unfiltered
 * modes once scratch has been sized by an identical earlier call; the hit total stays in hit_off[npat]. */
int fmx_query_batch_device(const fmx_index *idx, const fmx_query *q, void *stream);
/* Host buffers (pinned ones make the copies asynchronous): chunked H2D / kernels / D2H pipeline.  *total_hits
 * (nullable) receives hit_off[npat].  More hits than `capacity`: FMX_ERR_CAPACITY with counts / hit_off filled. */
int fmx_query_batch(const fmx_index *idx, const fmx_query *q, uint64_t *total_hits);

/* ---------------------------------------------------------------- multi-GPU
 * Queries are independent, so FMIndex / RLFMIndex shard by query with the index REPLICATED per GPU and no exchange
 * (SURVEY.md 8e).  FMIndexMultiPieces may instead be PARTITIONED BY PIECE: every GPU indexes a contiguous group of
 * pieces, every pattern is answered by every GPU, and the per-pattern hit counts and hit lists are combined
 * (multi_pieces.rs:188-223: a pattern without \0 cannot span two pieces, so counts add and the match sets are
 * disjoint).  Match sets, counts, positions and piece ids equal the single index's; the reference's iteration ORDER
 * is reproduced only by the replicated form (hits of a pattern come partition-major here).
 *
 * fmx_csr_merge_device: the combine step on ONE device, for callers that move the parts themselves (one process
 * per GPU with NCCL: fm-index_b200/partitioned.py).  Part r is a CSR in its own coordinates -- hit_off (npat + 1),
 * positions, piece_ids (nullable), all `width` bytes per entry, resident on `device` -- plus the text offset and the
 * first piece of its partition.  Output is a uint64 CSR of at most `capacity` hits; hit_off[npat] is the total. */
typedef struct fmx_csr_part {
    const void *hit_off;
    const void *positions;
    const void *piece_ids;
    uint64_t position_base;
    uint64_t piece_base;
} fmx_csr_part;
int fmx_csr_merge_device(int device, const fmx_csr_part *parts, int nparts, uint64_t npat, uint32_t width,
                         uint64_t *d_hit_off, uint64_t *d_positions, uint64_t *d_piece_ids, uint64_t capacity,
                         void *stream);

/* A group of GPUs of one process behind one handle.
 *   FMX_GROUP_REPLICATE  the index is built once and uploaded to every device; a batch is cut into contiguous
 *                        shards, one per device (fmx_query_batch on each, in parallel); results come back in input
 *                        order, bit-identical to the single index -- iteration order included.
 *   FMX_GROUP_BY_PIECE   (FMX_KIND_MULTI) the pieces are split into contiguous, length-balanced groups, one index
 *                        per device; every device answers the whole batch; the parts are gathered on the first
 *                        device over NVLink peer copies and merged there by fmx_csr_merge_device.  Patterns
 *                        containing \0 are rejected (they could span the partition).
 * fmx_group_query_batch takes the same descriptor as fmx_query_batch (host buffers; out_width 8; out_s / out_e
 * are not available BY_PIECE: SA rows of different partitions are unrelated). */
typedef struct fmx_group fmx_group;
typedef enum fmx_group_mode { FMX_GROUP_REPLICATE = 0, FMX_GROUP_BY_PIECE = 1 } fmx_group_mode;
int fmx_group_create(const int *device_ids, int ndev, int group_mode, const void *text, uint64_t n, uint32_t char_width,
                     uint64_t max_character, int kind, int level, int index_mode, fmx_group **out);
void fmx_group_free(fmx_group *g);
int fmx_group_size(const fmx_group *g);
const fmx_index *fmx_group_index(const fmx_group *g, int k);
uint64_t fmx_group_len(const fmx_group *g);           /* SearchIndex::len of the whole text */
uint64_t fmx_group_pieces_count(const fmx_group *g);  /* pieces of the whole text */
int fmx_group_query_batch(const fmx_group *g, const fmx_query *q, uint64_t *total_hits);

/* ---------------------------------------------------------------- extraction
 * Batched Match::iter_chars_backward / iter_chars_forward taken k characters deep
 * -- src/wrapper.rs:143-183, 229-235.  rows are SA rows (the `i` of a Match).
 * out is nrows*k characters of fmx_index_char_width(idx) bytes each (bytes for u8 indexes), row-major;
 * out_len[r] is the number of characters produced
 * (always k backward; forward stops early on MultiPieces when fl_map is None,
 * multi_pieces.rs:171-181).  Unproduced bytes are 0. */
int fmx_extract_batch(const fmx_index *idx, const uint64_t *rows, uint64_t nrows, uint32_t k,
                      int forward, uint8_t *out, uint32_t *out_len);
int fmx_extract_batch_device(const fmx_index *idx, const uint64_t *d_rows, uint64_t nrows,
                             uint32_t k, int forward, uint8_t *d_out, uint32_t *d_out_len,
                             void *stream);

/* Backend primitives over a batch of rows (crate-private seam src/backend.rs:5-40),
 * exposed for parity tests: op 0 = get_l, 1 = lf_map, 2 = get_f, 3 = fl_map (UINT64_MAX = None),
 * 4 = get_sa, 5 = piece_id via the reference's literal LF walk. */
int fmx_rows_op(const fmx_index *idx, int op, const uint64_t *rows, uint64_t nrows, uint64_t *out);
/* lf_map2(c[k], i[k]) for a batch (backend.rs:16); c holds nrows characters of fmx_index_char_width(idx) bytes each. */
int fmx_lf_map2_batch(const fmx_index *idx, const uint8_t *c, const uint64_t *i, uint64_t nrows,
                      uint64_t *out);

/* ---------------------------------------------------------------- measurement helpers */
/* Work counters of the last search / locate-fill launch (executed backward-search
 * iterations, executed LF steps): the roofline numerator.  Synchronises the stream. */
int fmx_last_work(const fmx_index *idx, void *stream, uint64_t *search_steps, uint64_t *lf_steps);
/* Index requests issued by the phased kernels of the last query (option "count_work"): lane-level loads of table
 * entries, rank blocks, suffix-array / inverse entries and text sectors by the search, and of suffix-array entries
 * by the emit kernels.  Synchronises the stream. */
int fmx_last_requests(const fmx_index *idx, void *stream, uint64_t *search_requests, uint64_t *emit_requests);
/* Milliseconds of the phases of the last fmx_query_batch_device call made with option "phase_timing" = 1
 * (CUDA events between the library's own launches): [0] seed / k_search, [1] steps, [2] verify, [3] second steps
 * pass, [4] counts + offsets scan, [5] emit / locate.  Synchronises the stream. */
int fmx_last_phase_ms(const fmx_index *idx, void *stream, float *ms, int n);
/* Peak random gather rate microbenchmark over `bytes` of device memory (independent random
 * loads of load_bytes = 32 (one sector), 64 or 128 bytes each; no dependent chain): the
 * random-access roofline denominator.  Returns 32-byte sectors per second. */
int fmx_random_gather_bench(int device, uint64_t bytes, uint64_t nloads, int iters,
                            uint32_t load_bytes, double *sectors_per_s);
/* number of kernel launches issued by this library since load */
uint64_t fmx_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* FMX_H */
