// fmx_layout.h -- the device layout ("blob") shared by the host builder and the kernels.
//
// One contiguous allocation in HBM: BlobHeader followed by 256-byte aligned sections.
//
// RANK BLOCKS ("RB192").  Every rank-able bit vector (each wavelet-matrix level, and the
// RLFM run-start vectors b / bp) is an array of 32-byte blocks
//       word 0    : u32  number of 1-bits in all preceding blocks of this vector
//       word 1    : byte 1 = popcount(P0), byte 2 = popcount(P0) + popcount(P1)  (bytes 0, 3 = 0)
//       word 2..7 : three 64-bit payload words P0, P1, P2 (192 payload bits, LSB first)
// so rank1(pos) -- and the bit at pos, for access -- costs exactly ONE 32-byte HBM sector
// (one 256-bit load) plus ONE masked 64-bit popcount: the in-block sub-counts replace the
// up-to-seven masked popcounts a flat payload needs (the first kernels were ALU-bound on
// exactly that when the index fits L2).  The reference's dependency (vers-vecs RsVec) keeps
// bits, 512-bit block counts and 8192-bit super-block counts in three separate arrays, i.e.
// up to three cache lines per rank.  u32 counts bound the text length to n < 2^32 (every
// BASELINE config; the largest is 3*10^9).
//
// WAVELET MATRIX.  L = Text::max_bits() levels (text.rs:61-63), level 0 = most significant
// bit, zeros stably partitioned before ones -- the structure of vers' WaveletMatrix, which
// the reference stores its BWT in (fm_index.rs:44-58).  rank(i, c) walks ONE position down
// the levels (the walk of position 0 is a per-symbol constant, folded into `adj`):
//       lf_map2(c, i) = adj[c] + walk_c(i),   adj[c] = cs[c] - walk_c(0)      (mod 2^32)
//
// QUATERNARY LEVEL ("Q4").  When max_character <= 4 (DNA coded 0..4: every DNA config of
// BASELINE.json) and the sequence holds at most FMX_MAX_EXC zeros, the L = 3 binary levels
// collapse into ONE level of arity 4: 32-byte blocks
//       word 0..3 : u32 occurrences of the 2-bit codes 0..3 in all preceding blocks
//       word 4..7 : 64 two-bit codes (code = symbol - 1) as two 64-bit PLANES: words 4,5 hold the low
//                   bit of codes 0..63, words 6,7 the high bit.  "positions holding code c" is then
//                   (p0 ^ a) & (p1 ^ b) with a, b in {0, ~0}: one masked 64-bit popcount per rank
//                   (interleaved 2-bit fields needed four; k_search was issue-bound on them when the
//                   index fits L2)
// so lf_map2(c, i) = cs[c] + rank(i, c) costs ONE sector instead of three.
//
// QUATERNARY WAVELET MATRIX ("WM4", the default for every other alphabet).  The same 32-byte
// blocks, one level per TWO bits of the symbol (ceil(L/2) levels, most significant digit first),
// symbols stably partitioned by digit into four groups starting at qoff[l][0..3].  A walk is
//       pos <- qoff[l][d] + rank_d(level l, pos),   d = digit l of c
// i.e. HALF the sectors of the binary matrix (4 instead of 8 for bytes).  adj[] as above.
//
// PER-SYMBOL BIT VECTORS ("SYM", the default for every non-DNA alphabet while it fits the HBM
// budget).  One RB192 bit vector PER SYMBOL (bit i of vector c = [seq[i] == c]), concatenated in
// SEC_LEVEL0 (vector c starts at block c * sym_nblk), plus the raw sequence bytes in SEC_LEVEL0+1:
//       lf_map2(c, i) = cs[c] + rank1(vec_c, i)          ONE sector, like Q4
//       access(i)     = raw[i]                           ONE request
// Space is cs_len * n / 6 bytes (43 GB for a 1 GB byte text): a deliberate trade of the B200's
// 180 GB of HBM for requests, because the random-request rate -- not bandwidth, not capacity -- is
// what bounds this workload (profiles/r01b_random_access_study.md: ~45 G DRAM-missing requests/s
// whatever their size).  Above the budget (FMX_SYM_BUDGET_MB, default 49152) the builder falls back to WM4.
//
// SEED-AND-VERIFY TAIL (FM and MultiPieces kinds; SEC_TEXT, SEC_ISA, SEC_VSA).  Once backward search has narrowed the range
// to ONE row, the remaining characters can only match one place: the row is located, the rest of the
// pattern is compared with the text directly, and the row the reference's loop would end in is
// ISA[pos - matched].  DENSE form (default while 9 n bytes fit FMX_VERIFY_BUDGET_MB, 32 GiB): the full
// suffix array and its inverse, so the tail costs three or four memory requests -- SA[s], one or two text
// lines, ISA[q] -- instead of one per remaining character, with no data-dependent loop.  SAMPLED form
// (larger texts with the SYM layout): the caller's suffix-array samples (level <= 3) and an ISA sampled
// every 4 positions, each followed by a short LF walk; it pays only where an LF step is expensive.
// The tail consumes exactly the characters that match; the ordinary loop then takes the next one, so a
// mismatch empties the range -- and an invalid character raises the error -- exactly as in the reference,
// with the same (s, e) and step count.  MultiPieces: the comparison stops in front of a \0 of the text.  These are
// acceleration structures of the SEARCH, like the k-mer tables: locate still walks to the samples of the
// level the caller asked for.
//
// WIDE ALPHABETS ("WIDE": max_character > 255, i.e. texts of u16 / u32 / u64 characters, character.rs:38-42).
// The binary wavelet matrix with L = floor(log2 max_character) + 1 <= 32 levels, all levels in SEC_LEVEL0 one after the
// other (level l starts at block l * wide_nblk), zeros per level in SEC_WZEROS, and cs / adj -- max_character + 1 words
// each, as in the reference (sais.rs:9-19) -- read from global memory instead of being staged in shared memory.
// adj[c] is only filled for symbols that occur; for the others rank(., c) = 0 and the kernels return cs[c] without
// touching the matrix (cs[c + 1] == cs[c] marks them).  Texts of wide characters whose max_character is <= 255 take
// the u8 layouts above; only the pattern / extraction character width (header char_width) differs.
//
// Q4 DETAILS.  The rare symbol 0
// (the \0 terminators: 1 for a single text, one per piece for MultiPieces) is stored as code 0
// and its positions are listed in SEC_EXC (sorted; staged in shared memory), which corrects
// rank(., 1) and answers rank(., 0) / access exactly.  Results are identical to the wavelet
// matrix; only the number of sectors per probe changes (profiles/: search is bound by the
// count of lane-sector requests in L1TEX when the index fits L2 and by the random-sector rate
// of HBM when it does not -- both scale with probes per step).
#pragma once
#include <stdint.h>
#include <vector_types.h>  // uint4 (CUDA toolkit header, host-safe)

#define FMX_BLOB_MAGIC 0x3030324258584d46ull /* "FMXXB200" little endian-ish tag */
#define FMX_BLOB_VERSION 7u
#define FMX_MAX_LEVELS 8
#define FMX_RB_BITS 192u
#define FMX_MAX_EXC 1024u   /* Q4 layout: at most this many \0 symbols in the sequence */
#define FMX_LAYOUT_WAVELET 0u
#define FMX_LAYOUT_QUAT 1u
#define FMX_LAYOUT_WM4 2u
#define FMX_LAYOUT_SYM 3u
#define FMX_LAYOUT_WIDE 4u
#define FMX_MAX_WIDE_LEVELS 32
#define FMX_MAX_QLEVELS 4
#define FMX_SECTION_ALIGN 256u

enum FmxSection : uint32_t {
    SEC_LEVEL0 = 0,  // .. SEC_LEVEL0 + 7 : wavelet levels (RB192); Q4: SEC_LEVEL0 = the Q4 blocks; SYM: vectors, +1 = raw bytes
    SEC_ADJ = 8,     // u32[cs_len]   adj[c] = cs[c] - walk_c(0)
    SEC_CS = 9,      // u32[cs_len+1] cs[c] (sais.rs:21-32), cs[cs_len] = n (FM/MULTI) or runs (RLFM)
    SEC_SA = 10,     // u32[((n-1)>>level)+1]  sampled suffix array, sa[i << level]  (sample.rs:33-37)
    SEC_DOC = 11,    // u32[ndoc]     multi_pieces.rs:53-79
    SEC_PIECE_END = 12,  // u32[ndoc]  text position of the k-th \0 (piece k is [end[k-1]+1, end[k]])
    SEC_RL_B = 13,   // RB192 over n bits: run starts in L order  (rlfmi.rs:40-67)
    SEC_RL_BP = 14,  // RB192 over n bits: run starts in F order  (rlfmi.rs:70-83)
    SEC_RL_BSEL = 15,   // u32[runs+1]  select1(b, j), [runs] = n
    SEC_RL_BPSEL = 16,  // u32[runs+1]  select1(bp, j), [runs] = n
    SEC_EXC = 17,       // u32[nexc]  Q4 layout: sorted positions whose symbol is 0
    SEC_TEXT = 18,      // u8[n]      the text itself (verify path, FM kind)
    SEC_ISA = 19,       // u32[ceil(n / 2^isa_level)]  row of the suffix starting at text position k << isa_level
    SEC_VSA = 20,       // u32[n]     the FULL suffix array, for the verify path only (locate keeps the caller's level)
    SEC_WZEROS = 21,    // u32[levels]  WIDE layout: zeros per wavelet level
    SEC_COUNT = 22
};

struct FmxSectionEntry {
    uint64_t offset;  // from the start of the blob
    uint64_t bytes;
};

struct FmxBlobHeader {
    uint64_t magic;
    uint32_t version;
    uint32_t kind;           // fmx_kind
    uint64_t n;              // text length incl. the trailing \0 (SearchIndex::len)
    uint64_t seq_len;        // length of the wavelet-matrix sequence: n (FM/MULTI) or runs (RLFM)
    uint32_t levels;         // L
    uint32_t max_character;
    uint32_t cs_len;         // max_character + 1
    uint32_t has_locate;
    uint32_t sa_level;       // effective level (0 if n <= 2^level, sample.rs:28-31)
    uint32_t sa_word_size;   // floor(log2 n) + 1: width the reference packs samples with
    uint64_t sa_count;
    uint64_t ndoc;           // pieces_count
    uint64_t first_row;      // sa_idx_first_text (multi_pieces.rs:23)
    uint64_t runs;
    uint64_t zeros[FMX_MAX_LEVELS];  // zeros per level
    uint64_t total_bytes;
    FmxSectionEntry sec[SEC_COUNT];
    uint32_t layout;  // FMX_LAYOUT_WAVELET | FMX_LAYOUT_QUAT | FMX_LAYOUT_WM4 | FMX_LAYOUT_SYM | FMX_LAYOUT_WIDE
    uint32_t nexc;    // Q4: number of zeros in the sequence
    uint32_t qlevels; // WM4: ceil(levels / 2)
    uint32_t sym_nblk; // SYM: RB192 blocks per symbol vector (seq_len / 192 + 1)
    uint32_t verify;     // 1: SEC_TEXT / SEC_ISA / SEC_SA present for the seed-and-verify tail of k_search
    uint32_t isa_level;  // ISA sampling: every 2^isa_level-th text position (0 with the dense structures)
    uint32_t vsa_level;  // sampling level of the suffix-array samples the verify path walks to (0 = SEC_VSA / dense)
    uint32_t char_width; // bytes per character of the text the index was built from (1, 2, 4, 8): width of patterns and extracted characters
    uint64_t qoff[FMX_MAX_QLEVELS][4];  // WM4: start of digit group d at level l
    uint64_t reserved[4];
};

// What the kernels see (passed by value as a __grid_constant__ parameter).
struct FmxDev {
    const uint4 *lv[FMX_MAX_LEVELS];
    uint32_t zeros[FMX_MAX_LEVELS];
    const uint32_t *adj;
    const uint32_t *cs;
    const uint32_t *sa;
    const uint32_t *doc;
    const uint32_t *piece_end;
    const uint4 *rl_b;
    const uint4 *rl_bp;
    const uint32_t *rl_bsel;
    const uint32_t *rl_bpsel;
    const uint32_t *exc;
    uint32_t n;
    uint32_t seq_len;
    uint32_t levels;
    uint32_t max_character;
    uint32_t cs_len;
    uint32_t kind;
    uint32_t has_locate;
    uint32_t sa_level;
    uint32_t ndoc;
    uint32_t first_row;
    uint32_t runs;
    uint32_t layout;
    uint32_t nexc;
    uint32_t qlevels;
    uint32_t qoff[FMX_MAX_QLEVELS * 4];
    uint32_t sym_nblk;
    const uint8_t *raw;  // SYM: the sequence itself
    const uint8_t *text; // verify path: the text
    const uint32_t *isa; // verify path: sampled inverse suffix array
    uint32_t verify;
    uint32_t isa_level;
    const uint32_t *vsa;  // verify path: suffix-array samples of level vsa_level (the full array when dense)
    uint32_t vsa_level;
    const uint32_t *wzeros;  // WIDE: zeros per level
    uint32_t wide_nblk;      // WIDE: RB192 blocks per level (seq_len / 192 + 1)
    uint32_t cw_shift;       // log2(char_width)
};
