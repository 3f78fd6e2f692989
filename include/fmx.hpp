// fmx.hpp -- C++ host-side mirror of the reference crate's public API (ajalab/fm-index 0.3.1,
// src/frontend.rs, src/text.rs) over the C ABI in fmx.h.  Header-only; link libfmx_b200.so.
//
// The reference is Rust; no Rust toolchain exists in the build image, so this facade is the compiled
// host side of the drop-in (INTEGRATION.md shows the Rust binding).  Names, argument meaning and
// error behaviour follow the crate:
//
//     Text text(bytes);                                  // Text::new           (text.rs:28-33)
//     auto t2 = Text::with_max_character(bytes, 4);      //                     (text.rs:44-49)
//     FMIndexWithLocate index(text, 2);                  // ::new(&text, level) (frontend.rs:205-218)
//     Search s = index.search("dolor");                  //                     (frontend.rs:291-296)
//     s.count();  s.search("x");                         // refine              (frontend.rs:70-83)
//     for (Match& m : s.iter_matches()) m.locate();      // ascending SA rows   (wrapper.rs:203-217)
//     m.iter_chars_backward(16); m.iter_chars_forward(20); m.piece_id();
//     index.search_batch(patterns)                       // the batched entry the GPU path exists for
//     index.query_batch(patterns)                        // counts + CSR matches from ONE fused call (fmx_query_batch)
//     IndexGroup g(text, FMX_KIND_MULTI, 2, {0, 1}, FMX_GROUP_BY_PIECE);  g.query_batch(patterns);   // several GPUs
//
// Differences forced by the language: iterators are returned as vectors (iter_chars_* take the
// number of characters, like `.take(k)`), errors are exceptions (fmx::InvalidText = Error::InvalidText;
// a pattern character above max_character throws std::out_of_range where the crate panics).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "fmx.h"

namespace fmx {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
struct InvalidText : Error {  // Error::InvalidText (src/error.rs:3-6)
    explicit InvalidText(const std::string &m) : Error(FMX_ERR_INVALID_TEXT, "invalid text: " + m) {}
};

inline void check(int rc) {
    if (rc == FMX_OK) return;
    std::string msg = fmx_last_error();
    if (rc == FMX_ERR_INVALID_TEXT) throw InvalidText(msg);
    if (rc == FMX_ERR_PATTERN_CHAR) throw std::out_of_range(msg);  // the crate panics (fm_index.rs:94)
    throw Error(rc, msg);
}

using PieceId = uint64_t;  // src/piece.rs

// src/text.rs:10-64 (u8 characters)
class Text {
  public:
    explicit Text(std::vector<uint8_t> text) : text_(std::move(text)), max_character_(255) {}
    explicit Text(const std::string &text) : text_(text.begin(), text.end()), max_character_(255) {}
    static Text with_max_character(std::vector<uint8_t> text, uint8_t max_character) {
        Text t(std::move(text));
        t.max_character_ = max_character;
        return t;
    }
    const std::vector<uint8_t> &text() const { return text_; }
    uint8_t max_character() const { return max_character_; }

  private:
    std::vector<uint8_t> text_;
    uint8_t max_character_;
};

class IndexBase;

// MatchWithLocate / MatchWithPieceId (frontend.rs:85-104)
class Match {
  public:
    uint64_t row() const { return i_; }
    uint64_t locate() const;                                      // wrapper.rs:238-242
    PieceId piece_id() const;                                     // wrapper.rs:244-248
    std::vector<uint8_t> iter_chars_forward(uint32_t k) const;    // wrapper.rs:229-231, first k items
    std::vector<uint8_t> iter_chars_backward(uint32_t k) const;   // wrapper.rs:233-235, first k items

  private:
    friend class Search;
    const IndexBase *index_ = nullptr;
    uint64_t i_ = 0, position_ = UINT64_MAX, piece_ = UINT64_MAX;
};

// Search (frontend.rs:65-83) over wrapper.rs:14-23
class Search {
  public:
    Search search(const std::string &pattern) const;   // prepends `pattern` (wrapper.rs:99-124)
    Search search(const std::vector<uint8_t> &pattern) const;
    uint64_t count() const { return e_ - s_; }        // wrapper.rs:132-134 (ignores the prefix filter)
    std::pair<uint64_t, uint64_t> get_range() const { return {s_, e_}; }
    std::vector<Match> iter_matches() const;           // wrapper.rs:137-139, 203-217

  private:
    friend class IndexBase;
    const IndexBase *index_ = nullptr;
    int mode_ = FMX_SEARCH;
    uint64_t s_ = 0, e_ = 0;
    bool fresh_ = true;
};

// result of a batched search: SA ranges + (on demand) CSR hit lists
struct SearchBatch {
    std::vector<uint64_t> s, e;
    std::vector<uint64_t> hit_off, positions, piece_ids;
    uint64_t count(size_t p) const { return e[p] - s[p]; }
};

// result of the fused batched query: Search::count per pattern + the matches as CSR (reference iteration order)
struct QueryBatch {
    std::vector<uint64_t> s, e;                           // only with with_ranges
    std::vector<uint64_t> counts, hit_off, positions, piece_ids;
    uint64_t total_hits = 0;
};

// fills a fmx_query over byte patterns with 64-bit outputs and runs `call`; a capacity guess that is too small is
// retried once with the exact size the first call reported (FMX_ERR_CAPACITY leaves counts / hit_off filled)
// `offsets`: hit_off is wanted even without positions (a piece-partitioned group derives its counts from it)
template <class Call>
inline QueryBatch run_query_batch(Call call, const std::vector<std::string> &patterns, int mode, bool locate, bool piece_ids,
                                  bool with_ranges, bool offsets = false) {
    std::vector<uint8_t> flat;
    std::vector<uint64_t> off{0};
    for (auto &p : patterns) {
        flat.insert(flat.end(), p.begin(), p.end());
        off.push_back(flat.size());
    }
    const size_t n = patterns.size();
    QueryBatch b;
    b.counts.resize(n);
    if (with_ranges) {
        b.s.resize(n);
        b.e.resize(n);
    }
    if (locate || offsets) b.hit_off.resize(n + 1);
    uint64_t cap = locate ? (2 * n > 1024 ? 2 * n : 1024) : 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        if (locate) b.positions.resize(cap);
        if (piece_ids) b.piece_ids.resize(cap);
        fmx_query q{};
        q.mode = mode;
        q.patterns = flat.data();
        q.pat_off = off.data();
        q.npat = n;
        q.out_width = 8;
        q.out_s = with_ranges ? b.s.data() : nullptr;
        q.out_e = with_ranges ? b.e.data() : nullptr;
        q.counts = b.counts.data();
        q.hit_off = (locate || offsets) ? b.hit_off.data() : nullptr;
        q.positions = locate ? b.positions.data() : nullptr;
        q.piece_ids = piece_ids ? b.piece_ids.data() : nullptr;
        q.capacity = cap;
        uint64_t total = 0;
        const int rc = call(&q, &total);
        b.total_hits = total;
        if (rc == FMX_ERR_CAPACITY && attempt == 0 && locate) {
            cap = total;
            continue;
        }
        check(rc);
        break;
    }
    if (locate) b.positions.resize(b.total_hits);
    if (piece_ids) b.piece_ids.resize(b.total_hits);
    return b;
}

class IndexBase {
  public:
    IndexBase(const IndexBase &) = delete;
    IndexBase &operator=(const IndexBase &) = delete;
    ~IndexBase() { fmx_index_free(h_); }
    uint64_t len() const { return fmx_index_len(h_); }                  // frontend.rs:35-39
    uint64_t heap_size() const { return fmx_index_device_bytes(h_); }   // frontend.rs:41-44 (device bytes)
    Search search(const std::string &p) const { return start(FMX_SEARCH).search(p); }
    Search search(const std::vector<uint8_t> &p) const { return start(FMX_SEARCH).search(p); }
    // batched: many patterns, one fused launch sequence; with_locate fills the CSR hit lists
    SearchBatch search_batch(const std::vector<std::string> &patterns, bool with_locate = false, int mode = FMX_SEARCH) const {
        std::vector<uint8_t> flat;
        std::vector<uint64_t> off{0};
        for (auto &p : patterns) {
            flat.insert(flat.end(), p.begin(), p.end());
            off.push_back(flat.size());
        }
        SearchBatch b;
        size_t n = patterns.size();
        b.s.resize(n);
        b.e.resize(n);
        check(fmx_search_batch(h_, mode, flat.data(), off.data(), 0, n, nullptr, nullptr, b.s.data(), b.e.data()));
        if (with_locate) {
            b.hit_off.resize(n + 1);
            uint64_t *pos = nullptr, *pid = nullptr;
            bool multi = fmx_index_kind(h_) == FMX_KIND_MULTI;
            int po = mode == FMX_SEARCH_PREFIX || mode == FMX_SEARCH_EXACT;
            check(fmx_locate_batch(h_, po, b.s.data(), b.e.data(), n, b.hit_off.data(), &pos, multi ? &pid : nullptr));
            uint64_t total = b.hit_off[n];
            if (pos) b.positions.assign(pos, pos + total);
            if (pid) b.piece_ids.assign(pid, pid + total);
            fmx_free(pos);
            fmx_free(pid);
        }
        return b;
    }
    // the same batch through ONE fused call (fmx_query_batch: counts, CSR hit offsets, positions, piece ids; no SA
    // ranges unless asked for, which lets an HBM-rich index skip the inverse-suffix-array request)
    QueryBatch query_batch(const std::vector<std::string> &patterns, int mode = FMX_SEARCH, bool with_ranges = false) const {
        const fmx_index *h = h_;
        return run_query_batch([h](const fmx_query *q, uint64_t *total) { return fmx_query_batch(h, q, total); }, patterns, mode,
                               has_locate(), is_multi() && has_locate(), with_ranges);
    }
    const fmx_index *handle() const { return h_; }
    bool has_locate() const { return fmx_index_has_locate(h_) != 0; }
    bool is_multi() const { return fmx_index_kind(h_) == FMX_KIND_MULTI; }

  protected:
    IndexBase(const Text &text, int kind, int level, int device) {
        check(fmx_index_build(text.text().data(), text.text().size(), 1, text.max_character(), kind, level, device, &h_));
    }
    Search start(int mode) const {
        Search s;
        s.index_ = this;
        s.mode_ = mode;
        return s;
    }
    fmx_index *h_ = nullptr;
    friend class Search;
    friend class Match;
};

// the six index types (frontend.rs:110-193, 195-267)
struct FMIndex : IndexBase {
    explicit FMIndex(const Text &t, int device = 0) : IndexBase(t, FMX_KIND_FM, FMX_LEVEL_COUNT_ONLY, device) {}
};
struct FMIndexWithLocate : IndexBase {
    FMIndexWithLocate(const Text &t, size_t level, int device = 0) : IndexBase(t, FMX_KIND_FM, (int)level, device) {}
};
struct RLFMIndex : IndexBase {
    explicit RLFMIndex(const Text &t, int device = 0) : IndexBase(t, FMX_KIND_RLFM, FMX_LEVEL_COUNT_ONLY, device) {}
};
struct RLFMIndexWithLocate : IndexBase {
    RLFMIndexWithLocate(const Text &t, size_t level, int device = 0) : IndexBase(t, FMX_KIND_RLFM, (int)level, device) {}
};
// SearchIndexWithMultiPieces (frontend.rs:47-63)
struct MultiPiecesBase : IndexBase {
    using IndexBase::IndexBase;
    Search search_prefix(const std::string &p) const { return start(FMX_SEARCH_PREFIX).search(p); }
    Search search_suffix(const std::string &p) const { return start(FMX_SEARCH_SUFFIX).search(p); }
    Search search_exact(const std::string &p) const { return start(FMX_SEARCH_EXACT).search(p); }
    uint64_t pieces_count() const { return fmx_index_pieces_count(h_); }

  protected:
    MultiPiecesBase(const Text &t, int level, int device) : IndexBase(t, FMX_KIND_MULTI, level, device) {}
};
struct FMIndexMultiPieces : MultiPiecesBase {
    explicit FMIndexMultiPieces(const Text &t, int device = 0) : MultiPiecesBase(t, FMX_LEVEL_COUNT_ONLY, device) {}
};
struct FMIndexMultiPiecesWithLocate : MultiPiecesBase {
    FMIndexMultiPiecesWithLocate(const Text &t, size_t level, int device = 0) : MultiPiecesBase(t, (int)level, device) {}
};

// fmx_group: several GPUs of this process behind one handle (fmx.h, "multi-GPU").
//   FMX_GROUP_REPLICATE  the index copied to every device, batches cut into shards; input order and the reference's
//                        iteration order preserved
//   FMX_GROUP_BY_PIECE   FMIndexMultiPieces: pieces partitioned over the devices (device ids may repeat), every member
//                        answers the whole batch, parts merged on the first device; match SETS, counts, positions and
//                        piece ids of the whole text (multi_pieces.rs:188-223), hits partition-major.  This is also how a
//                        MultiPieces text of 2^32 symbols or more is served: every partition stays below 2^32, positions
//                        are u64 in the coordinates of the whole text.
class IndexGroup {
  public:
    IndexGroup(const Text &text, int kind, int level, const std::vector<int> &devices, int group_mode = FMX_GROUP_REPLICATE,
               int index_mode = FMX_MODE_AUTO)
        : kind_(kind), level_(level) {
        check(fmx_group_create(devices.data(), (int)devices.size(), group_mode, text.text().data(), text.text().size(), 1,
                               text.max_character(), kind, level, index_mode, &g_));
    }
    IndexGroup(const IndexGroup &) = delete;
    IndexGroup &operator=(const IndexGroup &) = delete;
    ~IndexGroup() { fmx_group_free(g_); }
    int size() const { return fmx_group_size(g_); }
    uint64_t len() const { return fmx_group_len(g_); }
    uint64_t pieces_count() const { return fmx_group_pieces_count(g_); }
    QueryBatch query_batch(const std::vector<std::string> &patterns, int mode = FMX_SEARCH) const {
        const fmx_group *g = g_;
        const bool locate = level_ >= 0;
        return run_query_batch([g](const fmx_query *q, uint64_t *total) { return fmx_group_query_batch(g, q, total); }, patterns, mode,
                               locate, locate && kind_ == FMX_KIND_MULTI, false, /*offsets=*/true);
    }

  private:
    fmx_group *g_ = nullptr;
    int kind_, level_;
};

// ---- Search
inline Search Search::search(const std::vector<uint8_t> &pattern) const {
    Search r = *this;
    uint64_t off[2] = {0, pattern.size()};
    const uint64_t *is = fresh_ ? nullptr : &s_, *ie = fresh_ ? nullptr : &e_;
    check(fmx_search_batch(index_->h_, mode_, pattern.data(), off, 0, 1, is, ie, &r.s_, &r.e_));
    r.fresh_ = false;
    return r;
}
inline Search Search::search(const std::string &pattern) const {
    return search(std::vector<uint8_t>(pattern.begin(), pattern.end()));
}
inline std::vector<Match> Search::iter_matches() const {
    std::vector<Match> out;
    if (e_ <= s_) return out;
    const bool prefix_only = mode_ == FMX_SEARCH_PREFIX || mode_ == FMX_SEARCH_EXACT;
    std::vector<uint64_t> rows;
    for (uint64_t i = s_; i < e_; i++) rows.push_back(i);
    if (prefix_only) {  // wrapper.rs:208: keep rows whose L is \0
        std::vector<uint64_t> l(rows.size());
        check(fmx_rows_op(index_->h_, 0, rows.data(), rows.size(), l.data()));
        std::vector<uint64_t> kept;
        for (size_t k = 0; k < rows.size(); k++)
            if (l[k] == 0) kept.push_back(rows[k]);
        rows.swap(kept);
    }
    std::vector<uint64_t> pos, pid;
    if (index_->has_locate() && !rows.empty()) {
        uint64_t hit_off[2];
        uint64_t *p = nullptr, *d = nullptr;
        check(fmx_locate_batch(index_->h_, prefix_only, &s_, &e_, 1, hit_off, &p, index_->is_multi() ? &d : nullptr));
        if (p) pos.assign(p, p + hit_off[1]);
        if (d) pid.assign(d, d + hit_off[1]);
        fmx_free(p);
        fmx_free(d);
    }
    for (size_t k = 0; k < rows.size(); k++) {
        Match m;
        m.index_ = index_;
        m.i_ = rows[k];
        if (k < pos.size()) m.position_ = pos[k];
        if (k < pid.size()) m.piece_ = pid[k];
        out.push_back(m);
    }
    return out;
}

// ---- Match
inline uint64_t Match::locate() const {
    if (position_ != UINT64_MAX) return position_;
    uint64_t v = 0;
    check(fmx_rows_op(index_->h_, 4, &i_, 1, &v));
    return v;
}
inline PieceId Match::piece_id() const {
    if (piece_ != UINT64_MAX) return piece_;
    uint64_t v = 0;
    check(fmx_rows_op(index_->h_, 5, &i_, 1, &v));
    return v;
}
inline std::vector<uint8_t> Match::iter_chars_forward(uint32_t k) const {
    std::vector<uint8_t> out(k);
    uint32_t got = 0;
    check(fmx_extract_batch(index_->h_, &i_, 1, k, 1, out.data(), &got));
    out.resize(got);
    return out;
}
inline std::vector<uint8_t> Match::iter_chars_backward(uint32_t k) const {
    std::vector<uint8_t> out(k);
    uint32_t got = 0;
    check(fmx_extract_batch(index_->h_, &i_, 1, k, 0, out.data(), &got));
    out.resize(got);
    return out;
}

}  // namespace fmx
