// Compile-only check of include/fmx.hpp (tests/test_cpp_header.py, no GPU): every facade entry a maintainer would call,
// including the fused batched query and the multi-GPU group, must type-check against the C ABI in include/fmx.h.
#include "fmx.hpp"

uint64_t surface(const fmx::Text &t) {
    fmx::FMIndexMultiPiecesWithLocate ix(t, 2);
    fmx::Search s = ix.search("ab").search("c");
    uint64_t acc = s.count() + ix.len() + ix.heap_size() + ix.pieces_count();
    for (const fmx::Match &m : ix.search_prefix("a").iter_matches()) acc += m.locate() + m.piece_id() + m.iter_chars_forward(4).size();
    fmx::SearchBatch b = ix.search_batch({"ab", "c"}, true, FMX_SEARCH_SUFFIX);
    fmx::QueryBatch q = ix.query_batch({"ab", "c"}, FMX_SEARCH, true);
    acc += b.count(0) + q.total_hits + q.counts.size();
    fmx::IndexGroup by_piece(t, FMX_KIND_MULTI, 2, {0, 0}, FMX_GROUP_BY_PIECE, FMX_MODE_COMPACT);
    fmx::IndexGroup replicated(t, FMX_KIND_FM, 2, {0}, FMX_GROUP_REPLICATE);
    acc += by_piece.query_batch({"ab"}).total_hits + replicated.query_batch({"ab"}, FMX_SEARCH).total_hits;
    return acc + by_piece.size() + by_piece.len() + by_piece.pieces_count();
}
