// sais.hpp -- host-side suffix array construction for the index builder.
//
// Produces exactly what sais::build_suffix_array returns (reference src/suffix_array/sais.rs:115-144):
// the suffixes of the text in plain lexicographic order with \0 an ordinary, smallest symbol.
// Algorithm: SA-IS (Nong, Zhang, Chan 2010), written from the paper for templated index width
// (int32 when n < 2^31, halving the memory traffic of the induced-sorting passes).
// Construction is not the hot path; it feeds the device layout.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace fmx {

struct TypeBits {
    std::vector<uint64_t> w;
    explicit TypeBits(size_t n) : w((n + 63) / 64, 0) {}
    inline bool get(size_t i) const { return (w[i >> 6] >> (i & 63)) & 1; }
    inline void set(size_t i) { w[i >> 6] |= 1ull << (i & 63); }
    inline bool lms(size_t i) const { return i > 0 && get(i) && !get(i - 1); }
};

// text whose last symbol is already a unique smallest sentinel
template <class T>
struct PlainAcc {
    const T *t;
    inline int64_t operator()(int64_t i) const { return (int64_t)t[i]; }
};
// arbitrary byte text + a virtual sentinel at index n-1 (symbols shifted by one)
struct SentinelAcc {
    const uint8_t *t;
    int64_t n;
    inline int64_t operator()(int64_t i) const { return i == n - 1 ? 0 : (int64_t)t[i] + 1; }
};

// the same over 32-bit symbols (ranks of wide characters)
struct SentinelAcc32 {
    const uint32_t *t;
    int64_t n;
    inline int64_t operator()(int64_t i) const { return i == n - 1 ? 0 : (int64_t)t[i] + 1; }
};

template <class Acc, class I>
static void sais_buckets(const Acc &s, I n, std::vector<I> &cnt, std::vector<I> &bkt, bool end) {
    (void)s;
    (void)n;
    I sum = 0;
    for (size_t c = 0; c < cnt.size(); c++) {
        sum += cnt[c];
        bkt[c] = end ? sum : sum - cnt[c];
    }
}

template <class Acc, class I>
static void sais_induce(const Acc &s, const TypeBits &t, I *SA, I n, std::vector<I> &cnt,
                        std::vector<I> &bkt) {
    sais_buckets(s, n, cnt, bkt, false);
    for (I i = 0; i < n; i++) {
        I j = SA[i];
        if (j > 0 && !t.get((size_t)j - 1)) SA[bkt[s(j - 1)]++] = j - 1;
    }
    sais_buckets(s, n, cnt, bkt, true);
    for (I i = n - 1; i >= 0; i--) {
        I j = SA[i];
        if (j > 0 && t.get((size_t)j - 1)) SA[--bkt[s(j - 1)]] = j - 1;
    }
}

template <class Acc, class I>
static void sais_core(const Acc &s, I *SA, I n, I K) {
    if (n == 1) {
        SA[0] = 0;
        return;
    }
    TypeBits t((size_t)n);
    std::vector<I> cnt((size_t)K, 0), bkt((size_t)K, 0);
    for (I i = 0; i < n; i++) cnt[s(i)]++;
    t.set((size_t)n - 1);
    {
        int64_t c1 = s(n - 1);
        bool t1 = true;
        for (I i = n - 2; i >= 0; i--) {
            int64_t c0 = s(i);
            bool t0 = c0 < c1 || (c0 == c1 && t1);
            if (t0) t.set((size_t)i);
            c1 = c0;
            t1 = t0;
        }
    }
    // stage 1: sort the LMS substrings by induced sorting
    sais_buckets(s, n, cnt, bkt, true);
    for (I i = 0; i < n; i++) SA[i] = -1;
    for (I i = 1; i < n; i++)
        if (t.lms((size_t)i)) SA[--bkt[s(i)]] = i;
    sais_induce(s, t, SA, n, cnt, bkt);
    I n1 = 0;
    for (I i = 0; i < n; i++) {
        I p = SA[i];
        if (p > 0 && t.lms((size_t)p)) SA[n1++] = p;
    }
    for (I i = n1; i < n; i++) SA[i] = -1;
    I name = 0, prev = -1;
    for (I i = 0; i < n1; i++) {
        I pos = SA[i];
        bool diff = false;
        if (prev < 0) {
            diff = true;
        } else {
            for (I d = 0;; d++) {
                if (s(pos + d) != s(prev + d) || t.get((size_t)(pos + d)) != t.get((size_t)(prev + d))) {
                    diff = true;
                    break;
                }
                if (d > 0 && (t.lms((size_t)(pos + d)) || t.lms((size_t)(prev + d)))) break;
            }
        }
        if (diff) {
            name++;
            prev = pos;
        }
        SA[n1 + pos / 2] = name - 1;
    }
    {
        I j = n - 1;
        for (I i = n - 1; i >= n1; i--)
            if (SA[i] >= 0) SA[j--] = SA[i];
    }
    I *s1 = SA + n - n1;
    // stage 2: order the LMS suffixes (recursively if names collide)
    if (name < n1) {
        PlainAcc<I> r{s1};
        sais_core<PlainAcc<I>, I>(r, SA, n1, name);
    } else {
        for (I i = 0; i < n1; i++) SA[s1[i]] = i;
    }
    // stage 3: induce the order of all suffixes from the sorted LMS suffixes
    {
        I j = 0;
        for (I i = 1; i < n; i++)
            if (t.lms((size_t)i)) s1[j++] = i;
    }
    for (I i = 0; i < n1; i++) SA[i] = s1[SA[i]];
    for (I i = n1; i < n; i++) SA[i] = -1;
    sais_buckets(s, n, cnt, bkt, true);
    for (I i = n1 - 1; i >= 0; i--) {
        I j = SA[i];
        SA[i] = -1;
        SA[--bkt[s(j)]] = j;
    }
    sais_induce(s, t, SA, n, cnt, bkt);
}

// SA of text[0..n) into out[0..n) (I = int32_t or int64_t).  `unique_sentinel` says the last
// symbol is a unique minimum (single-piece texts), which lets the text be used in place.
template <class I>
static void suffix_array_bytes(const uint8_t *text, uint64_t n, I *out, bool unique_sentinel) {
    if (n == 0) return;
    if (n == 1) {
        out[0] = 0;
        return;
    }
    if (unique_sentinel) {
        PlainAcc<uint8_t> a{text};
        sais_core<PlainAcc<uint8_t>, I>(a, out, (I)n, (I)256);
    } else {
        std::vector<I> tmp(n + 1);
        SentinelAcc a{text, (int64_t)n + 1};
        sais_core<SentinelAcc, I>(a, tmp.data(), (I)(n + 1), (I)257);
        std::memcpy(out, tmp.data() + 1, n * sizeof(I));
    }
}

}  // namespace fmx
