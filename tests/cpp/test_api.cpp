// C++ facade test: the reference's README doctest (README.md:31-86), examples/multi_pieces.rs:34-88
// and the mississippi SA ranges (rlfmi.rs:314-319), written against include/fmx.hpp the way the
// reference's tests are written against the crate.  Needs a GPU.
#include <algorithm>
#include <cstdio>
#include <string>

#include "fmx.hpp"

#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                      \
        }                                                                  \
    } while (0)

static std::string str(const std::vector<uint8_t> &v) { return std::string(v.begin(), v.end()); }

int main() {
    using namespace fmx;
    {
        std::string raw =
            "Lorem ipsum dolor sit amet, consectetur adipiscing elit, sed do eiusmod tempor incididunt ut labore et dolore magna aliqua."
            "Ut enim ad minim veniam, quis nostrud exercitation ullamco laboris nisi ut aliquip ex ea commodo consequat."
            "Duis aute irure dolor in reprehenderit in voluptate velit esse cillum dolore eu fugiat nulla pariatur."
            "Excepteur sint occaecat cupidatat non proident, sunt in culpa qui officia deserunt mollit anim id est laborum.";
        raw.push_back('\0');
        Text text(raw);
        FMIndexWithLocate index(text, 2);
        Search search = index.search("dolor");
        CHECK(search.count() == 4);
        std::vector<uint64_t> positions;
        for (auto &m : search.iter_matches()) positions.push_back(m.locate());
        CHECK((positions == std::vector<uint64_t>{246, 12, 300, 103}));
        auto prefix = search.iter_matches().front().iter_chars_backward(16);
        std::reverse(prefix.begin(), prefix.end());
        CHECK(str(prefix) == "Duis aute irure ");
        CHECK(str(search.iter_matches()[3].iter_chars_forward(20)) == "dolore magna aliqua.");
        CHECK(index.len() == raw.size() && index.heap_size() > 0);
        auto b = index.search_batch({"dolor", "ipsum", "zzz", ""}, true);
        CHECK(b.count(0) == 4 && b.count(1) == 1 && b.count(2) == 0 && b.count(3) == raw.size());
        CHECK((std::vector<uint64_t>(b.positions.begin(), b.positions.begin() + 4) == std::vector<uint64_t>{246, 12, 300, 103}));
        CHECK(search.search("m ").count() == 2 && index.search("or").search("dol").count() == 4);  // "ipsum dolor", "cillum dolore"
    }
    {
        std::string raw = std::string("mississippi") + '\0';
        for (int rl = 0; rl < 2; rl++) {
            Text t(raw);
            std::pair<uint64_t, uint64_t> r[4];
            const char *pats[4] = {"iss", "ppi", "si", "ssi"};
            if (rl) {
                RLFMIndex index(t);
                for (int k = 0; k < 4; k++) r[k] = index.search(pats[k]).get_range();
            } else {
                FMIndex index(t);
                for (int k = 0; k < 4; k++) r[k] = index.search(pats[k]).get_range();
            }
            using R = std::pair<uint64_t, uint64_t>;
            CHECK((r[0] == R(3, 5) && r[1] == R(7, 8)));
            CHECK((r[2] == R(8, 10) && r[3] == R(10, 12)));
        }
    }
    {
        std::string raw =
            "Twinkle, twinkle, little star,\nHow I wonder what you are!\nUp above the world so high,\nLike a diamond in the sky.\n"
            "Twinkle, twinkle, little star,\nHow I wonder what you are!\n";
        raw.push_back('\0');
        raw += "When the blazing sun is gone,\nWhen he nothing shines upon,\nThen you show your little light,\n"
               "Twinkle, twinkle, all the night.\nTwinkle, twinkle, little star,\nHow I wonder what you are!\n";
        raw.push_back('\0');
        raw += "Then the traveller in the dark,\nThanks you for your tiny spark;\nHe could not see which way to go,\n"
               "If you did not twinkle so.\nTwinkle, twinkle, little star,\nHow I wonder what you are!\n";
        raw.push_back('\0');
        Text text(raw);
        FMIndexMultiPiecesWithLocate index(text, 2);
        CHECK(index.search("star").count() == 4 && index.pieces_count() == 3);
        std::vector<uint64_t> ids;
        for (auto &m : index.search("How I wonder").iter_matches()) ids.push_back(m.piece_id());
        std::sort(ids.begin(), ids.end());
        CHECK((ids == std::vector<uint64_t>{0, 0, 1, 2}));
        auto dark = index.search(" in the dark").iter_matches();
        CHECK(dark.size() == 1);
        auto pre = str(dark[0].iter_chars_backward(9));
        CHECK(pre == "rellevart");
        std::vector<std::string> suc;
        for (auto &m : index.search("ing ").iter_matches()) {
            auto s = str(m.iter_chars_forward(40));
            suc.push_back(s.substr(0, s.find(',')));
        }
        CHECK((suc == std::vector<std::string>{"ing shines upon", "ing sun is gone"}));
        ids.clear();
        for (auto &m : index.search_prefix("Twinkle").iter_matches()) ids.push_back(m.piece_id());
        CHECK((ids == std::vector<uint64_t>{0}));
        ids.clear();
        for (auto &m : index.search_suffix("what you are!\n").iter_matches()) ids.push_back(m.piece_id());
        std::sort(ids.begin(), ids.end());
        CHECK((ids == std::vector<uint64_t>{0, 1, 2}));
    }
    {
        bool threw = false;
        try {
            std::string bad("\0abc\0", 5);
            FMIndex index{Text(bad)};
        } catch (const InvalidText &e) {
            threw = std::string(e.what()).find("must not start with zero") != std::string::npos;
        }
        CHECK(threw);
        FMIndex small(Text::with_max_character({1, 2, 3, 4, 1, 2, 0}, 4));
        threw = false;
        try {
            small.search(std::vector<uint8_t>{1, 5});
        } catch (const std::out_of_range &) {
            threw = true;
        }
        CHECK(threw);
        CHECK(small.search(std::vector<uint8_t>{1, 2}).count() == 2);
    }
    std::puts("cpp facade ok");
    return 0;
}
