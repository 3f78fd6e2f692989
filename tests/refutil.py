"""Test helpers restating the reference's own test utilities (tests/testutil/mod.rs,
src/testutil.rs).  rand's StdRng stream cannot be reproduced without the crate, so
the random tests are reproduced as a METHOD (differential vs a naive scan) with
numpy's PCG64 seeded explicitly."""
import numpy as np

TEXT_README = (
    b"Lorem ipsum dolor sit amet, consectetur adipiscing elit, sed do eiusmod tempor incididunt ut labore et dolore magna aliqua."
    b"Ut enim ad minim veniam, quis nostrud exercitation ullamco laboris nisi ut aliquip ex ea commodo consequat."
    b"Duis aute irure dolor in reprehenderit in voluptate velit esse cillum dolore eu fugiat nulla pariatur."
    b"Excepteur sint occaecat cupidatat non proident, sunt in culpa qui officia deserunt mollit anim id est laborum."
    b"\0"
)  # README.md:35-41

TEXT_TWINKLE = (
    b"Twinkle, twinkle, little star,\n"
    b"How I wonder what you are!\n"
    b"Up above the world so high,\n"
    b"Like a diamond in the sky.\n"
    b"Twinkle, twinkle, little star,\n"
    b"How I wonder what you are!\n\0"
    b"When the blazing sun is gone,\n"
    b"When he nothing shines upon,\n"
    b"Then you show your little light,\n"
    b"Twinkle, twinkle, all the night.\n"
    b"Twinkle, twinkle, little star,\n"
    b"How I wonder what you are!\n\0"
    b"Then the traveller in the dark,\n"
    b"Thanks you for your tiny spark;\n"
    b"He could not see which way to go,\n"
    b"If you did not twinkle so.\n"
    b"Twinkle, twinkle, little star,\n"
    b"How I wonder what you are!\n\0"
)  # examples/multi_pieces.rs:5-28


def build_text(rng, length, alphabet_size, multi_pieces):
    """tests/testutil/mod.rs:7-32 + :109-113: no leading zero, no double zeros, one trailing zero."""
    def gen():
        v = int(rng.integers(0, 256)) % alphabet_size
        return v if multi_pieces else v + 1

    text = [0] * length
    if length == 1:
        return bytes(text)
    prev_zero = True
    for i in range(length - 1):
        c = gen()
        if prev_zero:
            while c == 0:
                c = gen()
        prev_zero = c == 0
        text[i] = c
    while text[length - 2] == 0:
        text[length - 2] = gen()
    return bytes(text)


def naive_search(text: bytes, pattern: bytes, prefix_only=False, suffix_only=False):
    """tests/testutil/mod.rs:56-87 (NaiveSearchIndex::do_search): list of (position, piece_id)."""
    res = []
    piece_id = 0
    n, m = len(text), len(pattern)
    for i in range(0, n - m + 1):
        if text[i] == 0:
            piece_id += 1
        if ((not prefix_only or i == 0 or text[i - 1] == 0)
                and (not suffix_only or i + m == n or text[i + m] == 0)
                and text[i:i + m] == pattern):
            res.append((i, piece_id))
    return res


def naive_suffix_array(text: bytes):
    """src/testutil.rs:25-35."""
    return sorted(range(len(text)), key=lambda i: text[i:])


def random_cases(seed, texts, patterns, text_size_max, alphabet_size, level_max, pattern_size_max, multi_pieces):
    """tests/testutil/mod.rs:95-143 (TestRunner::run) as a generator of (text, level, [patterns])."""
    rng = np.random.default_rng(seed)
    for _ in range(texts):
        text_size = int(rng.integers(2, text_size_max + 1))
        text = build_text(rng, text_size, alphabet_size, multi_pieces)
        level = int(rng.integers(0, level_max + 1))
        pats = []
        for _ in range(patterns):
            psm = min(pattern_size_max, text_size)
            psize = 1 if psm == 1 else int(rng.integers(1, psm))
            pats.append(bytes(int(rng.integers(0, 256)) % (alphabet_size - 1) + 1 for _ in range(psize)))
        yield text, level, pats
