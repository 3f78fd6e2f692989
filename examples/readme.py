"""The reference's README example (README.md:35-85), line for line, on the CUDA engine through the Python mirror of the
crate's API.  Needs a B200 (there is no CPU fallback):   python examples/readme.py"""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fmx_pkg

fmx = fmx_pkg.load()

text = (
    b"Lorem ipsum dolor sit amet, consectetur adipiscing elit, sed do eiusmod tempor incididunt ut labore et dolore magna aliqua."
    b"Ut enim ad minim veniam, quis nostrud exercitation ullamco laboris nisi ut aliquip ex ea commodo consequat."
    b"Duis aute irure dolor in reprehenderit in voluptate velit esse cillum dolore eu fugiat nulla pariatur."
    b"Excepteur sint occaecat cupidatat non proident, sunt in culpa qui officia deserunt mollit anim id est laborum."
    b"\0"
)
# let text = Text::new(text);  let index = FMIndexWithLocate::new(&text, 2).unwrap();
index = fmx.FMIndexWithLocate.new(fmx.Text.new(text), 2)

# Count the number of occurrences.
search = index.search("dolor")
n = search.count()
assert n == 4

# List the position of all occurrences (the reference's iteration order).
positions = [m.locate() for m in search.iter_matches()]
assert positions == [246, 12, 300, 103]

# Extract preceding characters from a search position.
first = next(iter(search.iter_matches()))
prefix = bytes(reversed(list(itertools.islice(first.iter_chars_backward(), 16))))
assert prefix == b"Duis aute irure "

# Extract succeeding characters from a search position.
fourth = list(search.iter_matches())[3]
postfix = bytes(itertools.islice(fourth.iter_chars_forward(), 20))
assert postfix == b"dolore magna aliqua."

# Search can be chained backward.
assert search.search("ipsum ").count() == 1

# The batched entry the GPU path exists for: many patterns, one fused call (counts + CSR hit lists).
r = index.query_batch([b"dolor", b"ipsum dolor", b"zzz"], counts=True)
assert list(r["counts"]) == [4, 1, 0] and list(r["positions"][:4]) == positions
print("README example ok:", n, "matches at", positions)
