"""The reference's examples/multi_pieces.rs (lines 5-88) on the CUDA engine through the Python mirror of the crate's API.
Needs a B200 (there is no CPU fallback):   python examples/multi_pieces.py"""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fmx_pkg

fmx = fmx_pkg.load()

text = (
    b"Twinkle, twinkle, little star,\nHow I wonder what you are!\nUp above the world so high,\nLike a diamond in the sky.\n"
    b"Twinkle, twinkle, little star,\nHow I wonder what you are!\n\0"
    b"When the blazing sun is gone,\nWhen he nothing shines upon,\nThen you show your little light,\n"
    b"Twinkle, twinkle, all the night.\nTwinkle, twinkle, little star,\nHow I wonder what you are!\n\0"
    b"Then the traveller in the dark,\nThanks you for your tiny spark;\nHe could not see which way to go,\n"
    b"If you did not twinkle so.\nTwinkle, twinkle, little star,\nHow I wonder what you are!\n\0"
)
index = fmx.FMIndexMultiPiecesWithLocate.new(fmx.Text.new(text), 2)

# Count the number of occurrences.
assert index.search("star").count() == 4

# List the pieces that contain the pattern.
pieces = sorted(int(m.piece_id()) for m in index.search("How I wonder").iter_matches())
assert pieces == [0, 0, 1, 2]

# Extract preceding / succeeding characters from a search position.
pre = [bytes(itertools.takewhile(lambda c: c != ord(" "), m.iter_chars_backward())) for m in index.search(" in the dark").iter_matches()]
assert pre == [b"rellevart"]
suc = [bytes(itertools.takewhile(lambda c: c != ord(","), m.iter_chars_forward())) for m in index.search("ing ").iter_matches()]
assert suc == [b"ing shines upon", b"ing sun is gone"]

# Search for a pattern that is a prefix / suffix of a piece.
assert sorted(int(m.piece_id()) for m in index.search_prefix("Twinkle").iter_matches()) == [0]
assert sorted(int(m.piece_id()) for m in index.search_suffix("what you are!\n").iter_matches()) == [0, 1, 2]

# The same index partitioned BY PIECE over the GPUs of this process (here: three partitions on device 0): counts, match
# sets and piece ids of the whole text -- the form a text beyond one GPU's memory, or beyond 2^32 symbols, takes.
group = fmx.IndexGroup(fmx.Text.new(text), fmx.KIND_MULTI, 2, [0, 0, 0], fmx.GROUP_BY_PIECE)
r = group.query_batch([b"How I wonder", b"star"], piece_ids=True, counts=True)
assert list(r["counts"]) == [4, 4] and sorted(int(p) for p in r["piece_ids"][:4]) == [0, 0, 1, 2]
print("multi_pieces example ok: pieces", pieces)
