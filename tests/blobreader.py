"""CPU reader of the device-layout blob (fm-index_b200/csrc/fmx_layout.h).  TEST CODE ONLY:
it lets the host builder be checked against the oracle without a GPU.  Slow, pure numpy/Python."""
import struct

import numpy as np

SEC_LEVEL0, SEC_ADJ, SEC_CS, SEC_SA, SEC_DOC, SEC_PIECE_END, SEC_RL_B, SEC_RL_BP, SEC_RL_BSEL, SEC_RL_BPSEL, SEC_COUNT = \
    0, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17
RB_BITS = 224
M32 = 0xFFFFFFFF


class RBVec:
    def __init__(self, raw):
        self.w = np.frombuffer(raw, dtype=np.uint32).reshape(-1, 8)

    def rank1(self, pos):
        b, r = divmod(pos, RB_BITS)
        c = int(self.w[b, 0])
        for k in range(7):
            if r >= 32 * (k + 1):
                c += bin(int(self.w[b, 1 + k])).count("1")
            elif r > 32 * k:
                c += bin(int(self.w[b, 1 + k]) & ((1 << (r - 32 * k)) - 1)).count("1")
        return c

    def bit(self, pos):
        b, r = divmod(pos, RB_BITS)
        return (int(self.w[b, 1 + (r >> 5)]) >> (r & 31)) & 1


class Blob:
    def __init__(self, raw: np.ndarray):
        raw = bytes(raw)
        self.raw = raw
        f = struct.unpack_from("<QIIQQIIIIIIQQQQ", raw, 0)
        (self.magic, self.version, self.kind, self.n, self.seq_len, self.levels, self.max_character, self.cs_len,
         self.has_locate, self.sa_level, self.sa_word_size, self.sa_count, self.ndoc, self.first_row, self.runs) = f
        o = struct.calcsize("<QIIQQIIIIIIQQQQ")
        self.zeros = struct.unpack_from("<8Q", raw, o)
        o += 64
        (self.total_bytes,) = struct.unpack_from("<Q", raw, o)
        o += 8
        self.sec = [struct.unpack_from("<QQ", raw, o + 16 * k) for k in range(SEC_COUNT)]
        assert self.total_bytes == len(raw)
        self.lv = [RBVec(self._sec(SEC_LEVEL0 + l)) for l in range(self.levels)]
        self.adj = np.frombuffer(self._sec(SEC_ADJ), dtype=np.uint32)
        self.cs = np.frombuffer(self._sec(SEC_CS), dtype=np.uint32)
        self.sa = np.frombuffer(self._sec(SEC_SA), dtype=np.uint32)
        self.doc = np.frombuffer(self._sec(SEC_DOC), dtype=np.uint32)
        self.piece_end = np.frombuffer(self._sec(SEC_PIECE_END), dtype=np.uint32)
        if self.kind == 1:
            self.b = RBVec(self._sec(SEC_RL_B))
            self.bp = RBVec(self._sec(SEC_RL_BP))
            self.bsel = np.frombuffer(self._sec(SEC_RL_BSEL), dtype=np.uint32)
            self.bpsel = np.frombuffer(self._sec(SEC_RL_BPSEL), dtype=np.uint32)

    def _sec(self, k):
        off, nb = self.sec[k]
        assert off % 256 == 0
        return self.raw[off:off + nb]

    # the same arithmetic the kernels do (kernels.cuh), in Python
    def walk(self, c, pos):
        L = self.levels
        for l in range(L):
            ones = self.lv[l].rank1(pos)
            pos = self.zeros[l] + ones if (c >> (L - 1 - l)) & 1 else pos - ones
        return pos

    def access_walk(self, pos):
        L, c = self.levels, 0
        for l in range(L):
            ones, bit = self.lv[l].rank1(pos), self.lv[l].bit(pos)
            c = (c << 1) | bit
            pos = self.zeros[l] + ones if bit else pos - ones
        return c, pos

    def _zero_rule(self, i, rank):
        return rank + 1 if i < self.first_row else (0 if i == self.first_row else rank)

    def lf_map2(self, c, i):
        if self.kind == 1:
            j = self.b.rank1(i)
            starts = self.b.bit(i) if i < self.n else 0
            h = j if starts else j - 1
            hc, _ = self.access_walk(h)
            nr = (int(self.adj[c]) + self.walk(c, j)) & M32
            t = int(self.bpsel[nr])
            return t if hc != c else t + i - int(self.bsel[j])
        w = (int(self.adj[c]) + self.walk(c, i)) & M32
        if self.kind == 2 and c == 0:
            return self._zero_rule(i, w)
        return w

    def lf_step(self, i):
        if self.kind == 1:
            j = self.b.rank1(i)
            h = j if self.b.bit(i) else j - 1
            c, q = self.access_walk(h)
            nr = (int(self.adj[c]) + q + (j - h)) & M32
            return c, int(self.bpsel[nr]) + i - int(self.bsel[j])
        c, w = self.access_walk(i)
        w = (w + int(self.adj[c])) & M32
        if self.kind == 2 and c == 0:
            return c, self._zero_rule(i, w)
        return c, w

    def get_sa(self, i):
        steps, mask = 0, (1 << self.sa_level) - 1
        while i & mask:
            _, i = self.lf_step(i)
            steps += 1
        return (int(self.sa[i >> self.sa_level]) + steps) % self.n

    def search(self, pat, mode=0):
        s, e = 0, (self.ndoc if mode in (2, 3) else self.n)
        for c in reversed(pat):
            s, e = self.lf_map2(c, s), self.lf_map2(c, e)
            if s == e:
                break
        return s, e
