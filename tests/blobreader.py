"""CPU reader of the device-layout blob (fm-index_b200/csrc/fmx_layout.h).  TEST CODE ONLY:
it lets the host builder be checked against the oracle without a GPU.  Slow, pure numpy/Python."""
import struct

import numpy as np

SEC_LEVEL0, SEC_ADJ, SEC_CS, SEC_SA, SEC_DOC, SEC_PIECE_END, SEC_RL_B, SEC_RL_BP, SEC_RL_BSEL, SEC_RL_BPSEL, SEC_EXC, SEC_TEXT, SEC_ISA, \
    SEC_VSA, SEC_WZEROS, SEC_COUNT = 0, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22
RB_BITS = 192
M32 = 0xFFFFFFFF


class RBVec:
    def __init__(self, raw):
        self.w = np.frombuffer(raw, dtype=np.uint32).reshape(-1, 8)

    def rank1(self, pos):
        b, r = divmod(pos, RB_BITS)
        j, x = r >> 6, r & 63
        word = int(self.w[b, 2 + 2 * j]) | (int(self.w[b, 3 + 2 * j]) << 32)
        return int(self.w[b, 0]) + ((int(self.w[b, 1]) >> (8 * j)) & 0xFF) + bin(word & ((1 << x) - 1)).count("1")

    def bit(self, pos):
        b, r = divmod(pos, RB_BITS)
        return (int(self.w[b, 2 + (r >> 5)]) >> (r & 31)) & 1


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
        o += 16 * SEC_COUNT
        self.layout, self.nexc, self.qlevels, self.sym_nblk = struct.unpack_from("<IIII", raw, o)
        o += 16
        self.verify, self.isa_level, self.vsa_level, self.char_width = struct.unpack_from("<IIII", raw, o)
        o += 16
        self.qoff = [struct.unpack_from("<4Q", raw, o + 32 * l) for l in range(4)]
        assert self.total_bytes == len(raw)
        if self.layout == 4:   # WIDE: all levels in SEC_LEVEL0, zeros per level in SEC_WZEROS, cs / adj of max_character + 1 words
            nblk = self.seq_len // RB_BITS + 1
            all_lv = self._sec(SEC_LEVEL0)
            assert len(all_lv) == self.levels * nblk * 32
            self.lv = [RBVec(all_lv[l * nblk * 32:(l + 1) * nblk * 32]) for l in range(self.levels)]
            self.zeros = [int(v) for v in np.frombuffer(self._sec(SEC_WZEROS), dtype=np.uint32)]
            assert len(self.zeros) == self.levels and self.char_width in (2, 4, 8) and self.max_character > 255
        elif self.layout == 3:
            assert self.sym_nblk == self.seq_len // RB_BITS + 1
            allv = np.frombuffer(self._sec(SEC_LEVEL0), dtype=np.uint32).reshape(self.cs_len, self.sym_nblk, 8)
            self.symv = [RBVec(allv[c].tobytes()) for c in range(self.cs_len)]
            self.rawseq = np.frombuffer(self._sec(SEC_LEVEL0 + 1), dtype=np.uint8)
            self.lv = []
        elif self.layout == 2:
            assert self.qlevels == (self.levels + 1) // 2
            self.q4l = [np.frombuffer(self._sec(SEC_LEVEL0 + l), dtype=np.uint32).reshape(-1, 8) for l in range(self.qlevels)]
            self.lv = []
        elif self.layout == 1:
            self.q4 = np.frombuffer(self._sec(SEC_LEVEL0), dtype=np.uint32).reshape(-1, 8)
            self.exc = [int(v) for v in np.frombuffer(self._sec(SEC_EXC), dtype=np.uint32)]
            assert len(self.exc) == self.nexc
            self.lv = []
        else:
            self.lv = [RBVec(self._sec(SEC_LEVEL0 + l)) for l in range(self.levels)]
        self.adj = np.frombuffer(self._sec(SEC_ADJ), dtype=np.uint32)
        self.cs = np.frombuffer(self._sec(SEC_CS), dtype=np.uint32)
        self.sa = np.frombuffer(self._sec(SEC_SA), dtype=np.uint32)
        self.doc = np.frombuffer(self._sec(SEC_DOC), dtype=np.uint32)
        self.piece_end = np.frombuffer(self._sec(SEC_PIECE_END), dtype=np.uint32)
        if self.verify:
            self.text = np.frombuffer(self._sec(SEC_TEXT), dtype=np.uint8)
            self.isa = np.frombuffer(self._sec(SEC_ISA), dtype=np.uint32)
            self.vsa = np.frombuffer(self._sec(SEC_VSA), dtype=np.uint32) if self.sec[SEC_VSA][1] else self.sa
        if self.kind == 1:
            self.b = RBVec(self._sec(SEC_RL_B))
            self.bp = RBVec(self._sec(SEC_RL_BP))
            self.bsel = np.frombuffer(self._sec(SEC_RL_BSEL), dtype=np.uint32)
            self.bpsel = np.frombuffer(self._sec(SEC_RL_BPSEL), dtype=np.uint32)

    def _sec(self, k):
        off, nb = self.sec[k]
        assert off % 256 == 0
        return self.raw[off:off + nb]

    # ---- Q4 layout
    def q4_code(self, i):
        b, t = divmod(i, 64)
        return ((int(self.q4[b, 4 + (t >> 5)]) >> (t & 31)) & 1) | (((int(self.q4[b, 6 + (t >> 5)]) >> (t & 31)) & 1) << 1)

    def q4_rank(self, i, c):
        import bisect
        x = bisect.bisect_left(self.exc, i)
        if c == 0:
            return x
        code = c - 1
        b, r = divmod(i, 64)
        v = int(self.q4[b, code]) + sum(1 for t in range(r) if self.q4_code(b * 64 + t) == code)
        return v - x if code == 0 else v

    def q4_access(self, i):
        import bisect
        k = bisect.bisect_left(self.exc, i)
        if k < len(self.exc) and self.exc[k] == i:
            return 0
        return self.q4_code(i) + 1

    # cs[c] + rank(i, c) and (seq[i], cs + rank) for either layout
    def seq_lf(self, c, i):
        if self.layout == 3:
            return int(self.cs[c]) + self.symv[c].rank1(i)
        if self.layout == 1:
            return int(self.cs[c]) + self.q4_rank(i, c)
        if self.layout == 4 and int(self.cs[c + 1]) == int(self.cs[c]):   # a symbol that does not occur (kernels.cuh sym_absent)
            return int(self.cs[c])
        return (int(self.adj[c]) + self.walk(c, i)) & M32

    def seq_access_lf(self, i):
        if self.layout == 3:
            c = int(self.rawseq[i])
            return c, int(self.cs[c]) + self.symv[c].rank1(i)
        if self.layout == 1:
            c = self.q4_access(i)
            return c, int(self.cs[c]) + self.q4_rank(i, c)
        c, w = self.access_walk(i)
        return c, (w + int(self.adj[c])) & M32

    # ---- WM4 layout: quaternary wavelet matrix
    def w4_code(self, l, i):
        b, t = divmod(i, 64)
        return ((int(self.q4l[l][b, 4 + (t >> 5)]) >> (t & 31)) & 1) | (((int(self.q4l[l][b, 6 + (t >> 5)]) >> (t & 31)) & 1) << 1)

    def w4_down(self, l, d, pos):
        b, r = divmod(pos, 64)
        return self.qoff[l][d] + int(self.q4l[l][b, d]) + sum(1 for t in range(r) if self.w4_code(l, b * 64 + t) == d)

    # the same arithmetic the kernels do (kernels.cuh), in Python
    def walk(self, c, pos):
        if self.layout == 2:
            for l in range(self.qlevels):
                pos = self.w4_down(l, (c >> (2 * (self.qlevels - 1 - l))) & 3, pos)
            return pos
        L = self.levels
        for l in range(L):
            ones = self.lv[l].rank1(pos)
            pos = self.zeros[l] + ones if (c >> (L - 1 - l)) & 1 else pos - ones
        return pos

    def access_walk(self, pos):
        if self.layout == 2:
            c = 0
            for l in range(self.qlevels):
                d = self.w4_code(l, pos)
                c = (c << 2) | d
                pos = self.w4_down(l, d, pos)
            return c, pos
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
            hc, _ = self.seq_access_lf(h)
            t = int(self.bpsel[self.seq_lf(c, j)])
            return t if hc != c else t + i - int(self.bsel[j])
        w = self.seq_lf(c, i)
        if self.kind == 2 and c == 0:
            return self._zero_rule(i, w)
        return w

    def lf_step(self, i):
        if self.kind == 1:
            j = self.b.rank1(i)
            h = j if self.b.bit(i) else j - 1
            c, q = self.seq_access_lf(h)
            return c, int(self.bpsel[q + (j - h)]) + i - int(self.bsel[j])
        c, w = self.seq_access_lf(i)
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


    # ---- the seed-and-verify tail of k_search (kernels.cuh: verify_tail), in Python
    @property
    def VERIFY_MIN(self):
        return 6 if self.isa_level == 0 else 10

    def _verify_tail(self, pat, k, s):
        """range [s, s+1), characters pat[0:k] still to consume -> (row, characters consumed); consumed 0 = nothing done"""
        row, st, mask = s, 0, (1 << self.vsa_level) - 1
        while row & mask:
            _, row = self.lf_step(row)
            st += 1
        pos = (int(self.vsa[row >> self.vsa_level]) + st) % self.n
        if pos < k:
            return s, 0
        matched = 0
        while matched < k:
            t = int(self.text[pos - 1 - matched])
            if t == 0 or pat[k - 1 - matched] != t:
                break
            matched += 1
        if matched == 0:
            return s, 0
        q = pos - matched
        step = 1 << self.isa_level
        q4 = (q + step - 1) // step * step
        if q4 >= self.n:
            r, q4 = 0, self.n - 1          # the suffix "\0" is row 0
        else:
            r = int(self.isa[q4 >> self.isa_level])
        for _ in range(q4 - q):
            _, r = self.lf_step(r)
        return r, matched

    def search_verify(self, pat, mode=0):
        """SearchWrapper::search with the verify tail (search_one in kernels.cuh); -> (s, e, steps executed)"""
        s, e, it, k = 0, (self.ndoc if mode in (2, 3) else self.n), 0, len(pat)
        armed = bool(self.verify) and self.kind != 1
        while k > 0:
            if armed and e - s == 1 and k >= self.VERIFY_MIN:
                s, got = self._verify_tail(pat, k, s)
                armed = got != 0
                if got:
                    e, it, k = s + 1, it + got, k - got
                    continue
            s, e = self.lf_map2(pat[k - 1], s), self.lf_map2(pat[k - 1], e)
            it += 1
            k -= 1
            if s == e:
                break
        return s, e, it
