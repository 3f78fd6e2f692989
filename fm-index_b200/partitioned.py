"""Piece-partitioned MultiPieces: the one place the query path has an exchange step.

The pieces of a multi-piece text are split into contiguous groups, one group per GPU; every rank
builds an independent FMIndexMultiPiecesWithLocate over the concatenation of ITS pieces, every
pattern is answered by every rank, and the per-pattern hit counts / hit lists are combined with
NCCL collectives (`torch.distributed`, one process per GPU):

    counts      : all_gather of one int64 per pattern per rank  -> exclusive scans give CSR offsets
    hit lists   : all_gather of the (padded) local position / piece-id lists -> scattered into CSR

A pattern without \\0 cannot span two pieces, so counts add exactly and the union of the local hit
lists (positions shifted by the shard's text offset, piece ids by its first piece) is exactly the
reference's match set (multi_pieces.rs:188-223).  What a partitioned index canNOT reproduce is the
reference's iteration ORDER (global SA-row order depends on text outside the shard): hits of one
pattern come shard-major, and inside a shard in that shard's SA-row order.  Use the replicated index
when order parity matters.  Patterns containing \\0 are rejected here for the same reason.

The local engine is pluggable so the host logic is testable on CPU with the gloo backend
(tests/test_partitioned.py uses the oracle as the engine); on GPUs the engine is the CUDA index.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def piece_bounds(text: np.ndarray):
    """(starts, ends) of the pieces of a multi-piece text; piece k is text[starts[k]:ends[k]] + \\0."""
    ends = np.flatnonzero(text == 0)
    starts = np.concatenate([[0], ends[:-1] + 1])
    return starts, ends


def partition_pieces(lengths, world: int):
    """Contiguous groups of pieces, balanced by symbol count: returns `world` (first, last+1) ranges."""
    lengths = np.asarray(lengths, dtype=np.int64)
    d = len(lengths)
    if d < world:
        raise ValueError(f"{d} pieces cannot be partitioned over {world} ranks")
    csum = np.concatenate([[0], np.cumsum(lengths)])
    total = int(csum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(csum, target))
        k = min(max(k, cuts[-1] + 1), d - (world - r))  # every rank keeps at least one piece
        # choose the nearer of k-1 / k to the ideal cut
        if k - 1 > cuts[-1] and abs(csum[k - 1] - target) < abs(csum[k] - target):
            k -= 1
        cuts.append(k)
    cuts.append(d)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


class CudaEngine:
    """Local engine = the CUDA index of this package (results come back as device tensors)."""

    def __init__(self, shard_text: np.ndarray, level: int, max_character: int, device: int):
        from . import FMIndexMultiPiecesWithLocate, Text

        self.index = FMIndexMultiPiecesWithLocate.new(Text.with_max_character(shard_text, max_character), level,
                                                      device=device)
        self.device = torch.device("cuda", device)

    def search_locate(self, patterns: torch.Tensor, mode: int = 0):
        """-> (hit_off[npat+1], positions, piece_ids) as int64 tensors on the device"""
        import ctypes as C

        L, h = self.index._L, self.index._h
        pats = patterns.to(self.device).contiguous()
        npat, m = pats.shape
        st = torch.cuda.current_stream(self.device)
        sp = C.c_void_p(st.cuda_stream) if st.cuda_stream else None
        d_s = torch.empty(npat, dtype=torch.int64, device=self.device)
        d_e = torch.empty_like(d_s)
        d_off = torch.empty(npat + 1, dtype=torch.int64, device=self.device)
        total = C.c_uint64(0)
        prefix_only = 1 if mode in (1, 3) else 0

        def chk(rc):
            if rc != 0:
                raise RuntimeError(L.fmx_last_error().decode())

        chk(L.fmx_search_batch_device(h, mode, pats.data_ptr(), None, m, npat, None, None, d_s.data_ptr(), d_e.data_ptr(), sp))
        chk(L.fmx_locate_count_device(h, prefix_only, d_s.data_ptr(), d_e.data_ptr(), npat, d_off.data_ptr(), C.byref(total), sp))
        n = int(total.value)
        pos = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        pid = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        chk(L.fmx_locate_fill_device(h, prefix_only, d_s.data_ptr(), d_e.data_ptr(), npat, d_off.data_ptr(), n,
                                     pos.data_ptr(), pid.data_ptr(), sp))
        chk(L.fmx_search_check(h, sp))
        return d_off, pos[:n], pid[:n]


class PartitionedMultiPieces:
    """FMIndexMultiPiecesWithLocate partitioned by piece over the ranks of a process group."""

    def __init__(self, text, level: int, max_character: int = 255, group=None, engine_factory=None, device=None):
        text = np.ascontiguousarray(np.asarray(text, dtype=np.uint8))
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        starts, ends = piece_bounds(text)
        if len(ends) == 0 or ends[-1] != text.size - 1:
            raise ValueError("the given text must end with exactly one zero character")
        self.pieces_total = len(ends)
        self.len_total = int(text.size)
        self.ranges = partition_pieces(ends - starts + 1, self.world)
        first, last = self.ranges[self.rank]
        self.first_piece = first
        self.base_offset = int(starts[first])
        shard = text[self.base_offset:int(ends[last - 1]) + 1]
        if engine_factory is None:
            dev = torch.cuda.current_device() if device is None else device
            engine_factory = lambda t: CudaEngine(t, level, max_character, dev)  # noqa: E731
        self.engine = engine_factory(shard)
        self.shard_len = int(shard.size)

    def len(self):
        return self.len_total

    def pieces_count(self):
        return self.pieces_total

    def _all_gather(self, t: torch.Tensor):
        if self.world == 1:
            return [t]
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        return out

    def search_locate(self, patterns, mode: int = 0):
        """Every rank passes the SAME patterns ([npat, m] uint8).  Returns, on every rank,
        (counts[npat], hit_off[npat+1], positions, piece_ids): global text positions and global piece
        ids in CSR form, hits of one pattern ordered shard-major."""
        pats = torch.as_tensor(patterns, dtype=torch.uint8)
        if bool((pats == 0).any()):
            raise ValueError("patterns containing \\0 are not supported by the piece-partitioned index")
        off, pos, pid = self.engine.search_locate(pats, mode)
        dev = off.device
        npat = pats.shape[0]
        pos = pos + self.base_offset      # shard-local -> global text position
        pid = pid + self.first_piece      # shard-local -> global piece id
        counts = (off[1:] - off[:-1]).contiguous()
        # ---- exchange 1: per-pattern counts of every rank
        all_counts = torch.stack(self._all_gather(counts))          # [world, npat]
        total = all_counts.sum(dim=0)
        hit_off = torch.zeros(npat + 1, dtype=torch.int64, device=dev)
        hit_off[1:] = torch.cumsum(total, dim=0)
        rank_prefix = torch.cumsum(all_counts, dim=0) - all_counts  # hits of lower ranks, per pattern
        # ---- exchange 2: the hit lists themselves (padded to the longest)
        sizes = all_counts.sum(dim=1)
        cap = max(int(sizes.max().item()), 1)

        def padded(t):
            buf = torch.zeros(cap, dtype=torch.int64, device=dev)
            buf[: t.numel()] = t
            return buf

        g_pos = self._all_gather(padded(pos))
        g_pid = self._all_gather(padded(pid))
        n_total = int(hit_off[-1].item())
        positions = torch.empty(n_total, dtype=torch.int64, device=dev)
        pieces = torch.empty(n_total, dtype=torch.int64, device=dev)
        ar = torch.arange(npat, device=dev)
        for r in range(self.world):
            cr = all_counts[r]
            nr = int(sizes[r].item())
            if nr == 0:
                continue
            owner = torch.repeat_interleave(ar, cr)                 # pattern of every local hit of rank r
            local_off = torch.cumsum(cr, dim=0) - cr
            within = torch.arange(nr, device=dev) - local_off[owner]
            dest = hit_off[:-1][owner] + rank_prefix[r][owner] + within
            positions[dest] = g_pos[r][:nr]
            pieces[dest] = g_pid[r][:nr]
        return total, hit_off, positions, pieces
