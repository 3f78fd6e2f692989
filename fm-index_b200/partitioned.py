"""Piece-partitioned MultiPieces, one process per GPU: the one place the query path has an exchange step
(BASELINE config 4).

The pieces of a multi-piece text are split into contiguous, length-balanced groups, one group per rank; every rank
builds an independent FMIndexMultiPiecesWithLocate over the concatenation of ITS pieces and answers EVERY pattern
of the batch (one `fmx_query_batch_device` call: 32-bit CSR of local positions and local piece ids).  The parts are
then GATHERED ON ONE RANK over NCCL (NVLink / NVSwitch) and merged there by one CUDA kernel:

    hit offsets : `dist.gather` of npat + 1 uint32 per rank (fixed size)
    hit lists   : exact-size point-to-point transfers (`batch_isend_irecv` = grouped ncclSend / ncclRecv); the
                  root learns the sizes from the gathered offsets -- nothing is padded, only the root receives
    merge       : `fmx_csr_merge_device` -- one scan over the interleaved per-(pattern, rank) counts + one scatter
                  kernel that shifts positions / piece ids into the coordinates of the whole text

A pattern without \\0 cannot span two pieces, so counts add exactly and the union of the parts' hit lists is exactly
the reference's match set (multi_pieces.rs:188-223).  What a partitioned index canNOT reproduce is the reference's
iteration ORDER (global SA-row order depends on text outside the partition): hits of one pattern come rank-major,
inside a rank in that partition's SA-row order.  Use the replicated index when order parity matters.  Patterns
containing \\0 are rejected for the same reason.

The local engine and the merge are pluggable so that the host logic (partitioning, offsets, exchange) is testable
on CPU with the gloo backend (tests/test_partitioned.py: the oracle as the engine, a vectorised host merge); on
GPUs the engine is the CUDA index and the merge is the C-ABI kernel.  `fmx_group_create(FMX_GROUP_BY_PIECE)`
(include/fmx.h) is the same scheme for several GPUs of ONE process, with peer copies in place of NCCL.
"""
from __future__ import annotations

import ctypes as C

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
    """Local engine = the CUDA index of this package; results are int32 device tensors (n < 2^32: 4 bytes per
    offset / position / piece id on the wire)."""

    def __init__(self, shard_text: np.ndarray, level: int, max_character: int, device: int, mode: int = 0):
        from . import FMIndexMultiPiecesWithLocate, Text

        self.index = FMIndexMultiPiecesWithLocate.new(Text.with_max_character(shard_text, max_character), level,
                                                      device=device, mode=mode)
        self.device = torch.device("cuda", device)
        self._cap = 0
        self._pos = self._pid = None

    def search_locate(self, patterns: torch.Tensor, mode: int = 0):
        """-> (hit_off[npat+1], positions, piece_ids): int32 tensors on the device (values are uint32)"""
        from . import _lib

        L, h = self.index._L, self.index._h
        pats = patterns.to(self.device).contiguous()
        npat, m = pats.shape
        # the library launches on the stream it is handed: always torch's CURRENT stream, so that the work is
        # ordered after whatever produced `pats` (cudaStreamLegacy stands for the legacy default stream)
        st = torch.cuda.current_stream(self.device)
        sp = C.c_void_p(st.cuda_stream if st.cuda_stream else 1)
        d_off = torch.empty(npat + 1, dtype=torch.int32, device=self.device)
        if self._cap < npat + 1024:
            self._cap = 2 * npat + 1024
            self._pos = torch.empty(self._cap, dtype=torch.int32, device=self.device)
            self._pid = torch.empty(self._cap, dtype=torch.int32, device=self.device)
        for _ in range(2):
            q = _lib.Query()
            q.mode, q.patterns, q.fixed_len, q.npat, q.out_width = mode, pats.data_ptr(), m, npat, 4
            q.hit_off, q.positions, q.piece_ids, q.capacity = d_off.data_ptr(), self._pos.data_ptr(), self._pid.data_ptr(), self._cap
            if L.fmx_query_batch_device(h, C.byref(q), sp) != 0:
                raise RuntimeError(L.fmx_last_error().decode())
            n = int(d_off[-1].item()) & 0xFFFFFFFF          # the one host synchronisation of the local step
            if n <= self._cap:
                break
            self._cap = n + 1024                               # more hits than room: once more with the exact size
            self._pos = torch.empty(self._cap, dtype=torch.int32, device=self.device)
            self._pid = torch.empty(self._cap, dtype=torch.int32, device=self.device)
        if L.fmx_search_check(h, sp) != 0:
            raise IndexError(L.fmx_last_error().decode())
        return d_off, self._pos[:n], self._pid[:n]


def merge_parts_cuda(parts, bases, npat: int, device: torch.device):
    """fmx_csr_merge_device over device-resident int32 parts -> (hit_off int64[npat+1], positions, piece_ids)"""
    from . import _lib

    L = _lib.load()
    R = len(parts)
    arr = (_lib.CsrPart * R)()
    total = 0
    for r, ((off, pos, pid), (pb, db)) in enumerate(zip(parts, bases)):
        arr[r].hit_off, arr[r].positions, arr[r].piece_ids = off.data_ptr(), pos.data_ptr(), pid.data_ptr()
        arr[r].position_base, arr[r].piece_base = pb, db
        total += pos.numel()
    hit_off = torch.empty(npat + 1, dtype=torch.int64, device=device)
    positions = torch.empty(max(total, 1), dtype=torch.int64, device=device)
    pieces = torch.empty(max(total, 1), dtype=torch.int64, device=device)
    st = torch.cuda.current_stream(device)
    rc = L.fmx_csr_merge_device(device.index or 0, arr, R, npat, 4, hit_off.data_ptr(), positions.data_ptr(),
                                pieces.data_ptr(), total, C.c_void_p(st.cuda_stream if st.cuda_stream else 1))
    if rc != 0:
        raise RuntimeError(L.fmx_last_error().decode())
    return hit_off, positions[:total], pieces[:total]


def merge_parts_host(parts, bases, npat: int, device):
    """the same merge with torch ops (CPU tensors; the gloo tests): one scan over the interleaved counts, one
    gather per part -- no per-rank host synchronisation"""
    R = len(parts)
    offs = [p[0].to(torch.int64) & 0xFFFFFFFF for p in parts]
    cnt = torch.stack([o[1:] - o[:-1] for o in offs], dim=1).reshape(-1)          # [npat * R], pattern-major
    slot = torch.zeros(npat * R + 1, dtype=torch.int64)
    slot[1:] = torch.cumsum(cnt, dim=0)
    total = int(slot[-1])
    hit_off = torch.cat([slot[0:npat * R:R], slot[-1:]]) if npat else slot[-1:].clone()
    positions = torch.empty(total, dtype=torch.int64)
    pieces = torch.empty(total, dtype=torch.int64)
    ar = torch.arange(npat)
    for r, ((off, pos, pid), (pb, db)) in enumerate(zip(parts, bases)):
        c = offs[r][1:] - offs[r][:-1]
        owner = torch.repeat_interleave(ar, c)
        within = torch.arange(pos.numel()) - offs[r][:-1][owner]
        dest = slot[owner * R + r] + within
        positions[dest] = (pos.to(torch.int64) & 0xFFFFFFFF) + pb
        pieces[dest] = (pid.to(torch.int64) & 0xFFFFFFFF) + db
    return hit_off, positions, pieces


class PartitionedMultiPieces:
    """FMIndexMultiPiecesWithLocate partitioned by piece over the ranks of a process group."""

    def __init__(self, text, level: int, max_character: int = 255, group=None, engine_factory=None, device=None, mode: int = 0):
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
        # (text offset, first piece) of every rank's partition: all ranks know all of them
        self.bases = [(int(starts[a]), int(a)) for a, _ in self.ranges]
        first, last = self.ranges[self.rank]
        self.first_piece = first
        self.base_offset = int(starts[first])
        shard = text[self.base_offset:int(ends[last - 1]) + 1]
        if engine_factory is None:
            dev = torch.cuda.current_device() if device is None else device
            engine_factory = lambda t: CudaEngine(t, level, max_character, dev, mode)  # noqa: E731
        self.engine = engine_factory(shard)
        self.shard_len = int(shard.size)

    def len(self):
        return self.len_total

    def pieces_count(self):
        return self.pieces_total

    def search_locate(self, patterns, mode: int = 0, dst: int = 0):
        """Every rank passes the SAME patterns ([npat, m] uint8).  Rank `dst` returns
        (counts[npat], hit_off[npat+1], positions, piece_ids): global text positions and global piece ids in CSR
        form, hits of one pattern ordered rank-major; every other rank returns None."""
        pats = torch.as_tensor(patterns, dtype=torch.uint8)
        if bool((pats == 0).any()):
            raise ValueError("patterns containing \\0 are not supported by the piece-partitioned index")
        off, pos, pid = self.engine.search_locate(pats, mode)
        dev = off.device
        npat = pats.shape[0]
        merge = merge_parts_cuda if dev.type == "cuda" else merge_parts_host
        if self.world == 1:
            hit_off, positions, pieces = merge([(off, pos, pid)], self.bases, npat, dev)
            return hit_off[1:] - hit_off[:-1], hit_off, positions, pieces
        # ---- exchange 1: the hit offsets of every rank, gathered on dst (fixed size)
        root = self.rank == dst
        g_off = [torch.empty_like(off) for _ in range(self.world)] if root else None
        dist.gather(off, g_off, dst=dst, group=self.group)
        # ---- exchange 2: the hit lists themselves, exact sizes, point to point
        if not root:
            ops = []
            if pos.numel():
                ops = [dist.P2POp(dist.isend, pos.contiguous(), dst, self.group), dist.P2POp(dist.isend, pid.contiguous(), dst, self.group)]
            for w in (dist.batch_isend_irecv(ops) if ops else []):
                w.wait()
            return None
        sizes = [int(v) & 0xFFFFFFFF for v in torch.stack([o[-1] for o in g_off]).tolist()]   # one host read
        parts, ops = [], []
        for r in range(self.world):
            if r == self.rank:
                parts.append((off, pos, pid))
                continue
            rp = torch.empty(sizes[r], dtype=pos.dtype, device=dev)
            rd = torch.empty(sizes[r], dtype=pid.dtype, device=dev)
            if sizes[r]:
                ops += [dist.P2POp(dist.irecv, rp, r, self.group), dist.P2POp(dist.irecv, rd, r, self.group)]
            parts.append((g_off[r], rp, rd))
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
        hit_off, positions, pieces = merge(parts, self.bases, npat, dev)
        return hit_off[1:] - hit_off[:-1], hit_off, positions, pieces
