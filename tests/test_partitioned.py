"""Piece-partitioned MultiPieces: the N > 1 path, world_size 2 and 3 over gloo on CPU.
The local engine is the oracle (test infrastructure); what is tested is the partitioning, the offset
arithmetic, the gather of the offsets, the exact-size point-to-point transfers of the hit lists and the merge
arithmetic of fm-index_b200/partitioned.py (on GPUs the merge is the CUDA kernel fmx_csr_merge_device, which
tests/test_gpu_parity.py compares with the same host merge)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fmx_pkg
from oracle import oracle as orc

fmx = fmx_pkg.load()
from fm_index_b200 import partitioned as part  # noqa: E402


class OracleEngine:
    def __init__(self, shard, level, mc):
        self.o = orc.OracleIndex(shard, orc.MULTI, level=level, max_character=mc)

    def search_locate(self, pats, mode=0):
        p = pats.numpy()
        npat, m = p.shape
        flat, off = p.reshape(-1), np.arange(npat + 1, dtype=np.uint64) * np.uint64(m)
        s, e = self.o.search_batch(flat, off, mode)
        hoff, pos, pid = self.o.locate_batch(s, e, prefix_only=mode in (1, 3), want_piece_ids=True)
        t = lambda a: torch.from_numpy(a.astype(np.int64))  # noqa: E731
        return t(hoff), t(pos), t(pid)


def make_case(seed=5, pieces=7):
    rng = np.random.default_rng(seed)
    text = np.concatenate([np.append(rng.integers(1, 5, int(l), dtype=np.uint8), np.uint8(0))
                           for l in rng.integers(200, 3000, pieces)])
    starts = rng.integers(0, text.size - 12, 400)
    pats = np.stack([text[s:s + 6] if k & 1 else rng.integers(1, 5, 6, dtype=np.uint8) for k, s in enumerate(starts)])
    pats = pats[(pats != 0).all(axis=1)]
    return text, pats


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        text, pats = make_case()
        idx = part.PartitionedMultiPieces(text, 2, 4, engine_factory=lambda t: OracleEngine(t, 2, 4))
        full = orc.OracleIndex(text, orc.MULTI, level=2, max_character=4)
        ok = idx.pieces_count() == full.pieces_count() and idx.len() == full.len()
        for mode in (0, 1, 2, 3):
            res = idx.search_locate(pats, mode)
            if rank != 0:                       # gathered on rank 0 only: nobody else receives anything
                ok &= res is None
                continue
            counts, hoff, pos, pid = res
            flat, off = pats.reshape(-1), np.arange(pats.shape[0] + 1, dtype=np.uint64) * np.uint64(pats.shape[1])
            s, e = full.search_batch(flat, off, mode)
            rh, rp, rd = full.locate_batch(s, e, prefix_only=mode in (1, 3), want_piece_ids=True)
            ok &= np.array_equal(hoff.numpy(), rh.astype(np.int64))
            for k in range(pats.shape[0]):
                a, b = int(rh[k]), int(rh[k + 1])
                got = sorted(zip(pos[a:b].tolist(), pid[a:b].tolist()))
                exp = sorted(zip(rp[a:b].tolist(), rd[a:b].tolist()))
                ok &= got == exp
            ok &= int(counts.sum()) == int(rh[-1])
        q.put((rank, bool(ok), idx.ranges))
    finally:
        dist.destroy_process_group()


def test_partition_pieces_balanced_and_contiguous():
    lens = [248, 242, 198, 190, 181, 171, 159, 145, 138, 133, 135, 133, 114, 107, 102, 90, 83, 80, 58, 64, 46, 50, 156, 57]
    for world in (1, 2, 3, 4, 8):
        r = part.partition_pieces(lens, world)
        assert r[0][0] == 0 and r[-1][1] == len(lens)
        assert all(r[k][1] == r[k + 1][0] and r[k][1] > r[k][0] for k in range(world - 1))
        loads = [sum(lens[a:b]) for a, b in r]
        assert max(loads) <= 1.6 * sum(lens) / world
    with pytest.raises(ValueError):
        part.partition_pieces([5, 5], 3)


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_multi_pieces_gloo(world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(r[2] == res[0][2] for r in res) and len(res[0][2]) == world


def test_partitioned_rejects_zero_in_pattern_and_single_rank():
    text, pats = make_case(seed=9, pieces=3)
    idx = part.PartitionedMultiPieces(text, 1, 4, engine_factory=lambda t: OracleEngine(t, 1, 4))
    counts, hoff, pos, pid = idx.search_locate(pats)
    full = orc.OracleIndex(text, orc.MULTI, level=1, max_character=4)
    flat, off = pats.reshape(-1), np.arange(pats.shape[0] + 1, dtype=np.uint64) * np.uint64(pats.shape[1])
    s, e = full.search_batch(flat, off)
    rh, rp, rd = full.locate_batch(s, e, want_piece_ids=True)
    assert np.array_equal(pos.numpy(), rp.astype(np.int64)) and np.array_equal(pid.numpy(), rd.astype(np.int64))
    bad = pats.copy()
    bad[3, 2] = 0
    with pytest.raises(ValueError):
        idx.search_locate(bad)
