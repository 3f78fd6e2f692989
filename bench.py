#!/usr/bin/env python
"""bench.py -- count + locate throughput of the B200 FM-index query engine.

One "step" = one pass of the hot path (backward search of every pattern, then locate of every
match) over one batch of synthetic patterns.  Default workload = the north-star target of
BASELINE.json: FMIndexWithLocate, sampling level 2, 100M 32-mers (50 % sampled from the text, 50 %
uniform random) over 1 GB synthetic DNA.  Prints ONE JSON line (see the contract in the task description).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
                    [--npat P] [--by-piece]

Other workloads (the remaining BASELINE configs and the north-star target) are selected with
--workload; `--by-piece` runs the MultiPieces workload piece-partitioned with the NCCL gather.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FM, RLFM, MULTI = 0, 1, 2
# name: dict(kind, n, npat, m (0 = ragged 8..64), sigma, mc, level, desc)
WORKLOADS = {
    "cfg2_dna100m": dict(kind=FM, n=100_000_000, npat=1_000_000, m=32, sigma=4, mc=4, level=2,
                         desc="BASELINE configs[1]: FMIndexWithLocate level 2, 1M 32-mers over 100 MB synthetic DNA"),
    "target_dna1g": dict(kind=FM, n=1_000_000_000, npat=100_000_000, m=32, sigma=4, mc=4, level=2,
                         desc="north-star target: 100M 32-mers over 1 GB synthetic DNA, level 2"),
    "cfg1_dna1m": dict(kind=FM, n=1_000_000, npat=10_000, m=20, sigma=4, mc=4, level=2,
                       desc="BASELINE configs[0] shape (CPU-runnable): 10k 20-mers over 1 MB DNA"),
    "dna16m": dict(kind=FM, n=16_000_000, npat=1_000_000, m=32, sigma=4, mc=4, level=2, desc="small smoke workload"),
    "cfg3_rlfm": dict(kind=RLFM, n=64 * (16 << 20), npat=10_000_000, m=32, sigma=4, mc=4, level=2, copies=64,
                      desc="BASELINE configs[2]: RLFMIndexWithLocate, 64 mutated copies (0.1% SNPs) of a 16 Mi haplotype, "
                           "10M 32-mers sampled from the text"),
    "cfg3_rlfm_256m": dict(kind=RLFM, n=16 * (16 << 20), npat=2_000_000, m=32, sigma=4, mc=4, level=2, copies=16,
                           desc="BASELINE configs[2] scaled: 16 mutated copies of a 16 Mi haplotype, 2M 32-mers"),
    "cfg4_multi": dict(kind=MULTI, n=3_000_000_000, npat=10_000_000, m=32, sigma=4, mc=4, level=2, pieces=24,
                       desc="BASELINE configs[3]: FMIndexMultiPiecesWithLocate over 24 chromosome-like pieces, 3 GB"),
    "cfg4_multi_480m": dict(kind=MULTI, n=480_000_000, npat=4_000_000, m=32, sigma=4, mc=4, level=2, pieces=24,
                            desc="BASELINE configs[3] scaled: 24 chromosome-like pieces, 480 MB total"),
    "cfg5_bytes1g": dict(kind=FM, n=1_000_000_000, npat=100_000_000, m=0, sigma=255, mc=255, level=2,
                         desc="BASELINE configs[4]: byte alphabet, FMIndexWithLocate over 1 GB, 100M patterns of length 8-64"),
}
CHROM = [248, 242, 198, 190, 181, 171, 159, 145, 138, 133, 135, 133, 114, 107, 102, 90, 83, 80, 58, 64, 46, 50, 156, 57]


def _s64(c: int) -> int:
    return c - (1 << 64) if c >= (1 << 63) else c


def _lsr(x, k: int):
    """logical right shift of an int64 torch tensor"""
    return (x >> k) & ((1 << (64 - k)) - 1)


def mix64(x):
    """splitmix64 finaliser on int64 torch tensors (wrapping arithmetic): a self-contained,
    reproducible generator that runs identically on CPU and GPU."""
    x = x + _s64(0x9E3779B97F4A7C15)
    x = (x ^ _lsr(x, 30)) * _s64(0xBF58476D1CE4E5B9)
    x = (x ^ _lsr(x, 27)) * _s64(0x94D049BB133111EB)
    return x ^ _lsr(x, 31)


def gen_symbols(lo: int, hi: int, sigma: int, seed: int, device):
    import torch

    r = mix64(torch.arange(lo, hi, dtype=torch.int64, device=device) + seed * 0x1000000000)
    return (_lsr(r, 33) % sigma + 1).to(torch.uint8)


def gen_text(n: int, sigma: int, seed: int, device="cpu"):
    """n symbols uniform over 1..=sigma followed by the \\0 terminator (torch uint8 tensor on `device`)."""
    import torch

    out = torch.empty(n + 1, dtype=torch.uint8, device=device)
    chunk = 1 << 26
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        out[lo:hi] = gen_symbols(lo, hi, sigma, seed, device)
    out[n] = 0
    return out


def gen_text_for(w: dict, device="cpu"):
    """the text of a workload (torch uint8 on `device`)"""
    import torch

    n, sigma = w["n"], w["sigma"]
    if w["kind"] == RLFM:  # mutated copies of one haplotype (config 3): seeds 5 (base), 6.. (copies)
        copies = w["copies"]
        base_len = n // copies
        base = gen_symbols(0, base_len, sigma, 5, device)
        out = torch.empty(n + 1, dtype=torch.uint8, device=device)
        for c in range(copies):
            r = mix64(torch.arange(0, base_len, dtype=torch.int64, device=device) + (6 + c) * 0x1000000000)
            mut = (_lsr(r, 20) % 1000) == 0                       # 0.1 % substitutions
            sub = (_lsr(r, 40) % sigma + 1).to(torch.uint8)
            out[c * base_len:(c + 1) * base_len] = torch.where(mut, sub, base)
        out[n] = 0
        return out
    if w["kind"] == MULTI:  # chromosome-like pieces, each followed by \0 (config 4)
        total = sum(CHROM[: w["pieces"]])
        lens = [max(1000, int(n * c / total)) for c in CHROM[: w["pieces"]]]
        out = torch.empty(sum(lens) + len(lens), dtype=torch.uint8, device=device)
        pos = 0
        for k, ln in enumerate(lens):
            out[pos:pos + ln] = gen_symbols(0, ln, sigma, 100 + k, device)
            out[pos + ln] = 0
            pos += ln + 1
        return out
    return gen_text(n, sigma, 3, device)


def gen_patterns(text, npat: int, m: int, sigma: int, seed: int, all_sampled=False):
    """even patterns: substrings at uniform random offsets (>= 1 hit); odd: uniform random.
    `text` is a torch uint8 tensor; returns (patterns [npat, m] uint8, starts [npat] int64) on its device."""
    import torch

    device = text.device
    n = text.numel() - 1
    pats = torch.empty((npat, m), dtype=torch.uint8, device=device)
    starts = torch.empty(npat, dtype=torch.int64, device=device)
    cols = torch.arange(m, dtype=torch.int64, device=device)[None, :]
    chunk = 1 << 21
    for lo in range(0, npat, chunk):
        hi = min(npat, lo + chunk)
        k = torch.arange(lo, hi, dtype=torch.int64, device=device)
        st = _lsr(mix64(k + seed * 0x1000000000), 1) % (n - m)
        samp = text[st[:, None] + cols]
        if all_sampled:
            pats[lo:hi] = samp
        else:
            rnd = mix64(k[:, None] * 64 + cols + (seed + 1) * 0x1000000000)
            blk = (_lsr(rnd, 33) % sigma + 1).to(torch.uint8)
            pats[lo:hi] = torch.where((k % 2 == 0)[:, None], samp, blk)
        starts[lo:hi] = st
    return pats, starts


def gen_ragged_patterns(text, npat: int, sigma: int, seed: int, lo_len=8, hi_len=64):
    """lengths uniform in lo_len..=hi_len; even patterns sampled from the text, odd uniform random.
    returns (flat uint8, offsets int64[npat+1])"""
    import torch

    device = text.device
    n = text.numel() - 1
    k = torch.arange(npat, dtype=torch.int64, device=device)
    lens = _lsr(mix64(k + (seed + 7) * 0x1000000000), 8) % (hi_len - lo_len + 1) + lo_len
    off = torch.zeros(npat + 1, dtype=torch.int64, device=device)
    off[1:] = torch.cumsum(lens, dim=0)
    total = int(off[-1].item())
    flat = torch.empty(total, dtype=torch.uint8, device=device)
    st = _lsr(mix64(k + seed * 0x1000000000), 1) % (n - hi_len)
    chunk = 1 << 20
    for a in range(0, npat, chunk):
        b = min(npat, a + chunk)
        lo_b, hi_b = int(off[a].item()), int(off[b].item())
        owner = torch.repeat_interleave(torch.arange(a, b, device=device), lens[a:b])
        within = torch.arange(lo_b, hi_b, device=device) - off[owner]
        samp = text[st[owner] + within]
        rnd = mix64(torch.arange(lo_b, hi_b, dtype=torch.int64, device=device) + (seed + 1) * 0x1000000000)
        blk = (_lsr(rnd, 33) % sigma + 1).to(torch.uint8)
        flat[lo_b:hi_b] = torch.where(owner % 2 == 0, samp, blk)
    return flat, off


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons while the measured section runs (NVML)."""

    def __init__(self, dev_index: int):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = dev_index
            if vis:
                try:
                    phys = int(vis.split(",")[dev_index])
                except Exception:
                    phys = dev_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)),
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, v in names.items():
                    if bits & v:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        if self.ok:
            self.join(timeout=2)
        return {
            "sm_mhz": float(np.median(self.samples)) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def source_hash():
    """sha256 over the kernel and API sources: ties an ncu capture to the code it measured"""
    import hashlib

    h = hashlib.sha256()
    d = os.path.join(ROOT, "fm-index_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh", ".h", ".hpp", ".cpp")):
            h.update(name.encode())
            h.update(open(os.path.join(d, name), "rb").read())
    return h.hexdigest()[:16]


def ncu_capture(workload: str, npat: int, mode: str):
    """Per-step DRAM bytes and L2 read requests measured under ncu (profiles/ncu_traffic.json), used ONLY when the
    capture was taken from exactly these sources, this workload, this batch size and this index mode.
    Nothing is scaled: a stale or mismatching capture gives None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[f"{workload}:{mode}"]
        if t.get("source_hash") != source_hash() or int(t.get("npat", -1)) != int(npat):
            return None
        return t
    except Exception:
        return None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def _gpu_numa_node(torch, dev: int):
    try:
        pr = torch.cuda.get_device_properties(dev)
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/numa_node"
        return int(open(path).read().strip())
    except Exception:
        return -1


def _node_cores(node: int):
    out = []
    for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


def pin_rank_to_numa(local: int, world: int):
    """One rank per GPU: keep the rank's threads (and the pinned buffers it first-touches) on the cores of its GPU's
    NUMA node when the box says which that is; the ranks that share a node split its cores evenly.
    -> number of cores this rank may use (None: affinity left alone)."""
    try:
        import torch

        cores = sorted(os.sched_getaffinity(0))
        nodes = [_gpu_numa_node(torch, d) for d in range(world)]
        mine = nodes[local]
        pool = cores
        if mine >= 0:
            nc = [c for c in _node_cores(mine) if c in cores]
            if nc:
                pool = nc
        peers = [d for d in range(world) if nodes[d] == mine] if mine >= 0 else list(range(world))
        share = max(1, len(pool) // max(1, len(peers)))
        k = peers.index(local)
        sel = pool[k * share:(k + 1) * share] or pool
        os.sched_setaffinity(0, sel)
        return len(sel)
    except Exception:
        return None


def run_cpu_path(oracle_index, flat, off, nthreads):
    """the restated CPU path (oracle, OpenMP over patterns): returns (seconds, s, e, hit_off, pos)"""
    t0 = time.perf_counter()
    s, e = oracle_index.search_batch(flat, off, nthreads=nthreads)
    hit_off, pos, _ = oracle_index.locate_batch(s, e, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return dt, s, e, hit_off, pos


def host_patterns(w, text_t, npat, seed):
    """(flat uint8 ndarray, offsets uint64 ndarray) of the first npat patterns of the workload, on the host"""
    if w["m"]:
        p = gen_patterns(text_t, npat, w["m"], w["sigma"], seed, all_sampled=w["kind"] == RLFM)[0].cpu().numpy()
        return p.reshape(-1), np.arange(npat + 1, dtype=np.uint64) * np.uint64(w["m"])
    flat, off = gen_ragged_patterns(text_t, npat, w["sigma"], seed)
    return flat.cpu().numpy(), off.cpu().numpy().astype(np.uint64)


def config_of(name, w, text_len, npat, extra=None):
    cfg = {"workload": name, "description": w["desc"], "text_len": text_len, "patterns_per_step_per_gpu": npat,
           "pattern_len": w["m"] if w["m"] else "8..64 (ragged)",
           "pattern_mix": "all sampled from the text" if w["kind"] == RLFM else "50% sampled from text / 50% uniform random",
           "index": ["FMIndexWithLocate", "RLFMIndexWithLocate", "FMIndexMultiPiecesWithLocate"][w["kind"]],
           "sampling_level": w["level"], "wavelet_levels": int(w["mc"]).bit_length()}
    if extra:
        cfg.update(extra)
    return cfg


def oracle_for(w, text, device=None):
    """The oracle's index of the workload's text.  GB-scale texts: its single-threaded SA-IS takes ~5 min per GB,
    so -- when a GPU is there -- it is handed the GPU-built suffix array, which it VERIFIES with its own
    linear-time checker (orc_check_suffix_array) before building everything else itself.  Construction is not the
    measured path."""
    from oracle import oracle as orc

    sa_src = "oracle SA-IS"
    if text.size >= (1 << 27) and device is not None:
        import fmx_pkg

        fmx = fmx_pkg.load()
        sa, _ = fmx.suffix_array_device(text, w["mc"], device)
        oi = orc.OracleIndex(text, w["kind"], level=w["level"], max_character=w["mc"], sa=sa)
        del sa
        sa_src = "GPU-built SA, verified by the oracle's linear-time checker"
    else:
        oi = orc.OracleIndex(text, w["kind"], level=w["level"], max_character=w["mc"])
    return oi, sa_src


def reference_arm(args, name, w):
    """--impl reference: the reference's CPU implementation of the path.  The crate cannot be built
    here (no Rust toolchain, vers-vecs un-vendored), so this is the oracle port of it, run on all
    host threads over a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nthreads = host_threads()
    text_t = gen_text_for(w)
    npat = args.npat or w["npat"]
    sample = min(npat, args.cpu_sample)
    flat, off = host_patterns(w, text_t, sample, 4)
    text = text_t.numpy()
    t0 = time.perf_counter()
    device = None
    try:
        import torch

        if torch.cuda.is_available():
            device = 0
    except Exception:
        device = None
    if args.oracle_own_sa:
        device = None
    oracle_index, sa_src = oracle_for(w, text, device)
    build_s = time.perf_counter() - t0
    times, hits = [], 0
    for it in range(args.warmup + args.steps):
        dt, s, e, hit_off, pos = run_cpu_path(oracle_index, flat, off, nthreads)
        hits = int(hit_off[-1])
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": "count+locate queries/s", "value": value, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic", "config": config_of(name, w, int(text.size), sample),
        "located_hits_per_s": hits / (ms * 1e-3),
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": nthreads, "kind": "port",
                         "sample": f"first {sample} patterns of the workload, count+locate, {args.steps} passes",
                         "cpu_model": cpu_model(), "index_build_s": round(build_s, 1), "index_suffix_array": sa_src},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


PHASES = ["query kernel (rich: k_query_fused[_defer] = table lookup + rank steps + verify in one launch; compact: k_search; "
          "option search_phased=1: k_ph_seed)", "steps (k_ph_steps; search_phased=1 only)", "verify (k_ph_verify; search_phased=1 only)",
          "steps, second pass (+ widen)", "counts + offsets scan (rich, emit_fused: k_offsets_emit also writes the positions of ranges of <= 4 rows)",
          "emit / locate (rich, emit_fused: k_emit_big only)"]


class DeviceRun:
    """One index + one device-resident batch: the timed step and its variants, through fmx_query_batch_device."""

    def __init__(self, fmx, L, index, d_pat, d_off, m, npat, stream):
        import torch

        from fm_index_b200 import _lib

        self.torch, self.L, self.index, self.h = torch, L, index, index._h
        self.d_pat, self.d_off, self.m, self.npat, self.stream = d_pat, d_off, m, npat, stream
        self.sp = C.c_void_p(stream.cuda_stream)
        self.Q = _lib.Query
        self.d_cnt = torch.empty(npat, dtype=torch.int64, device="cuda")
        self.d_hoff = torch.empty(npat + 1, dtype=torch.int64, device="cuda")
        self.d_pos = torch.empty(1024, dtype=torch.int64, device="cuda")
        self.cap = 1024
        self.hits = 0

    def chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.fmx_last_error().decode())

    def query(self, counts=False, locate=True, rows=None):
        q = self.Q()
        q.mode, q.packed_bits, q.patterns = 0, 0, self.d_pat.data_ptr()
        q.pat_off = self.d_off.data_ptr() if self.d_off is not None else None
        q.fixed_len, q.npat, q.out_width = self.m, self.npat, 8
        if rows is not None:
            q.out_s, q.out_e = rows[0].data_ptr(), rows[1].data_ptr()
        if counts:
            q.counts = self.d_cnt.data_ptr()
        if locate:
            q.hit_off, q.positions, q.capacity = self.d_hoff.data_ptr(), self.d_pos.data_ptr(), self.cap
        self.chk(self.L.fmx_query_batch_device(self.h, C.byref(q), self.sp))

    def size_outputs(self):
        """first pass: learn the hit count, size the position buffer, run once more with room for every hit"""
        torch = self.torch
        self.query()
        self.chk(self.L.fmx_search_check(self.h, self.sp))
        self.hits = int(self.d_hoff[-1].item())
        self.cap = self.hits + 1024
        self.d_pos = torch.empty(self.cap, dtype=torch.int64, device="cuda")
        self.query()
        torch.cuda.synchronize()

    def timed(self, fn, steps, flush, graph=True):
        """-> (per-step ms list, launches per step, replayed as a graph?)"""
        torch, L, stream = self.torch, self.L, self.stream
        g, per = None, None
        fn()
        torch.cuda.synchronize()
        if graph:
            try:
                l0 = L.fmx_launch_count()
                gg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gg, stream=stream):
                    fn()
                per = L.fmx_launch_count() - l0
                gg.replay()
                torch.cuda.synchronize()
                g = gg
            except Exception as ex:  # pragma: no cover
                print(f"CUDA graph capture failed ({ex}); timing plain launches", file=sys.stderr)
                torch.cuda.synchronize()
            torch.cuda.set_stream(stream)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(steps)]
        l0 = L.fmx_launch_count()
        for k in range(steps):
            flush.zero_()  # L2 flush between timed iterations
            ev[k][0].record(stream)
            if g is not None:
                g.replay()
            else:
                fn()
            ev[k][1].record(stream)
        torch.cuda.synchronize()
        launches = (L.fmx_launch_count() - l0) / steps if g is None else per
        return [ev[k][0].elapsed_time(ev[k][1]) for k in range(steps)], int(launches), g is not None

    def phases(self, fn, flush, passes=3):
        """live per-phase durations: CUDA events the library records between its own launches (option phase_timing)"""
        acc = np.zeros(len(PHASES))
        self.index.set_option("phase_timing", 1)
        buf = (C.c_float * len(PHASES))()
        for _ in range(passes):
            flush.zero_()
            fn()
            self.chk(self.L.fmx_last_phase_ms(self.h, self.sp, buf, len(PHASES)))
            acc += np.array(list(buf))
        self.index.set_option("phase_timing", 0)
        return acc / passes

    def work(self, fn):
        """one counting pass: executed search iterations, LF steps, index requests issued"""
        self.index.set_option("count_work", 1)
        fn()
        a, b = self.index.last_work(self.sp)
        r0, r1 = C.c_uint64(0), C.c_uint64(0)
        self.chk(self.L.fmx_last_requests(self.h, self.sp, C.byref(r0), C.byref(r1)))
        self.index.set_option("count_work", 0)
        return a, b, r0.value, r1.value


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="target_dna1g", choices=sorted(WORKLOADS))
    ap.add_argument("--npat", type=int, default=0, help="override the workload's patterns per step")
    ap.add_argument("--cpu-sample", type=int, default=200_000, help="patterns in the CPU baseline sample")
    ap.add_argument("--by-piece", action="store_true", help="MultiPieces workloads: partition the pieces over the GPUs")
    ap.add_argument("--mode", default="auto", choices=["auto", "rich", "compact"], help="index mode of the headline run")
    ap.add_argument("--option", action="append", default=[], help="key=value tuning option (fmx_index_set_option), for A/B runs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gather-peak", action="store_true")
    ap.add_argument("--no-compact", action="store_true", help="skip the compact-mode index measured beside the headline one")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extract", action="store_true", help="skip the extraction (iter_chars_*) measurement")
    ap.add_argument("--extract-rows", type=int, default=4_000_000)
    ap.add_argument("--no-by-piece", action="store_true", help="N > 1 default run: skip the config-4 piece-partitioned measurement")
    ap.add_argument("--oracle-own-sa", action="store_true",
                    help="CPU baseline: build the suffix array with the oracle's own SA-IS even for GB-scale texts")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return reference_arm(args, args.workload, w)

    import torch
    import torch.distributed as dist

    import fmx_pkg

    fmx = fmx_pkg.load()
    L = fmx.load_library()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.by_piece:
        return by_piece_bench(args, w, fmx, rank, world, local)
    ncores = pin_rank_to_numa(local, world)
    if ncores:
        os.environ["OMP_NUM_THREADS"] = str(ncores)   # the host builder's OpenMP passes (torchrun forces 1)

    npat = args.npat or w["npat"]
    kind, mc, level, m = w["kind"], w["mc"], w["level"], w["m"]
    # synthetic text and patterns are generated on the GPU (same integer arithmetic as on the CPU)
    d_text = gen_text_for(w, device="cuda")
    # index replicated on every GPU; each rank answers its own batch (weak scaling, no collective
    # on the query path)
    seed = 4 + 1000 * rank
    if m:
        d_pat, _ = gen_patterns(d_text, npat, m, w["sigma"], seed, all_sampled=kind == RLFM)
        d_off = None
    else:
        d_pat, d_off = gen_ragged_patterns(d_text, npat, w["sigma"], seed)
    text = d_text.cpu().numpy()
    del d_text
    torch.cuda.empty_cache()
    cls = [fmx.FMIndexWithLocate, fmx.RLFMIndexWithLocate, fmx.FMIndexMultiPiecesWithLocate][kind]
    mode_id = {"auto": fmx.MODE_AUTO, "rich": fmx.MODE_RICH, "compact": fmx.MODE_COMPACT}[args.mode]
    t0 = time.perf_counter()
    index = cls.new(fmx.Text.with_max_character(text, mc), level, device=local, mode=mode_id)
    build_s = time.perf_counter() - t0
    for kv in args.option:
        k_, v_ = kv.split("=")
        index.set_option(k_, int(v_))
    mode_name = {fmx.MODE_RICH: "rich", fmx.MODE_COMPACT: "compact"}[index.mode()]
    # a real (non-default) stream: the library launches on the stream it is handed, and the CUDA
    # events below must sit on that same stream
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    run = DeviceRun(fmx, L, index, d_pat, d_off, m, npat, stream)
    run.size_outputs()
    hits = run.hits

    def step():
        # the timed step: count + locate of every pattern, one asynchronous call (the hit total never leaves the device)
        run.query(counts=False, locate=True)

    def count_step():
        run.query(counts=True, locate=False)

    for _ in range(args.warmup):
        flush.zero_()
        step()
    torch.cuda.synchronize()

    # ---- timed region (device-resident inputs)
    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    wall0 = time.perf_counter()
    t_step, launches_per_step, graphed = run.timed(step, args.steps, flush)
    wall = time.perf_counter() - wall0
    t_count, _, _ = run.timed(count_step, args.steps, flush)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total, ms_count_total = float(np.sum(t_step)), float(np.sum(t_count))
    if world > 1:
        tt = torch.tensor([ms_total, ms_count_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, ms_count_total = [float(v) for v in tt.tolist()]
        hh = torch.tensor([hits], device="cuda", dtype=torch.int64)
        dist.all_reduce(hh, op=dist.ReduceOp.SUM)
        hits_all = int(hh.item())
    else:
        hits_all = hits
    ms_per_step = ms_total / args.steps
    ms_count = ms_count_total / args.steps
    value = world * npat / (ms_per_step * 1e-3)

    # ---- end to end through the host-buffer C ABI: pinned host inputs, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e = e2e_measure(args, fmx, L, index, d_pat, d_off, m, npat, hits, mc, world, dist, torch)
    clocks = sampler.stop()

    # ---- N > 1: BASELINE config 4 (MultiPieces, 24 pieces / 3 GB) partitioned by piece over the same ranks, with the NCCL
    # gather -- the one multi-GPU form with an exchange step -- measured in the same driver run
    by_piece = None
    if world > 1 and args.workload == "target_dna1g" and not args.no_by_piece:
        torch.cuda.empty_cache()
        try:
            by_piece = by_piece_measure(args, WORKLOADS["cfg4_multi"], fmx, rank, world, local, with_replicated=False)
        except Exception as ex:  # pragma: no cover
            by_piece = {"error": str(ex)}
        torch.cuda.set_stream(stream)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- what the step is made of, live: per-phase durations, executed work, requests issued
    phase_ms = run.phases(step, flush)
    search_steps, lf_steps, req_search, req_emit = run.work(step)
    dominant = int(np.argmax(phase_ms))
    peak, peak_src = measured_peaks()
    gp = None
    if not args.no_gather_peak:
        try:
            gp = fmx.random_gather_peak(local, nbytes=4 << 30, nloads=1 << 28, iters=3)
        except Exception as ex:  # pragma: no cover
            print(f"random gather peak failed: {ex}", file=sys.stderr)
    roofline = roofline_block(args, w, index, mode_name, npat, hits, ms_per_step, ms_count, phase_ms, dominant, search_steps,
                              lf_steps, req_search, req_emit, peak, peak_src, gp)

    # ---- extraction walks (wrapper.rs:143-183) from random rows of the same index
    extraction = None
    if not args.no_extract and world == 1:
        try:
            extraction = extraction_measure(args, L, index, int(text.size), int(mc).bit_length(), flush, stream, gp)
        except Exception as ex:  # pragma: no cover
            extraction = {"error": str(ex)[:300]}

    # ---- the same batch on a COMPACT index (reference-sized: rank structure + samples + L2-resident table)
    compact = None
    if not args.no_compact and mode_name == "rich" and world == 1:
        try:
            compact = compact_measure(args, fmx, L, cls, text, mc, level, local, d_pat, d_off, m, npat, stream, flush, run, gp)
        except Exception as ex:  # pragma: no cover
            compact = {"error": str(ex)}

    # ---- CPU baseline beside it: the oracle port on all host threads, bounded sample; also the parity check
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_baseline_and_parity(args, w, text, local, run, d_pat, d_off, m, npat, e2e)

    if e2e is not None:
        e2e.pop("_bytes_form_results", None)
    line = {
        "metric": "count+locate queries/s", "value": value, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": config_of(args.workload, w, int(text.size), npat, {
            "index_mode": mode_name, "index_device_bytes": index.heap_size(), "device_layout": index.layout_name(),
            "index_bytes_per_text_symbol": round(index.heap_size() / max(1, int(text.size)), 2),
            "kmer_table_k": [index.kmer_k(False), index.kmer_k(True)],
            "kmer_table_entry_bytes": int(L.fmx_index_kmer_entry_bytes(index._h)),
            "l2": "flushed between timed iterations (512 MiB memset)",
            "step": "fmx_query_batch_device (hit offsets + positions)" + (" replayed as one CUDA graph" if graphed else ""),
            "parallelism": f"index replicated x{world}, query batches sharded", "index_build_s": round(build_s, 1),
            **({"options": args.option} if args.option else {})}),
        "count_queries_per_s": world * npat / (ms_count * 1e-3),
        "located_hits_per_s": hits_all / (ms_per_step * 1e-3),
        "hits_per_step": hits_all,
        "gpu_launches": int(launches_per_step * args.steps),
        "clocks": clocks,
        "roofline": roofline,
        "wall_ms_per_step_incl_flush": wall / args.steps * 1e3,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if extraction is not None:
        line["extraction"] = extraction
    if compact is not None:
        line["compact_mode"] = compact
    if by_piece is not None:
        line["cfg4_by_piece"] = by_piece
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def roofline_block(args, w, index, mode_name, npat, hits, ms_step, ms_count, phase_ms, dominant, search_steps, lf_steps,
                   req_search, req_emit, peak, peak_src, gp):
    """What the hardware did (frac) and, separately, how much less it had to do than the reference's structure
    (algorithmic_gain).

    frac = max(DRAM bytes moved / step time / HBM peak, L2 read requests / step time / measured random-request peak).
    DRAM bytes and L2 requests are per-step sums over the step's kernels from an ncu capture of exactly these
    sources, this workload, batch size and index mode (profiles/ncu_traffic.json, keyed by a source hash; nothing is
    scaled).  When no such capture exists, `traffic` is null and frac falls back to the requests the kernels counted
    themselves (lane-level index loads, an upper bound on their L2 requests), said so in `frac_source`."""
    kind, mc = w["kind"], w["mc"]
    Lw = int(mc).bit_length()
    P = index.sectors_per_rank()
    ref_per_lf2 = Lw if kind != RLFM else 2 * Lw + 3
    dev_per_lf2 = P + (3 if kind == RLFM else 0)
    alg_bytes_search = 32.0 * 2 * ref_per_lf2 * search_steps
    # locate: the reference walks to a sample (mean 2^level - 1 steps of L probes) + the sample itself
    mean_walk = (1 << w["level"]) - 1
    alg_bytes_locate = 32.0 * hits * ((Lw if kind != RLFM else Lw + 3) * (lf_steps / hits if lf_steps and hits else mean_walk) + 1)
    cap = ncu_capture(args.workload, npat, mode_name)
    t = ms_step * 1e-3
    names = ["k_ph_seed", "k_ph_steps", "k_ph_verify", "k_ph_steps (second pass)", "scan (+ k_offsets_emit)", "k_emit_small + k_emit_big / k_locate_*"]
    opts = dict(kv.split("=") for kv in args.option)
    if not req_search or mode_name != "rich" or kind == RLFM:
        names[0] = "k_search"
    elif int(opts.get("search_phased", 2)) == 2:
        # fmx_api.cu search_phased: the block-local second pass runs with the 16-byte table entries (fused_defer = 1) or always (= 2)
        defer = int(opts.get("fused_defer", 1))
        entry16 = int(index._L.fmx_index_kmer_entry_bytes(index._h)) == 16
        names[0] = "k_query_fused_defer" if (defer == 2 or (defer == 1 and entry16)) and not int(opts.get("order_by_length", 0)) else "k_query_fused"
    r = {"bound": "hbm", "kernel": names[dominant], "peak": peak, "unit": "GB/s", "peak_source": peak_src,
         "step_ms": ms_step, "phase_ms": {PHASES[k]: round(float(phase_ms[k]), 4) for k in range(len(PHASES))},
         "phase_ms_note": "CUDA events the library records between its own launches, separate untimed-region passes with the same L2 flush",
         "kernel_share_of_step": float(phase_ms[dominant] / max(1e-9, float(np.sum(phase_ms))))}
    fr_req = None
    if gp:
        r["random_access"] = {"peak_requests_per_s": gp,
                              "how": "independent uniform-random 32 B loads over a 4 GiB buffer, best of 3, measured in this run "
                                     "(every load is one DRAM- and TLB-missing L2 request; profiles/r02_random_access_study.md)"}
    if cap:
        traffic = float(cap["dram_bytes_per_step"])
        reqs = float(cap["l2_read_requests_per_step"])
        r["traffic"] = traffic
        r["achieved"] = traffic / t / 1e9
        fr_bw = r["achieved"] / peak
        r["traffic_source"] = cap.get("source")
        r["l2_read_requests_per_step"] = reqs
        r["requests_per_pattern"] = reqs / npat
        r["dram_bytes_per_pattern"] = traffic / npat
        r["active_lanes_per_instruction"] = cap.get("active_lanes_per_instruction")
        r["ncu_kernels"] = [{k: v for k, v in kd.items() if k in ("kernel", "ns_under_ncu", "l2_read_requests", "active_lanes_per_instruction")}
                            for kd in cap.get("kernels", [])]
        if gp:
            fr_req = reqs / t / gp
            r["random_access"].update({"achieved_requests_per_s": reqs / t, "frac": fr_req})
            issued = float(req_search + req_emit)
            if issued:
                # the part of those requests that is RANDOM (index loads the kernels count themselves; the rest streams
                # patterns, ranges and offsets): the conservative reading of the same fraction
                r["random_access"].update({"index_requests_per_step": issued, "index_requests_per_pattern": issued / npat,
                                           "index_requests_frac": issued / t / gp,
                                           "index_requests_frac_in_query_kernel": float(req_search) / (phase_ms[0] * 1e-3) / gp if phase_ms[0] > 0 else None})
        r["frac"] = max(fr_bw, fr_req or 0.0)
        r["frac_source"] = "max(DRAM bytes / time / HBM peak = %.3f, L2 read requests / time / random-request peak = %s); ncu capture of these sources" % (
            fr_bw, "%.3f" % fr_req if fr_req is not None else "n/a")
    else:
        r["traffic"] = None
        # k_search (compact indexes, RLFM) carries no request counters: the emit kernel's alone would say nothing about the step
        issued = float(req_search + req_emit) if req_search else 0.0
        r["issued_index_requests_per_step"] = issued
        r["requests_per_pattern"] = issued / npat if issued else None
        est_bytes = issued * 64.0   # every index load carries .L2::64B: a miss fills 64 bytes
        r["achieved"] = est_bytes / t / 1e9 if issued else alg_bytes_search / (ms_count * 1e-3) / 1e9
        if issued and gp:
            fr_req = issued / t / gp
            r["random_access"].update({"achieved_requests_per_s": issued / t, "frac": fr_req})
        r["frac"] = max(r["achieved"] / peak, fr_req or 0.0) if issued else None
        r["frac_source"] = ("no ncu capture of these exact sources / batch: requests counted by the kernels themselves (lane-level index "
                            "loads, an upper bound on L2 requests) x 64 B fills" if issued else
                            "no ncu capture and no request counters on this path")
    r["algorithmic_gain"] = {
        "definition": "SURVEY.md 8(d): bytes the REFERENCE's structure would touch for the same answers -- 32 B x 2 range ends x L "
                      "wavelet levels x executed search iterations (RLFM: 2L+3 probes per lf_map2), locate 32 B x (L x walk + 1) per "
                      "hit -- over the step time; far above the HBM peak because the tables, the one-sector layout and the "
                      "suffix-array structures remove most of those probes.  Not a utilisation figure.",
        "reference_sectors_per_lf_map2": ref_per_lf2, "device_sectors_per_lf_map2": dev_per_lf2,
        "executed_search_steps": int(search_steps), "executed_lf_steps": int(lf_steps),
        "search_bytes": alg_bytes_search, "locate_bytes": alg_bytes_locate,
        "GBps": (alg_bytes_search + alg_bytes_locate) / t / 1e9,
        "x_hbm_peak": (alg_bytes_search + alg_bytes_locate) / t / 1e9 / peak}
    return r


def e2e_measure(args, fmx, L, index, d_pat, d_off, m, npat, hits, mc, world, dist, torch):
    """fmx_query_batch with pinned HOST buffers: H2D of the step's patterns and D2H of its results inside the timed
    region.  Two forms: byte patterns with 64-bit outputs (what a caller holding the crate's &[u8] patterns passes)
    and, for alphabets of <= 4 symbols, 2-bit packed patterns with 32-bit outputs (8 B per 32-mer over PCIe)."""
    from fm_index_b200 import _lib

    h = index._h
    out = {}

    def run_form(packed):
        W = 4 if packed else 8
        dt = torch.int32 if packed else torch.int64
        if packed:
            # pack on the GPU (fmx_query's layout: character k of a pattern in bits [2k, 2k+2) of its words, as c - 1)
            wpp = (2 * m + 63) // 64
            shifts = (torch.arange(32, device="cuda", dtype=torch.int64) * 2)
            h_in = torch.empty((npat, wpp), dtype=torch.int64).pin_memory()
            chunk = 1 << 21
            for lo in range(0, npat, chunk):
                hi = min(npat, lo + chunk)
                codes = torch.zeros((hi - lo, wpp * 32), dtype=torch.int64, device="cuda")
                codes[:, :m] = d_pat[lo:hi].to(torch.int64) - 1
                h_in[lo:hi].copy_((codes.view(hi - lo, wpp, 32) << shifts[None, None, :]).sum(dim=2))  # disjoint fields: sum == or
                del codes
        else:
            h_in = torch.empty(d_pat.shape, dtype=torch.uint8).pin_memory()
            h_in.copy_(d_pat)
        h_off_in = None
        if d_off is not None:
            h_off_in = torch.empty(npat + 1, dtype=torch.int64).pin_memory()
            h_off_in.copy_(d_off)
        h_hoff = torch.empty(npat + 1, dtype=dt).pin_memory()
        cap = int(hits) + 1024
        h_pos = torch.empty(cap, dtype=dt).pin_memory()
        q = _lib.Query()
        q.mode, q.packed_bits, q.patterns = 0, 2 if packed else 0, h_in.data_ptr()
        q.pat_off = h_off_in.data_ptr() if h_off_in is not None else None
        q.fixed_len, q.npat, q.out_width = m, npat, W
        q.hit_off, q.positions, q.capacity = h_hoff.data_ptr(), h_pos.data_ptr(), cap
        total = C.c_uint64(0)

        def call():
            rc = L.fmx_query_batch(h, C.byref(q), C.byref(total))
            if rc != 0:
                raise RuntimeError(L.fmx_last_error().decode())
            return int(total.value)

        for _ in range(2):
            call()
        steps = max(3, min(args.steps, 10))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            nh = call()
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / steps
        if world > 1:
            tt = torch.tensor([sec], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            sec = float(tt.item())
        h2d = int(h_in.numel() * h_in.element_size()) + (8 * (npat + 1) if h_off_in is not None else 0)
        d2h = W * (npat + 1) + W * nh
        res = {"value": world * npat / sec, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": sec * 1e3, "pcie_GBps_in": h2d / sec / 1e9, "pcie_GBps_out": d2h / sec / 1e9,
               "_keep": (h_hoff, h_pos)}
        # The floor under this number: the SAME bytes moved by plain pinned copies, both directions at once, on every rank
        # at the same time, no kernels.  What is left between the floor and ms_per_step is the library's to remove; the
        # floor itself is the box (PCIe links, host memory, and -- with N ranks -- whatever the GPUs share on the host side).
        try:
            src_in = h_in.view(torch.uint8).reshape(-1)
            d_in = torch.empty(src_in.numel(), dtype=torch.uint8, device="cuda")
            n_out1, n_out2 = W * (npat + 1), W * nh
            d_o1 = torch.empty(n_out1, dtype=torch.uint8, device="cuda")
            d_o2 = torch.empty(max(1, n_out2), dtype=torch.uint8, device="cuda")
            # separate pinned destinations: h_hoff / h_pos hold the call's results, which the parity check reads later
            h_o1 = torch.empty(n_out1, dtype=torch.uint8).pin_memory()
            h_o2 = torch.empty(max(1, n_out2), dtype=torch.uint8).pin_memory()
            s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

            def copies():
                with torch.cuda.stream(s_in):
                    d_in.copy_(src_in, non_blocking=True)
                with torch.cuda.stream(s_out):
                    h_o1.copy_(d_o1, non_blocking=True)
                    h_o2.copy_(d_o2, non_blocking=True)

            for _ in range(2):
                copies()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                copies()
            torch.cuda.synchronize()
            fsec = (time.perf_counter() - t0) / steps
            if world > 1:
                tt = torch.tensor([fsec], device="cuda", dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                fsec = float(tt.item())
            res["copy_floor"] = {"ms_per_step": fsec * 1e3, "GBps_in": h2d / fsec / 1e9, "GBps_out": d2h / fsec / 1e9,
                                 "frac": fsec / sec,
                                 "what": "the step's H2D and D2H bytes as plain pinned cudaMemcpyAsync on two streams, all ranks at once; "
                                         "frac = floor / e2e step (1.0 = the call costs nothing beyond the copies)"}
            del d_in, d_o1, d_o2, h_o1, h_o2
        except Exception as ex:  # the floor is a diagnostic; the e2e number stands without it
            res["copy_floor"] = {"error": str(ex)[:200]}
        return res

    bytes_form = run_form(False)
    best = bytes_form
    best_api = "fmx_query_batch: byte patterns in (1 B per character), uint64 CSR hit offsets + positions out"
    if m and mc <= 4 and m <= 4096:
        packed_form = run_form(True)
        out["packed_vs_bytes"] = {"bytes_form_queries_per_s": bytes_form["value"], "bytes_form_ms": bytes_form["ms_per_step"],
                                  "bytes_form_h2d": bytes_form["h2d_bytes_per_step"], "bytes_form_d2h": bytes_form["d2h_bytes_per_step"]}
        best = packed_form
        best_api = ("fmx_query_batch: 2-bit packed patterns in (fmx_query.packed_bits = 2: 8 B per 32-mer), uint32 CSR hit offsets + "
                    "positions out")
    keep = bytes_form.pop("_keep")
    if best is not bytes_form:
        best.pop("_keep")
    out.update(best)
    out["api"] = best_api + " (pinned buffers; chunked H2D / kernels / D2H pipeline, one host wait per chunk)"
    out["_bytes_form_results"] = keep
    return out


def extraction_measure(args, L, index, n, Lw, flush, stream, gp, k=32):
    """Match::iter_chars_backward / iter_chars_forward taken k characters deep from random rows (fmx_extract_batch_device).
    Two kernels answer it: k_extract_text (HBM-rich indexes: characters read from the text at SA[row]) and k_extract
    (every index: one LF step per character backward, one FL step = C-array search + select forward).  Both are timed on
    this index when it holds the text; results are compared byte for byte."""
    import torch

    h = index._h
    sp = C.c_void_p(stream.cuda_stream)
    nrows = int(min(args.extract_rows, max(1, n)))
    g = torch.Generator(device="cuda")
    g.manual_seed(4242)
    d_rows = torch.randint(0, n, (nrows,), generator=g, device="cuda", dtype=torch.int64)
    outs = {}

    def run(forward, d_out, d_len):
        rc = L.fmx_extract_batch_device(h, d_rows.data_ptr(), nrows, k, int(forward), d_out.data_ptr(), d_len.data_ptr(), sp)
        if rc != 0:
            raise RuntimeError(L.fmx_last_error().decode())

    def timed(forward, reps=5):
        d_out = torch.empty((nrows, k), dtype=torch.uint8, device="cuda")
        d_len = torch.empty(nrows, dtype=torch.int32, device="cuda")
        run(forward, d_out, d_len)
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            run(forward, d_out, d_len)
            e1.record(stream)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return float(np.mean(ms)), d_out, d_len

    res = {"rows": nrows, "chars_per_row": k,
           "algorithmic_bytes_per_char": {"backward": 32 * Lw, "forward": 32 * Lw,
                                          "note": "SURVEY.md 8(d): one sector per wavelet level per character"}}
    variants = [("walk", 0)]
    has_text = False
    try:
        index.set_option("extract_text", 1)
        has_text = index.mode() == 2 or bool(L.fmx_index_has_text(h))
    except Exception:
        has_text = False
    if has_text:
        variants.append(("text", 1))
    for name, opt in variants:
        index.set_option("extract_text", opt)
        v = {}
        for forward in (False, True):
            ms, d_out, d_len = timed(forward)
            chars = int(d_len.sum().item())
            key = "forward" if forward else "backward"
            v[key] = {"ms": ms, "chars_per_s": chars / (ms * 1e-3), "rows_per_s": nrows / (ms * 1e-3)}
            outs[(name, forward)] = (d_out, d_len)
        res["k_extract_text (text at SA[row])" if name == "text" else "k_extract (LF / FL steps)"] = v
    index.set_option("extract_text", 1)
    if has_text:
        res["same_characters_both_kernels"] = bool(all(
            torch.equal(outs[("walk", f)][0], outs[("text", f)][0]) and torch.equal(outs[("walk", f)][1], outs[("text", f)][1])
            for f in (False, True)))
        # hardware view of the text kernel: one suffix-array request + the text sectors of k characters per row
        ms = res["k_extract_text (text at SA[row])"]["backward"]["ms"]
        req = nrows * (1 + (k + 31 + 31) // 32)
        res["text_kernel_requests_per_s_upper_bound"] = req / (ms * 1e-3)
        if gp:
            res["text_kernel_frac_of_random_request_peak"] = req / (ms * 1e-3) / gp
    return res


def compact_measure(args, fmx, L, cls, text, mc, level, local, d_pat, d_off, m, npat, stream, flush, rich_run, gp):
    """the same batch, same call, on an index built with FMX_MODE_COMPACT"""
    import torch

    t0 = time.perf_counter()
    cidx = cls.new(fmx.Text.with_max_character(text, mc), level, device=local, mode=fmx.MODE_COMPACT)
    build_s = time.perf_counter() - t0
    run = DeviceRun(fmx, L, cidx, d_pat, d_off, m, npat, stream)
    run.size_outputs()
    same = bool(run.hits == rich_run.hits and torch.equal(run.d_hoff, rich_run.d_hoff) and
                torch.equal(run.d_pos[: run.hits], rich_run.d_pos[: run.hits]))
    steps = max(3, min(args.steps, 10))

    def step():
        run.query(counts=False, locate=True)

    def count_step():
        run.query(counts=True, locate=False)

    t_step, launches, _ = run.timed(step, steps, flush)
    t_count, _, _ = run.timed(count_step, steps, flush)
    ms, msc = float(np.mean(t_step)), float(np.mean(t_count))
    search_steps, lf_steps, _, _ = run.work(step)
    out = {"value": npat / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms, "count_queries_per_s": npat / (msc * 1e-3),
           "located_hits_per_s": run.hits / (ms * 1e-3), "index_device_bytes": cidx.heap_size(),
           "index_bytes_per_text_symbol": round(cidx.heap_size() / max(1, int(text.size)), 2),
           "kmer_table_k": [cidx.kmer_k(False), cidx.kmer_k(True)], "index_build_s": round(build_s, 1),
           "executed_search_steps": int(search_steps), "executed_lf_steps": int(lf_steps),
           "same_results_as_rich_index": same,
           "what": "FMX_MODE_COMPACT: rank structure + the caller's SA samples + an L2-resident k-mer table; k_search + LF-walk locate"}
    cap = ncu_capture(args.workload, npat, "compact")
    if cap:
        out["traffic"] = cap["dram_bytes_per_step"]
        out["l2_read_requests_per_step"] = cap["l2_read_requests_per_step"]
        out["dram_frac_of_hbm_peak"] = cap["dram_bytes_per_step"] / (ms * 1e-3) / 1e9 / measured_peaks()[0]
        if gp:
            out["request_frac_of_random_peak"] = cap["l2_read_requests_per_step"] / (ms * 1e-3) / gp
    del run, cidx
    torch.cuda.empty_cache()
    return out


def cpu_baseline_and_parity(args, w, text, local, run, d_pat, d_off, m, npat, e2e):
    import torch

    kind = w["kind"]
    nthreads = host_threads()
    sample = min(npat, args.cpu_sample)
    tb = time.perf_counter()
    oracle_index, sa_src = oracle_for(w, text, None if args.oracle_own_sa else local)
    obuild = time.perf_counter() - tb
    if m:
        s_flat = d_pat[:sample].cpu().numpy().reshape(-1)
        s_off = np.arange(sample + 1, dtype=np.uint64) * np.uint64(m)
    else:
        s_off = d_off[: sample + 1].cpu().numpy().astype(np.uint64)
        s_flat = d_pat[: int(s_off[-1])].cpu().numpy()
    best = None
    for _ in range(2):
        dt, s, e, ohoff, opos = run_cpu_path(oracle_index, s_flat, s_off, nthreads)
        best = dt if best is None else min(best, dt)
    # the device step's results for the sample, plus the exact SA ranges from a rows-wanted query
    d_s = torch.empty(npat, dtype=torch.int64, device="cuda")
    d_e = torch.empty(npat, dtype=torch.int64, device="cuda")
    run.query(counts=True, locate=True, rows=(d_s, d_e))
    torch.cuda.synchronize()
    g_s = d_s[:sample].cpu().numpy().view(np.uint64)
    g_e = d_e[:sample].cpu().numpy().view(np.uint64)
    g_cnt = run.d_cnt[:sample].cpu().numpy().view(np.uint64)
    g_hoff = run.d_hoff[: sample + 1].cpu().numpy().view(np.uint64)
    g_pos = run.d_pos[: int(g_hoff[-1])].cpu().numpy().view(np.uint64)
    parity = bool(np.array_equal(g_s, s) and np.array_equal(g_e, e) and np.array_equal(g_hoff, ohoff) and
                  np.array_equal(g_pos, opos) and np.array_equal(g_cnt, np.where(e > s, e - s, 0)))
    e2e_parity = None
    if e2e is not None and "_bytes_form_results" in e2e:
        h_hoff, h_pos = e2e.pop("_bytes_form_results")
        e_hoff = h_hoff[: sample + 1].numpy().view(np.uint64)
        e_pos = h_pos[: int(e_hoff[-1])].numpy().view(np.uint64)
        e2e_parity = bool(np.array_equal(e_hoff, ohoff) and np.array_equal(e_pos, opos))
    cpu = {"value": sample / best, "unit": "queries/s", "cores": nthreads, "kind": "port",
           "sample": f"first {sample} patterns of the workload, count+locate, best of 2",
           "located_hits_per_s": int(ohoff[-1]) / best, "cpu_model": cpu_model(),
           "index_build_s": round(obuild, 1), "index_suffix_array": sa_src,
           "gpu_matches_oracle_on_sample": parity, "e2e_call_matches_oracle_on_sample": e2e_parity}
    if not parity or e2e_parity is False:
        print(f"PARITY FAILURE: results differ from the oracle on the sample (device call: {parity}, e2e call: {e2e_parity})", file=sys.stderr)
    return cpu


def by_piece_measure(args, w, fmx, rank, world, local, npat=None, with_replicated=True):
    """MultiPieces partitioned by piece (BASELINE config 4): every rank indexes its pieces and answers EVERY pattern;
    the parts are gathered on rank 0 over NCCL (offsets: gather; hit lists: exact-size send / recv) and merged there
    by fmx_csr_merge_device (fm-index_b200/partitioned.py).  Strong scaling: the batch is the same on every rank.
    Beside it, on rank 0: the replicated (whole-text) index answering the same batch, and 1/world of it (what each
    rank of a query-sharded run would do) -- which form wins is stated in the result.  -> dict on rank 0, else None"""
    import torch
    import torch.distributed as dist

    from fm_index_b200 import partitioned as part

    if w["kind"] != MULTI:
        raise SystemExit("--by-piece needs a MultiPieces workload (cfg4_multi*)")
    npat = npat or args.npat or w["npat"]
    steps = max(3, min(args.steps, 10))
    d_text = gen_text_for(w, device="cuda")
    d_pat, _ = gen_patterns(d_text, npat, w["m"], w["sigma"], 4)          # the SAME batch on every rank
    keep = (d_pat != 0).all(dim=1)                                          # \0 cannot cross the partition
    dropped = int(npat - int(keep.sum().item()))
    d_pat = d_pat[keep].contiguous()
    npat = int(d_pat.shape[0])
    text = d_text.cpu().numpy()
    del d_text
    torch.cuda.empty_cache()
    t0 = time.perf_counter()
    idx, build_err = None, None
    try:
        idx = part.PartitionedMultiPieces(text, w["level"], w["mc"], device=local)
    except Exception as ex:  # pragma: no cover
        build_err = str(ex)
    build_s = time.perf_counter() - t0
    if world > 1:   # nobody enters the exchange unless every rank has its partition's index
        okf = torch.tensor([0 if build_err else 1], device="cuda", dtype=torch.int32)
        dist.all_reduce(okf, op=dist.ReduceOp.MIN)
        if int(okf.item()) == 0:
            return {"error": "index build failed on a rank: " + (build_err or "(another rank)")} if rank == 0 else None
    elif build_err:
        return {"error": build_err}
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    L = fmx.load_library()
    res = None
    for _ in range(3):
        res = idx.search_locate(d_pat)
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = L.fmx_launch_count()
    for k in range(steps):
        ev[k][0].record(stream)
        res = idx.search_locate(d_pat)
        ev[k][1].record(stream)
    torch.cuda.synchronize()
    launches = L.fmx_launch_count() - launches0
    ms_total = float(sum(ev[k][0].elapsed_time(ev[k][1]) for k in range(steps)))
    if world > 1:
        tt = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    ms = ms_total / steps
    if rank != 0:
        return None
    counts, hoff, pos, pid = res
    hits = int(hoff[-1].item())
    # every gathered hit really holds its pattern, in the right piece
    ends = np.flatnonzero(text == 0)
    p_h, o_h = pos.cpu().numpy(), np.repeat(np.arange(npat), np.diff(hoff.cpu().numpy()))
    pats_h = d_pat.cpu().numpy()
    sel = np.random.default_rng(0).integers(0, max(hits, 1), min(hits, 200_000))
    got = text[p_h[sel][:, None] + np.arange(w["m"])[None, :]]
    ok = bool(np.array_equal(got, pats_h[o_h[sel]]) and
              np.array_equal(pid.cpu().numpy()[sel], np.searchsorted(ends, p_h[sel], side="left")))
    out = {"metric": "count+locate queries/s", "value": npat / (ms * 1e-3), "unit": "queries/s", "n_gpus": world,
           "steps": steps, "ms_per_step": ms, "scaling": "strong", "located_hits_per_s": hits / (ms * 1e-3),
           "hits_per_step": hits, "gpu_launches": int(launches), "patterns": npat, "patterns_dropped_for_zero": dropped,
           "partitions": [list(r) for r in idx.ranges], "index_build_s": round(build_s, 1),
           "partition_index_device_bytes": idx.engine.index.heap_size(),
           "exchange": "dist.gather of uint32 hit offsets to rank 0 + exact-size batch_isend_irecv of uint32 positions / piece ids "
                       "(NCCL); merge = fmx_csr_merge_device (one scan + one scatter kernel)",
           "gathered_hits_verified_against_text": ok}
    if with_replicated:
        try:
            del idx
            torch.cuda.empty_cache()
            t0 = time.perf_counter()
            full = fmx.FMIndexMultiPiecesWithLocate.new(fmx.Text.with_max_character(text, w["mc"]), w["level"], device=local)
            fb = time.perf_counter() - t0
            flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
            run = DeviceRun(fmx, L, full, d_pat, None, w["m"], npat, stream)
            run.size_outputs()
            same = bool(run.hits == hits and torch.equal(run.d_hoff, hoff.to(run.d_hoff.dtype)))
            t_all, _, _ = run.timed(lambda: run.query(), steps, flush)
            share = max(1, npat // world)
            run_s = DeviceRun(fmx, L, full, d_pat[:share].contiguous(), None, w["m"], share, stream)
            run_s.size_outputs()
            t_share, _, _ = run_s.timed(lambda: run_s.query(), steps, flush)
            ms_all, ms_share = float(np.mean(t_all)), float(np.mean(t_share))
            out["replicated_index"] = {
                "index_device_bytes": full.heap_size(), "index_build_s": round(fb, 1), "same_counts_as_partitioned": same,
                "one_gpu_whole_batch_ms": ms_all, "one_gpu_whole_batch_queries_per_s": npat / (ms_all * 1e-3),
                "one_rank_share_ms": ms_share,
                "query_sharded_estimate_queries_per_s": npat / (ms_share * 1e-3),
                "note": f"the whole-text index fits one GPU; sharding the QUERIES over {world} GPUs costs each rank the 'share' time "
                        "(no exchange), partitioning the PIECES costs every rank the whole batch plus the gather",
                "winner": "replicated index, queries sharded" if ms_share < ms else "pieces partitioned"}
        except Exception as ex:  # pragma: no cover
            out["replicated_index"] = {"error": str(ex)}
    return out


def by_piece_bench(args, w, fmx, rank, world, local):
    import torch.distributed as dist

    out = by_piece_measure(args, w, fmx, rank, world, local)
    if rank == 0:
        line = {"metric": out["metric"], "value": out["value"], "unit": out["unit"], "n_gpus": world, "steps": out["steps"],
                "warmup": 3, "ms_per_step": out["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "config": config_of(args.workload, w, w["n"], out["patterns"], {"parallelism": f"pieces partitioned over {world} GPUs"}),
                "by_piece": out}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
