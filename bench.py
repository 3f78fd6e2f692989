#!/usr/bin/env python
"""bench.py -- count + locate throughput of the B200 FM-index query engine.

One "step" = one pass of the hot path (backward search of every pattern, then locate of every
match) over one batch of synthetic patterns.  Default workload = BASELINE.json configs[1]:
FMIndexWithLocate, sampling level 2, 1M 32-mers (50 % sampled from the text, 50 % uniform random)
over 100 MB synthetic DNA.  Prints ONE JSON line (see the contract in the task description).

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


def ncu_traffic(workload: str, npat: int):
    """DRAM bytes per k_search launch measured under ncu (profiles/ncu_traffic.json), scaled by batch size."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[workload]
        return t["k_search_dram_bytes"] * npat / t["npat"], t
    except Exception:
        return None, None


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


def reference_arm(args, name, w):
    """--impl reference: the reference's CPU implementation of the path.  The crate cannot be built
    here (no Rust toolchain, vers-vecs un-vendored), so this is the oracle port of it, run on all
    host threads over a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as orc

    nthreads = host_threads()
    text_t = gen_text_for(w)
    npat = args.npat or w["npat"]
    sample = min(npat, args.cpu_sample)
    flat, off = host_patterns(w, text_t, sample, 4)
    text = text_t.numpy()
    t0 = time.perf_counter()
    oracle_index = orc.OracleIndex(text, w["kind"], level=w["level"], max_character=w["mc"])
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
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32/u64 integer",
        "data": "synthetic", "config": config_of(name, w, int(text.size), sample),
        "located_hits_per_s": hits / (ms * 1e-3),
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": nthreads, "kind": "port",
                         "sample": f"first {sample} patterns of the workload, count+locate, {args.steps} passes",
                         "cpu_model": cpu_model(), "index_build_s": round(build_s, 1)},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2_dna100m", choices=sorted(WORKLOADS))
    ap.add_argument("--npat", type=int, default=0, help="override the workload's patterns per step")
    ap.add_argument("--cpu-sample", type=int, default=200_000, help="patterns in the CPU baseline sample")
    ap.add_argument("--by-piece", action="store_true", help="MultiPieces workloads: partition the pieces over the GPUs")
    ap.add_argument("--option", action="append", default=[], help="key=value tuning option (fmx_index_set_option), for A/B runs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gather-peak", action="store_true")
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
    t0 = time.perf_counter()
    index = cls.new(fmx.Text.with_max_character(text, mc), level, device=local)
    build_s = time.perf_counter() - t0
    for kv in args.option:
        k_, v_ = kv.split("=")
        index.set_option(k_, int(v_))
    h = index._h
    # a real (non-default) stream: the library launches on the stream it is handed, and the CUDA
    # events below must sit on that same stream
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    assert stream.cuda_stream != 0

    d_s = torch.empty(npat, dtype=torch.int64, device="cuda")
    d_e = torch.empty(npat, dtype=torch.int64, device="cuda")
    d_hoff = torch.empty(npat + 1, dtype=torch.int64, device="cuda")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    p_off = d_off.data_ptr() if d_off is not None else None

    def chk(rc):
        if rc != 0:
            raise RuntimeError(L.fmx_last_error().decode())

    d_pos = [torch.empty(1, dtype=torch.int64, device="cuda")]
    total = C.c_uint64(0)

    def search():
        chk(L.fmx_search_batch_device(h, 0, d_pat.data_ptr(), p_off, m, npat, None, None, d_s.data_ptr(),
                                      d_e.data_ptr(), sp))

    def locate():
        chk(L.fmx_locate_count_device(h, 0, d_s.data_ptr(), d_e.data_ptr(), npat, d_hoff.data_ptr(), C.byref(total), sp))
        if d_pos[0].numel() < total.value:
            d_pos[0] = torch.empty(total.value + 1024, dtype=torch.int64, device="cuda")
        chk(L.fmx_locate_fill_device(h, 0, d_s.data_ptr(), d_e.data_ptr(), npat, d_hoff.data_ptr(), total.value,
                                     d_pos[0].data_ptr(), None, sp))

    # first pass through the two-phase API learns the hit count (sizes the position buffer)
    search()
    locate()
    chk(L.fmx_search_check(h, sp))
    hits = total.value
    cap_dev = int(hits) + 1024
    if d_pos[0].numel() < cap_dev:
        d_pos[0] = torch.empty(cap_dev, dtype=torch.int64, device="cuda")

    def step():
        # the timed step: backward search of every pattern, then locate of every match; both calls
        # are asynchronous (the hit total never leaves the device), so the host only enqueues
        search()
        chk(L.fmx_locate_batch_device(h, 0, d_s.data_ptr(), d_e.data_ptr(), npat, d_hoff.data_ptr(),
                                      d_pos[0].data_ptr(), None, cap_dev, sp))

    for _ in range(args.warmup):
        flush.zero_()
        step()
    search_steps, lf_steps = index.last_work(sp)
    torch.cuda.synchronize()
    # the step is a fixed launch sequence: replay it as a CUDA graph when capture works
    graph, launches_per_step = None, None
    try:
        l0 = L.fmx_launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            step()
        launches_per_step = L.fmx_launch_count() - l0
        g.replay()
        torch.cuda.synchronize()
        graph = g
    except Exception as ex:  # pragma: no cover
        print(f"CUDA graph capture failed ({ex}); timing plain launches", file=sys.stderr)
        torch.cuda.synchronize()
    torch.cuda.set_stream(stream)

    # ---- timed region (device-resident inputs)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    ev_s = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = L.fmx_launch_count()
    sampler.start()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations
        ev[k][0].record(stream)
        if graph is not None:
            graph.replay()
        else:
            step()
        ev[k][1].record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    launches = L.fmx_launch_count() - launches0 if graph is None else launches_per_step * args.steps
    # the dominant kernel on its own (same flush, same stream): its launch duration for the roofline
    for k in range(args.steps):
        flush.zero_()
        ev_s[k][0].record(stream)
        search()
        ev_s[k][1].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_step = np.array([ev[k][0].elapsed_time(ev[k][1]) for k in range(args.steps)])
    t_search = np.array([ev_s[k][0].elapsed_time(ev_s[k][1]) for k in range(args.steps)])
    ms_total = float(t_step.sum())
    if world > 1:
        tt = torch.tensor([ms_total, float(t_search.sum())], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, ms_search_total = [float(v) for v in tt.tolist()]
        hh = torch.tensor([hits], device="cuda", dtype=torch.int64)
        dist.all_reduce(hh, op=dist.ReduceOp.SUM)
        hits_all = int(hh.item())
    else:
        ms_search_total = float(t_search.sum())
        hits_all = hits
    ms_locate_total = max(ms_total - ms_search_total, 0.0)
    ms_per_step = ms_total / args.steps
    value = world * npat / (ms_per_step * 1e-3)

    # ---- end to end through the host-buffer C ABI: pinned host inputs, H2D + D2H inside the timed region
    h_pat = torch.empty(d_pat.shape, dtype=torch.uint8).pin_memory()
    h_pat.copy_(d_pat)
    h_off = None
    if d_off is not None:
        h_off = torch.empty(npat + 1, dtype=torch.int64).pin_memory()
        h_off.copy_(d_off)
    h_s = torch.empty(npat, dtype=torch.int64).pin_memory()
    h_e = torch.empty(npat, dtype=torch.int64).pin_memory()
    h_hoff = torch.empty(npat + 1, dtype=torch.int64).pin_memory()
    cap = int(hits) + 1024
    h_pos = torch.empty(cap, dtype=torch.int64).pin_memory()
    nhits = C.c_uint64(0)

    def e2e_step():
        # the call a user makes: host patterns in, SA ranges + CSR hit lists out (one fused C-ABI call)
        # (SA ranges are internal state of the crate's Search object -- get_range is test-only,
        # wrapper.rs:126-129 -- so the user-visible result is counts = diff(hit_off) and the positions)
        chk(L.fmx_search_locate_batch(h, 0, h_pat.data_ptr(), None if h_off is None else h_off.data_ptr(), m, npat,
                                      None, None, h_hoff.data_ptr(), h_pos.data_ptr(), None, cap, C.byref(nhits)))
        return int(nhits.value)

    for _ in range(2):
        e2e_step()
    e2e_steps = max(3, min(args.steps, 10))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        nh = e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    clocks = sampler.stop()
    if world > 1:
        tt = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    h2d = int(h_pat.numel()) + (8 * (npat + 1) if h_off is not None else 0)
    d2h = 8 * (npat + 1) + 8 * nh

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (k_search).
    # `achieved` follows SURVEY.md 8(d): algorithmic bytes = one 32-byte sector per wavelet-level rank probe of
    # the REFERENCE's structure -- 32 B x 2 range ends x L levels x executed search iterations (RLFM: 2L + 3
    # probes per lf_map2) -- divided by the kernel's measured duration.  The device layout needs fewer sectors
    # (1 per rank for Q4 / SYM) and the k-mer tables memoise the first iterations, so this can exceed the
    # streaming peak; the same count for the device layout, the DRAM bytes ncu measured (`traffic`) and the
    # L2-request rate against the measured random-request peak are reported beside it.
    Lw = int(mc).bit_length()
    P = index.sectors_per_rank()
    ref_per_lf2 = Lw if kind != RLFM else 2 * Lw + 3
    dev_per_lf2 = P + (3 if kind == RLFM else 0)
    peak, peak_src = measured_peaks()
    ms_search = ms_search_total / args.steps
    ms_locate = max(ms_locate_total / args.steps, 1e-9)
    alg_bytes_search = 32.0 * 2 * ref_per_lf2 * search_steps
    dev_bytes_search = 32.0 * 2 * dev_per_lf2 * search_steps
    alg_bytes_locate = 32.0 * ((Lw if kind != RLFM else Lw + 3) * lf_steps + hits)
    achieved = alg_bytes_search / (ms_search * 1e-3) / 1e9
    traffic, tsrc = ncu_traffic(args.workload, npat)
    beyond_l2 = index.heap_size() > (400 << 20)
    roofline = {"bound": "hbm", "kernel": "k_search", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes_search, "kernel_ms": ms_search,
                "definition": "SURVEY.md 8(d): 32 B x 2 range ends x L wavelet levels x executed search iterations "
                              "(RLFM: 2L+3 probes per lf_map2) / kernel duration (CUDA events on the launch stream)",
                "reference_sectors_per_lf_map2": ref_per_lf2,
                "device_layout": {"sectors_per_lf_map2": dev_per_lf2, "bytes_per_launch": dev_bytes_search,
                                  "achieved": dev_bytes_search / (ms_search * 1e-3) / 1e9,
                                  "frac": dev_bytes_search / (ms_search * 1e-3) / 1e9 / peak},
                "note": ("index larger than L2: bound by the rate of DRAM-missing L2 requests (see random_access), "
                         "not by bandwidth" if beyond_l2 else
                         "index fits the 126 MB L2: bound by instruction issue / L1TEX request rate, not HBM (see traffic)")
                        + "; frac > 1 means fewer bytes moved than the reference structure's sector model needs",
                "executed_search_steps": int(search_steps), "executed_lf_steps": int(lf_steps),
                "locate_kernel": {"achieved": alg_bytes_locate / (ms_locate * 1e-3) / 1e9, "phase_ms": ms_locate,
                                  "frac": alg_bytes_locate / (ms_locate * 1e-3) / 1e9 / peak}}
    if traffic:
        roofline["traffic_frac_of_peak"] = traffic / (ms_search * 1e-3) / 1e9 / peak
        roofline["traffic_source"] = tsrc.get("source")
    gp = None
    if not args.no_gather_peak:
        try:
            gp = fmx.random_gather_peak(local, nbytes=4 << 30, nloads=1 << 28, iters=3)
        except Exception as ex:  # pragma: no cover
            roofline["random_access"] = {"error": str(ex)}
    if gp:
        ra = {"peak_requests_per_s": gp,
              "how": "independent uniform-random 32 B loads over a 4 GiB buffer, best of 3 (every load is one "
                     "DRAM-missing L2 request; profiles/r01b_random_access_study.md)"}
        if tsrc and tsrc.get("k_search", {}).get("l2_read_requests"):
            req = tsrc["k_search"]["l2_read_requests"] * npat / tsrc["npat"]
            ra.update({"l2_read_requests_per_launch": req, "achieved_requests_per_s": req / (ms_search * 1e-3),
                       "frac": req / (ms_search * 1e-3) / gp,
                       "requests_source": "lts__t_requests_srcunit_tex_op_read.sum of the ncu capture, scaled by batch size"})
        roofline["random_access"] = ra

    # ---- CPU baseline beside it: the oracle port on all host threads, bounded sample; also the parity check
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import oracle as orc

        nthreads = host_threads()
        sample = min(npat, args.cpu_sample)
        tb = time.perf_counter()
        sa_src = "oracle SA-IS"
        if text.size >= (1 << 27) and not args.oracle_own_sa:
            # GB-scale texts: the oracle's single-threaded SA-IS takes ~5 min per GB of box time.  Hand it
            # the GPU-built suffix array instead; the oracle VERIFIES it with its own linear-time checker
            # (orc_check_suffix_array) before using it, and builds everything else itself.
            sa, _ = fmx.suffix_array_device(text, mc, local)
            oracle_index = orc.OracleIndex(text, kind, level=level, max_character=mc, sa=sa)
            del sa
            sa_src = "GPU-built SA, verified by the oracle's linear-time checker"
        else:
            oracle_index = orc.OracleIndex(text, kind, level=level, max_character=mc)
        obuild = time.perf_counter() - tb
        if m:
            s_flat = h_pat[:sample].numpy().reshape(-1)
            s_off = np.arange(sample + 1, dtype=np.uint64) * np.uint64(m)
        else:
            s_off = h_off[: sample + 1].numpy().astype(np.uint64)
            s_flat = h_pat.numpy()[: int(s_off[-1])]
        best = None
        for _ in range(2):
            dt, s, e, ohoff, opos = run_cpu_path(oracle_index, s_flat, s_off, nthreads)
            best = dt if best is None else min(best, dt)
        g_s = d_s[:sample].cpu().numpy().view(np.uint64)
        g_e = d_e[:sample].cpu().numpy().view(np.uint64)
        g_hoff = d_hoff[: sample + 1].cpu().numpy().view(np.uint64)
        g_pos = d_pos[0][: int(g_hoff[-1])].cpu().numpy().view(np.uint64)
        parity = bool(np.array_equal(g_s, s) and np.array_equal(g_e, e) and np.array_equal(g_hoff, ohoff)
                      and np.array_equal(g_pos, opos))
        # the end-to-end (host-buffer) call must agree too: its CSR offsets and positions for the same sample
        e_hoff = h_hoff[: sample + 1].numpy().view(np.uint64)
        e_pos = h_pos[: int(e_hoff[-1])].numpy().view(np.uint64)
        e2e_parity = bool(np.array_equal(e_hoff, ohoff) and np.array_equal(e_pos, opos))
        parity = parity and e2e_parity
        cpu = {"value": sample / best, "unit": "queries/s", "cores": nthreads, "kind": "port",
               "sample": f"first {sample} patterns of the workload, count+locate, best of 2",
               "located_hits_per_s": int(ohoff[-1]) / best, "cpu_model": cpu_model(),
               "index_build_s": round(obuild, 1), "index_suffix_array": sa_src,
               "gpu_matches_oracle_on_sample": parity, "e2e_call_matches_oracle_on_sample": e2e_parity}
        if not parity:
            print("PARITY FAILURE: GPU results differ from the oracle on the sample", file=sys.stderr)

    line = {
        "metric": "count+locate queries/s", "value": value, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32/u64 integer", "data": "synthetic",
        "config": config_of(args.workload, w, int(text.size), npat, {
            "index_device_bytes": index.heap_size(), "device_layout": index.layout_name(),
            "l2": "flushed between timed iterations (512 MiB memset)",
            "step": "fmx_search_batch_device + fmx_locate_batch_device" + (" replayed as one CUDA graph" if graph is not None else ""),
            "parallelism": f"index replicated x{world}, query batches sharded", "index_build_s": round(build_s, 1),
            **({"options": args.option} if args.option else {})}),
        "count_queries_per_s": world * npat / (ms_search_total / args.steps * 1e-3),
        "located_hits_per_s": hits_all / (ms_per_step * 1e-3),
        "hits_per_step": hits_all,
        "e2e": {"value": world * npat / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3, "api": "fmx_search_locate_batch: host patterns in, CSR hit offsets (counts) + positions out (pinned buffers; chunked H2D/kernel/D2H pipeline)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "wall_ms_per_step_incl_flush": wall / args.steps * 1e3,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def by_piece_bench(args, w, fmx, rank, world, local):
    """MultiPieces partitioned by piece: every rank indexes its pieces, answers every pattern, and the
    per-pattern counts and hit lists are combined with NCCL all_gathers (fm-index_b200/partitioned.py)."""
    import torch
    import torch.distributed as dist

    from fm_index_b200 import partitioned as part

    if w["kind"] != MULTI:
        raise SystemExit("--by-piece needs a MultiPieces workload (cfg4_multi*)")
    npat = args.npat or w["npat"]
    d_text = gen_text_for(w, device="cuda")
    d_pat, _ = gen_patterns(d_text, npat, w["m"], w["sigma"], 4)          # the SAME batch on every rank
    keep = (d_pat != 0).all(dim=1)                                          # \0 cannot cross the partition
    d_pat = d_pat[keep].contiguous()
    npat = int(d_pat.shape[0])
    text = d_text.cpu().numpy()
    del d_text
    torch.cuda.empty_cache()
    t0 = time.perf_counter()
    idx = part.PartitionedMultiPieces(text, w["level"], w["mc"], device=local)
    build_s = time.perf_counter() - t0
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    L = fmx.load_library()
    for _ in range(args.warmup):
        counts, hoff, pos, pid = idx.search_locate(d_pat)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = L.fmx_launch_count()
    sampler.start()
    for k in range(args.steps):
        ev[k][0].record(stream)
        counts, hoff, pos, pid = idx.search_locate(d_pat)
        ev[k][1].record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = L.fmx_launch_count() - launches0
    ms_total = float(sum(ev[k][0].elapsed_time(ev[k][1]) for k in range(args.steps)))
    if world > 1:
        tt = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    ms = ms_total / args.steps
    hits = int(hoff[-1].item())
    # property check: every gathered hit really holds its pattern, in the right piece
    ends = np.flatnonzero(text == 0)
    p_h, o_h = pos.cpu().numpy(), np.repeat(np.arange(npat), np.diff(hoff.cpu().numpy()))
    pats_h = d_pat.cpu().numpy()
    sel = np.random.default_rng(0).integers(0, max(hits, 1), min(hits, 200_000))
    got = text[p_h[sel][:, None] + np.arange(w["m"])[None, :]]
    ok = bool(np.array_equal(got, pats_h[o_h[sel]]) and
              np.array_equal(pid.cpu().numpy()[sel], np.searchsorted(ends, p_h[sel], side="left")))
    if rank == 0:
        line = {"metric": "count+locate queries/s", "value": npat / (ms * 1e-3), "unit": "queries/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "u32/u64 integer", "data": "synthetic",
                "config": config_of(args.workload, w, int(text.size), npat, {
                    "parallelism": f"pieces partitioned over {world} GPUs {idx.ranges}; every pattern answered by every "
                                   "rank; counts + hit lists combined by NCCL all_gather",
                    "index_build_s": round(build_s, 1)}),
                "located_hits_per_s": hits / (ms * 1e-3), "hits_per_step": hits, "gpu_launches": int(launches),
                "clocks": clocks, "gathered_hits_verified_against_text": ok}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
