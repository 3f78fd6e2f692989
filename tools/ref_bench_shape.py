"""The reference's OWN criterion bench shape (benches/common.rs:5-27, benches/count.rs:23-26, benches/locate.rs:32-35):
text = 50 000 symbols from {'0','1'} with P('0') = prob, plus \\0, Text::with_max_character(text, b'1') (6 wavelet
levels); patterns = all 256 binary strings of length 8; throughput unit = patterns.  Timed here for the CPU port
(oracle/, ONE thread, as criterion runs) so that it can be read beside the crate's published figures (CHANGES.md:45-88,
BASELINE.md section 1) -- the sanity anchor of every `cpu_baseline` in the bench lines -- and, when a GPU is present, for
the CUDA path through the C ABI on the same inputs (results compared).

    python tools/ref_bench_shape.py [--reps 200]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc

PUBLISHED = {  # patterns/s, default build / -C target-cpu=native (CHANGES.md:50-61, :77-88; hardware unstated, 1 thread)
    ("count", "FM", 0.005): (3.6e6, 5.2e6), ("count", "FM", 0.05): (2.1e6, 3.2e6), ("count", "FM", 0.5): (2.1e6, 2.8e6),
    ("count", "RLFM", 0.005): (1.0752e6, 2.1e6), ("count", "RLFM", 0.05): (586.3e3, 1.2269e6), ("count", "RLFM", 0.5): (506.1e3, 988.8e3),
    ("locate", "FM", 1): (79.1e3, 93.6e3), ("locate", "FM", 2): (30.1e3, 35.2e3), ("locate", "FM", 3): (13.8e3, 16.0e3),
    ("locate", "RLFM", 1): (28.0e3, 48.5e3), ("locate", "RLFM", 2): (10.0e3, 17.7e3), ("locate", "RLFM", 3): (4.3e3, 7.8e3),
}


def text_of(prob, seed):
    rng = np.random.default_rng(seed)
    body = np.where(rng.random(50_000) < prob, ord("0"), ord("1")).astype(np.uint8)
    return np.append(body, np.uint8(0))


def patterns():
    pats = np.empty((256, 8), dtype=np.uint8)
    for v in range(256):
        for j in range(8):
            pats[v, j] = ord("1") if (v >> j) & 1 else ord("0")
    return pats


def best_of(fn, reps):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=200)
    args = ap.parse_args()
    pats = patterns()
    flat, off = pats.reshape(-1), np.arange(257, dtype=np.uint64) * 8
    gpu = None
    try:
        import torch

        if torch.cuda.is_available():
            import fmx_pkg

            gpu = fmx_pkg.load()
    except Exception:
        gpu = None
    for kind_name, kind in (("FM", orc.FM), ("RLFM", orc.RLFM)):
        for prob in (0.005, 0.05, 0.5):
            text = text_of(prob, 1)
            ix = orc.OracleIndex(text, kind, level=None, max_character=ord("1"))
            s, e = ix.search_batch(flat, off, nthreads=1)
            dt = best_of(lambda: ix.search_batch(flat, off, nthreads=1), args.reps)
            pub = PUBLISHED[("count", kind_name, prob)]
            line = {"bench": "count", "index": kind_name + "Index", "prob": prob, "cpu_port_patterns_per_s_1_thread": 256 / dt,
                    "crate_published_patterns_per_s": {"default_build": pub[0], "target_cpu_native": pub[1]},
                    "port_over_crate_default_build": 256 / dt / pub[0], "matches": int((e - s).sum())}
            if gpu:
                cls = gpu.FMIndex if kind == orc.FM else gpu.RLFMIndex
                gi = cls.new(gpu.Text.with_max_character(text, ord("1")))
                b = gi.search_batch(pats)
                line["gpu_equals_port"] = bool(np.array_equal(b.s, s) and np.array_equal(b.e, e))
            print(json.dumps(line), flush=True)
        text = text_of(0.5, 1)
        for level in (1, 2, 3):
            ix = orc.OracleIndex(text, kind, level=level, max_character=ord("1"))
            s, e = ix.search_batch(flat, off, nthreads=1)

            def run():
                s_, e_ = ix.search_batch(flat, off, nthreads=1)
                return ix.locate_batch(s_, e_, nthreads=1)

            hoff, pos, _ = run()
            dt = best_of(run, max(5, args.reps // 10))
            pub = PUBLISHED[("locate", kind_name, level)]
            line = {"bench": "locate", "index": kind_name + "IndexWithLocate", "level": level, "cpu_port_patterns_per_s_1_thread": 256 / dt,
                    "cpu_port_hits_per_s_1_thread": int(hoff[-1]) / dt,
                    "crate_published_patterns_per_s": {"default_build": pub[0], "target_cpu_native": pub[1]},
                    "port_over_crate_default_build": 256 / dt / pub[0], "hits": int(hoff[-1])}
            if gpu:
                cls = gpu.FMIndexWithLocate if kind == orc.FM else gpu.RLFMIndexWithLocate
                gi = cls.new(gpu.Text.with_max_character(text, ord("1")), level)
                r = gi.query_batch(pats, capacity=int(hoff[-1]) + 8)
                line["gpu_equals_port"] = bool(np.array_equal(r["hit_off"], hoff) and np.array_equal(r["positions"], pos))
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
