"""A/B of ONE tuning option (fmx_index_set_option) on one index and one device-resident batch:

    python tools/ab_option.py --option emit_fused --values 0,1 [--workload target_dna1g] [--steps 10]

Per value: the bench's timed step (fmx_query_batch_device: hit offsets + positions, L2 flushed between iterations, CUDA
events on the launching stream) and whether hit offsets and positions equal the first value's over the WHOLE batch."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import fmx_pkg

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="target_dna1g")
ap.add_argument("--npat", type=int, default=0)
ap.add_argument("--option", required=True)
ap.add_argument("--values", default="0,1")
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
fmx = fmx_pkg.load()
L = fmx.load_library()
w = bench.WORKLOADS[args.workload]
npat = args.npat or w["npat"]
kind, mc, level, m = w["kind"], w["mc"], w["level"], w["m"]
torch.cuda.set_device(0)
d_text = bench.gen_text_for(w, device="cuda")
if m:
    d_pat, _ = bench.gen_patterns(d_text, npat, m, w["sigma"], 4, all_sampled=kind == bench.RLFM)
    d_off = None
else:
    d_pat, d_off = bench.gen_ragged_patterns(d_text, npat, w["sigma"], 4)
text = d_text.cpu().numpy()
del d_text
torch.cuda.empty_cache()
cls = [fmx.FMIndexWithLocate, fmx.RLFMIndexWithLocate, fmx.FMIndexMultiPiecesWithLocate][kind]
index = cls.new(fmx.Text.with_max_character(text, mc), level, device=0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
run = bench.DeviceRun(fmx, L, index, d_pat, d_off, m, npat, stream)
run.size_outputs()
ref_off = ref_pos = None
results = []
for rep in range(2):                       # every value twice, interleaved: drift shows up as a difference between the passes
    for v in [int(x) for x in args.values.split(",")]:
        index.set_option(args.option, v)
        for _ in range(3):
            flush.zero_()
            run.query()
        torch.cuda.synchronize()
        run.d_hoff.zero_()
        run.d_pos.zero_()
        ms, _, _ = run.timed(run.query, args.steps, flush, graph=False)
        phase = run.phases(run.query, flush)
        same = None
        if ref_off is None:
            ref_off, ref_pos = run.d_hoff.clone(), run.d_pos[: run.hits].clone()
        else:
            same = bool(torch.equal(run.d_hoff, ref_off)) and bool(torch.equal(run.d_pos[: run.hits], ref_pos))
        r = {"option": args.option, "value": v, "pass": rep, "ms_per_step": sum(ms) / len(ms), "ms_min": min(ms), "ms_max": max(ms),
             "queries_per_s": npat / (sum(ms) / len(ms) * 1e-3), "same_offsets_and_positions_as_first": same,
             "phase_ms": [round(float(x), 4) for x in phase], "workload": args.workload, "npat": npat, "hits": run.hits}
        results.append(r)
        print(json.dumps(r), flush=True)
ok = all(r["same_offsets_and_positions_as_first"] in (None, True) for r in results)
by = {}
for r in results:
    by.setdefault(r["value"], []).append(r["ms_per_step"])
mean = {v: sum(x) / len(x) for v, x in by.items()}
best = min(mean, key=mean.get)
print(json.dumps({"option": args.option, "mean_ms": mean, "best": best, "all_equal": ok}), flush=True)
