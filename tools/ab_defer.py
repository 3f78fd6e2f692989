"""A/B of the fused query kernel's variants (option fused_defer, phased.cuh) on ONE index and ONE device-resident batch:

    python tools/ab_defer.py [--workload target_dna1g] [--variants 4,3,5,6,7,0] [--per-sm 64,1,2,4] [--steps 10]

--per-sm: grid sizes as blocks per SM; 64 = the default cap, 1 / 2 / 4 = that many times the variant's RESIDENT blocks.

Per variant: the bench's timed step (fmx_query_batch_device: hit offsets + positions, L2 flushed between iterations,
CUDA events on the launching stream) and whether its hit offsets and positions equal the first variant's over the
WHOLE batch.  Prints one JSON line per variant and a last line naming the fastest one."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import fmx_pkg

NAMES = {0: "k_query_fused (no second pass)", 1: "default instance", 2: "second pass always", 3: "6 blocks/SM, look every round",
         4: "5 blocks/SM, look every round", 5: "5 blocks/SM, look every 2 rounds", 6: "6 blocks/SM, look every 2 rounds",
         7: "5 blocks/SM, look every 4 rounds"}

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="target_dna1g")
ap.add_argument("--npat", type=int, default=0)
ap.add_argument("--variants", default="4,3,5,6,7,0")
ap.add_argument("--per-sm", default="64,1,2,4")
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
sms = torch.cuda.get_device_properties(0).multi_processor_count
RESIDENT = {0: 5, 1: 5, 2: 5, 3: 6, 4: 5, 5: 5, 6: 6, 7: 5}
for v in [int(x) for x in args.variants.split(",")]:
  for mult in [int(x) for x in args.per_sm.split(",")]:
    per_sm = 64 if mult == 64 else mult * RESIDENT[v]
    index.set_option("fused_defer", v)
    index.set_option("query_blocks", 0 if mult == 64 else per_sm * sms)
    for _ in range(3):
        flush.zero_()
        run.query()
    torch.cuda.synchronize()
    run.d_hoff.zero_()
    run.d_pos.zero_()
    ms, _, _ = run.timed(run.query, args.steps, flush, graph=False)
    same = None
    if ref_off is None:
        ref_off, ref_pos = run.d_hoff.clone(), run.d_pos[: run.hits].clone()
    else:
        same = bool(torch.equal(run.d_hoff, ref_off)) and bool(torch.equal(run.d_pos[: run.hits], ref_pos))
    r = {"fused_defer": v, "blocks_per_sm": per_sm, "what": NAMES.get(v), "ms_per_step": sum(ms) / len(ms), "ms_min": min(ms), "ms_max": max(ms),
         "queries_per_s": npat / (sum(ms) / len(ms) * 1e-3), "same_offsets_and_positions_as_first_variant": same,
         "workload": args.workload, "npat": npat, "hits": run.hits, "entry_bytes": int(L.fmx_index_kmer_entry_bytes(index._h))}
    results.append(r)
    print(json.dumps(r), flush=True)
ok = [r for r in results if r["same_offsets_and_positions_as_first_variant"] in (None, True)]
best = min(ok, key=lambda r: r["ms_per_step"])
print(json.dumps({"best": best["fused_defer"], "best_blocks_per_sm": best["blocks_per_sm"], "best_ms": best["ms_per_step"],
                  "all_equal": len(ok) == len(results)}), flush=True)
