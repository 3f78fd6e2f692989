"""One timed-step of a bench workload (fmx_query_batch_device: hit offsets + positions) between cudaProfilerStart/Stop, for ncu:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum \
        --clock-control none --profile-from-start off -o gpurun_out/step_X -f python tools/prof_step.py --workload X [--mode compact]
    (or --set full --import-source on for a source-level look at one kernel)

Index build, k-mer table build and warm-up run OUTSIDE the profiled range, so the report holds exactly the launches of
ONE step at the bench's own batch size.  Prints the step's own (un-profiled) timing, the source hash of the kernels and
the batch, which tools/ncu_traffic.py stores next to the counters so that bench.py can refuse a stale capture."""
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
ap.add_argument("--mode", default="auto", choices=["auto", "rich", "compact"])
ap.add_argument("--option", action="append", default=[], help="key=value index option (fmx_index_set_option)")
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
mode_id = {"auto": fmx.MODE_AUTO, "rich": fmx.MODE_RICH, "compact": fmx.MODE_COMPACT}[args.mode]
index = cls.new(fmx.Text.with_max_character(text, mc), level, device=0, mode=mode_id)
for kv in args.option:
    k, v = kv.split("=")
    index.set_option(k, int(v))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
run = bench.DeviceRun(fmx, L, index, d_pat, d_off, m, npat, stream)
run.size_outputs()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ms = []
for _ in range(3):
    flush.zero_()
    ev[0].record(stream)
    run.query()
    ev[1].record(stream)
    torch.cuda.synchronize()
    ms.append(ev[0].elapsed_time(ev[1]))
print(json.dumps({"workload": args.workload, "npat": npat, "hits": run.hits, "step_ms": min(ms),
                  "mode": {fmx.MODE_RICH: "rich", fmx.MODE_COMPACT: "compact"}[index.mode()], "source_hash": bench.source_hash(),
                  "index_bytes": index.heap_size(), "kmer_k": [index.kmer_k(False), index.kmer_k(True)],
                  "options": args.option}), flush=True)
flush.zero_()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run.query()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
