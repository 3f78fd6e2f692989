"""One search+locate step of a bench workload between cudaProfilerStart/Stop, for ncu:

    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -k regex:'k_search|k_locate' -o gpurun_out/prof_X python tools/prof_step.py --workload X [--npat N]

Index build, k-mer table build and warm-up run OUTSIDE the profiled range, so the report holds exactly
the launches of one step.  Prints the step's own (un-profiled) timings first for reference."""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import fmx_pkg

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2_dna100m")
ap.add_argument("--npat", type=int, default=0)
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
index = cls.new(fmx.Text.with_max_character(text, mc), level, device=0)
for kv in args.option:
    k, v = kv.split("=")
    index.set_option(k, int(v))
h = index._h
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
sp = C.c_void_p(stream.cuda_stream)
d_s = torch.empty(npat, dtype=torch.int64, device="cuda")
d_e = torch.empty(npat, dtype=torch.int64, device="cuda")
d_hoff = torch.empty(npat + 1, dtype=torch.int64, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
p_off = d_off.data_ptr() if d_off is not None else None
total = C.c_uint64(0)


def chk(rc):
    if rc != 0:
        raise RuntimeError(L.fmx_last_error().decode())


def search():
    chk(L.fmx_search_batch_device(h, 0, d_pat.data_ptr(), p_off, m, npat, None, None, d_s.data_ptr(), d_e.data_ptr(), sp))


search()
chk(L.fmx_locate_count_device(h, 0, d_s.data_ptr(), d_e.data_ptr(), npat, d_hoff.data_ptr(), C.byref(total), sp))
cap = int(total.value) + 1024
d_pos = torch.empty(cap, dtype=torch.int64, device="cuda")


def locate():
    chk(L.fmx_locate_batch_device(h, 0, d_s.data_ptr(), d_e.data_ptr(), npat, d_hoff.data_ptr(), d_pos.data_ptr(),
                                  None, cap, sp))


ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for _ in range(3):
    flush.zero_()
    ev[0].record(stream)
    search()
    ev[1].record(stream)
    locate()
    ev[2].record(stream)
torch.cuda.synchronize()
work = index.last_work(sp)
print(json.dumps({"workload": args.workload, "npat": npat, "hits": int(total.value), "search_ms": ev[0].elapsed_time(ev[1]),
                  "locate_ms": ev[1].elapsed_time(ev[2]), "search_steps": int(work[0]), "lf_steps": int(work[1]),
                  "index_bytes": index.heap_size(), "sectors_per_rank": index.sectors_per_rank(),
                  "kmer_k": [int(L.fmx_index_kmer_k(h, 0)), int(L.fmx_index_kmer_k(h, 1))]}), flush=True)
flush.zero_()
torch.cuda.synchronize()
torch.cuda.profiler.start()
search()
locate()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
