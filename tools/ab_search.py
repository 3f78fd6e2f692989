"""A/B timing of search-kernel variants on one index (run on the GPU box).
usage: python tools/ab_search.py [--workload cfg2_dna100m] [--npat N] [--reps 5]"""
import argparse, ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, fmx_pkg

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2_dna100m")
ap.add_argument("--npat", type=int, default=0)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--variants", default="v1,v1k,v1kK")
ap.add_argument("--budgets", default="")
ap.add_argument("--bps", default="")
ap.add_argument("--gran", default="")
args = ap.parse_args()
fmx = fmx_pkg.load(); L = fmx.load_library()
w = bench.WORKLOADS[args.workload]
n, m, sigma, mc, level = w["n"], w["m"], w["sigma"], w["mc"], w["level"]
npat = args.npat or w["npat"]
torch.cuda.set_device(0)
d_text = bench.gen_text_for(w, device="cuda")
d_pat, _ = bench.gen_patterns(d_text, npat, m, sigma, 4)
text = d_text.cpu().numpy(); del d_text
t0 = time.time()
index = fmx.FMIndexWithLocate.new(fmx.Text.with_max_character(text, mc), level)
print("build s", round(time.time() - t0, 1), flush=True)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); sp = C.c_void_p(st.cuda_stream)
d_s = torch.empty(npat, dtype=torch.int64, device="cuda"); d_e = torch.empty_like(d_s)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
ref = None
rows = []
variants = [(v, None) for v in args.variants.split(",")]
grans = [int(g) for g in args.gran.split(",")] if args.gran else [None]
if args.bps:
    variants += [("v2k", int(b)) for b in args.bps.split(",")]
for name, bps, gran in [(n_, b_, g_) for g_ in grans for (n_, b_) in variants]:
    if gran:
        index.set_option("l2_fetch_granularity", gran)
    index.set_option("search_persistent", name.startswith("v2"))
    index.set_option("kmer", "k" in name)
    index.set_option("kmer_big", "K" in name)
    index.set_option("bucket", 1 if name.endswith("b") else 0)
    if bps:
        index.set_option("persist_blocks_per_sm", bps)
    ts = []
    for r in range(args.reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        rc = L.fmx_search_batch_device(index._h, 0, d_pat.data_ptr(), None, m, npat, None, None, d_s.data_ptr(), d_e.data_ptr(), sp)
        assert rc == 0, L.fmx_last_error()
        e1.record(st); torch.cuda.synchronize()
        if r >= 2: ts.append(e0.elapsed_time(e1))
    steps, _ = index.last_work(sp)
    cur = (d_s.clone(), d_e.clone())
    same = True if ref is None else bool(torch.equal(cur[0], ref[0]) and torch.equal(cur[1], ref[1]))
    if ref is None: ref = cur
    row = {"variant": name, "blocks_per_sm": bps, "l2_fetch_granularity": gran, "ms_best": min(ts), "ms_mean": float(np.mean(ts)),
           "Gq_per_s": npat / min(ts) / 1e6, "steps": steps, "identical_to_first": same}
    rows.append(row); print(row, flush=True)
for mb in [int(x) for x in args.budgets.split(",") if x]:
    t0 = time.time(); index.set_option("kmer_budget_mb", mb); torch.cuda.synchronize(); tb = time.time() - t0
    index.set_option("search_persistent", 0); index.set_option("kmer", 1); index.set_option("kmer_big", 1); index.set_option("bucket", 0)
    ts = []
    for r in range(args.reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        rc = L.fmx_search_batch_device(index._h, 0, d_pat.data_ptr(), None, m, npat, None, None, d_s.data_ptr(), d_e.data_ptr(), sp)
        assert rc == 0, L.fmx_last_error()
        e1.record(st); torch.cuda.synchronize()
        if r >= 2: ts.append(e0.elapsed_time(e1))
    steps, _ = index.last_work(sp)
    same = bool(torch.equal(d_s, ref[0]) and torch.equal(d_e, ref[1]))
    row = {"variant": "budget", "kmer_budget_mb": mb, "big_k": index.kmer_k(True), "table_build_s": round(tb, 2), "ms_best": min(ts),
           "Gq_per_s": npat / min(ts) / 1e6, "steps": steps, "identical_to_first": same, "device_bytes": index.heap_size()}
    rows.append(row); print(row, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open(f"gpurun_out/ab_search_{args.workload}.json", "w"), indent=1)
