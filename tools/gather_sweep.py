"""Random-gather roofline sweep: sectors/s vs buffer size and load width (run on the GPU box)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fmx_pkg
fmx = fmx_pkg.load()
rows = []
for mb in (32, 96, 256, 1024, 4096, 16384):
    for width in (32, 64, 128):
        r = fmx.random_gather_peak(0, nbytes=mb << 20, nloads=1 << 27, iters=3, load_bytes=width)
        rows.append({"buffer_MiB": mb, "load_bytes": width, "gsectors_per_s": r / 1e9, "GBps": r * 32 / 1e9})
        print(rows[-1], flush=True)
json.dump(rows, open("gpurun_out/gather_sweep.json", "w"), indent=1)
