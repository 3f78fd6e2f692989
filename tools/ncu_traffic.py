"""Rebuild profiles/ncu_traffic.json from ncu captures of tools/prof_step.py (run where ncu is installed; no GPU needed).

usage: python tools/ncu_traffic.py gpurun_out/step_X.ncu-rep:gpurun_out/step_X.log ...

Every capture holds exactly the launches of ONE step at the bench's batch size.  Per capture it stores, under the key
"<workload>:<index mode>": the per-step SUMS of DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and of L2 read
requests from the SMs (lts__t_requests_srcunit_tex_op_read.sum), the per-kernel breakdown (time under ncu, bytes, requests,
warp instructions, active lanes per instruction), the batch size and the SOURCE HASH of the kernels that were measured.
bench.py uses an entry only when hash, workload, batch size and mode all match its own run; nothing is ever scaled."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = {"dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
        "lts__t_requests_srcunit_tex_op_read.sum": "l2_read_requests", "lts__t_sectors_srcunit_tex_op_read.sum": "l2_read_sectors",
        "smsp__inst_executed.sum": "warp_instructions", "smsp__thread_inst_executed.sum": "thread_instructions",
        "gpu__time_duration.sum": "ns_under_ncu"}


def main():
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        out = json.load(open(path))
    except Exception:
        out = {}
    out = {k: v for k, v in out.items() if ":" in k}   # round-1 entries (scaled constants) are gone
    for arg in sys.argv[1:]:
        rep, log = arg.split(":")
        meta = {}
        for line in open(log):
            if line.startswith("{"):
                meta = json.loads(line)
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        h = rows[0]
        ki = h.index("Kernel Name")
        kernels, tot = [], {v: 0.0 for v in COLS.values()}
        for r in rows[2:]:
            d = {"kernel": r[ki].split("(")[0].replace("void ", "").replace("fmx::", "")}
            for col, key in COLS.items():
                if col in h:
                    d[key] = float(r[h.index(col)].replace(",", ""))
                    tot[key] += d[key]
            if d.get("warp_instructions"):
                d["active_lanes_per_instruction"] = round(d.get("thread_instructions", 0) / d["warp_instructions"], 2)
            kernels.append(d)
        entry = {"workload": meta.get("workload"), "mode": meta.get("mode"), "npat": meta.get("npat"), "hits": meta.get("hits"),
                 "source_hash": meta.get("source_hash"), "options": meta.get("options"),
                 "step_ms_unprofiled": meta.get("step_ms"),
                 "dram_bytes_per_step": tot["dram_read_bytes"] + tot["dram_write_bytes"],
                 "l2_read_requests_per_step": tot["l2_read_requests"], "warp_instructions_per_step": tot["warp_instructions"],
                 "active_lanes_per_instruction": round(tot["thread_instructions"] / max(1.0, tot["warp_instructions"]), 2),
                 "kernels": kernels,
                 "source": f"{os.path.basename(rep)}: ncu --clock-control none --profile-from-start off, one step of tools/prof_step.py"}
        out[f"{meta.get('workload')}:{meta.get('mode')}"] = entry
    json.dump(out, open(path, "w"), indent=1)
    for k, v in out.items():
        print(k, v["npat"], v["source_hash"], f"{v['dram_bytes_per_step'] / 1e9:.2f} GB", f"{v['l2_read_requests_per_step'] / 1e6:.1f} M requests",
              v["active_lanes_per_instruction"], "lanes")


if __name__ == "__main__":
    main()
