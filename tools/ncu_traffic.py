"""Rebuild profiles/ncu_traffic.json from the `ncu --set full` captures of tools/prof_step.py
(run where ncu is installed; no GPU needed).

usage: python tools/ncu_traffic.py workload=gpurun_out/prof_X.ncu-rep:gpurun_out/prof_X.log ...

Per workload it records, for the k_search and k_locate launches of ONE step: measured DRAM bytes
(dram__bytes_read.sum + dram__bytes_write.sum), L2 read requests from the SMs
(lts__t_requests_srcunit_tex_op_read.sum) and the duration under ncu, together with the batch size
of the capture so that bench.py can scale them to its own batch."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {"dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
        "lts__t_requests_srcunit_tex_op_read.sum": "l2_read_requests",
        "lts__t_sectors_srcunit_tex_op_read.sum": "l2_read_sectors",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "l1_load_sectors",
        "smsp__inst_executed.sum": "warp_instructions", "gpu__time_duration.sum": "ns_under_ncu"}


def main():
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        out = json.load(open(path))
    except Exception:
        out = {}
    for arg in sys.argv[1:]:
        name, rest = arg.split("=")
        rep, log = rest.split(":")
        meta = {}
        for line in open(log):
            if line.startswith("{"):
                meta = json.loads(line)
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True,
                             text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        h = rows[0]
        ki = h.index("Kernel Name")
        entry = {"npat": meta.get("npat"), "hits": meta.get("hits"), "source": f"{os.path.basename(rep)} (ncu --set full, one step of tools/prof_step.py)",
                 "capture": meta}
        for r in rows[2:]:
            kname = "k_search" if "k_search" in r[ki] else ("k_locate" if "k_locate" in r[ki] else None)
            if not kname:
                continue
            d = {}
            for col, key in WANT.items():
                if col in h:
                    d[key] = float(r[h.index(col)].replace(",", ""))
            d["dram_bytes"] = d.get("dram_read_bytes", 0) + d.get("dram_write_bytes", 0)
            entry[kname] = d
        # keys bench.py has always read
        if "k_search" in entry:
            entry["k_search_dram_bytes"] = entry["k_search"]["dram_bytes"]
            entry["k_search_ms_under_ncu"] = entry["k_search"]["ns_under_ncu"] / 1e6
        out[name] = entry
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps({k: {kk: vv for kk, vv in v.items() if kk in ("npat", "k_search_dram_bytes", "k_search_ms_under_ncu")} for k, v in out.items()}, indent=1))


if __name__ == "__main__":
    main()
