"""Extract a curated per-kernel summary from an .ncu-rep (run where ncu is installed; no GPU needed).
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/out.txt ["header line"]"""
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__sectors_read.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
]
PREFIX = ("smsp__average_warps_issue_stalled", "smsp__average_warp_latency_issue_stalled", "smsp__warp_issue_stalled")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    head = sys.argv[3] if len(sys.argv) > 3 else ""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, u = rows[0], rows[1]
    name_i = h.index("Kernel Name")
    with open(out, "w") as f:
        if head:
            f.write("# " + head + "\n")
        f.write(f"# source: {rep} (ncu --set full --clock-control none)\n")
        for v in rows[2:]:
            f.write(f"\n== {v[name_i]}\n")
            stalls = []
            for i, n in enumerate(h):
                if n in KEEP:
                    f.write(f"{n} = {v[i]} {u[i]}\n")
                elif n.startswith(PREFIX) and n.endswith(".ratio") and "not_issued" not in n:
                    try:
                        stalls.append((float(v[i].replace(",", "")), n))
                    except ValueError:
                        pass
            for val, n in sorted(stalls, reverse=True)[:8]:
                f.write(f"{n} = {val}\n")


if __name__ == "__main__":
    main()
