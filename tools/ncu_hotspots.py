"""Top SASS instructions by warp-stall samples for the kernels of an .ncu-rep (needs `--import-source on`;
run where ncu is installed, no GPU needed).
usage: python tools/ncu_hotspots.py gpurun_out/x.ncu-rep profiles/out.txt [top_n]"""
import csv
import subprocess
import sys


def main():
    rep, out = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    with open(out, "w") as f:
        f.write(f"# source: {rep} (ncu --set full --import-source on; --page source)\n")
        i = 0
        seen = set()
        while i < len(rows):
            if rows[i] and rows[i][0] == "Kernel Name":
                name = rows[i][1]
                hdr = rows[i + 1]
                j = i + 2
                body = []
                while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                    if len(rows[j]) == len(hdr):
                        body.append(rows[j])
                    j += 1
                si, ai = hdr.index("Source"), hdr.index("Address")
                st = hdr.index("Warp Stall Sampling (All Samples)")
                ie = hdr.index("Instructions Executed")
                at = hdr.index("Avg. Threads Executed")
                total = sum(int(r[st] or 0) for r in body) or 1
                insts = sum(int(r[ie] or 0) for r in body)
                if (name, insts, total) in seen:  # the page lists a kernel once per view
                    i = j
                    continue
                seen.add((name, insts, total))
                f.write(f"\n== {name}\n# {len(body)} SASS instructions, {insts} warp instructions executed, {total} stall samples\n")
                f.write("# share of stall samples | warp instructions executed | avg threads | SASS (offset)\n")
                base = int(body[0][ai], 16) if body else 0
                for r in sorted(body, key=lambda r: -int(r[st] or 0))[:top]:
                    f.write(f"{100.0 * int(r[st] or 0) / total:5.1f} %  {int(r[ie] or 0):>12}  {r[at]:>5}  {r[si].strip()}  (+0x{int(r[ai], 16) - base:x})\n")
                i = j
            else:
                i += 1


if __name__ == "__main__":
    main()
