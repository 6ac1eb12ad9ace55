"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export by source line: stall samples and executed warp
instructions, top N lines.  usage: python profiles/hot_lines.py report.ncu-rep kernel_regex [N]"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    cur_file, hdr, agg = None, None, {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            i_s, i_n = hdr.index("# Samples"), hdr.index("Instructions Executed")
            try:
                agg[(cur_file, int(r[0]))] = (int(r[i_s]), int(r[i_n]), r[1].strip()[:110])
            except ValueError:
                pass
    tot_s = sum(v[0] for v in agg.values()) or 1
    tot_i = sum(v[1] for v in agg.values()) or 1
    print(f"kernel {kern}: {tot_s} stall samples, {tot_i} warp instructions over {len(agg)} source lines")
    print("| file:line | samples % | instr % | source |\n|---|---|---|---|")
    for (f, l), (s, n, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"| {f}:{l} | {100 * s / tot_s:.1f} | {100 * n / tot_i:.1f} | `{src}` |")


if __name__ == "__main__":
    main()
